// nudge_b200 — state serialisation: the caller-owned PODs of nudge.h:73-129 (BodyData, ColliderData, BodyConnections,
// ContactCache; widened layout of include/nudge_b200.h) dumped from / loaded into the device-resident context as one flat file.
// Doubles as a checkpoint and makes a parity failure reproducible: tests dump the state they disagree on and tools/nb_replay steps it
// headless.  Included at the end of nb_api.cu.
//
// File layout (little endian): header { char magic[8] = "NBSTATE1"; u32 version = 1; u32 bodies, boxes, spheres, connections, cache;
// u32 solver_mode; u32 reserved[8]; } followed by the arrays in this order, each padded to 16 bytes:
//   transforms[bodies] (32 B) | properties[bodies] (16) | momentum[bodies] (32) | idle[bodies] (1)
//   box_tags[boxes] (4) | box_data[boxes] (16) | box_transforms[boxes] (32)
//   sphere_tags[spheres] (4) | sphere_data[spheres] (4) | sphere_transforms[spheres] (32)
//   connections[connections] (8) | cache_tags[cache] (8) | cache_features[cache] (4) | cache_data[cache] (16)
#pragma once

struct NbStateHeader { char magic[8]; u32 version, bodies, boxes, spheres, connections, cache, solver_mode, reserved[8]; };

static bool state_put(FILE* f, const void* dev, size_t bytes, std::vector<unsigned char>& tmp, cudaStream_t st) {
	const size_t padded = (bytes + 15) & ~(size_t)15;
	tmp.assign(padded, 0);
	if (bytes && cudaMemcpyAsync(tmp.data(), dev, bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess) return false;
	if (cudaStreamSynchronize(st) != cudaSuccess) return false;
	return padded == 0 || fwrite(tmp.data(), 1, padded, f) == padded;
}
static bool state_get(FILE* f, void* dev, size_t bytes, std::vector<unsigned char>& tmp, cudaStream_t st) {
	const size_t padded = (bytes + 15) & ~(size_t)15;
	tmp.resize(padded);
	if (padded && fread(tmp.data(), 1, padded, f) != padded) return false;
	if (bytes && cudaMemcpyAsync(dev, tmp.data(), bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) return false;
	return cudaStreamSynchronize(st) == cudaSuccess;
}

extern "C" {

int nb_save_state(nb_context* ctx, const char* path, void* stream) {
	cudaStream_t st = (cudaStream_t)stream;
	{ int jr = join_uploads(ctx, st); if (jr) return jr; }
	u32 c[CNT__COUNT];
	int r = get_counts(ctx, c, stream); if (r) return r;
	FILE* f = fopen(path, "wb");
	if (!f) { ctx->error = std::string("cannot open ") + path; return NB_ERR_ARGUMENT; }
	NbStateHeader h; memset(&h, 0, sizeof(h));
	memcpy(h.magic, "NBSTATE1", 8); h.version = 1;
	h.bodies = ctx->B; h.boxes = ctx->nboxes; h.spheres = ctx->nspheres; h.connections = ctx->nconn; h.cache = c[CNT_CACHE]; h.solver_mode = (u32)ctx->solver_mode;
	std::vector<unsigned char> tmp;
	bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
	ok = ok && state_put(f, ctx->xf, sizeof(nb_transform) * (size_t)h.bodies, tmp, st) && state_put(f, ctx->props, sizeof(nb_body_properties) * (size_t)h.bodies, tmp, st)
		&& state_put(f, ctx->mom, sizeof(nb_body_momentum) * (size_t)h.bodies, tmp, st) && state_put(f, ctx->idle, h.bodies, tmp, st)
		&& state_put(f, ctx->box_tags, 4 * (size_t)h.boxes, tmp, st) && state_put(f, ctx->box_data, sizeof(nb_box_collider) * (size_t)h.boxes, tmp, st) && state_put(f, ctx->box_xf, sizeof(nb_transform) * (size_t)h.boxes, tmp, st)
		&& state_put(f, ctx->sph_tags, 4 * (size_t)h.spheres, tmp, st) && state_put(f, ctx->sph_data, sizeof(nb_sphere_collider) * (size_t)h.spheres, tmp, st) && state_put(f, ctx->sph_xf, sizeof(nb_transform) * (size_t)h.spheres, tmp, st)
		&& state_put(f, ctx->conn, sizeof(nb_body_pair) * (size_t)h.connections, tmp, st)
		&& state_put(f, ctx->cache_tags, 8 * (size_t)h.cache, tmp, st) && state_put(f, ctx->cache_features, 4 * (size_t)h.cache, tmp, st) && state_put(f, ctx->cache_data, 16 * (size_t)h.cache, tmp, st);
	ok = (fclose(f) == 0) && ok;
	if (!ok) { ctx->error = std::string("write failed: ") + path; return NB_ERR_CUDA; }
	return NB_OK;
}

// Sizes in a state file (to size an nb_config before nb_create): counts[0..4] = bodies, boxes, spheres, connections, cache entries.
int nb_state_info(const char* path, uint32_t counts[5]) {
	FILE* f = fopen(path, "rb");
	if (!f) return NB_ERR_ARGUMENT;
	NbStateHeader h;
	bool ok = fread(&h, sizeof(h), 1, f) == 1 && !memcmp(h.magic, "NBSTATE1", 8) && h.version == 1;
	fclose(f);
	if (!ok) return NB_ERR_ARGUMENT;
	counts[0] = h.bodies; counts[1] = h.boxes; counts[2] = h.spheres; counts[3] = h.connections; counts[4] = h.cache;
	return NB_OK;
}

int nb_load_state(nb_context* ctx, const char* path, void* stream) {
	cudaStream_t st = (cudaStream_t)stream;
	{ int jr = join_uploads(ctx, st); if (jr) return jr; }
	FILE* f = fopen(path, "rb");
	if (!f) { ctx->error = std::string("cannot open ") + path; return NB_ERR_ARGUMENT; }
	NbStateHeader h;
	if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "NBSTATE1", 8) || h.version != 1) { fclose(f); ctx->error = "not an nb state file"; return NB_ERR_ARGUMENT; }
	const nb_config& c = ctx->cfg;
	if (h.bodies > c.max_bodies || h.boxes > c.max_boxes || h.spheres > c.max_spheres || h.connections > c.max_connections || h.cache > c.max_contacts) {
		fclose(f); ctx->error = "state file exceeds the context's capacities"; return NB_ERR_CAPACITY;
	}
	std::vector<unsigned char> tmp;
	std::vector<u32> tags;
	bool ok = state_get(f, ctx->xf, sizeof(nb_transform) * (size_t)h.bodies, tmp, st) && state_get(f, ctx->props, sizeof(nb_body_properties) * (size_t)h.bodies, tmp, st)
		&& state_get(f, ctx->mom, sizeof(nb_body_momentum) * (size_t)h.bodies, tmp, st) && state_get(f, ctx->idle, h.bodies, tmp, st);
	u32 maxtag = 1;
	ok = ok && state_get(f, ctx->box_tags, 4 * (size_t)h.boxes, tmp, st);
	if (ok) for (u32 i = 0; i < h.boxes; ++i) maxtag = std::max(maxtag, reinterpret_cast<const u32*>(tmp.data())[i]);
	ok = ok && state_get(f, ctx->box_data, sizeof(nb_box_collider) * (size_t)h.boxes, tmp, st) && state_get(f, ctx->box_xf, sizeof(nb_transform) * (size_t)h.boxes, tmp, st)
		&& state_get(f, ctx->sph_tags, 4 * (size_t)h.spheres, tmp, st);
	if (ok) for (u32 i = 0; i < h.spheres; ++i) maxtag = std::max(maxtag, reinterpret_cast<const u32*>(tmp.data())[i]);
	ok = ok && state_get(f, ctx->sph_data, sizeof(nb_sphere_collider) * (size_t)h.spheres, tmp, st) && state_get(f, ctx->sph_xf, sizeof(nb_transform) * (size_t)h.spheres, tmp, st)
		&& state_get(f, ctx->conn, sizeof(nb_body_pair) * (size_t)h.connections, tmp, st)
		&& state_get(f, ctx->cache_tags, 8 * (size_t)h.cache, tmp, st) && state_get(f, ctx->cache_features, 4 * (size_t)h.cache, tmp, st) && state_get(f, ctx->cache_data, 16 * (size_t)h.cache, tmp, st);
	fclose(f);
	if (!ok) { ctx->error = std::string("read failed: ") + path; return NB_ERR_CUDA; }
	ctx->B = h.bodies; ctx->nboxes = h.boxes; ctx->nspheres = h.spheres; ctx->nconn = h.connections;
	ctx->tagbits = bits_for((u64)maxtag + 1);
	ctx->kbits = bits_for(std::max(1u, h.boxes + h.spheres));
	CK(cudaMemcpyAsync(ctx->counts + CNT_CACHE, &h.cache, 4, cudaMemcpyHostToDevice, st));
	CK(cudaStreamSynchronize(st));
	if (h.solver_mode == NB_SOLVER_PARITY || h.solver_mode == NB_SOLVER_THROUGHPUT) ctx->solver_mode = (int)h.solver_mode;
	return NB_OK;
}

}  // extern "C"
