// nudge_b200 — user constraint rows on the device.
//
// The reference has no joints; it marks the three places where an application adds its own (example/main.cpp:285 "Custom constraints
// should be added as body connections", :288 "Custom contacts can be added here", :316 "Custom constraint impulses should be applied
// here", i.e. after every apply_impulses sweep).  With the state resident in HBM the application cannot run that code on the host
// between two sweeps without a round trip per sweep, so the hook becomes data: an array of generic velocity-constraint ROWS
// (one scalar constraint each: Jacobian, bias, impulse bounds, accumulated impulse) that the library applies, in the caller's
// order, right after every contact sweep — sequential impulses exactly as a host loop at example/main.cpp:316 would do them:
//
//     eff    = 1 / (J M^-1 J^T + softness)              M^-1 from BodyProperties, rotated like nudge.cpp:4182-4199
//     delta  = -eff * (J v + bias + softness * impulse)
//     impulse' = clamp(impulse + delta, lo, hi);  v += M^-1 J^T (impulse' - impulse)
//
// Order: rows run in upload order wherever two rows share a body (the host assigns every row the level 1 + max(level of the
// previous row on either body); rows of one level touch disjoint bodies and run in parallel, levels run one after the other), so the
// result equals the sequential loop.  Islands: connect the bodies of a joint with nb_upload_connections (example/main.cpp:285).
// Included at the end of nb_api.cu.
#pragma once
#include <map>

struct RowsDev { nb_constraint_row* rows; u32* level_off; u32 n, levels; };

__global__ void __launch_bounds__(NB_BLOCK) k_user_rows(nb_constraint_row* rows, u32 begin, u32 end, nb_body_momentum* momentum, const nb_body_properties* props, const float4* inertia, int warm) {
	for (u32 i = begin + blockIdx.x * blockDim.x + threadIdx.x; i < end; i += gridDim.x * blockDim.x) {
		nb_constraint_row r = rows[i];
		const u32 a = r.a, b = r.b;
		const float ma = a ? props[a].mass_inverse : 0.0f, mb = b ? props[b].mass_inverse : 0.0f;
		float4 Ad = inertia[2*a], Ao = inertia[2*a + 1], Bd = inertia[2*b], Bo = inertia[2*b + 1];   // (xx,yy,zz,-) (xy,xz,yz,-)
		if (!a) { Ad = make_float4(0, 0, 0, 0); Ao = Ad; }
		if (!b) { Bd = make_float4(0, 0, 0, 0); Bo = Bd; }
		// I^-1 J_ang
		const float iax = Ad.x*r.ang_a[0] + Ao.x*r.ang_a[1] + Ao.y*r.ang_a[2], iay = Ao.x*r.ang_a[0] + Ad.y*r.ang_a[1] + Ao.z*r.ang_a[2], iaz = Ao.y*r.ang_a[0] + Ao.z*r.ang_a[1] + Ad.z*r.ang_a[2];
		const float ibx = Bd.x*r.ang_b[0] + Bo.x*r.ang_b[1] + Bo.y*r.ang_b[2], iby = Bo.x*r.ang_b[0] + Bd.y*r.ang_b[1] + Bo.z*r.ang_b[2], ibz = Bo.y*r.ang_b[0] + Bo.z*r.ang_b[1] + Bd.z*r.ang_b[2];
		float4* pa = reinterpret_cast<float4*>(momentum + a); float4* pb = reinterpret_cast<float4*>(momentum + b);
		float4 al = pa[0], aw = pa[1], bl = pb[0], bw = pb[1];
		float delta;
		if (warm) delta = r.impulse;
		else {
			const float k = ma * (r.lin_a[0]*r.lin_a[0] + r.lin_a[1]*r.lin_a[1] + r.lin_a[2]*r.lin_a[2]) + (r.ang_a[0]*iax + r.ang_a[1]*iay + r.ang_a[2]*iaz)
						  + mb * (r.lin_b[0]*r.lin_b[0] + r.lin_b[1]*r.lin_b[1] + r.lin_b[2]*r.lin_b[2]) + (r.ang_b[0]*ibx + r.ang_b[1]*iby + r.ang_b[2]*ibz) + r.softness;
			const float jv = r.lin_a[0]*al.x + r.lin_a[1]*al.y + r.lin_a[2]*al.z + r.ang_a[0]*aw.x + r.ang_a[1]*aw.y + r.ang_a[2]*aw.z
						   + r.lin_b[0]*bl.x + r.lin_b[1]*bl.y + r.lin_b[2]*bl.z + r.ang_b[0]*bw.x + r.ang_b[1]*bw.y + r.ang_b[2]*bw.z;
			const float eff = k > 0.0f ? 1.0f / k : 0.0f;
			float next = r.impulse - eff * (jv + r.bias + r.softness * r.impulse);
			next = fminf(fmaxf(next, r.lo), r.hi);
			delta = next - r.impulse;
			rows[i].impulse = next;
		}
		if (a) { al.x += ma * r.lin_a[0] * delta; al.y += ma * r.lin_a[1] * delta; al.z += ma * r.lin_a[2] * delta; aw.x += iax * delta; aw.y += iay * delta; aw.z += iaz * delta; pa[0] = al; pa[1] = aw; }
		if (b) { bl.x += mb * r.lin_b[0] * delta; bl.y += mb * r.lin_b[1] * delta; bl.z += mb * r.lin_b[2] * delta; bw.x += ibx * delta; bw.y += iby * delta; bw.z += ibz * delta; pb[0] = bl; pb[1] = bw; }
	}
}

static int launch_user_rows(nb_context* ctx, int warm, cudaStream_t st) {
	for (u32 l = 0; l < ctx->urow_levels; ++l) {
		const u32 b = ctx->urow_level_off[l], e = ctx->urow_level_off[l + 1];
		if (e > b) { k_user_rows<<<GRID(e - b), NB_BLOCK, 0, st>>>(ctx->urows, b, e, ctx->mom, ctx->props, ctx->inertia, warm); ++ctx->launches; }
	}
	CK(cudaGetLastError());
	return NB_OK;
}

extern "C" {

// Replaces the set of user rows (host pointer; n = 0 removes them).  Rows stay in effect, with their accumulated impulses carried
// from step to step (warm start), until the next upload.  Synchronises the stream.
int nb_upload_constraint_rows(nb_context* ctx, const nb_constraint_row* rows, uint32_t n, void* stream) {
	cudaStream_t st = (cudaStream_t)stream;
	CK(cudaStreamSynchronize(st));
	if (n > ctx->urow_cap) {
		nb_constraint_row* p = nullptr;
		ALLOC(p, (size_t)n + n / 2 + 64);
		ctx->urows = p; ctx->urow_cap = n + n / 2 + 64;
	}
	ctx->urow_order.resize(n); ctx->urow_level_off.clear(); ctx->urow_n = n; ctx->urow_levels = 0; ++ctx->urow_version;
	if (!n) return NB_OK;
	std::vector<u32> level(n);
	std::map<u32, u32> last;   // body -> level of the last row that touched it
	u32 levels = 0;
	for (u32 i = 0; i < n; ++i) {
		const u32 a = rows[i].a, b = rows[i].b;
		if (a >= ctx->cfg.max_bodies || b >= ctx->cfg.max_bodies) { ctx->error = "constraint row references a body beyond max_bodies"; ctx->urow_n = 0; return NB_ERR_ARGUMENT; }
		u32 l = 0;
		if (a) { auto it = last.find(a); if (it != last.end()) l = std::max(l, it->second + 1); }
		if (b) { auto it = last.find(b); if (it != last.end()) l = std::max(l, it->second + 1); }
		level[i] = l; if (a) last[a] = l; if (b) last[b] = l;
		levels = std::max(levels, l + 1);
	}
	if (levels > 4096) { ctx->error = "constraint rows form a dependency chain longer than 4096 levels: reorder them (e.g. red-black along a chain)"; ctx->urow_n = 0; return NB_ERR_CAPACITY; }
	ctx->urow_level_off.assign(levels + 1, 0);
	for (u32 i = 0; i < n; ++i) ++ctx->urow_level_off[level[i] + 1];
	for (u32 l = 0; l < levels; ++l) ctx->urow_level_off[l + 1] += ctx->urow_level_off[l];
	std::vector<u32> cursor(ctx->urow_level_off.begin(), ctx->urow_level_off.end() - 1);
	std::vector<nb_constraint_row> sorted(n);
	for (u32 i = 0; i < n; ++i) { const u32 d = cursor[level[i]]++; sorted[d] = rows[i]; ctx->urow_order[i] = d; }
	CK(cudaMemcpy(ctx->urows, sorted.data(), sizeof(nb_constraint_row) * n, cudaMemcpyHostToDevice));
	ctx->urow_levels = levels;
	return NB_OK;
}

// Copies the rows back in the caller's order (the accumulated impulses are the result).  Synchronises.
int nb_download_constraint_rows(nb_context* ctx, nb_constraint_row* rows, uint32_t n, void* stream) {
	if (n != ctx->urow_n) { ctx->error = "row count differs from the uploaded set"; return NB_ERR_ARGUMENT; }
	if (!n) return NB_OK;
	CK(cudaStreamSynchronize((cudaStream_t)stream));
	std::vector<nb_constraint_row> sorted(n);
	CK(cudaMemcpy(sorted.data(), ctx->urows, sizeof(nb_constraint_row) * n, cudaMemcpyDeviceToHost));
	for (u32 i = 0; i < n; ++i) rows[i] = sorted[ctx->urow_order[i]];
	return NB_OK;
}

}  // extern "C"
