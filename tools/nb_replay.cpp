// nb_replay — headless replay of a dumped simulation state through the C ABI (include/nudge_b200.h).
//
//   nb_replay state.bin [--steps N] [--iterations I] [--dt T] [--gravity G] [--damping D] [--solver parity|throughput]
//                       [--graph 0|1] [--dump out.bin] [--quiet]
//
// Loads a state written by nb_save_state (BodyData / ColliderData / BodyConnections / ContactCache PODs, nudge.h:73-129 widened),
// runs N sub-steps of example/main.cpp:274-328 on the GPU and prints one line per step: contact / pair / batch counts, overflow flags
// and an FNV-1a hash of the transforms — two builds, two machines or two solver settings can be compared line by line, and a parity
// failure found by the tests can be reproduced outside Python.  No CPU path: without a GPU nb_create fails and so does this tool.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include "../include/nudge_b200.h"

static uint64_t fnv1a(const void* p, size_t n) {
	const unsigned char* b = (const unsigned char*)p;
	uint64_t h = 1469598103934665603ull;
	for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
	return h;
}

int main(int argc, char** argv) {
	if (argc < 2) { fprintf(stderr, "usage: %s state.bin [--steps N] [--iterations I] [--dt T] [--gravity G] [--damping D] [--solver parity|throughput] [--graph 0|1] [--dump out.bin] [--quiet]\n", argv[0]); return 2; }
	const char* path = argv[1];
	unsigned steps = 10, iterations = 8; float dt = 1.0f / 120.0f, gravity = 9.82f, damping = 0.25f; int graph = 1, quiet = 0;
	const char* dump = nullptr; const char* solver = nullptr;
	for (int i = 2; i < argc; ++i) {
		std::string a = argv[i];
		auto next = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "missing value after %s\n", a.c_str()); exit(2); } return argv[++i]; };
		if (a == "--steps") steps = (unsigned)atoi(next());
		else if (a == "--iterations") iterations = (unsigned)atoi(next());
		else if (a == "--dt") dt = (float)atof(next());
		else if (a == "--gravity") gravity = (float)atof(next());
		else if (a == "--damping") damping = (float)atof(next());
		else if (a == "--solver") solver = next();
		else if (a == "--graph") graph = atoi(next());
		else if (a == "--dump") dump = next();
		else if (a == "--quiet") quiet = 1;
		else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
	}
	uint32_t n[5];
	if (nb_state_info(path, n) != NB_OK) { fprintf(stderr, "%s: not an nb state file\n", path); return 1; }
	nb_config cfg; memset(&cfg, 0, sizeof(cfg));
	cfg.max_bodies = n[0]; cfg.max_boxes = n[1]; cfg.max_spheres = n[2]; cfg.max_connections = n[3] ? n[3] : 1;
	cfg.max_contacts = n[4] > 24u * n[0] ? n[4] : 24u * n[0];
	if (cfg.max_contacts < 1024) cfg.max_contacts = 1024;
	cfg.max_pairs = 0; cfg.device = 0;
	nb_context* ctx = nullptr;
	if (nb_create(&cfg, &ctx) != NB_OK) { fprintf(stderr, "nb_create: %s\n", ctx ? nb_last_error(ctx) : "failed"); return 1; }
	void* stream = graph ? nb_stream_create(ctx) : nullptr;    // nb_step replays a CUDA graph on a created stream, plain launches on the default one
	if (nb_load_state(ctx, path, stream) != NB_OK) { fprintf(stderr, "nb_load_state: %s\n", nb_last_error(ctx)); return 1; }
	if (solver && nb_set_solver_mode(ctx, !strcmp(solver, "throughput") ? NB_SOLVER_THROUGHPUT : NB_SOLVER_PARITY) != NB_OK) { fprintf(stderr, "%s\n", nb_last_error(ctx)); return 1; }
	std::vector<nb_transform> xf(n[0]);
	if (!quiet) printf("# %s: %u bodies, %u boxes, %u spheres, %u connections, %u cache entries; solver %s\n", path, n[0], n[1], n[2], n[3], n[4], nb_get_solver_mode(ctx) ? "throughput" : "parity");
	int rc = 0;
	for (unsigned s = 0; s < steps; ++s) {
		if (nb_step(ctx, dt, iterations, gravity, damping, stream) != NB_OK) { fprintf(stderr, "nb_step: %s\n", nb_last_error(ctx)); return 1; }
		nb_counts c;
		int r = nb_download_counts(ctx, &c, stream);
		if (r != NB_OK && r != NB_ERR_OVERFLOW) { fprintf(stderr, "nb_download_counts: %s\n", nb_last_error(ctx)); return 1; }
		if (nb_download_transforms(ctx, xf.data(), n[0], stream) != NB_OK || nb_stream_synchronize(ctx, stream) != NB_OK) { fprintf(stderr, "download: %s\n", nb_last_error(ctx)); return 1; }
		if (c.overflow) rc = 3;
		if (!quiet || s + 1 == steps)
			printf("step %u contacts %u pairs %u batches %u active %u cache %u overflow %u xf %016llx\n", s, c.contacts, c.pairs, c.batches, c.active, c.cache, c.overflow,
				   (unsigned long long)fnv1a(xf.data(), sizeof(nb_transform) * xf.size()));
	}
	if (dump && nb_save_state(ctx, dump, stream) != NB_OK) { fprintf(stderr, "nb_save_state: %s\n", nb_last_error(ctx)); return 1; }
	if (stream) nb_stream_destroy(ctx, stream);
	nb_destroy(ctx);
	return rc;
}
