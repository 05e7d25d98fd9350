// TEST INFRASTRUCTURE ONLY — runs the reference's demo application headless and records what it computes.
//
// The translation unit IS the reference's example/main.cpp, unmodified and compiled where it lies (-I$(REF)), with the two GL headers
// it includes replaced by the recording stand-ins under oracle/gl_stub/.  Its main() builds the scene from rand() (example/main.cpp:
// 391-432), calls timer(0) once and ends in glutMainLoop(), which lands in demo_capture_loop() below: render() and timer() (= simulate(),
// two sub-steps of 20 iterations, example/main.cpp:274-334) are called frame by frame.  Records are appended to a flat binary file:
// the INITIAL state (before the first simulate(): frame = 0xffffffff, no matrices) and, at the frames named on the command line, the
// state the demo owns plus the model matrices its render() loaded (example/main.cpp:224-268).  Frame f is the state after f + 1 calls
// of simulate().
//
//   u32 magic 'NDMO', u32 frame, u32 bodies, u32 boxes, u32 spheres, u32 matrices,
//   transforms[bodies] (32 B) | momentum[bodies] (32) | properties[bodies] (16) | idle[bodies] (1, padded to 4)
//   box transforms[boxes] (32) | box sizes[boxes] (16) | box tags (u16, padded to 4)
//   sphere transforms[spheres] (32) | radii[spheres] (4) | sphere tags (u16, padded to 4) | matrices[matrices] (64)
//
// Usage: DEMO_CAPTURE_OUT=path DEMO_CAPTURE_FRAMES=0,40 _ref/demo_capture     (tests/golden/make_demo_golden.py)
#include "example/main.cpp"
#include <vector>

static FILE* capture_file;

static void put(const void* p, size_t bytes) {
	static const char zero[4] = {};
	fwrite(p, 1, bytes, capture_file);
	if (bytes & 3) fwrite(zero, 1, 4 - (bytes & 3), capture_file);
}

static void record(unsigned frame, const float* matrices, unsigned count) {
	const unsigned head[6] = { 0x4f4d444eu, frame, bodies.count, colliders.boxes.count, colliders.spheres.count, count };
	put(head, sizeof(head));
	put(bodies.transforms, sizeof(nudge::Transform) * bodies.count);
	put(bodies.momentum, sizeof(nudge::BodyMomentum) * bodies.count);
	put(bodies.properties, sizeof(nudge::BodyProperties) * bodies.count);
	put(bodies.idle_counters, bodies.count);
	put(colliders.boxes.transforms, sizeof(nudge::Transform) * colliders.boxes.count);
	put(colliders.boxes.data, sizeof(nudge::BoxCollider) * colliders.boxes.count);
	put(colliders.boxes.tags, 2 * colliders.boxes.count);
	put(colliders.spheres.transforms, sizeof(nudge::Transform) * colliders.spheres.count);
	put(colliders.spheres.data, sizeof(nudge::SphereCollider) * colliders.spheres.count);
	put(colliders.spheres.tags, 2 * colliders.spheres.count);
	put(matrices, 64 * (size_t)count);
}

void demo_capture_before_simulate() {
	if (capture_file) return;
	// the demo turns FTZ/DAZ on for speed (example/main.cpp:338-339); the parity harness runs every implementation with denormals kept
	// (SURVEY.md section 8c), so the recorded frames are computed that way too — an MXCSR setting, not a change to the demo
	// (DEMO_CAPTURE_KEEP_FTZ=1 leaves the demo's own setting in place: the second fixture, pinned by the oracle's FTZ mode)
	if (!getenv("DEMO_CAPTURE_KEEP_FTZ")) {
		_MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_OFF);
		_MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_OFF);
	}
	const char* out = getenv("DEMO_CAPTURE_OUT");
	if (!out) { fprintf(stderr, "set DEMO_CAPTURE_OUT and DEMO_CAPTURE_FRAMES\n"); exit(2); }
	capture_file = fopen(out, "wb");
	if (!capture_file) { perror(out); exit(2); }
	record(0xffffffffu, nullptr, 0);
}

void demo_capture_loop() {
	const char* list = getenv("DEMO_CAPTURE_FRAMES");
	std::vector<unsigned> frames;
	for (const char* p = list ? list : "0"; *p; ) { frames.push_back((unsigned)strtoul(p, const_cast<char**>(&p), 10)); if (*p == ',') ++p; }
	const unsigned K = colliders.boxes.count + colliders.spheres.count;
	std::vector<float> matrices(16 * (size_t)K);
	unsigned last = 0;
	for (size_t i = 0; i < frames.size(); ++i) last = frames[i] > last ? frames[i] : last;
	for (unsigned frame = 0; frame <= last; ++frame) {
		gl_stub().frame = matrices.data(); gl_stub().capacity = K; gl_stub().count = 0;
		glut_stub().display();
		bool wanted = false;
		for (size_t i = 0; i < frames.size(); ++i) wanted |= frames[i] == frame;
		if (wanted) record(frame, matrices.data(), (unsigned)gl_stub().count);
		glut_stub().timer(0);
	}
	fclose(capture_file);
	exit(0);
}
