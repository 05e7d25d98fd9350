// TEST INFRASTRUCTURE ONLY — a headless application loop over the reference's nudge.h API (the shape of example/main.cpp:274-328:
// collide -> user gravity/damping loop -> read_cached_impulses -> setup_contact_constraints -> N x apply_impulses ->
// update_cached_impulses -> write_cached_impulses -> advance), with its own scene builder (ground box + boxes + spheres, a small LCG
// instead of libc rand()).  Compiled twice by oracle/Makefile against the reference's unmodified nudge.h:
//   _ref/headless_ref   linked with the reference's nudge.cpp (CPU)
//   _ref/headless_gpu   linked with nudge_b200's drop-in (libnudge_compat.so): same seven calls, host pointers, GPU inside
// Both print an FNV-1a hash of all transforms after the run (must be identical: tests/test_gpu_parity.py) and the measured steps/s
// (bench.py reports the drop-in's number as `dropin_seven_call_steps_per_s`).  `worlds` > 1 steps that many independent worlds from
// as many threads at once — the reference is re-entrant on disjoint data and so must the drop-in be.
//   _ref/headless_resident   -DNB_RESIDENT: the same program with step() replaced by integration/nudge_gpu.h (nudge::gpu::World::simulate =
//                        nb_step on the resident state, the fast path of INTEGRATION.md section 2), linked with libnudge_b200.so only
//   usage: headless_example <boxes> <spheres> <steps> <iterations> [worlds]
#include <nudge.h>
#ifdef NB_RESIDENT
#include "../integration/nudge_gpu.h"   // the reference-side binding over the C ABI: state resident in HBM, nb_step per sub-step
#endif
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static void* aligned(size_t bytes) { void* p = nullptr; if (posix_memalign(&p, 64, bytes ? bytes : 64)) abort(); memset(p, 0, bytes); return p; }

struct Lcg { uint64_t s; float next() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (float)((s >> 40) & 0xffffff) / 16777216.0f; } };

struct World {
	nudge::Arena arena; nudge::BodyData bodies; nudge::ColliderData colliders; nudge::ContactData contacts; nudge::ContactCache cache; nudge::ActiveBodies active;
	unsigned max_bodies;
#ifdef NB_RESIDENT
	nudge::gpu::World gpu;
#endif

	unsigned add_body(const float pos[3], float mass, const float inertia[3]) {
		unsigned b = bodies.count++;
		nudge::Transform t = {}; t.position[0] = pos[0]; t.position[1] = pos[1]; t.position[2] = pos[2]; t.rotation[3] = 1.0f;
		bodies.transforms[b] = t;
		bodies.properties[b].mass_inverse = 1.0f / mass;
		for (int k = 0; k < 3; ++k) bodies.properties[b].inertia_inverse[k] = 1.0f / inertia[k];
		memset(&bodies.momentum[b], 0, sizeof(nudge::BodyMomentum));
		bodies.idle_counters[b] = 0;
		return b;
	}
	void build(unsigned n_boxes, unsigned n_spheres, uint64_t seed) {
		max_bodies = n_boxes + n_spheres + 1;
		const unsigned max_contacts = 24 * max_bodies + 1024;
		arena.size = (size_t)64 << 20; arena.data = aligned(arena.size);
		bodies.idle_counters = (uint8_t*)aligned(max_bodies); bodies.transforms = (nudge::Transform*)aligned(sizeof(nudge::Transform) * max_bodies);
		bodies.momentum = (nudge::BodyMomentum*)aligned(sizeof(nudge::BodyMomentum) * max_bodies); bodies.properties = (nudge::BodyProperties*)aligned(sizeof(nudge::BodyProperties) * max_bodies);
		bodies.count = 0;
		colliders.boxes.data = (nudge::BoxCollider*)aligned(sizeof(nudge::BoxCollider) * (n_boxes + 1)); colliders.boxes.tags = (uint16_t*)aligned(2 * (n_boxes + 1));
		colliders.boxes.transforms = (nudge::Transform*)aligned(sizeof(nudge::Transform) * (n_boxes + 1)); colliders.boxes.count = 0;
		colliders.spheres.data = (nudge::SphereCollider*)aligned(sizeof(nudge::SphereCollider) * (n_spheres + 1)); colliders.spheres.tags = (uint16_t*)aligned(2 * (n_spheres + 1));
		colliders.spheres.transforms = (nudge::Transform*)aligned(sizeof(nudge::Transform) * (n_spheres + 1)); colliders.spheres.count = 0;
		contacts.capacity = max_contacts; contacts.count = 0; contacts.sleeping_count = 0;
		contacts.bodies = (nudge::BodyPair*)aligned(sizeof(nudge::BodyPair) * max_contacts); contacts.data = (nudge::Contact*)aligned(sizeof(nudge::Contact) * max_contacts);
		contacts.tags = (uint64_t*)aligned(8 * max_contacts); contacts.sleeping_pairs = (uint32_t*)aligned(4 * max_contacts);
		cache.capacity = max_contacts; cache.count = 0; cache.data = (nudge::CachedContactImpulse*)aligned(sizeof(nudge::CachedContactImpulse) * max_contacts); cache.tags = (uint64_t*)aligned(8 * max_contacts);
		active.capacity = max_bodies; active.count = 0; active.indices = (uint16_t*)aligned(2 * max_bodies);
		// body 0: the static world with a ground box
		bodies.count = 1; bodies.transforms[0].rotation[3] = 1.0f; bodies.idle_counters[0] = 0;
		{
			unsigned c = colliders.boxes.count++;
			nudge::Transform t = {}; t.position[1] = -20.0f; t.rotation[3] = 1.0f; t.body = 0;
			colliders.boxes.transforms[c] = t; colliders.boxes.data[c].size[0] = 400.0f; colliders.boxes.data[c].size[1] = 10.0f; colliders.boxes.data[c].size[2] = 400.0f; colliders.boxes.tags[c] = (uint16_t)c;
		}
		Lcg r = { seed };
		for (unsigned i = 0; i < n_boxes; ++i) {
			float sx = r.next() + 0.5f, sy = r.next() + 0.5f, sz = r.next() + 0.5f;
			float mass = 8.0f * sx * sy * sz, k = mass * (1.0f / 3.0f);
			float inertia[3] = { k * (sy * sy + sz * sz), k * (sx * sx + sz * sz), k * (sx * sx + sy * sy) };
			float pos[3] = { r.next() * 10.0f - 5.0f, r.next() * 60.0f, r.next() * 10.0f - 5.0f };
			unsigned b = add_body(pos, mass, inertia);
			unsigned c = colliders.boxes.count++;
			nudge::Transform t = {}; t.rotation[3] = 1.0f; t.body = b;
			colliders.boxes.transforms[c] = t; colliders.boxes.data[c].size[0] = sx; colliders.boxes.data[c].size[1] = sy; colliders.boxes.data[c].size[2] = sz; colliders.boxes.tags[c] = (uint16_t)c;
		}
		for (unsigned i = 0; i < n_spheres; ++i) {
			float rad = r.next() + 0.5f, mass = 4.18879f * rad * rad * rad, k = 0.4f * mass * rad * rad;
			float inertia[3] = { k, k, k };
			float pos[3] = { r.next() * 10.0f - 5.0f, r.next() * 60.0f, r.next() * 10.0f - 5.0f };
			unsigned b = add_body(pos, mass, inertia);
			unsigned c = colliders.spheres.count++;
			nudge::Transform t = {}; t.rotation[3] = 1.0f; t.body = b;
			colliders.spheres.transforms[c] = t; colliders.spheres.data[c].radius = rad; colliders.spheres.tags[c] = (uint16_t)(c + n_boxes + 1);
		}
	}
	void step(unsigned iterations, float time_step) {   // one sub-step of example/main.cpp:280-327
#ifdef NB_RESIDENT
		gpu.simulate(bodies, 1, iterations, 9.82f, 0.25f, time_step);   // gravity, damping and the seven stages on the device; state comes back
		return;
#endif
		nudge::Arena temporary = arena;
		nudge::BodyConnections connections = {};
		nudge::collide(&active, &contacts, bodies, colliders, connections, temporary);
		float damping = 1.0f - time_step * 0.25f;
		for (unsigned i = 0; i < active.count; ++i) {
			unsigned index = active.indices[i];
			bodies.momentum[index].velocity[1] -= 9.82f * time_step;
			for (int k = 0; k < 3; ++k) { bodies.momentum[index].velocity[k] *= damping; bodies.momentum[index].angular_velocity[k] *= damping; }
		}
		nudge::ContactImpulseData* ci = nudge::read_cached_impulses(cache, contacts, &temporary);
		nudge::ContactConstraintData* cc = nudge::setup_contact_constraints(active, contacts, bodies, ci, &temporary);
		for (unsigned i = 0; i < iterations; ++i) nudge::apply_impulses(cc, bodies);
		nudge::update_cached_impulses(cc, ci);
		nudge::write_cached_impulses(&cache, contacts, ci);
		nudge::advance(active, bodies, time_step);
	}
	uint64_t hash() const {
		const unsigned char* b = (const unsigned char*)bodies.transforms; size_t n = sizeof(nudge::Transform) * bodies.count;
		uint64_t h = 1469598103934665603ull;
		for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
		return h;
	}
};

int main(int argc, char** argv) {
	if (argc < 5) { fprintf(stderr, "usage: %s <boxes> <spheres> <steps> <iterations> [worlds]\n", argv[0]); return 2; }
	unsigned n_boxes = atoi(argv[1]), n_spheres = atoi(argv[2]), steps = atoi(argv[3]), iterations = atoi(argv[4]), n_worlds = argc > 5 ? atoi(argv[5]) : 1;
	std::vector<World> worlds(n_worlds);
	for (unsigned w = 0; w < n_worlds; ++w) worlds[w].build(n_boxes, n_spheres, 12345 + 77 * w);
#ifdef NB_RESIDENT
	try { for (unsigned w = 0; w < n_worlds; ++w) worlds[w].gpu.create(worlds[w].bodies, worlds[w].colliders, worlds[w].max_bodies, n_boxes + 1, n_spheres + 1); }
	catch (const std::exception& e) { fprintf(stderr, "nudge_b200: %s\n", e.what()); return 3; }
#endif
	const float dt = 1.0f / 120.0f;
	auto run = [&](unsigned w) { for (unsigned s = 0; s < steps; ++s) worlds[w].step(iterations, dt); };
	auto t0 = std::chrono::steady_clock::now();
	if (n_worlds == 1) run(0);
	else { std::vector<std::thread> th; for (unsigned w = 0; w < n_worlds; ++w) th.emplace_back(run, w); for (auto& t : th) t.join(); }
	double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	for (unsigned w = 0; w < n_worlds; ++w) {
		unsigned n_contacts = worlds[w].contacts.count, n_cache = worlds[w].cache.count;
#ifdef NB_RESIDENT
		{ nb_counts c = worlds[w].gpu.counts(); n_contacts = c.contacts; n_cache = c.cache; }
#endif
		printf("world %u bodies %u contacts %u cache %u hash %016llx\n", w, worlds[w].bodies.count, n_contacts, n_cache, (unsigned long long)worlds[w].hash());
	}
	printf("steps_per_s %.3f (%u steps of %u world(s) in %.3f s)\n", steps / sec, steps, n_worlds, sec);
	return 0;
}
