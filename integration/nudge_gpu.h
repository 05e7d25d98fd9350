// nudge_gpu.h — the reference-side binding of INTEGRATION.md section 2 as a real header: what a maintainer of an application built on
// rasmusbarr/nudge adds to run the simulation step on the B200 with the state resident in HBM.  It includes the REFERENCE's nudge.h
// (for the caller-owned structs, nudge.h:29-129) and nudge_b200's C ABI, and needs neither the CUDA toolkit nor nvcc: streams come from
// nb_stream_create.  The application keeps its arrays and its loop; `simulate()` replaces the body of example/main.cpp:274-328.
//
//   nudge::gpu::World w;  w.create(bodies, colliders, max_bodies, max_boxes, max_spheres);
//   every frame:  w.simulate(bodies, 2, 20);            // transforms (and momentum, idle counters) are back in the caller's arrays
//   after host-side edits of the bodies:  w.upload_bodies(bodies);
//
// Errors: the reference asserts; this throws std::runtime_error with nb_last_error (no GPU -> error at create, never a CPU fallback).
#pragma once
#include <nudge.h>
#include "nudge_b200.h"
#include <stdexcept>
#include <string>
#include <vector>

namespace nudge { namespace gpu {

class World {
public:
	World() : ctx_(nullptr), stream_(nullptr) {}
	~World() { destroy(); }
	World(const World&) = delete;
	World& operator=(const World&) = delete;

	void create(const BodyData& bodies, const ColliderData& colliders, unsigned max_bodies, unsigned max_boxes, unsigned max_spheres, int device = 0) {
		destroy();
		nb_config cfg = { max_bodies, max_boxes, max_spheres, /*max_connections*/ 1, /*max_pairs: default*/ 0, /*max_contacts: default*/ 0, device };
		nb_context* c = nullptr;
		const int r = nb_create(&cfg, &c);
		if (r != NB_OK) { std::string why = c ? nb_last_error(c) : "nb_create failed"; if (c) nb_destroy(c); throw std::runtime_error(why); }
		ctx_ = c;
		stream_ = nb_stream_create(ctx_);            // a created stream: nb_step replays its CUDA graph there
		if (!stream_) fail("nb_stream_create");
		upload_colliders(colliders);
		upload_bodies(bodies);
	}
	void destroy() {
		if (ctx_) { if (stream_) nb_stream_destroy(ctx_, stream_); nb_destroy(ctx_); }
		ctx_ = nullptr; stream_ = nullptr;
	}

	// Transform / BodyProperties / BodyMomentum rows are byte-identical to the C ABI's (include/nudge_b200.h): passed through unchanged.
	void upload_bodies(const BodyData& b) {
		nb_body_data hb = { reinterpret_cast<nb_transform*>(b.transforms), reinterpret_cast<nb_body_properties*>(b.properties),
		                    reinterpret_cast<nb_body_momentum*>(b.momentum), b.idle_counters, b.count };
		check(nb_upload_bodies(ctx_, &hb, stream_), "nb_upload_bodies");
	}
	// Collider tags are uint16 in the reference (nudge.h:84-98) and uint32 here: widened once.
	void upload_colliders(const ColliderData& c) {
		box_tags_.assign(c.boxes.tags, c.boxes.tags + c.boxes.count);
		sphere_tags_.assign(c.spheres.tags, c.spheres.tags + c.spheres.count);
		nb_collider_data hc = { { box_tags_.data(), reinterpret_cast<nb_box_collider*>(c.boxes.data), reinterpret_cast<nb_transform*>(c.boxes.transforms), c.boxes.count },
		                        { sphere_tags_.data(), reinterpret_cast<nb_sphere_collider*>(c.spheres.data), reinterpret_cast<nb_transform*>(c.spheres.transforms), c.spheres.count } };
		check(nb_upload_colliders(ctx_, &hc, stream_), "nb_upload_colliders");
		check(nb_stream_synchronize(ctx_, stream_), "nb_stream_synchronize");   // the widened tag vectors may be reused by the caller's next upload
	}

	// example/main.cpp:274-328 with the demo's constants as defaults; the caller's arrays receive the new state.
	void simulate(BodyData& bodies, unsigned steps = 2, unsigned iterations = 20, float gravity = 9.82f, float damping = 0.25f, float frame = 1.0f / 60.0f) {
		const float dt = frame / static_cast<float>(steps);
		for (unsigned n = 0; n < steps; ++n) check(nb_step(ctx_, dt, iterations, gravity, damping, stream_), "nb_step");
		download_bodies(bodies);
	}
	void download_bodies(BodyData& b) {
		nb_body_data hb = { reinterpret_cast<nb_transform*>(b.transforms), reinterpret_cast<nb_body_properties*>(b.properties),
		                    reinterpret_cast<nb_body_momentum*>(b.momentum), b.idle_counters, b.count };
		check(nb_download_bodies(ctx_, &hb, stream_), "nb_download_bodies");        // synchronises
	}
	// What the demo's draw loops compute (example/main.cpp:224-268): one column-major model matrix per collider, boxes first.
	unsigned matrices(float* out, unsigned capacity, bool out_is_device_pointer = false) {
		uint32_t n = 0;
		check(nb_instance_matrices(ctx_, out, capacity, out_is_device_pointer ? 1 : 0, &n, stream_), "nb_instance_matrices");
		return n;
	}
	nb_counts counts() { nb_counts c; const int r = nb_download_counts(ctx_, &c, stream_); if (r != NB_OK && r != NB_ERR_OVERFLOW) fail("nb_download_counts"); return c; }
	nb_context* context() const { return ctx_; }
	void* stream() const { return stream_; }

private:
	void check(int r, const char* what) { if (r != NB_OK) fail(what); }
	[[noreturn]] void fail(const char* what) { throw std::runtime_error(std::string(what) + ": " + (ctx_ ? nb_last_error(ctx_) : "no context")); }
	nb_context* ctx_; void* stream_;
	std::vector<uint32_t> box_tags_, sphere_tags_;
};

}}  // namespace nudge::gpu
