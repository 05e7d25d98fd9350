// TEST INFRASTRUCTURE ONLY — never linked into the shipped library.
//
// C-ABI shim around the UNMODIFIED reference (/root/reference/nudge.cpp, compiled in place by
// oracle/Makefile into oracle/_ref/).  It exposes the reference's seven entry points
// (nudge.h:134-146) to ctypes and lets tests read the reference's opaque solver structs
// (nudge.cpp:4011-4019 ContactImpulseData, nudge.cpp:4160-4168 ContactConstraintData,
// nudge.cpp:903-964 ContactConstraintV / ContactConstraintStateV) through mirror declarations.
// No reference source is copied; only struct *layouts* are re-declared so that fields can be read.
#include "nudge.h"
#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>

using namespace nudge;

namespace {
// Mirror of nudge.cpp:4011-4019.
struct ImpulseMirror {
	uint32_t* sorted_contacts;
	CachedContactImpulse* culled_data;
	uint64_t* culled_tags;
	unsigned culled_count;
	CachedContactImpulse* data;
};

// Mirror of nudge.cpp:903-958 for the 8-wide (AVX2) build: 2x uint16[8] then 39 float[8].
struct ConstraintVMirror {
	uint16_t a[8];
	uint16_t b[8];
	float f[39][8];
};

// Mirror of nudge.cpp:960-964.
struct StateVMirror {
	float normal[8];
	float friction_x[8];
	float friction_y[8];
};

// Mirror of nudge.cpp:4160-4168.
struct ConstraintDataMirror {
	unsigned contact_count;
	void* momentum_to_velocity;
	uint32_t* constraint_to_contact;
	ConstraintVMirror* constraints;
	StateVMirror* constraint_states;
	unsigned constraint_batches;
};
}

extern "C" {

int ref_simd_width() {
#ifdef __AVX2__
	return 8;
#else
	return 4;
#endif
}

void ref_set_ftz_daz(int on) {
	if (on) {
		_MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
		_MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
	}
	else {
		_MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_OFF);
		_MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_OFF);
	}
}

void ref_collide(ActiveBodies* active, ContactData* contacts, const BodyData* bodies, const ColliderData* colliders,
				 const BodyConnections* connections, void* arena_data, uintptr_t arena_size) {
	Arena temporary = { arena_data, arena_size };
	collide(active, contacts, *bodies, *colliders, *connections, temporary);
}

void* ref_read_cached_impulses(const ContactCache* cache, const ContactData* contacts, Arena* memory) {
	return read_cached_impulses(*cache, *contacts, memory);
}

void ref_write_cached_impulses(ContactCache* cache, const ContactData* contacts, void* impulses) {
	write_cached_impulses(cache, *contacts, (ContactImpulseData*)impulses);
}

void* ref_setup_contact_constraints(const ActiveBodies* active, const ContactData* contacts, const BodyData* bodies, void* impulses, Arena* memory) {
	return setup_contact_constraints(*active, *contacts, *bodies, (ContactImpulseData*)impulses, memory);
}

void ref_apply_impulses(void* constraints, const BodyData* bodies) {
	apply_impulses((ContactConstraintData*)constraints, *bodies);
}

void ref_update_cached_impulses(void* constraints, void* impulses) {
	update_cached_impulses((ContactConstraintData*)constraints, (ContactImpulseData*)impulses);
}

void ref_advance(const ActiveBodies* active, const BodyData* bodies, float time_step) {
	advance(*active, *bodies, time_step);
}

// ---- introspection of the opaque structs (read-only) ----

void ref_impulses_get(void* impulses, unsigned contact_count, uint32_t* sorted_contacts, float* data /*[n][4]*/,
					  unsigned* culled_count, uint64_t* culled_tags, float* culled_data, unsigned culled_capacity) {
	ImpulseMirror* m = (ImpulseMirror*)impulses;
	if (sorted_contacts) memcpy(sorted_contacts, m->sorted_contacts, sizeof(uint32_t)*contact_count);
	if (data) memcpy(data, m->data, sizeof(CachedContactImpulse)*contact_count);
	if (culled_count) *culled_count = m->culled_count;
	unsigned n = m->culled_count < culled_capacity ? m->culled_count : culled_capacity;
	if (culled_tags) memcpy(culled_tags, m->culled_tags, sizeof(uint64_t)*n);
	if (culled_data) memcpy(culled_data, m->culled_data, sizeof(CachedContactImpulse)*n);
}

unsigned ref_constraints_batches(void* constraints) {
	return ((ConstraintDataMirror*)constraints)->constraint_batches;
}

// Flattens batches into per-lane arrays: lane i = batch*8 + l.
// rows: [lanes][39] in the member order of nudge.cpp:907-957; states: [lanes][3].
void ref_constraints_get(void* constraints, uint32_t* constraint_to_contact, uint32_t* a, uint32_t* b, float* rows, float* states) {
	ConstraintDataMirror* m = (ConstraintDataMirror*)constraints;
	unsigned lanes = m->constraint_batches*8;
	if (constraint_to_contact) memcpy(constraint_to_contact, m->constraint_to_contact, sizeof(uint32_t)*lanes);
	for (unsigned i = 0; i < lanes; ++i) {
		unsigned bt = i >> 3, l = i & 7;
		if (a) a[i] = m->constraints[bt].a[l];
		if (b) b[i] = m->constraints[bt].b[l];
		if (rows)
			for (unsigned k = 0; k < 39; ++k)
				rows[i*39 + k] = m->constraints[bt].f[k][l];
		if (states) {
			states[i*3 + 0] = m->constraint_states[bt].normal[l];
			states[i*3 + 1] = m->constraint_states[bt].friction_x[l];
			states[i*3 + 2] = m->constraint_states[bt].friction_y[l];
		}
	}
}

// ---- whole step as example/main.cpp:274-328 runs it (one sub-step), used for CPU timing and trajectories ----
// phase_seconds[6]: collide, cache-read, setup, solve, cache-write(+update), advance.  May be null.
void ref_step(ActiveBodies* active, ContactData* contacts, const BodyData* bodies, const ColliderData* colliders,
			  const BodyConnections* connections, ContactCache* cache, void* arena_data, uintptr_t arena_size,
			  float time_step, unsigned iterations, float gravity, float damping_base, double* phase_seconds) {
	typedef std::chrono::steady_clock clk;
	Arena temporary = { arena_data, arena_size };
	clk::time_point t0 = clk::now();
	collide(active, contacts, *bodies, *colliders, *connections, temporary);
	clk::time_point t1 = clk::now();

	// User code between collide and the solver: example/main.cpp:291-305.
	float damping = 1.0f - time_step*damping_base;
	for (unsigned i = 0; i < active->count; ++i) {
		unsigned index = active->indices[i];
		bodies->momentum[index].velocity[1] -= gravity * time_step;
		bodies->momentum[index].velocity[0] *= damping;
		bodies->momentum[index].velocity[1] *= damping;
		bodies->momentum[index].velocity[2] *= damping;
		bodies->momentum[index].angular_velocity[0] *= damping;
		bodies->momentum[index].angular_velocity[1] *= damping;
		bodies->momentum[index].angular_velocity[2] *= damping;
	}
	clk::time_point t2 = clk::now();
	ContactImpulseData* impulses = read_cached_impulses(*cache, *contacts, &temporary);
	clk::time_point t3 = clk::now();
	ContactConstraintData* constraints = setup_contact_constraints(*active, *contacts, *bodies, impulses, &temporary);
	clk::time_point t4 = clk::now();
	for (unsigned i = 0; i < iterations; ++i)
		apply_impulses(constraints, *bodies);
	clk::time_point t5 = clk::now();
	update_cached_impulses(constraints, impulses);
	write_cached_impulses(cache, *contacts, impulses);
	clk::time_point t6 = clk::now();
	advance(*active, *bodies, time_step);
	clk::time_point t7 = clk::now();

	if (phase_seconds) {
		phase_seconds[0] += std::chrono::duration<double>(t1 - t0).count();
		phase_seconds[1] += std::chrono::duration<double>(t3 - t2).count();
		phase_seconds[2] += std::chrono::duration<double>(t4 - t3).count();
		phase_seconds[3] += std::chrono::duration<double>(t5 - t4).count();
		phase_seconds[4] += std::chrono::duration<double>(t6 - t5).count();
		phase_seconds[5] += std::chrono::duration<double>(t7 - t6).count() + std::chrono::duration<double>(t2 - t1).count();
	}
}

// ---- host rcpps / rsqrtps samples, used to validate the LUT model of SURVEY.md §0.5 ----
void ref_rcp(const float* x, float* y, unsigned n) {
	for (unsigned i = 0; i < n; ++i) y[i] = _mm_cvtss_f32(_mm_rcp_ps(_mm_set1_ps(x[i])));
}

void ref_rsqrt(const float* x, float* y, unsigned n) {
	for (unsigned i = 0; i < n; ++i) y[i] = _mm_cvtss_f32(_mm_rsqrt_ps(_mm_set1_ps(x[i])));
}

}
