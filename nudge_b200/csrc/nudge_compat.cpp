// nudge_b200 — literal drop-in for the reference's C++ API: the seven free functions of nudge.h:134-146 with the same
// names, the same by-value POD arguments (nudge.h:29-129, uint16 indices) and HOST pointers, implemented on the GPU
// path through the C ABI (include/nudge_b200.h).  An application or the reference's own tests/main.cpp compiled against
// the reference's nudge.h links against libnudge_compat.so instead of nudge.cpp and runs unmodified.
//
// Semantics kept from the reference: every call is synchronous; outputs are written into the caller's arrays; the caller
// may edit momentum / contacts between any two calls (example/main.cpp:285-305, 316), so every call uploads the arrays it
// reads and downloads the ones it writes.  That makes this shim a correctness/compatibility layer, not the fast path —
// the resident C ABI (nb_step) is.  The returned opaque pointers are carved from the caller's Arena like nudge.cpp:4022,
// 4174 do, and only reference the shim's context.  There is no CPU fallback: without a GPU the first call aborts.
//
// The reference is re-entrant on disjoint data (no globals), so the shim keeps ONE device context PER WORLD:
// a world is identified by the caller's BodyData::transforms array (collide / advance) and, from collide on, by its ContactData::data
// array (read_cached_impulses); the opaque handles carry their world.  Two worlds can be stepped from two threads.
#include "../../include/nudge_b200.h"
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <map>
#include <mutex>

// The interface structs, re-declared with the reference's names and layout (nudge.h:29-129) so that the mangled symbols match.
namespace nudge {
	struct Arena { void* data; uintptr_t size; };
	struct Transform { float position[3]; uint32_t body; float rotation[4]; };
	struct BodyProperties { float inertia_inverse[3]; float mass_inverse; };
	struct BodyMomentum { float velocity[3]; float unused0; float angular_velocity[3]; float unused1; };
	struct SphereCollider { float radius; };
	struct BoxCollider { float size[3]; float unused; };
	struct Contact { float position[3]; float penetration; float normal[3]; float friction; };
	struct BodyPair { uint16_t a; uint16_t b; };
	struct ContactData { Contact* data; BodyPair* bodies; uint64_t* tags; uint32_t capacity; uint32_t count; uint32_t* sleeping_pairs; uint32_t sleeping_count; };
	struct ColliderData {
		struct { uint16_t* tags; BoxCollider* data; Transform* transforms; uint32_t count; } boxes;
		struct { uint16_t* tags; SphereCollider* data; Transform* transforms; uint32_t count; } spheres;
	};
	struct BodyData { Transform* transforms; BodyProperties* properties; BodyMomentum* momentum; uint8_t* idle_counters; uint32_t count; };
	struct BodyConnections { BodyPair* data; uint32_t count; };
	struct CachedContactImpulse { float impulse[3]; float unused; };
	struct ContactCache { uint64_t* tags; CachedContactImpulse* data; uint32_t capacity; uint32_t count; };
	struct ActiveBodies { uint16_t* indices; uint32_t capacity; uint32_t count; };
	struct ContactImpulseData { uint64_t magic; void* world; };
	struct ContactConstraintData { uint64_t magic; void* world; };

	void collide(ActiveBodies* active_bodies, ContactData* contacts, BodyData bodies, ColliderData colliders, BodyConnections body_connections, Arena temporary);
	ContactImpulseData* read_cached_impulses(ContactCache contact_cache, ContactData contacts, Arena* memory);
	void write_cached_impulses(ContactCache* contact_cache, ContactData contacts, ContactImpulseData* contact_impulses);
	ContactConstraintData* setup_contact_constraints(ActiveBodies active_bodies, ContactData contacts, BodyData bodies, ContactImpulseData* contact_impulses, Arena* memory);
	void apply_impulses(ContactConstraintData* data, BodyData bodies);
	void update_cached_impulses(ContactConstraintData* data, ContactImpulseData* contact_impulses);
	void advance(ActiveBodies active_bodies, BodyData bodies, float time_step);
}

namespace {

struct Shim {
	nb_context* ctx = nullptr;
	uint32_t cap_bodies = 0, cap_boxes = 0, cap_spheres = 0, cap_conn = 0, cap_contacts = 0;
	std::vector<uint32_t> box_tags, sphere_tags, active32, features;
	std::vector<nb_body_pair> pairs32, conn32;
	std::vector<uint64_t> tags64, sleeping64;
	std::mutex lock;   // one world is still single-threaded, like the reference
};

// worlds by the identity of the caller's arrays
static std::mutex g_registry_lock;
static std::map<const void*, Shim*> g_by_bodies, g_by_contacts;

static Shim& world_of_bodies(const void* transforms) {
	std::lock_guard<std::mutex> l(g_registry_lock);
	Shim*& s = g_by_bodies[transforms];
	if (!s) s = new Shim();
	return *s;
}
static void bind_contacts(const void* contact_rows, Shim* s) {
	std::lock_guard<std::mutex> l(g_registry_lock);
	g_by_contacts[contact_rows] = s;
}
static Shim& world_of_contacts(const void* contact_rows) {
	std::lock_guard<std::mutex> l(g_registry_lock);
	auto it = g_by_contacts.find(contact_rows);
	if (it == g_by_contacts.end() || !it->second) { fprintf(stderr, "nudge_b200 compat: read_cached_impulses on a ContactData that no collide() call has filled\n"); abort(); }
	return *it->second;
}

static void die(Shim& g, const char* what) {
	fprintf(stderr, "nudge_b200 compat: %s: %s\n", what, g.ctx ? nb_last_error(g.ctx) : "no context");
	abort();  // the reference asserts on its failure paths (nudge.cpp:1000-1039, 3010, 4118); there is no CPU fallback
}
#define NBCK(call) do { if ((call) != NB_OK) die(g, #call); } while (0)

static uint32_t grow(uint32_t have, uint32_t need, uint32_t floor_) { uint32_t c = have ? have : floor_; while (c < need) c *= 2; return c; }

static void ensure(Shim& g, uint32_t bodies, uint32_t boxes, uint32_t spheres, uint32_t conn, uint32_t contacts) {
	if (g.ctx && bodies <= g.cap_bodies && boxes <= g.cap_boxes && spheres <= g.cap_spheres && conn <= g.cap_conn && contacts <= g.cap_contacts) return;
	if (g.ctx) nb_destroy(g.ctx);
	g.cap_bodies = grow(g.cap_bodies, bodies, 64); g.cap_boxes = grow(g.cap_boxes, boxes, 64); g.cap_spheres = grow(g.cap_spheres, spheres, 64);
	g.cap_conn = grow(g.cap_conn, conn, 16); g.cap_contacts = grow(g.cap_contacts, contacts, 4096);
	nb_config cfg = { g.cap_bodies, g.cap_boxes, g.cap_spheres, g.cap_conn, 0, g.cap_contacts, 0 };
	if (nb_create(&cfg, &g.ctx) != NB_OK) die(g, "nb_create");
}

static void upload_bodies(Shim& g, const nudge::BodyData& b) {
	nb_body_data hb = { (nb_transform*)b.transforms, (nb_body_properties*)b.properties, (nb_body_momentum*)b.momentum, b.idle_counters, b.count };
	NBCK(nb_upload_bodies(g.ctx, &hb, nullptr));
}

static void upload_contacts(Shim& g, const nudge::ContactData& c, const nudge::ActiveBodies* active) {
	// uint64 tag = feature | A<<32 | B<<48 (nudge.cpp:2089, 2108)  ->  tags = A | B<<32, features
	g.tags64.resize(c.count); g.features.resize(c.count); g.pairs32.resize(c.count);
	for (uint32_t i = 0; i < c.count; ++i) {
		uint64_t t = c.tags[i];
		g.features[i] = (uint32_t)t; g.tags64[i] = ((t >> 32) & 0xffff) | (((t >> 48) & 0xffff) << 32);
		g.pairs32[i].a = c.bodies[i].a; g.pairs32[i].b = c.bodies[i].b;
	}
	g.sleeping64.resize(c.sleeping_count);
	for (uint32_t i = 0; i < c.sleeping_count; ++i) g.sleeping64[i] = (uint64_t)(c.sleeping_pairs[i] & 0xffff) | ((uint64_t)(c.sleeping_pairs[i] >> 16) << 32);
	nb_contact_data hc = { (nb_contact*)c.data, g.pairs32.data(), g.tags64.data(), g.features.data(), c.count, c.count, g.sleeping64.data(), c.sleeping_count };
	nb_active_bodies ha = { nullptr, 0, 0 };
	if (active) {
		g.active32.assign(active->indices, active->indices + active->count);
		ha.indices = g.active32.data(); ha.capacity = ha.count = active->count;
	}
	NBCK(nb_upload_contacts(g.ctx, &hc, active ? &ha : nullptr, nullptr));
}

template<class T>
static T* arena_alloc(nudge::Arena* arena, uintptr_t alignment) {  // same bump discipline as nudge.cpp:990-1025
	uintptr_t p = ((uintptr_t)arena->data + alignment - 1) & ~(alignment - 1);
	uintptr_t end = (uintptr_t)arena->data + arena->size;
	if (p + sizeof(T) > end) { fprintf(stderr, "nudge_b200 compat: arena out of memory\n"); abort(); }
	arena->data = (void*)(p + sizeof(T)); arena->size = end - (p + sizeof(T));
	return (T*)p;
}

}  // namespace

namespace nudge {

void collide(ActiveBodies* active_bodies, ContactData* contacts, BodyData bodies, ColliderData colliders, BodyConnections body_connections, Arena) {
	contacts->count = 0; contacts->sleeping_count = 0; active_bodies->count = 0;  // nudge.cpp:3001-3003
	Shim& g = world_of_bodies(bodies.transforms);
	std::lock_guard<std::mutex> guard(g.lock);
	bind_contacts(contacts->data, &g);
	ensure(g, bodies.count, colliders.boxes.count, colliders.spheres.count, body_connections.count, contacts->capacity);
	upload_bodies(g, bodies);
	g.box_tags.assign(colliders.boxes.tags, colliders.boxes.tags + colliders.boxes.count);
	g.sphere_tags.assign(colliders.spheres.tags, colliders.spheres.tags + colliders.spheres.count);
	nb_collider_data hc = { { g.box_tags.data(), (nb_box_collider*)colliders.boxes.data, (nb_transform*)colliders.boxes.transforms, colliders.boxes.count },
							{ g.sphere_tags.data(), (nb_sphere_collider*)colliders.spheres.data, (nb_transform*)colliders.spheres.transforms, colliders.spheres.count } };
	NBCK(nb_upload_colliders(g.ctx, &hc, nullptr));
	g.conn32.resize(body_connections.count);
	for (uint32_t i = 0; i < body_connections.count; ++i) { g.conn32[i].a = body_connections.data[i].a; g.conn32[i].b = body_connections.data[i].b; }
	nb_body_connections hconn = { g.conn32.data(), body_connections.count };
	NBCK(nb_upload_connections(g.ctx, &hconn, nullptr));
	NBCK(nb_collide(g.ctx, nullptr));

	uint32_t cap = contacts->capacity;
	g.tags64.resize(cap); g.features.resize(cap); g.pairs32.resize(cap); g.sleeping64.resize((size_t)cap + 16); g.active32.resize(bodies.count);
	nb_contact_data out = { (nb_contact*)contacts->data, g.pairs32.data(), g.tags64.data(), g.features.data(), cap, 0, g.sleeping64.data(), 0 };
	nb_active_bodies oa = { g.active32.data(), bodies.count, 0 };
	NBCK(nb_download_contacts(g.ctx, &out, &oa, nullptr));
	contacts->count = out.count;
	for (uint32_t i = 0; i < out.count; ++i) {
		contacts->bodies[i].a = (uint16_t)g.pairs32[i].a; contacts->bodies[i].b = (uint16_t)g.pairs32[i].b;
		contacts->tags[i] = (uint64_t)g.features[i] | ((g.tags64[i] & 0xffff) << 32) | ((g.tags64[i] >> 32) << 48);
	}
	contacts->sleeping_count = out.sleeping_count;
	for (uint32_t i = 0; i < out.sleeping_count; ++i)  // never touched when nothing sleeps: the reference's tests pass a null array (tests/main.cpp:191-196)
		contacts->sleeping_pairs[i] = (uint32_t)(g.sleeping64[i] & 0xffff) | ((uint32_t)(g.sleeping64[i] >> 32) << 16);
	active_bodies->count = oa.count;
	for (uint32_t i = 0; i < oa.count; ++i) active_bodies->indices[i] = (uint16_t)g.active32[i];
}

ContactImpulseData* read_cached_impulses(ContactCache contact_cache, ContactData contacts, Arena* memory) {
	ContactImpulseData* h = arena_alloc<ContactImpulseData>(memory, 64);
	Shim& g = world_of_contacts(contacts.data);
	std::lock_guard<std::mutex> guard(g.lock);
	h->magic = 0x6e62696d70756c73ull; h->world = &g;
	ensure(g, g.cap_bodies, g.cap_boxes, g.cap_spheres, g.cap_conn, contact_cache.count > contacts.count ? contact_cache.count : contacts.count);
	upload_contacts(g, contacts, nullptr);
	g.tags64.resize(contact_cache.count); g.features.resize(contact_cache.count);
	for (uint32_t i = 0; i < contact_cache.count; ++i) {
		uint64_t t = contact_cache.tags[i];
		g.features[i] = (uint32_t)t; g.tags64[i] = ((t >> 32) & 0xffff) | (((t >> 48) & 0xffff) << 32);
	}
	nb_contact_cache hc = { g.tags64.data(), g.features.data(), (nb_cached_impulse*)contact_cache.data, contact_cache.count, contact_cache.count };
	NBCK(nb_upload_cache(g.ctx, &hc, nullptr));
	NBCK(nb_read_cached_impulses(g.ctx, nullptr));
	return h;
}

ContactConstraintData* setup_contact_constraints(ActiveBodies active_bodies, ContactData contacts, BodyData bodies, ContactImpulseData* contact_impulses, Arena* memory) {
	ContactConstraintData* h = arena_alloc<ContactConstraintData>(memory, 64);
	Shim& g = *static_cast<Shim*>(contact_impulses->world);
	std::lock_guard<std::mutex> guard(g.lock);
	h->magic = 0x6e62636f6e737472ull; h->world = &g;
	upload_bodies(g, bodies);  // the caller applied gravity to momentum on the host (example/main.cpp:291-305)
	// the caller may have edited contact rows (friction, penetration, normals ...) or the active list since read_cached_impulses
	// (example/main.cpp:288): the constraint rows are built from what it passes NOW, like nudge.cpp:4350-4561 reads `contacts` here.
	// Count and tags must be the ones read_cached_impulses saw (the reference's sorted order indexes them, nudge.cpp:4044).
	upload_contacts(g, contacts, &active_bodies);
	NBCK(nb_setup_contact_constraints(g.ctx, nullptr));
	NBCK(nb_download_momentum(g.ctx, (nb_body_momentum*)bodies.momentum, bodies.count, nullptr));  // unused0 + warm start (nudge.cpp:4198, 4626-4632)
	nb_body_data none = { nullptr, nullptr, nullptr, nullptr, 0 };
	NBCK(nb_download_bodies(g.ctx, &none, nullptr));  // synchronise
	return h;
}

void apply_impulses(ContactConstraintData* data, BodyData bodies) {
	Shim& g = *static_cast<Shim*>(data->world);
	std::lock_guard<std::mutex> guard(g.lock);
	NBCK(nb_upload_momentum(g.ctx, (const nb_body_momentum*)bodies.momentum, bodies.count, nullptr));  // custom impulses may have been applied (example/main.cpp:316)
	NBCK(nb_apply_impulses(g.ctx, 1, nullptr));
	NBCK(nb_download_momentum(g.ctx, (nb_body_momentum*)bodies.momentum, bodies.count, nullptr));
	nb_body_data none = { nullptr, nullptr, nullptr, nullptr, 0 };
	NBCK(nb_download_bodies(g.ctx, &none, nullptr));
}

void update_cached_impulses(ContactConstraintData* data, ContactImpulseData*) {
	Shim& g = *static_cast<Shim*>(data->world);
	std::lock_guard<std::mutex> guard(g.lock);
	NBCK(nb_update_cached_impulses(g.ctx, nullptr));
}

void write_cached_impulses(ContactCache* contact_cache, ContactData contacts, ContactImpulseData* contact_impulses) {
	Shim& g = *static_cast<Shim*>(contact_impulses->world);
	std::lock_guard<std::mutex> guard(g.lock);
	NBCK(nb_write_cached_impulses(g.ctx, nullptr));
	uint32_t cap = contact_cache->capacity;
	g.tags64.resize(cap); g.features.resize(cap);
	nb_contact_cache hc = { g.tags64.data(), g.features.data(), (nb_cached_impulse*)contact_cache->data, cap, 0 };
	NBCK(nb_download_cache(g.ctx, &hc, nullptr));  // fails (aborts) when capacity is too small, like the assert of nudge.cpp:4118
	contact_cache->count = hc.count;
	for (uint32_t i = 0; i < hc.count; ++i)
		contact_cache->tags[i] = (uint64_t)g.features[i] | ((g.tags64[i] & 0xffff) << 32) | ((g.tags64[i] >> 32) << 48);
	(void)contacts;
}

void advance(ActiveBodies active_bodies, BodyData bodies, float time_step) {
	Shim& g = world_of_bodies(bodies.transforms);
	std::lock_guard<std::mutex> guard(g.lock);
	if (!g.ctx) ensure(g, bodies.count, 1, 1, 1, 1024);   // advance() on a world that never collided: still valid in the reference
	upload_bodies(g, bodies);
	g.active32.assign(active_bodies.indices, active_bodies.indices + active_bodies.count);
	nb_active_bodies ha = { g.active32.data(), active_bodies.count, active_bodies.count };
	NBCK(nb_upload_contacts(g.ctx, nullptr, &ha, nullptr));
	NBCK(nb_advance(g.ctx, time_step, nullptr));
	nb_body_data hb = { (nb_transform*)bodies.transforms, (nb_body_properties*)bodies.properties, (nb_body_momentum*)bodies.momentum, bodies.idle_counters, bodies.count };
	NBCK(nb_download_bodies(g.ctx, &hb, nullptr));
}

}
