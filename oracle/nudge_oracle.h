/* TEST INFRASTRUCTURE ONLY — CPU restatement of rasmusbarr/nudge with widened (32-bit) indices.
 *
 * Same seven entry points as nudge.h:134-146, same SoA layout (nudge.h:29-129), except that every
 * index-carrying field is 32 bits wide:
 *   BodyPair{uint16 a,b}                  -> nbo_pair{uint32 a,b}
 *   collider tags uint16                  -> uint32
 *   contact tag uint64 = feature | A<<32 | B<<48   -> tags[i] = A | (uint64)B<<32  plus  features[i]
 *   sleeping pair uint32 = X | Y<<16      -> uint64 = X | (uint64)Y<<32
 *   ActiveBodies.indices uint16           -> uint32
 * Arithmetic is untouched.  Pinned bit-for-bit against oracle/_ref (the unmodified reference) on every
 * scene that fits the reference's limits: tests/test_oracle_vs_ref.py.
 */
#ifndef NUDGE_ORACLE_H
#define NUDGE_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float position[3]; uint32_t body; float rotation[4]; } nbo_transform;
typedef struct { float inertia_inverse[3]; float mass_inverse; } nbo_properties;
typedef struct { float velocity[3]; float unused0; float angular_velocity[3]; float unused1; } nbo_momentum;
typedef struct { float size[3]; float unused; } nbo_box;
typedef struct { float radius; } nbo_sphere;
typedef struct { float position[3]; float penetration; float normal[3]; float friction; } nbo_contact;
typedef struct { uint32_t a, b; } nbo_pair;
typedef struct { float impulse[3]; float unused; } nbo_impulse;

typedef struct {
	nbo_contact* data; nbo_pair* bodies; uint64_t* tags; uint32_t* features;
	uint32_t capacity, count;
	uint64_t* sleeping_pairs; uint32_t sleeping_count;
} nbo_contact_data;

typedef struct { uint32_t* tags; nbo_box* data; nbo_transform* transforms; uint32_t count; } nbo_boxes;
typedef struct { uint32_t* tags; nbo_sphere* data; nbo_transform* transforms; uint32_t count; } nbo_spheres;
typedef struct { nbo_boxes boxes; nbo_spheres spheres; } nbo_collider_data;
typedef struct { nbo_transform* transforms; nbo_properties* properties; nbo_momentum* momentum; uint8_t* idle_counters; uint32_t count; } nbo_body_data;
typedef struct { nbo_pair* data; uint32_t count; } nbo_connections;
typedef struct { uint64_t* tags; uint32_t* features; nbo_impulse* data; uint32_t capacity, count; } nbo_contact_cache;
typedef struct { uint32_t* indices; uint32_t capacity, count; } nbo_active_bodies;

typedef struct nbo_impulse_data nbo_impulse_data;       /* nudge.cpp:4011-4019 */
typedef struct nbo_constraint_data nbo_constraint_data; /* nudge.cpp:4160-4168 */

void nbo_collide(nbo_active_bodies* active, nbo_contact_data* contacts, const nbo_body_data* bodies, const nbo_collider_data* colliders, const nbo_connections* connections);
nbo_impulse_data* nbo_read_cached_impulses(const nbo_contact_cache* cache, const nbo_contact_data* contacts);
void nbo_write_cached_impulses(nbo_contact_cache* cache, const nbo_contact_data* contacts, nbo_impulse_data* impulses);
nbo_constraint_data* nbo_setup_contact_constraints(const nbo_active_bodies* active, const nbo_contact_data* contacts, const nbo_body_data* bodies, nbo_impulse_data* impulses);
void nbo_apply_impulses(nbo_constraint_data* data, const nbo_body_data* bodies);
void nbo_update_cached_impulses(nbo_constraint_data* data, nbo_impulse_data* impulses);
void nbo_advance(const nbo_active_bodies* active, const nbo_body_data* bodies, float time_step);
void nbo_free_impulses(nbo_impulse_data*);
void nbo_free_constraints(nbo_constraint_data*);

/* Introspection for stage-by-stage parity tests. */
int nbo_last_overflow(void);                             /* 1 if the last nbo_collide ran out of contact capacity */
uint32_t nbo_last_pair_count(void);                      /* broadphase pairs of the last nbo_collide (after the sort, before islands) */
void nbo_last_pairs(uint32_t* lo, uint32_t* hi);          /* nudge.cpp:3493-3498: lo = later in Morton order, hi = earlier */
void nbo_last_morton_order(uint32_t* sorted_indices);     /* nudge.cpp:3165-3172 */
void nbo_impulses_get(nbo_impulse_data*, uint32_t contact_count, uint32_t* sorted_contacts, float* data, uint32_t* culled_count,
					  uint64_t* culled_tags, uint32_t* culled_features, float* culled_data, uint32_t culled_capacity);
uint32_t nbo_constraints_batches(nbo_constraint_data*);
void nbo_constraints_get(nbo_constraint_data*, uint32_t* constraint_to_contact, uint32_t* a, uint32_t* b, float* rows39, float* states3);
void nbo_rcp(const float* x, float* y, uint32_t n);
void nbo_rsqrt(const float* x, float* y, uint32_t n);
void nbo_set_ftz_daz(int on);

#ifdef __cplusplus
}
#endif
#endif
