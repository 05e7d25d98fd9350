// TEST INFRASTRUCTURE ONLY — never linked into the shipped library, never timed as the product.
//
// Scalar CPU restatement of rasmusbarr/nudge (reference: /root/reference/nudge.cpp @ bb00cb29) with
// widened 32-bit indices, see nudge_oracle.h.  Every function cites the reference lines it follows.
// Arithmetic is restated lane by lane with the reference's exact operation order; build with
// -ffp-contract=off so that FMAs appear exactly where the reference writes madd/msub.
// rcp/rsqrt call the host's rcpss/rsqrtss, i.e. the very instructions the reference executes.
//
// Differences from the reference, none of which change results on scenes the reference can run:
//   * indices are 32 bit (no 8192-collider / 65535-body limit);
//   * the O((K/8)^2) coarse all-pairs accelerator (nudge.cpp:3275-3399) is replaced by an exact
//     sort-and-sweep that produces the same pair SET; pair orientation and order are then rebuilt
//     from the Morton order exactly as nudge.cpp:3493-3498 defines them;
//   * scratch comes from std::vector instead of the caller's Arena.
// Parity status: PINNED against oracle/_ref by tests/test_oracle_vs_ref.py (bit-exact, every stage).
#include "nudge_oracle.h"
#include <immintrin.h>
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>

typedef uint32_t u32;
typedef uint64_t u64;

namespace {

// ---------- exact op semantics: nudge.cpp:233-339 (128-bit) / 541-647 (256-bit) ----------
static inline u32 asu(float x) { u32 u; memcpy(&u, &x, 4); return u; }
static inline float asf(u32 u) { float x; memcpy(&x, &u, 4); return x; }
static inline float min1(float x, float y) { return (y < x) ? y : x; }  // nudge.cpp:286-289: first operand on NaN
static inline float max1(float x, float y) { return (y > x) ? y : x; }  // nudge.cpp:291-294
static inline float rcp(float x) { return _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(x))); }      // nudge.cpp:300-302
static inline float rsqrt(float x) { return _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x))); }  // nudge.cpp:296-298
static inline float madd(float x, float y, float z) { return __builtin_fmaf(x, y, z); }     // nudge.cpp:270-276 (__FMA__)
static inline float msub(float x, float y, float z) { return __builtin_fmaf(x, y, -z); }    // nudge.cpp:278-284
static inline float absf(float x) { return asf(asu(x) & 0x7fffffffu); }                    // nudge.cpp:308-310
static inline float xorf(float x, u32 m) { return asf(asu(x) ^ m); }
static inline float negf(float x) { return xorf(x, 0x80000000u); }                         // nudge.cpp:63-65
static inline u32 signbit32(float x) { return asu(x) >> 31; }
static inline int toint(float x) { return _mm_cvttss_si32(_mm_set_ss(x)); }                 // nudge.cpp:336-338
static const u32 SIGN = 0x80000000u;

struct float3 { float x, y, z; };
struct Rot { float3 v; float s; };

static inline float3 f3(const float p[3]) { float3 r = { p[0], p[1], p[2] }; return r; }
static inline float3 add(float3 a, float3 b) { float3 r = { a.x + b.x, a.y + b.y, a.z + b.z }; return r; }
static inline float3 sub(float3 a, float3 b) { float3 r = { a.x - b.x, a.y - b.y, a.z - b.z }; return r; }
static inline float3 mul(float a, float3 b) { float3 r = { a * b.x, a * b.y, a * b.z }; return r; }
static inline float3 mul(float3 a, float b) { float3 r = { a.x * b, a.y * b, a.z * b }; return r; }
static inline float dot(float3 a, float3 b) { return a.x*b.x + a.y*b.y + a.z*b.z; }  // nudge.cpp:1104-1106
static inline float3 cross(float3 a, float3 b) {                                     // nudge.cpp:1112-1115, 836-840
	float3 v = { a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x };
	return v;
}
static inline Rot rot(const float q[4]) { Rot r = { { q[0], q[1], q[2] }, q[3] }; return r; }
static inline float3 rotate(Rot l, float3 r) {  // nudge.cpp:1117-1120
	float3 t = mul(2.0f, cross(l.v, r));
	return add(add(r, mul(l.s, t)), cross(l.v, t));
}
static inline Rot rotmul(Rot l, Rot r) {  // nudge.cpp:1122-1126
	float3 v = add(add(mul(r.v, l.s), mul(l.v, r.s)), cross(l.v, r.v));
	Rot q = { v, l.s*r.s - dot(l.v, r.v) };
	return q;
}
struct Mat3 { float3 c0, c1, c2; };
static inline Mat3 matrix(Rot q) {  // nudge.cpp:1142-1163
	float kx = q.v.x + q.v.x, ky = q.v.y + q.v.y, kz = q.v.z + q.v.z;
	float xx = kx*q.v.x, yy = ky*q.v.y, zz = kz*q.v.z;
	float xy = kx*q.v.y, xz = kx*q.v.z, yz = ky*q.v.z;
	float sx = kx*q.s, sy = ky*q.s, sz = kz*q.s;
	Mat3 m = {
		{ 1.0f - yy - zz, xy + sz, xz - sy },
		{ xy - sz, 1.0f - xx - zz, yz + sx },
		{ xz + sy, yz - sx, 1.0f - xx - yy },
	};
	return m;
}
static inline nbo_transform xfmul(const nbo_transform& l, const nbo_transform& r) {  // nudge.cpp:1165-1175
	float3 p = add(rotate(rot(l.rotation), f3(r.position)), f3(l.position));
	Rot q = rotmul(rot(l.rotation), rot(r.rotation));
	nbo_transform t = { { p.x, p.y, p.z }, r.body, { q.v.x, q.v.y, q.v.z, q.s } };
	return t;
}

struct Pair { u32 lo, hi; };  // nudge.cpp:3495: lo = later in Morton order, hi = earlier

static std::vector<Pair> g_pairs;
static std::vector<u32> g_order;

// ---------- Morton (nudge.cpp:2606-2645) ----------
static inline void dilate3(u32 x, unsigned offset, u32& lo32, u32& hi32) {
	u32 lo24 = x & 0xff, hi24 = (x >> 8) & 0xff;
	lo24 = (lo24 | (lo24 << 8)) & 0x0f00f00f; hi24 = (hi24 | (hi24 << 8)) & 0x0f00f00f;
	lo24 = (lo24 | (lo24 << 4)) & 0xc30c30c3; hi24 = (hi24 | (hi24 << 4)) & 0xc30c30c3;
	lo24 = (lo24 | (lo24 << 2)) & 0x49249249; hi24 = (hi24 | (hi24 << 2)) & 0x49249249;
	lo32 = (lo24 << offset) | (hi24 << (24 + offset));
	hi32 = hi24 >> (8 - offset);
}
static inline u64 morton48(u32 x, u32 y, u32 z) {
	u32 lx, hx, ly, hy, lz, hz;
	dilate3(x, 2, lx, hx); dilate3(y, 1, ly, hy); dilate3(z, 0, lz, hz);
	return (u64)(lx | ly | lz) | ((u64)(hx | hy | hz) << 32);
}

// ---------- union-find; only the resulting partition matters (nudge.cpp:3505-3661, 3793-3952) ----------
struct Sets {
	std::vector<u32> parent;
	explicit Sets(u32 n) : parent(n) { for (u32 i = 0; i < n; ++i) parent[i] = i; }
	u32 find(u32 x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; }
	void join(u32 a, u32 b) {
		if (!a || !b) return;  // body 0 is the static world and is ignored (nudge.cpp:3517-3519)
		a = find(a); b = find(b);
		if (a != b) parent[a < b ? b : a] = a < b ? a : b;
	}
};

static int g_overflow = 0;

struct Out {
	nbo_contact* data; nbo_pair* bodies; u64* tags; u32* features; u32 count; u32 capacity;
	void push(const float p[3], float pen, const float n[3], u32 a, u32 b, u64 tag, u32 feature) {
		if (count >= capacity) { g_overflow = 1; return; }  // the reference never checks; the oracle refuses to overrun test buffers
		nbo_contact c = { { p[0], p[1], p[2] }, pen, { n[0], n[1], n[2] }, 0.5f };  // friction fixed: nudge.cpp:2105,2456,2515,2598
		data[count] = c; bodies[count].a = a; bodies[count].b = b; tags[count] = tag; features[count] = feature; ++count;
	}
};

struct BoxCtx {
	const nbo_box* colliders; const nbo_transform* transforms; const u32* tags;
};

struct RelFrame {  // shared head of pass 1 and pass 2: nudge.cpp:1228-1268 / 1465-1505
	float m[9];    // a_to_b: m[r*3+c], rows vx, vy, vz
};

static inline void relative_rotation(const float* qa, const float* qb, RelFrame& f) {
	float ax = qa[0], ay = qa[1], az = qa[2], as = qa[3];
	float bx = qb[0], by = qb[1], bz = qb[2], bs = qb[3];
	float3 t = cross(float3{ bx, by, bz }, float3{ ax, ay, az });
	float rx = ax*bs - bx*as - t.x;
	float ry = ay*bs - by*as - t.y;
	float rz = az*bs - bz*as - t.z;
	float rs = ax*bx + ay*by + az*bz + as*bs;
	float kx = rx + rx, ky = ry + ry, kz = rz + rz;
	float xx = kx*rx, yy = ky*ry, zz = kz*rz, xy = kx*ry, xz = kx*rz, yz = ky*rz, sx = kx*rs, sy = ky*rs, sz = kz*rs;
	f.m[0] = 1.0f - yy - zz; f.m[1] = xy + sz; f.m[2] = xz - sy;
	f.m[3] = xy - sz; f.m[4] = 1.0f - xx - zz; f.m[5] = yz + sx;
	f.m[6] = xz + sy; f.m[7] = yz - sx; f.m[8] = 1.0f - xx - yy;
}

struct Cand { u32 a, b; float pen; u32 feature; };

// Pass 1 for one pair: nudge.cpp:1195-1410.  Returns false if a face separates the boxes.
static bool box_box_faces(const BoxCtx& c, u32 a, u32 b, Cand& out) {
	const nbo_transform& ta = c.transforms[a]; const nbo_transform& tb = c.transforms[b];
	RelFrame f; relative_rotation(ta.rotation, tb.rotation, f);
	float vx_x = absf(f.m[0]), vx_y = absf(f.m[1]), vx_z = absf(f.m[2]);
	float vy_x = absf(f.m[3]), vy_y = absf(f.m[4]), vy_z = absf(f.m[5]);
	float vz_x = absf(f.m[6]), vz_y = absf(f.m[7]), vz_z = absf(f.m[8]);
	const float* sa = c.colliders[a].size; const float* sb = c.colliders[b].size;
	float pax = sb[0] + vx_x*sa[0] + vy_x*sa[1] + vz_x*sa[2];
	float pay = sb[1] + vx_y*sa[0] + vy_y*sa[1] + vz_y*sa[2];
	float paz = sb[2] + vx_z*sa[0] + vy_z*sa[1] + vz_z*sa[2];
	float pbx = sa[0] + vx_x*sb[0] + vx_y*sb[1] + vx_z*sb[2];
	float pby = sa[1] + vy_x*sb[0] + vy_y*sb[1] + vy_z*sb[2];
	float pbz = sa[2] + vz_x*sb[0] + vz_y*sb[1] + vz_z*sb[2];
	float3 delta = sub(f3(ta.position), f3(tb.position));
	float3 qa = { ta.rotation[0], ta.rotation[1], ta.rotation[2] }; float3 qb = { tb.rotation[0], tb.rotation[1], tb.rotation[2] };
	float3 t = cross(qb, delta); t = add(t, t);
	float3 u = cross(qb, t);
	float3 a_off = { u.x + delta.x - tb.rotation[3]*t.x, u.y + delta.y - tb.rotation[3]*t.y, u.z + delta.z - tb.rotation[3]*t.z };
	pax -= absf(a_off.x); pay -= absf(a_off.y); paz -= absf(a_off.z);
	t = cross(delta, qa); t = add(t, t);
	u = cross(qa, t);
	float3 b_off = { u.x - delta.x - ta.rotation[3]*t.x, u.y - delta.y - ta.rotation[3]*t.y, u.z - delta.z - ta.rotation[3]*t.z };
	pbx -= absf(b_off.x); pby -= absf(b_off.y); pbz -= absf(b_off.z);
	float payz = min1(pay, paz), pbyz = min1(pby, pbz);
	float pa = min1(pax, payz), pb = min1(pbx, pbyz);
	float p = min1(pa, pb);
	u32 aface = (payz == pa ? 1u : 0u) + (paz == pa ? 1u : 0u);
	u32 bface = (pbyz == pb ? 1u : 0u) + (pbz == pb ? 1u : 0u);
	bool swap = pa == p;  // nudge.cpp:1381-1387
	if (!(p > 0.0f)) return false;
	out.pen = p; out.feature = swap ? aface : bface;
	out.a = swap ? b : a; out.b = swap ? a : b;
	return true;
}

// Pass 2 for one surviving pair: nudge.cpp:1432-2136.  Emits face contacts or defers an edge candidate.
// Returns 0 = separated, 1 = face contacts emitted, 2 = edge candidate in `edge`.
static int box_box_face_or_edge(const BoxCtx& c, const Cand& in, Out& out, Cand& edge) {
	u32 a = in.a, b = in.b;
	const nbo_transform& ta = c.transforms[a]; const nbo_transform& tb = c.transforms[b];
	RelFrame f; relative_rotation(ta.rotation, tb.rotation, f);
	const float* a_to_b = f.m;
	const float* sa = c.colliders[a].size; const float* sb = c.colliders[b].size;
	float3 delta = sub(f3(ta.position), f3(tb.position));
	float3 qa = { ta.rotation[0], ta.rotation[1], ta.rotation[2] };
	float3 t = cross(delta, qa); t = add(t, t);
	float3 u = cross(qa, t);
	float b_offset[3] = { u.x - delta.x - ta.rotation[3]*t.x, u.y - delta.y - ta.rotation[3]*t.y, u.z - delta.z - ta.rotation[3]*t.z };
	float face_penetration = in.pen;

	float epa[9], epb[9];
	for (unsigned i = 0; i < 3; ++i) {  // nudge.cpp:1578-1640
		float acx = a_to_b[0*3 + i], acy = a_to_b[1*3 + i], acz = a_to_b[2*3 + i];
		float bcx = a_to_b[i*3 + 0], bcy = a_to_b[i*3 + 1], bcz = a_to_b[i*3 + 2];
		float ac2x = acx*acx, ac2y = acy*acy, ac2z = acz*acz;
		float bc2x = bcx*bcx, bc2y = bcy*bcy, bc2z = bcz*bcz;
		float aacx = absf(acx), aacy = absf(acy), aacz = absf(acz);
		float abcx = absf(bcx), abcy = absf(bcy), abcz = absf(bcz);
		float ra[3] = { ac2y + ac2z, ac2z + ac2x, ac2x + ac2y };
		float rb[3] = { bc2y + bc2z, bc2z + bc2x, bc2x + bc2y };
		for (unsigned k = 0; k < 3; ++k) {  // rsqrt | cmp_le -> NaN for degenerate axes (nudge.cpp:1611-1619)
			ra[k] = asf(asu(rsqrt(ra[k])) | (ra[k] <= 1e-3f ? 0xffffffffu : 0u));
			rb[k] = asf(asu(rsqrt(rb[k])) | (rb[k] <= 1e-3f ? 0xffffffffu : 0u));
		}
		float pa0 = aacy*sa[2] + aacz*sa[1], pa1 = aacz*sa[0] + aacx*sa[2], pa2 = aacx*sa[1] + aacy*sa[0];
		float pb0 = abcy*sb[2] + abcz*sb[1], pb1 = abcz*sb[0] + abcx*sb[2], pb2 = abcx*sb[1] + abcy*sb[0];
		float o0 = absf(acy*b_offset[2] - acz*b_offset[1]);
		float o1 = absf(acz*b_offset[0] - acx*b_offset[2]);
		float o2 = absf(acx*b_offset[1] - acy*b_offset[0]);
		epa[i*3 + 0] = (pa0 - o0) * ra[0]; epa[i*3 + 1] = (pa1 - o1) * ra[1]; epa[i*3 + 2] = (pa2 - o2) * ra[2];
		epb[i*3 + 0] = pb0 * rb[0]; epb[i*3 + 1] = pb1 * rb[1]; epb[i*3 + 2] = pb2 * rb[2];
	}
	u32 a_edge = 0, b_edge = 0;
	float penetration = face_penetration;
	for (unsigned i = 0; i < 3; ++i)
		for (unsigned j = 0; j < 3; ++j) {  // nudge.cpp:1647-1657
			float p = epa[i*3 + j] + epb[j*3 + i];
			bool m = penetration > p;
			penetration = min1(penetration, p);
			if (m) { a_edge = j; b_edge = i; }
		}
	bool is_edge = face_penetration > penetration + 1e-3f;  // nudge.cpp:1659-1661
	bool overlapping = penetration > 0.0f;
	if (!overlapping) return 0;

	if (is_edge) {  // nudge.cpp:2116-2135
		u32 at = c.tags[a], bt = c.tags[b];
		edge.pen = penetration;
		edge.feature = at > bt ? a_edge | (b_edge << 16) : b_edge | (a_edge << 16);
		edge.a = at > bt ? a : b; edge.b = at > bt ? b : a;
		return 2;
	}

	// ---- face-face: nudge.cpp:1678-2112 ----
	u32 a_face = in.feature;
	float dirs[4] = { absf(a_to_b[a_face*3 + 0]), absf(a_to_b[a_face*3 + 1]), absf(a_to_b[a_face*3 + 2]), 0.0f };
	float c0[3] = { a_to_b[0], a_to_b[3], a_to_b[6] };
	float c1[3] = { a_to_b[1], a_to_b[4], a_to_b[7] };
	float c2[3] = { a_to_b[2], a_to_b[5], a_to_b[8] };
	float max_dir[4] = { max1(dirs[0], dirs[0]), max1(dirs[2], dirs[0]), max1(dirs[1], dirs[0]), max1(dirs[3], dirs[0]) };
	unsigned dir_mask = 0;
	for (unsigned k = 0; k < 4; ++k) dir_mask |= (dirs[k] >= max_dir[k] ? 1u : 0u) << k;
	for (unsigned k = 0; k < 3; ++k) { c0[k] *= sb[0]; c1[k] *= sb[1]; c2[k] *= sb[2]; }
	u32 b_face = 0;
	float cc[3], dx[3], dy[3];
	if (dir_mask & 4) { memcpy(cc, c2, 12); memcpy(dx, c0, 12); memcpy(dy, c1, 12); b_face = 2; }
	else if (dir_mask & 2) { memcpy(cc, c1, 12); memcpy(dx, c2, 12); memcpy(dy, c0, 12); b_face = 1; }
	else { memcpy(cc, c0, 12); memcpy(dx, c1, 12); memcpy(dy, c2, 12); }
	unsigned b_positive_face_bit = ((asu(b_offset[a_face]) ^ asu(cc[a_face])) >> 31) << a_face;
	unsigned b_offset_neg = (asu(b_offset[a_face]) >> 31) << a_face;
	if (!b_positive_face_bit) for (unsigned k = 0; k < 3; ++k) cc[k] = negf(cc[k]);
	for (unsigned k = 0; k < 3; ++k) cc[k] += b_offset[k];

	// quads rows: (a.size[k], c[k], dx[k], dy[k]), nudge.cpp:1759-1773
	unsigned X = (a_face + 1) % 3, Y = (a_face + 2) % 3, Z = a_face;
	float sx = sa[X], sy = sa[Y], cx = cc[X], cy = cc[Y];
	float d0 = dx[X], d1 = dx[Y], d2 = dy[X], d3 = dy[Y];  // dxy = (dx.X, dx.Y, dy.X, dy.Y)

	float support_x[16], support_y[16], support_z[16];
	u32 support_tags[16];
	unsigned mask;
	{
		static const u32 npnp[4] = { SIGN, 0, SIGN, 0 }, pnpn[4] = { 0, SIGN, 0, SIGN }, nnpp[4] = { SIGN, SIGN, 0, 0 };
		bool mask0[4], mask1[4];
		float k0 = cx*d3 - cy*d2, k1 = cx*d1 - cy*d0, k2 = d0*d3 - d1*d2;  // nudge.cpp:1814-1815
		float ox = k0, oy = k1, delta_max = absf(k2);
		float sd0 = d0*sy, sd1 = d1*sx, sd2 = d2*sy, sd3 = d3*sx;           // nudge.cpp:1821
		for (unsigned l = 0; l < 4; ++l) {
			float corner0x = xorf(sx, pnpn[l]), corner0y = xorf(sy, nnpp[l]);
			float corner1x = cx + xorf(d0, npnp[l]) + xorf(d2, nnpp[l]);
			float corner1y = cy + xorf(d1, npnp[l]) + xorf(d3, nnpp[l]);
			float delta_x = ox + xorf(sd2, nnpp[l]) + xorf(sd3, npnp[l]);
			float delta_y = oy + xorf(sd0, nnpp[l]) + xorf(sd1, npnp[l]);
			bool inside_x = absf(corner1x) <= sx, inside_y = absf(corner1y) <= sy;
			mask0[l] = max1(absf(delta_x), absf(delta_y)) <= delta_max;
			mask1[l] = inside_x && inside_y;
			support_x[l] = corner0x; support_y[l] = corner0y;
			support_x[4 + l] = corner1x; support_y[4 + l] = corner1y;
		}
		// Don't allow edge intersections if both vertices are inside: nudge.cpp:1834-1836
		bool pre[8] = { mask0[3] && mask0[1], mask0[2] && mask0[0], mask0[0] && mask0[1], mask0[2] && mask0[3],
						mask1[1] && mask1[0], mask1[3] && mask1[2], mask1[2] && mask1[0], mask1[3] && mask1[1] };
		unsigned edge_axis_near = 0, edge_axis_far = 0;
		bool mask_a[4], mask_b[4];
		{
			float dxy[4] = { d0, d1, d2, d3 };
			float rdxy[4] = { 1.0f/d0, 1.0f/d1, 1.0f/d2, 1.0f/d3 };  // nudge.cpp:1849
			static const unsigned i0022[4] = { 0, 0, 2, 2 }, i1133[4] = { 1, 1, 3, 3 }, i2200[4] = { 2, 2, 0, 0 }, i3311[4] = { 3, 3, 1, 1 };
			for (unsigned l = 0; l < 4; ++l) {
				float offset_x = dxy[i0022[l]], offset_y = dxy[i1133[l]];
				float pivot_x = cx + xorf(dxy[i2200[l]], npnp[l]);
				float pivot_y = cy + xorf(dxy[i3311[l]], npnp[l]);
				float pos_x = asf((asu(offset_x) & SIGN) | asu(sx));  // copy sign: nudge.cpp:1858-1859
				float pos_y = asf((asu(offset_y) & SIGN) | asu(sy));
				float rx = rdxy[i0022[l]], ry = rdxy[i1133[l]];
				float near_x = (pos_x + pivot_x) * rx, far_x = (pos_x - pivot_x) * rx;
				float near_y = (pos_y + pivot_y) * ry, far_y = (pos_y - pivot_y) * ry;
				float ea = min1(1.0f, near_x), eb = min1(1.0f, far_x);
				if (ea > near_y) edge_axis_near |= 1u << l;
				if (eb > far_y) edge_axis_far |= 1u << l;
				ea = min1(ea, near_y); eb = min1(eb, far_y);
				float ax = pivot_x - offset_x * ea, ay = pivot_y - offset_y * ea;
				float bx = pivot_x + offset_x * eb, by = pivot_y + offset_y * eb;
				bool m = (ea + eb) > 0.0f;                  // make sure -a < b
				mask_a[l] = !(ea == 1.0f) && m;             // _mm_cmpneq_ps is unordered: true on NaN (nudge.cpp:328-330,1886)
				mask_b[l] = !(eb == 1.0f) && m;
				support_x[8 + l] = ax; support_y[8 + l] = ay;
				support_x[12 + l] = bx; support_y[12 + l] = by;
			}
		}
		mask = 0;
		for (unsigned l = 0; l < 4; ++l) {
			mask |= (mask0[l] ? 1u : 0u) << l;
			mask |= (mask1[l] ? 1u : 0u) << (4 + l);
			mask |= ((!pre[l] && mask_a[l]) ? 1u : 0u) << (8 + l);
			mask |= ((!pre[4 + l] && mask_b[l]) ? 1u : 0u) << (12 + l);
		}

		// vertex / edge labels: nudge.cpp:1902-1970
		unsigned a_sign_face_bit = b_offset_neg ? (1u << a_face) : 0;
		unsigned b_sign_face_bit = b_positive_face_bit ? 0 : (1u << b_face);
		unsigned a_vertices = 0x12003624u >> (3 - a_face);
		unsigned b_vertices = 0x00122436u >> (3 - b_face);
		unsigned a_face_bits = 0xffff0000u | a_sign_face_bit;
		unsigned b_face_bits = 0x0000ffffu | (b_sign_face_bit << 16);
		support_tags[0] = ((a_vertices >> 0) & 0x7) | a_face_bits;
		support_tags[1] = ((a_vertices >> 8) & 0x7) | a_face_bits;
		support_tags[2] = ((a_vertices >> 16) & 0x7) | a_face_bits;
		support_tags[3] = ((a_vertices >> 24) & 0x7) | a_face_bits;
		support_tags[4] = ((b_vertices << 16) & 0x70000) | b_face_bits;
		support_tags[5] = ((b_vertices << 8) & 0x70000) | b_face_bits;
		support_tags[6] = ((b_vertices >> 0) & 0x70000) | b_face_bits;
		support_tags[7] = ((b_vertices >> 8) & 0x70000) | b_face_bits;
		unsigned winding = signbit32(d0) | (signbit32(d1) << 1) | (signbit32(d2) << 2) | (signbit32(d3) << 3);
		unsigned near_e[4], far_e[4];
		for (unsigned l = 0; l < 4; ++l) {
			unsigned yn = (edge_axis_near >> l) & 1;
			near_e[l] = yn*2 + ((winding >> ((l < 2 ? 0 : 2) + yn)) & 1);
		}
		winding ^= 0xf;
		for (unsigned l = 0; l < 4; ++l) {
			unsigned yf = (edge_axis_far >> l) & 1;
			far_e[l] = yf*2 + ((winding >> ((l < 2 ? 0 : 2) + yf)) & 1);
		}
		u64 a_edge_map = 0x1200362424003612llu >> (3 - a_face);
		u64 b_edge_map = 0x2400361212003624llu >> (3 - b_face);
		unsigned face_bits = a_sign_face_bit | (a_sign_face_bit << 8) | (b_sign_face_bit << 16) | (b_sign_face_bit << 24);
		for (unsigned l = 0; l < 4; ++l) {
			unsigned b_edge_l = ((unsigned)((b_edge_map >> (l << 4)) & 0x0707) << 16) | face_bits;
			support_tags[8 + l] = (unsigned)((a_edge_map >> (near_e[l] << 4)) & 0x0707) | b_edge_l;
			support_tags[12 + l] = (unsigned)((a_edge_map >> (far_e[l] << 4)) & 0x0707) | b_edge_l;
		}
	}

	// z-plane through face b: nudge.cpp:1973-2019
	float penetrations[16];
	{
		float dxt[3] = { dx[X], dx[Y], dx[Z] }, dyt[3] = { dy[X], dy[Y], dy[Z] }, ct[3] = { cc[X], cc[Y], cc[Z] };
		float zn0 = dxt[1]*dyt[2] - dxt[2]*dyt[1];
		float zn1 = dxt[2]*dyt[0] - dxt[0]*dyt[2];
		float zn2 = dxt[0]*dyt[1] - dxt[1]*dyt[0];
		float dt = ct[0]*zn0 + ct[1]*zn1 + ct[2]*zn2;
		float inv = 1.0f / zn2;
		float plane0 = negf(zn0) * inv, plane1 = negf(zn1) * inv, plane2 = dt * inv;
		u32 z_sign = b_offset_neg ? SIGN : 0;
		float penetration_offset = sa[Z];
		unsigned penetration_mask = 0;
		for (unsigned i = 0; i < 16; ++i) {
			float x = support_x[i], y = support_y[i];
			float z = x*plane0 + y*plane1 + plane2;
			float pen = penetration_offset - xorf(z, z_sign);
			z += pen * xorf(0.5f, z_sign);
			if (pen > 0.0f) penetration_mask |= 1u << i;
			penetrations[i] = pen; support_z[i] = z;
		}
		mask &= penetration_mask;
	}

	// a to world: nudge.cpp:2028-2056 (note the diagonal is built as -((p + q) - 1))
	float w0[3], w1[3], w2[3];
	{
		float qx = ta.rotation[0], qy = ta.rotation[1], qz = ta.rotation[2], qs = ta.rotation[3];
		float kx = qx + qx, ky = qy + qy, kz = qz + qz, ks = negf(qs + qs);
		w0[0] = negf((ky*qy + kz*qz) - 1.0f); w0[1] = (kx*qy + kz*qs) - 0.0f; w0[2] = (kx*qz + ks*qy) - 0.0f;
		w1[0] = (kx*qy + ks*qz) - 0.0f; w1[1] = negf((kz*qz + kx*qx) - 1.0f); w1[2] = (ky*qz + kx*qs) - 0.0f;
		w2[0] = (kx*qz + ky*qs) - 0.0f; w2[1] = (ky*qz + ks*qx) - 0.0f; w2[2] = negf((kx*qx + ky*qy) - 1.0f);
	}
	// support arrays are in a's face frame; local X/Y/Z = box axes (a_face+1)%3, (a_face+2)%3, a_face (nudge.cpp:2021-2026)
	unsigned afi = (a_face ^ 1) ^ (a_face >> 1);
	const float* sup[3] = { support_x, support_y, support_z };
	const float* spx = sup[(afi + 1) % 3]; const float* spy = sup[(afi + 2) % 3]; const float* spz = sup[afi];
	const float* wsel = a_face == 0 ? w0 : (a_face == 1 ? w1 : w2);
	float wn[3] = { wsel[0], wsel[1], wsel[2] };
	if (b_offset_neg) for (unsigned k = 0; k < 3; ++k) wn[k] = negf(wn[k]);
	u32 a_body = ta.body, b_body = tb.body;
	u32 a_tag = c.tags[a], b_tag = c.tags[b];
	unsigned tag_swap = 0;
	if (b_tag > a_tag) {  // nudge.cpp:2074-2087
		u32 tt = a_tag; a_tag = b_tag; b_tag = tt;
		u32 tb2 = a_body; a_body = b_body; b_body = tb2;
		tag_swap = 16;
		for (unsigned k = 0; k < 3; ++k) wn[k] = negf(wn[k]);
	}
	u64 high_tag = (u64)a_tag | ((u64)b_tag << 32);
	while (mask) {
		unsigned index = __builtin_ctz(mask);
		mask &= mask - 1;
		float wp[3];
		for (unsigned k = 0; k < 3; ++k)
			wp[k] = w0[k]*spx[index] + w1[k]*spy[index] + w2[k]*spz[index] + ta.position[k];
		u32 st = support_tags[index];
		u32 feature = tag_swap ? ((st >> 16) | (st << 16)) : st;  // nudge.cpp:2108 (for tag_swap = 0 both terms equal st)
		out.push(wp, penetrations[index], wn, a_body, b_body, high_tag, feature);
	}
	return 1;
}

// Pass 3 for one edge pair: nudge.cpp:2157-2479.
static void box_box_edge(const BoxCtx& c, const Cand& in, Out& out) {
	u32 a = in.a, b = in.b;
	const nbo_transform& ta = c.transforms[a]; const nbo_transform& tb = c.transforms[b];
	float ab[3][3], bb[3][3];
	for (int w = 0; w < 2; ++w) {
		const float* q = w ? tb.rotation : ta.rotation;
		float kx = q[0] + q[0], ky = q[1] + q[1], kz = q[2] + q[2];
		float xx = kx*q[0], yy = ky*q[1], zz = kz*q[2], xy = kx*q[1], xz = kx*q[2], yz = ky*q[2], sx = kx*q[3], sy = ky*q[3], sz = kz*q[3];
		float (*m)[3] = w ? bb : ab;
		m[0][0] = 1.0f - yy - zz; m[0][1] = xy + sz; m[0][2] = xz - sy;
		m[1][0] = xy - sz; m[1][1] = 1.0f - xx - zz; m[1][2] = yz + sx;
		m[2][0] = xz + sy; m[2][1] = yz - sx; m[2][2] = 1.0f - xx - yy;
	}
	u32 edge = in.feature;
	// blendv on shifted bits (NUDGE_NATIVE_BLENDV32 path, nudge.cpp:2257-2278)
	unsigned ua = (edge & 2) ? 2 : ((edge & 1) ? 1 : 0);
	unsigned ub = (edge & (2u << 16)) ? 2 : ((edge & (1u << 16)) ? 1 : 0);
	float3 u = { ab[ua][0], ab[ua][1], ab[ua][2] };
	float3 v = { bb[ub][0], bb[ub][1], bb[ub][2] };
	float3 n = cross(u, v);
	float3 delta = sub(f3(tb.position), f3(ta.position));
	u32 flip = asu(n.x*delta.x + n.y*delta.y + n.z*delta.z) & SIGN;
	n.x = xorf(n.x, flip); n.y = xorf(n.y, flip); n.z = xorf(n.z, flip);
	float sa[3] = { c.colliders[a].size[0], c.colliders[a].size[1], c.colliders[a].size[2] };
	float sb[3] = { c.colliders[b].size[0], c.colliders[b].size[1], c.colliders[b].size[2] };
	u32 asg[3], bsg[3];
	for (unsigned k = 0; k < 3; ++k) {
		asg[k] = asu(ab[k][0]*n.x + ab[k][1]*n.y + ab[k][2]*n.z) & SIGN;
		bsg[k] = asu(bb[k][0]*n.x + bb[k][1]*n.y + bb[k][2]*n.z) & SIGN;
	}
	u32 edge_x = (asg[0] >> 31) | ((bsg[0] ^ SIGN) >> 15);
	u32 edge_y = (asg[1] >> 30) | ((bsg[1] ^ SIGN) >> 14);
	u32 edge_z = (asg[2] >> 29) | ((bsg[2] ^ SIGN) >> 13);
	u32 elo = edge & 0xffff, ehi = edge >> 16;  // per-16-bit-lane (e + 1) + (e >> 1) = 1 << e for e in 0..2 (nudge.cpp:2381)
	u32 edge_w = (((elo + 1) + (elo >> 1)) & 0xffff) | ((((ehi + 1) + (ehi >> 1)) & 0xffff) << 16);
	u32 tag_hi = edge_x | edge_y | edge_z | edge_w;
	u32 tag_lo = tag_hi & ~edge_w;
	u32 tag = tag_lo | (tag_hi << 8);
	for (unsigned k = 0; k < 3; ++k) { sa[k] = xorf(sa[k], asg[k]); sb[k] = xorf(sb[k], bsg[k]); }
	for (unsigned k = 0; k < 3; ++k) for (unsigned j = 0; j < 3; ++j) { ab[k][j] *= sa[k]; bb[k][j] *= sb[k]; }
	float ca[3], cb[3], o[3];
	for (unsigned j = 0; j < 3; ++j) {
		ca[j] = ab[0][j] + ab[1][j] + ab[2][j] + ta.position[j];
		cb[j] = bb[0][j] + bb[1][j] + bb[2][j] - tb.position[j];  // negated on purpose (nudge.cpp:2428)
		o[j] = ca[j] + cb[j];
	}
	float ia = u.x*u.x + u.y*u.y + u.z*u.z;
	float ib = u.x*v.x + u.y*v.y + u.z*v.z;
	float ic = v.x*v.x + v.y*v.y + v.z*v.z;
	float id = o[0]*u.x + o[1]*u.y + o[2]*u.z;
	float ie = o[0]*v.x + o[1]*v.y + o[2]*v.z;
	float ir = 0.5f / (ia*ic - ib*ib);
	float s_a = (ib*ie - ic*id) * ir;
	float s_b = (ia*ie - ib*id) * ir;
	float p[3] = { (ca[0] - cb[0])*0.5f + u.x*s_a + v.x*s_b,
				   (ca[1] - cb[1])*0.5f + u.y*s_a + v.y*s_b,
				   (ca[2] - cb[2])*0.5f + u.z*s_a + v.z*s_b };
	float fn = rsqrt(n.x*n.x + n.y*n.y + n.z*n.z);  // nudge.cpp:842-847, 2453
	float nn[3] = { n.x*fn, n.y*fn, n.z*fn };
	out.push(p, in.pen, nn, ta.body, tb.body, (u64)c.tags[a] | ((u64)c.tags[b] << 32), tag);
}

static void box_box_collide(const std::vector<Pair>& pairs, const BoxCtx& c, Out& out) {  // nudge.cpp:1177-2487
	std::vector<Cand> stage1, edges;
	for (size_t i = 0; i < pairs.size(); ++i) {
		Cand cd;
		if (box_box_faces(c, pairs[i].lo, pairs[i].hi, cd)) stage1.push_back(cd);  // a = low half, b = high half (nudge.cpp:1202-1203)
	}
	for (size_t i = 0; i < stage1.size(); ++i) {
		Cand e;
		if (box_box_face_or_edge(c, stage1[i], out, e) == 2) edges.push_back(e);
	}
	for (size_t i = 0; i < edges.size(); ++i)
		box_box_edge(c, edges[i], out);
}

static bool sphere_sphere(const nbo_sphere& a, const nbo_sphere& b, const nbo_transform& ta, const nbo_transform& tb, float p[3], float& pen, float n[3]) {
	// nudge.cpp:2489-2521
	float r = a.radius + b.radius;
	float3 dp = sub(f3(tb.position), f3(ta.position));
	float l2 = dot(dp, dp);
	if (l2 > r*r) return false;
	float3 nn;
	float l = sqrtf(l2);
	if (l2 > 1e-4f) nn = mul(dp, 1.0f / l);
	else { nn.x = 1.0f; nn.y = 0.0f; nn.z = 0.0f; }
	float3 pp = add(f3(ta.position), mul(nn, l - b.radius));
	p[0] = pp.x; p[1] = pp.y; p[2] = pp.z; pen = r - l; n[0] = nn.x; n[1] = nn.y; n[2] = nn.z;
	return true;
}

static bool box_sphere(const nbo_box& a, const nbo_sphere& b, const nbo_transform& ta, const nbo_transform& tb, float p[3], float& pen, float n[3]) {
	// nudge.cpp:2523-2604
	Rot a_to_world = rot(ta.rotation);
	Rot world_to_a = a_to_world; world_to_a.v.x = -world_to_a.v.x; world_to_a.v.y = -world_to_a.v.y; world_to_a.v.z = -world_to_a.v.z;
	float3 offset_b = rotate(world_to_a, sub(f3(tb.position), f3(ta.position)));
	float dx = fabsf(offset_b.x), dy = fabsf(offset_b.y), dz = fabsf(offset_b.z);
	float w = a.size[0] + b.radius, h = a.size[1] + b.radius, d = a.size[2] + b.radius;
	if (dx >= w || dy >= h || dz >= d) return false;
	float3 nn; float penetration; float r = b.radius;
	unsigned outside_x = dx > a.size[0], outside_y = dy > a.size[1], outside_z = dz > a.size[2];
	if (outside_x + outside_y + outside_z >= 2) {
		float3 corner = {
			outside_x ? (offset_b.x > 0.0f ? a.size[0] : -a.size[0]) : offset_b.x,
			outside_y ? (offset_b.y > 0.0f ? a.size[1] : -a.size[1]) : offset_b.y,
			outside_z ? (offset_b.z > 0.0f ? a.size[2] : -a.size[2]) : offset_b.z,
		};
		float3 dp = sub(offset_b, corner);
		float l2 = dot(dp, dp);
		if (l2 > r*r) return false;
		float l = sqrtf(l2);
		float m = 1.0f / l;
		nn = mul(dp, m);
		penetration = r - l;
	}
	else if (w - dx < h - dy && w - dx < d - dz) { nn.x = offset_b.x > 0.0f ? 1.0f : -1.0f; nn.y = 0.0f; nn.z = 0.0f; penetration = w - dx; }
	else if (h - dy < d - dz) { nn.x = 0.0f; nn.y = offset_b.y > 0.0f ? 1.0f : -1.0f; nn.z = 0.0f; penetration = h - dy; }
	else { nn.x = 0.0f; nn.y = 0.0f; nn.z = offset_b.z > 0.0f ? 1.0f : -1.0f; penetration = d - dz; }
	float3 pp = sub(offset_b, mul(nn, r));
	pp = add(rotate(a_to_world, pp), f3(ta.position));
	nn = rotate(a_to_world, nn);
	p[0] = pp.x; p[1] = pp.y; p[2] = pp.z; pen = penetration; n[0] = nn.x; n[1] = nn.y; n[2] = nn.z;
	return true;
}

struct Key {  // widened 64-bit contact tag: major = pair (B then A), minor = feature
	u64 pair; u32 feature;
	bool operator<(const Key& o) const { return pair < o.pair || (pair == o.pair && feature < o.feature); }
	bool operator==(const Key& o) const { return pair == o.pair && feature == o.feature; }
};

}  // namespace

struct nbo_impulse_data {  // nudge.cpp:4011-4019
	std::vector<u32> sorted_contacts;
	std::vector<nbo_impulse> culled_data;
	std::vector<u64> culled_tags;
	std::vector<u32> culled_features;
	std::vector<nbo_impulse> data;
};

struct InertiaTransform { float xx, yy, zz, unused0, xy, xz, yz, unused1; };  // nudge.cpp:966-975

struct nbo_constraint_data {  // nudge.cpp:4160-4168
	u32 contact_count;
	std::vector<InertiaTransform> momentum_to_velocity;
	std::vector<u32> constraint_to_contact;  // [batches*8]
	std::vector<u32> a, b;                   // [batches*8]
	std::vector<float> rows;                 // [batches*8][39], member order of nudge.cpp:907-957
	std::vector<float> states;               // [batches*8][3]
	u32 batches;
};

enum {  // nudge.cpp:907-957
	PA_Z, PA_X, PA_Y, PB_Z, PB_X, PB_Y, N_X, U_X, V_X, N_Y, U_Y, V_Y, N_Z, U_Z, V_Z, BIAS, FRICTION, NVTNI, FC_X, FC_Y, FC_Z,
	NA_X, NA_Y, NA_Z, NB_X, NB_Y, NB_Z, UA_X, UA_Y, UA_Z, VA_X, VA_Y, VA_Z, UB_X, UB_Y, UB_Z, VB_X, VB_Y, VB_Z, ROW_FLOATS
};

extern "C" {

void nbo_set_ftz_daz(int on) {
	_MM_SET_FLUSH_ZERO_MODE(on ? _MM_FLUSH_ZERO_ON : _MM_FLUSH_ZERO_OFF);
	_MM_SET_DENORMALS_ZERO_MODE(on ? _MM_DENORMALS_ZERO_ON : _MM_DENORMALS_ZERO_OFF);
}

void nbo_rcp(const float* x, float* y, u32 n) { for (u32 i = 0; i < n; ++i) y[i] = rcp(x[i]); }
void nbo_rsqrt(const float* x, float* y, u32 n) { for (u32 i = 0; i < n; ++i) y[i] = rsqrt(x[i]); }

void nbo_collide(nbo_active_bodies* active_bodies, nbo_contact_data* contacts, const nbo_body_data* bodies_p, const nbo_collider_data* colliders_p, const nbo_connections* connections_p) {
	const nbo_body_data& bodies = *bodies_p; const nbo_collider_data& colliders = *colliders_p; const nbo_connections& body_connections = *connections_p;
	contacts->count = 0; contacts->sleeping_count = 0; active_bodies->count = 0;  // nudge.cpp:3001-3003
	g_overflow = 0;
	const u32 nboxes = colliders.boxes.count, nspheres = colliders.spheres.count, count = nboxes + nspheres;

	struct AABB { float mn[4], mx[4]; };
	std::vector<AABB> aabb(count);
	std::vector<nbo_transform> transforms(count);
	std::vector<u32> collider_tags(count), collider_bodies(count);

	for (u32 i = 0; i < nboxes; ++i) {  // nudge.cpp:3021-3054
		nbo_transform t = colliders.boxes.transforms[i];
		t = xfmul(bodies.transforms[t.body], t);
		Mat3 m = matrix(rot(t.rotation));
		const float* s = colliders.boxes.data[i].size;
		m.c0 = mul(m.c0, s[0]); m.c1 = mul(m.c1, s[1]); m.c2 = mul(m.c2, s[2]);
		float3 size = { fabsf(m.c0.x) + fabsf(m.c1.x) + fabsf(m.c2.x), fabsf(m.c0.y) + fabsf(m.c1.y) + fabsf(m.c2.y), fabsf(m.c0.z) + fabsf(m.c1.z) + fabsf(m.c2.z) };
		AABB b = { { t.position[0] - size.x, t.position[1] - size.y, t.position[2] - size.z, 0.0f }, { t.position[0] + size.x, t.position[1] + size.y, t.position[2] + size.z, 0.0f } };
		transforms[i] = t; aabb[i] = b;
		collider_tags[i] = colliders.boxes.tags[i]; collider_bodies[i] = colliders.boxes.transforms[i].body;
	}
	for (u32 i = 0; i < nspheres; ++i) {  // nudge.cpp:3056-3079
		nbo_transform t = colliders.spheres.transforms[i];
		t = xfmul(bodies.transforms[t.body], t);
		float r = colliders.spheres.data[i].radius;
		AABB b = { { t.position[0] - r, t.position[1] - r, t.position[2] - r, 0.0f }, { t.position[0] + r, t.position[1] + r, t.position[2] + r, 0.0f } };
		transforms[nboxes + i] = t; aabb[nboxes + i] = b;
		collider_tags[nboxes + i] = colliders.spheres.tags[i]; collider_bodies[nboxes + i] = colliders.spheres.transforms[i].body;
	}

	g_pairs.clear(); g_order.assign(count, 0);
	std::vector<Pair>& pairs = g_pairs;
	if (count) {
		// scene bounds over AABB mins and Morton scale: nudge.cpp:3087-3100
		float smin[4], smax[4];
		for (int k = 0; k < 4; ++k) smin[k] = smax[k] = aabb[0].mn[k];
		for (u32 i = 1; i < count; ++i)
			for (int k = 0; k < 4; ++k) { smin[k] = min1(smin[k], aabb[i].mn[k]); smax[k] = max1(smax[k], aabb[i].mn[k]); }
		float sc[4];
		for (int k = 0; k < 4; ++k) sc[k] = 65535.0f * rcp(smax[k] - smin[k]);
		float A[4] = { min1(sc[0], sc[2]), min1(sc[1], sc[2]), min1(sc[2], sc[0]), min1(sc[2], sc[1]) };
		float L[4] = { min1(A[0], A[1]), min1(A[1], A[0]), min1(A[2], A[3]), min1(A[3], A[2]) };
		float smin_scaled[3] = { smin[0] * L[0], smin[1] * L[1], smin[2] * L[2] };
		// The reference multiplies SoA lane j by L[j & 3] (nudge.cpp:3143-3145); all four are the same min.
		std::vector<u64> codes(count);
		for (u32 i = 0; i < count; ++i) {
			float s = L[i & 3];
			u32 x = (u32)toint(msub(aabb[i].mn[0], s, smin_scaled[0]));
			u32 y = (u32)toint(msub(aabb[i].mn[1], s, smin_scaled[1]));
			u32 z = (u32)toint(msub(aabb[i].mn[2], s, smin_scaled[2]));
			codes[i] = morton48(x, y, z);
		}
		std::vector<u32>& order = g_order;
		for (u32 i = 0; i < count; ++i) order[i] = i;
		std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) { return codes[a] < codes[b]; });  // nudge.cpp:3165 (stable LSD radix)
		std::vector<u32> rank(count);
		for (u32 i = 0; i < count; ++i) rank[order[i]] = i;

		// All strictly overlapping AABB pairs (the set nudge.cpp:3275-3489 produces), by sort-and-sweep on x.
		std::vector<u32> by_x(count);
		for (u32 i = 0; i < count; ++i) by_x[i] = i;
		std::sort(by_x.begin(), by_x.end(), [&](u32 a, u32 b) { return aabb[a].mn[0] < aabb[b].mn[0] || (aabb[a].mn[0] == aabb[b].mn[0] && a < b); });
		for (u32 ii = 0; ii < count; ++ii) {
			u32 i = by_x[ii];
			const AABB& A0 = aabb[i];
			for (u32 jj = ii + 1; jj < count; ++jj) {
				u32 j = by_x[jj];
				const AABB& B0 = aabb[j];
				if (!(A0.mx[0] > B0.mn[0])) break;
				bool hit = B0.mx[0] > A0.mn[0] && A0.mx[0] > B0.mn[0] && B0.mx[1] > A0.mn[1] && A0.mx[1] > B0.mn[1] && B0.mx[2] > A0.mn[2] && A0.mx[2] > B0.mn[2];
				if (hit) {
					Pair p = rank[i] < rank[j] ? Pair{ j, i } : Pair{ i, j };  // lo = later in Morton order (nudge.cpp:3495)
					pairs.push_back(p);
				}
			}
		}
		std::sort(pairs.begin(), pairs.end(), [](const Pair& a, const Pair& b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); });  // nudge.cpp:3498
	}
	std::vector<Pair> sorted_pairs_snapshot = pairs;

	// coarse islands: nudge.cpp:3500-3703
	std::vector<Pair> live;
	{
		Sets sets(bodies.count);
		for (u32 i = 0; i < body_connections.count; ++i) sets.join(body_connections.data[i].a, body_connections.data[i].b);
		for (size_t i = 0; i < pairs.size(); ++i) sets.join(collider_bodies[pairs[i].lo], collider_bodies[pairs[i].hi]);
		std::vector<uint8_t> act(bodies.count, 0);
		for (u32 i = 1; i < bodies.count; ++i) if (bodies.idle_counters[i] != 0xff) act[sets.find(i)] = 1;
		for (size_t i = 0; i < pairs.size(); ++i) {
			u32 a = collider_bodies[pairs[i].lo], b = collider_bodies[pairs[i].hi];
			if (a == b) continue;
			u32 set = a ? sets.find(a) : sets.find(b);  // sets[0] = 0 and OR of equal set ids (nudge.cpp:3663,3688)
			if (act[set]) live.push_back(pairs[i]);
			else {
				u64 ta = collider_tags[pairs[i].lo], tb = collider_tags[pairs[i].hi];
				if (contacts->sleeping_count < contacts->capacity) contacts->sleeping_pairs[contacts->sleeping_count++] = ta > tb ? ta | (tb << 32) : tb | (ta << 32);  // nudge.cpp:3697
				else g_overflow = 1;
			}
		}
	}

	// partition by shape type: nudge.cpp:3705-3751
	std::vector<Pair> bucket[4];
	for (size_t i = 0; i < live.size(); ++i) {
		unsigned ab = (live[i].lo >= nboxes ? 1 : 0) | (live[i].hi >= nboxes ? 2 : 0);
		bucket[ab].push_back(live[i]);
	}
	for (size_t i = 0; i < bucket[2].size(); ++i) std::swap(bucket[2][i].lo, bucket[2][i].hi);

	Out out = { contacts->data, contacts->bodies, contacts->tags, contacts->features, 0, contacts->capacity };
	BoxCtx bc = { colliders.boxes.data, transforms.data(), collider_tags.data() };
	box_box_collide(bucket[0], bc, out);  // nudge.cpp:3753

	for (int k = 1; k <= 2; ++k)  // nudge.cpp:3756-3769
		for (size_t i = 0; i < bucket[k].size(); ++i) {
			u32 a = bucket[k][i].hi, b = bucket[k][i].lo;  // a = box, b = sphere (global collider index)
			float p[3], n[3], pen;
			if (box_sphere(colliders.boxes.data[a], colliders.spheres.data[b - nboxes], transforms[a], transforms[b], p, pen, n))
				out.push(p, pen, n, transforms[a].body, transforms[b].body, (u64)collider_tags[a] | ((u64)collider_tags[b] << 32), 0);
		}
	for (size_t i = 0; i < bucket[3].size(); ++i) {  // nudge.cpp:3772-3786
		u32 a = bucket[3][i].hi, b = bucket[3][i].lo;
		float p[3], n[3], pen;
		if (sphere_sphere(colliders.spheres.data[a - nboxes], colliders.spheres.data[b - nboxes], transforms[a], transforms[b], p, pen, n))
			out.push(p, pen, n, transforms[a].body, transforms[b].body, (u64)collider_tags[a] | ((u64)collider_tags[b] << 32), 0);
	}
	contacts->count = out.count;

	// fine islands: nudge.cpp:3788-4006
	{
		Sets sets(bodies.count);
		for (u32 i = 0; i < body_connections.count; ++i) sets.join(body_connections.data[i].a, body_connections.data[i].b);
		for (u32 i = 0; i < contacts->count; ++i) sets.join(contacts->bodies[i].a, contacts->bodies[i].b);
		std::vector<uint8_t> act(bodies.count, 0);
		for (u32 i = 1; i < bodies.count; ++i) if (bodies.idle_counters[i] != 0xff) act[sets.find(i)] = 1;
		for (u32 i = 1; i < bodies.count; ++i) if (act[sets.find(i)]) active_bodies->indices[active_bodies->count++] = i;
		u32 removed = 0;
		for (u32 i = 0; i < contacts->count; ) {
			u32 a = contacts->bodies[i].a, b = contacts->bodies[i].b;
			u64 tag = contacts->tags[i];
			u32 span = 0;
			do { ++span; } while (i + span < contacts->count && contacts->tags[i + span] == tag);
			u32 set = a ? sets.find(a) : sets.find(b);
			if (act[set]) {
				for (u32 j = 0; j < span; ++j) {
					contacts->tags[i + j - removed] = contacts->tags[i + j]; contacts->features[i + j - removed] = contacts->features[i + j];
					contacts->data[i + j - removed] = contacts->data[i + j]; contacts->bodies[i + j - removed] = contacts->bodies[i + j];
				}
			}
			else { contacts->sleeping_pairs[contacts->sleeping_count++] = tag; removed += span; }
			i += span;
		}
		contacts->count -= removed;
	}
	std::sort(contacts->sleeping_pairs, contacts->sleeping_pairs + contacts->sleeping_count);  // nudge.cpp:4008
	g_pairs = sorted_pairs_snapshot;
}

int nbo_last_overflow(void) { return g_overflow; }
uint32_t nbo_last_pair_count(void) { return (uint32_t)g_pairs.size(); }
void nbo_last_pairs(uint32_t* lo, uint32_t* hi) { for (size_t i = 0; i < g_pairs.size(); ++i) { lo[i] = g_pairs[i].lo; hi[i] = g_pairs[i].hi; } }
void nbo_last_morton_order(uint32_t* s) { for (size_t i = 0; i < g_order.size(); ++i) s[i] = g_order[i]; }

nbo_impulse_data* nbo_read_cached_impulses(const nbo_contact_cache* cache_p, const nbo_contact_data* contacts_p) {  // nudge.cpp:4021-4108
	const nbo_contact_cache& contact_cache = *cache_p; const nbo_contact_data& contacts = *contacts_p;
	nbo_impulse_data* data = new nbo_impulse_data;
	data->sorted_contacts.resize(contacts.count);
	for (u32 i = 0; i < contacts.count; ++i) data->sorted_contacts[i] = i;
	// two stable 32-bit key sorts (low then high, nudge.cpp:4031-4043) == one stable sort on the whole tag
	std::stable_sort(data->sorted_contacts.begin(), data->sorted_contacts.end(), [&](u32 x, u32 y) {
		Key kx = { contacts.tags[x], contacts.features[x] }, ky = { contacts.tags[y], contacts.features[y] };
		return kx < ky;
	});
	data->data.resize(contacts.count);
	u32 cached_contact_offset = 0, sleeping_pair_offset = 0;
	for (u32 i = 0; i < contacts.count; ++i) {
		u32 index = data->sorted_contacts[i];
		Key tag = { contacts.tags[index], contacts.features[index] };
		nbo_impulse cached_impulse = {};
		while (cached_contact_offset < contact_cache.count) {
			Key cached_tag = { contact_cache.tags[cached_contact_offset], contact_cache.features[cached_contact_offset] };
			if (!(cached_tag < tag)) break;
			u64 cached_pair = cached_tag.pair;
			while (sleeping_pair_offset < contacts.sleeping_count && contacts.sleeping_pairs[sleeping_pair_offset] < cached_pair) ++sleeping_pair_offset;
			if (sleeping_pair_offset < contacts.sleeping_count && contacts.sleeping_pairs[sleeping_pair_offset] == cached_pair) {
				data->culled_data.push_back(contact_cache.data[cached_contact_offset]);
				data->culled_tags.push_back(contact_cache.tags[cached_contact_offset]);
				data->culled_features.push_back(contact_cache.features[cached_contact_offset]);
			}
			++cached_contact_offset;
		}
		if (cached_contact_offset < contact_cache.count) {
			Key cached_tag = { contact_cache.tags[cached_contact_offset], contact_cache.features[cached_contact_offset] };
			if (cached_tag == tag) cached_impulse = contact_cache.data[cached_contact_offset];
		}
		data->data[index] = cached_impulse;
	}
	for (; cached_contact_offset < contact_cache.count && sleeping_pair_offset < contacts.sleeping_count; ) {  // nudge.cpp:4085-4101
		u64 a = contact_cache.tags[cached_contact_offset], b = contacts.sleeping_pairs[sleeping_pair_offset];
		if (a < b) ++cached_contact_offset;
		else if (a == b) {
			data->culled_data.push_back(contact_cache.data[cached_contact_offset]);
			data->culled_tags.push_back(contact_cache.tags[cached_contact_offset]);
			data->culled_features.push_back(contact_cache.features[cached_contact_offset]);
			++cached_contact_offset;
		}
		else ++sleeping_pair_offset;
	}
	return data;
}

void nbo_write_cached_impulses(nbo_contact_cache* contact_cache, const nbo_contact_data* contacts_p, nbo_impulse_data* ci) {  // nudge.cpp:4110-4158
	const nbo_contact_data& contacts = *contacts_p;
	u32 culled_count = (u32)ci->culled_tags.size();
	contact_cache->count = contacts.count + culled_count;
	u32 i = 0, j = 0, k = 0;
	while (i < contacts.count && j < culled_count) {
		u32 index = ci->sorted_contacts[i];
		Key a = { contacts.tags[index], contacts.features[index] }, b = { ci->culled_tags[j], ci->culled_features[j] };
		if (a < b) { contact_cache->tags[k] = a.pair; contact_cache->features[k] = a.feature; contact_cache->data[k] = ci->data[index]; ++i; }
		else { contact_cache->tags[k] = b.pair; contact_cache->features[k] = b.feature; contact_cache->data[k] = ci->culled_data[j]; ++j; }
		++k;
	}
	for (; i < contacts.count; ++i, ++k) {
		u32 index = ci->sorted_contacts[i];
		contact_cache->tags[k] = contacts.tags[index]; contact_cache->features[k] = contacts.features[index]; contact_cache->data[k] = ci->data[index];
	}
	for (; j < culled_count; ++j, ++k) {
		contact_cache->tags[k] = ci->culled_tags[j]; contact_cache->features[k] = ci->culled_features[j]; contact_cache->data[k] = ci->culled_data[j];
	}
}

nbo_constraint_data* nbo_setup_contact_constraints(const nbo_active_bodies*, const nbo_contact_data* contacts_p, const nbo_body_data* bodies_p, nbo_impulse_data* contact_impulses) {
	// nudge.cpp:4170-4638
	const nbo_contact_data& contacts = *contacts_p; const nbo_body_data& bodies = *bodies_p;
	const std::vector<u32>& contact_order = contact_impulses->sorted_contacts;
	nbo_constraint_data* data = new nbo_constraint_data;
	data->contact_count = contacts.count;
	data->momentum_to_velocity.resize(bodies.count);
	InertiaTransform* momentum_to_velocity = data->momentum_to_velocity.data();
	for (u32 i = 0; i < bodies.count; ++i) {  // nudge.cpp:4182-4199
		Mat3 m = matrix(rot(bodies.transforms[i].rotation));
		const float* ii = bodies.properties[i].inertia_inverse;
		InertiaTransform t = {};
		t.xx = ii[0]*m.c0.x*m.c0.x + ii[1]*m.c1.x*m.c1.x + ii[2]*m.c2.x*m.c2.x;
		t.yy = ii[0]*m.c0.y*m.c0.y + ii[1]*m.c1.y*m.c1.y + ii[2]*m.c2.y*m.c2.y;
		t.zz = ii[0]*m.c0.z*m.c0.z + ii[1]*m.c1.z*m.c1.z + ii[2]*m.c2.z*m.c2.z;
		t.xy = ii[0]*m.c0.x*m.c0.y + ii[1]*m.c1.x*m.c1.y + ii[2]*m.c2.x*m.c2.y;
		t.xz = ii[0]*m.c0.x*m.c0.z + ii[1]*m.c1.x*m.c1.z + ii[2]*m.c2.x*m.c2.z;
		t.yz = ii[0]*m.c0.y*m.c0.z + ii[1]*m.c1.y*m.c1.z + ii[2]*m.c2.y*m.c2.z;
		momentum_to_velocity[i] = t;
		bodies.momentum[i].unused0 = bodies.properties[i].mass_inverse;
	}

	// Scheduler: nudge.cpp:4206-4340, restated literally with 32-bit body ids.
	struct Slot { u32 a[8], b[8], idx[8]; u32 filled; };
	std::vector<std::vector<u32> > batches;  // 8 contact indices each
	{
		static const unsigned bucket_count = 16;
		std::vector<Slot> vacant[bucket_count];
		for (u32 i = 0; i < contacts.count; ++i) {
			u32 index = contact_order[i];
			nbo_pair bp = contacts.bodies[index];
			std::vector<Slot>& v = vacant[i % bucket_count];
			u32 ca = bp.a ? bp.a : bp.b, cb = bp.b ? bp.b : bp.a;  // ignore dependencies on body 0
			size_t j = 0;
			for (;; ++j) {
				if (j == v.size()) break;
				bool conflict = false;
				for (u32 l = 0; l < v[j].filled; ++l)
					if (v[j].a[l] == ca || v[j].b[l] == ca || v[j].a[l] == cb || v[j].b[l] == cb) { conflict = true; break; }
				if (!conflict) break;
			}
			if (j == v.size()) { Slot s; s.filled = 0; v.push_back(s); }
			Slot& s = v[j];
			u32 lane = s.filled;
			s.idx[lane] = index; s.a[lane] = ca; s.b[lane] = cb; s.filled = lane + 1;
			if (lane == 0) continue;      // j == vacancy_count branch: new slot, nothing else to do
			if (lane == 7) {              // slot complete: emit, swap-remove (nudge.cpp:4294-4306)
				batches.push_back(std::vector<u32>(s.idx, s.idx + 8));
				v[j] = v.back(); v.pop_back();
			}
		}
		for (unsigned bkt = 0; bkt < bucket_count; ++bkt)  // leftovers, bucket-major; unset lanes repeat lane 0 (nudge.cpp:4316-4338)
			for (size_t j = 0; j < vacant[bkt].size(); ++j) {
				const Slot& s = vacant[bkt][j];
				std::vector<u32> idx(8);
				for (u32 l = 0; l < 8; ++l) idx[l] = l < s.filled ? s.idx[l] : s.idx[0];
				batches.push_back(idx);
			}
	}
	u32 nb = (u32)batches.size();
	data->batches = nb;
	data->constraint_to_contact.resize(nb*8); data->a.resize(nb*8); data->b.resize(nb*8);
	data->rows.assign((size_t)nb*8*ROW_FLOATS, 0.0f); data->states.assign((size_t)nb*8*3, 0.0f);
	const nbo_impulse* impulses = contact_impulses->data.data();

	for (u32 i = 0; i < nb; ++i) {  // nudge.cpp:4350-4633
		nbo_momentum ma[8], mb[8];
		float normal_impulse_l[8], fix_l[8], fiy_l[8], lin[8][3], aang[8][3], bang[8][3];
		for (u32 l = 0; l < 8; ++l) {
			u32 ci = batches[i][l];
			u32 lane = i*8 + l;
			data->constraint_to_contact[lane] = ci;
			const nbo_contact& ct = contacts.data[ci];
			u32 a = contacts.bodies[ci].a, b = contacts.bodies[ci].b;
			data->a[lane] = a; data->b[lane] = b;
			float* row = &data->rows[(size_t)lane*ROW_FLOATS];
			float position_x = ct.position[0], position_y = ct.position[1], position_z = ct.position[2], penetration = ct.penetration;
			float normal_x = ct.normal[0], normal_y = ct.normal[1], normal_z = ct.normal[2], friction = ct.friction;
			float a_mass_inverse = bodies.momentum[a].unused0, b_mass_inverse = bodies.momentum[b].unused0;
			float3 pa = { position_x - bodies.transforms[a].position[0], position_y - bodies.transforms[a].position[1], position_z - bodies.transforms[a].position[2] };
			float3 pb = { position_x - bodies.transforms[b].position[0], position_y - bodies.transforms[b].position[1], position_z - bodies.transforms[b].position[2] };
			const InertiaTransform& A = momentum_to_velocity[a]; const InertiaTransform& B = momentum_to_velocity[b];
			float3 n = { normal_x, normal_y, normal_z };
			float3 nat = cross(pa, n);
			float na_x = A.xx*nat.x + A.xy*nat.y + A.xz*nat.z;
			float na_y = A.xy*nat.x + A.yy*nat.y + A.yz*nat.z;
			float na_z = A.xz*nat.x + A.yz*nat.y + A.zz*nat.z;
			float3 nbt = cross(pb, n);
			float nb_x = B.xx*nbt.x + B.xy*nbt.y + B.xz*nbt.z;
			float nb_y = B.xy*nbt.x + B.yy*nbt.y + B.yz*nbt.z;
			float nb_z = B.xz*nbt.x + B.yz*nbt.y + B.zz*nbt.z;
			nat = cross(float3{ na_x, na_y, na_z }, pa);
			nbt = cross(float3{ nb_x, nb_y, nb_z }, pb);
			float rx = nat.x + nbt.x, ry = nat.y + nbt.y, rz = nat.z + nbt.z;
			float r_dot_n = rx*normal_x + ry*normal_y + rz*normal_z;
			float mass_inverse = a_mass_inverse + b_mass_inverse;
			float nvtni = mass_inverse + r_dot_n;
			bool nonzero = nvtni < 0.0f || nvtni > 0.0f;  // _CMP_NEQ_OQ: ordered (nudge.cpp:636-638, 4439)
			nvtni = nonzero ? (-1.0f / nvtni) : 0.0f;
			float bias = -2.0f * max1(penetration - 1e-3f, 0.0f) * nvtni;  // nudge.cpp:4442, constants 49-50
			float s = absf(normal_x);
			float u_x = normal_z*s;
			float u_y = u_x - normal_z;
			float u_z = madd(normal_x - normal_y, s, normal_y);
			u_x = negf(u_x);
			{ float f = rsqrt(u_x*u_x + u_y*u_y + u_z*u_z); u_x *= f; u_y *= f; u_z *= f; }
			float3 u = { u_x, u_y, u_z };
			float3 v = cross(u, n);
			float3 ua = cross(pa, u), va = cross(pa, v), ub = cross(pb, u), vb = cross(pb, v);
			float a_duu = A.xx*ua.x*ua.x + A.yy*ua.y*ua.y + A.zz*ua.z*ua.z;
			float a_dvv = A.xx*va.x*va.x + A.yy*va.y*va.y + A.zz*va.z*va.z;
			float a_duv = A.xx*ua.x*va.x + A.yy*ua.y*va.y + A.zz*ua.z*va.z;
			float a_suu = A.xy*ua.x*ua.y + A.xz*ua.x*ua.z + A.yz*ua.y*ua.z;
			float a_svv = A.xy*va.x*va.y + A.xz*va.x*va.z + A.yz*va.y*va.z;
			float a_suv = A.xy*(ua.x*va.y + ua.y*va.x) + A.xz*(ua.x*va.z + ua.z*va.x) + A.yz*(ua.y*va.z + ua.z*va.y);
			float b_duu = B.xx*ub.x*ub.x + B.yy*ub.y*ub.y + B.zz*ub.z*ub.z;
			float b_dvv = B.xx*vb.x*vb.x + B.yy*vb.y*vb.y + B.zz*vb.z*vb.z;
			float b_duv = B.xx*ub.x*vb.x + B.yy*ub.y*vb.y + B.zz*ub.z*vb.z;
			float b_suu = B.xy*ub.x*ub.y + B.xz*ub.x*ub.z + B.yz*ub.y*ub.z;
			float b_svv = B.xy*vb.x*vb.y + B.xz*vb.x*vb.z + B.yz*vb.y*vb.z;
			float b_suv = B.xy*(ub.x*vb.y + ub.y*vb.x) + B.xz*(ub.x*vb.z + ub.z*vb.x) + B.yz*(ub.y*vb.z + ub.z*vb.y);
			float friction_x = mass_inverse + a_duu + a_suu + a_suu + b_duu + b_suu + b_suu;
			float friction_y = mass_inverse + a_dvv + a_svv + a_svv + b_dvv + b_svv + b_svv;
			float friction_z = a_duv + a_duv + a_suv + a_suv + b_duv + b_duv + b_suv + b_suv;
			float ua_xt = A.xx*ua.x + A.xy*ua.y + A.xz*ua.z, ua_yt = A.xy*ua.x + A.yy*ua.y + A.yz*ua.z, ua_zt = A.xz*ua.x + A.yz*ua.y + A.zz*ua.z;
			float va_xt = A.xx*va.x + A.xy*va.y + A.xz*va.z, va_yt = A.xy*va.x + A.yy*va.y + A.yz*va.z, va_zt = A.xz*va.x + A.yz*va.y + A.zz*va.z;
			float ub_xt = B.xx*ub.x + B.xy*ub.y + B.xz*ub.z, ub_yt = B.xy*ub.x + B.yy*ub.y + B.yz*ub.z, ub_zt = B.xz*ub.x + B.yz*ub.y + B.zz*ub.z;
			float vb_xt = B.xx*vb.x + B.xy*vb.y + B.xz*vb.z, vb_yt = B.xy*vb.x + B.yy*vb.y + B.yz*vb.z, vb_zt = B.xz*vb.x + B.yz*vb.y + B.zz*vb.z;
			row[N_X] = normal_x; row[N_Y] = normal_y; row[N_Z] = normal_z;
			row[PA_X] = pa.x; row[PA_Y] = pa.y; row[PA_Z] = pa.z; row[PB_X] = pb.x; row[PB_Y] = pb.y; row[PB_Z] = pb.z;
			row[NVTNI] = nvtni; row[BIAS] = bias; row[FRICTION] = friction;
			row[U_X] = u.x; row[U_Y] = u.y; row[U_Z] = u.z; row[V_X] = v.x; row[V_Y] = v.y; row[V_Z] = v.z;
			row[FC_X] = friction_x; row[FC_Y] = friction_y; row[FC_Z] = friction_z;
			row[UA_X] = negf(ua_xt); row[UA_Y] = negf(ua_yt); row[UA_Z] = negf(ua_zt);
			row[VA_X] = negf(va_xt); row[VA_Y] = negf(va_yt); row[VA_Z] = negf(va_zt);
			row[NA_X] = negf(na_x); row[NA_Y] = negf(na_y); row[NA_Z] = negf(na_z);
			row[UB_X] = ub_xt; row[UB_Y] = ub_yt; row[UB_Z] = ub_zt; row[VB_X] = vb_xt; row[VB_Y] = vb_yt; row[VB_Z] = vb_zt;
			row[NB_X] = nb_x; row[NB_Y] = nb_y; row[NB_Z] = nb_z;

			// warm start: nudge.cpp:4563-4632 (gather now, scatter after all 8 lanes like the SIMD code)
			const float* cimp = impulses[ci].impulse;
			ma[l] = bodies.momentum[a]; mb[l] = bodies.momentum[b];
			float normal_impulse = max1(normal_x*cimp[0] + normal_y*cimp[1] + normal_z*cimp[2], 0.0f);
			float max_friction_impulse = normal_impulse * friction;
			float fix = u.x*cimp[0] + u.y*cimp[1] + u.z*cimp[2];
			float fiy = v.x*cimp[0] + v.y*cimp[1] + v.z*cimp[2];
			float fcs = fix*fix + fiy*fiy;
			fcs = rsqrt(fcs);
			fcs = fcs * max_friction_impulse;
			fcs = min1(1.0f, fcs);
			fix = fix * fcs; fiy = fiy * fcs;
			lin[l][0] = fix*u.x + fiy*v.x + normal_x * normal_impulse;
			lin[l][1] = fix*u.y + fiy*v.y + normal_y * normal_impulse;
			lin[l][2] = fix*u.z + fiy*v.z + normal_z * normal_impulse;
			aang[l][0] = fix*row[UA_X] + fiy*row[VA_X] + normal_impulse*row[NA_X];
			aang[l][1] = fix*row[UA_Y] + fiy*row[VA_Y] + normal_impulse*row[NA_Y];
			aang[l][2] = fix*row[UA_Z] + fiy*row[VA_Z] + normal_impulse*row[NA_Z];
			bang[l][0] = fix*row[UB_X] + fiy*row[VB_X] + normal_impulse*row[NB_X];
			bang[l][1] = fix*row[UB_Y] + fiy*row[VB_Y] + normal_impulse*row[NB_Y];
			bang[l][2] = fix*row[UB_Z] + fiy*row[VB_Z] + normal_impulse*row[NB_Z];
			normal_impulse_l[l] = normal_impulse; fix_l[l] = fix; fiy_l[l] = fiy;
		}
		for (u32 l = 0; l < 8; ++l) {
			for (int k = 0; k < 3; ++k) {
				ma[l].velocity[k] -= lin[l][k] * ma[l].unused0;
				ma[l].angular_velocity[k] += aang[l][k];
				mb[l].velocity[k] += lin[l][k] * mb[l].unused0;
				mb[l].angular_velocity[k] += bang[l][k];
			}
			float* st = &data->states[(size_t)(i*8 + l)*3];
			st[0] = normal_impulse_l[l]; st[1] = fix_l[l]; st[2] = fiy_l[l];
		}
		for (u32 l = 0; l < 8; ++l) bodies.momentum[data->a[i*8 + l]] = ma[l];
		for (u32 l = 0; l < 8; ++l) bodies.momentum[data->b[i*8 + l]] = mb[l];
	}
	return data;
}

void nbo_apply_impulses(nbo_constraint_data* data, const nbo_body_data* bodies_p) {  // nudge.cpp:4640-4855
	const nbo_body_data& bodies = *bodies_p;
	for (u32 i = 0; i < data->batches; ++i) {
		nbo_momentum ma[8], mb[8];
		for (u32 l = 0; l < 8; ++l) { ma[l] = bodies.momentum[data->a[i*8 + l]]; mb[l] = bodies.momentum[data->b[i*8 + l]]; }
		for (u32 l = 0; l < 8; ++l) {
			const float* c = &data->rows[(size_t)(i*8 + l)*ROW_FLOATS];
			float* st = &data->states[(size_t)(i*8 + l)*3];
			float a_velocity_x = ma[l].velocity[0], a_velocity_y = ma[l].velocity[1], a_velocity_z = ma[l].velocity[2], a_mass_inverse = ma[l].unused0;
			float a_angular_velocity_x = ma[l].angular_velocity[0], a_angular_velocity_y = ma[l].angular_velocity[1], a_angular_velocity_z = ma[l].angular_velocity[2];
			float b_velocity_x = mb[l].velocity[0], b_velocity_y = mb[l].velocity[1], b_velocity_z = mb[l].velocity[2], b_mass_inverse = mb[l].unused0;
			float b_angular_velocity_x = mb[l].angular_velocity[0], b_angular_velocity_y = mb[l].angular_velocity[1], b_angular_velocity_z = mb[l].angular_velocity[2];
			float pa_z = c[PA_Z], pa_x = c[PA_X], pa_y = c[PA_Y];
			float v_xa = madd(a_angular_velocity_y, pa_z, a_velocity_x);
			float v_ya = madd(a_angular_velocity_z, pa_x, a_velocity_y);
			float v_za = madd(a_angular_velocity_x, pa_y, a_velocity_z);
			float pb_z = c[PB_Z], pb_x = c[PB_X], pb_y = c[PB_Y];
			float v_xb = madd(b_angular_velocity_y, pb_z, b_velocity_x);
			float v_yb = madd(b_angular_velocity_z, pb_x, b_velocity_y);
			float v_zb = madd(b_angular_velocity_x, pb_y, b_velocity_z);
			v_xa = madd(b_angular_velocity_z, pb_y, v_xa);
			v_ya = madd(b_angular_velocity_x, pb_z, v_ya);
			v_za = madd(b_angular_velocity_y, pb_x, v_za);
			float n_x = c[N_X], fu_x = c[U_X], fv_x = c[V_X];
			v_xb = madd(a_angular_velocity_z, pa_y, v_xb);
			v_yb = madd(a_angular_velocity_x, pa_z, v_yb);
			v_zb = madd(a_angular_velocity_y, pa_x, v_zb);
			float n_y = c[N_Y], fu_y = c[U_Y], fv_y = c[V_Y];
			float v_x = v_xb - v_xa, v_y = v_yb - v_ya, v_z = v_zb - v_za;
			float t_z = n_x * v_x, t_x = v_x * fu_x, t_y = v_x * fv_x;
			float n_z = c[N_Z], fu_z = c[U_Z], fv_z = c[V_Z];
			float normal_bias = c[BIAS];
			float old_normal_impulse = st[0];
			float normal_factor = c[NVTNI];
			t_z = madd(n_y, v_y, t_z); t_x = madd(v_y, fu_y, t_x); t_y = madd(v_y, fv_y, t_y);
			normal_bias = normal_bias + old_normal_impulse;
			t_z = madd(n_z, v_z, t_z); t_x = madd(v_z, fu_z, t_x); t_y = madd(v_z, fv_z, t_y);
			float normal_impulse = madd(normal_factor, t_z, normal_bias);
			float t_xx = t_x*t_x, t_yy = t_y*t_y, t_xy = t_x*t_y;
			float tl2 = t_xx + t_yy;
			normal_impulse = max1(normal_impulse, 0.0f);
			t_x *= tl2; t_y *= tl2;
			st[0] = normal_impulse;
			float max_friction_impulse = normal_impulse * c[FRICTION];
			normal_impulse = normal_impulse - old_normal_impulse;
			float friction_factor = t_xx * c[FC_X];
			float linear_impulse_x = n_x * normal_impulse;
			friction_factor = madd(t_yy, c[FC_Y], friction_factor);
			float linear_impulse_y = n_y * normal_impulse;
			friction_factor = madd(t_xy, c[FC_Z], friction_factor);
			float linear_impulse_z = n_z * normal_impulse;
			friction_factor = rcp(friction_factor);
			a_angular_velocity_x = madd(c[NA_X], normal_impulse, a_angular_velocity_x);
			a_angular_velocity_y = madd(c[NA_Y], normal_impulse, a_angular_velocity_y);
			a_angular_velocity_z = madd(c[NA_Z], normal_impulse, a_angular_velocity_z);
			float old_friction_impulse_x = st[1], old_friction_impulse_y = st[2];
			friction_factor = min1(1e+6f, friction_factor);
			float friction_impulse_x = t_x*friction_factor, friction_impulse_y = t_y*friction_factor;
			friction_impulse_x = old_friction_impulse_x - friction_impulse_x;
			friction_impulse_y = old_friction_impulse_y - friction_impulse_y;
			float friction_clamp_scale = friction_impulse_x*friction_impulse_x + friction_impulse_y*friction_impulse_y;
			friction_clamp_scale = rsqrt(friction_clamp_scale);
			b_angular_velocity_x = madd(c[NB_X], normal_impulse, b_angular_velocity_x);
			b_angular_velocity_y = madd(c[NB_Y], normal_impulse, b_angular_velocity_y);
			b_angular_velocity_z = madd(c[NB_Z], normal_impulse, b_angular_velocity_z);
			friction_clamp_scale = friction_clamp_scale * max_friction_impulse;
			friction_clamp_scale = min1(1.0f, friction_clamp_scale);
			friction_impulse_x = friction_impulse_x * friction_clamp_scale;
			friction_impulse_y = friction_impulse_y * friction_clamp_scale;
			st[1] = friction_impulse_x; st[2] = friction_impulse_y;
			friction_impulse_x -= old_friction_impulse_x;
			friction_impulse_y -= old_friction_impulse_y;
			linear_impulse_x = madd(fu_x, friction_impulse_x, linear_impulse_x);
			linear_impulse_y = madd(fu_y, friction_impulse_x, linear_impulse_y);
			linear_impulse_z = madd(fu_z, friction_impulse_x, linear_impulse_z);
			linear_impulse_x = madd(fv_x, friction_impulse_y, linear_impulse_x);
			linear_impulse_y = madd(fv_y, friction_impulse_y, linear_impulse_y);
			linear_impulse_z = madd(fv_z, friction_impulse_y, linear_impulse_z);
			float a_mass_inverse_neg = negf(a_mass_inverse);
			a_velocity_x = madd(linear_impulse_x, a_mass_inverse_neg, a_velocity_x);
			a_velocity_y = madd(linear_impulse_y, a_mass_inverse_neg, a_velocity_y);
			a_velocity_z = madd(linear_impulse_z, a_mass_inverse_neg, a_velocity_z);
			a_angular_velocity_x = madd(c[UA_X], friction_impulse_x, a_angular_velocity_x);
			a_angular_velocity_y = madd(c[UA_Y], friction_impulse_x, a_angular_velocity_y);
			a_angular_velocity_z = madd(c[UA_Z], friction_impulse_x, a_angular_velocity_z);
			a_angular_velocity_x = madd(c[VA_X], friction_impulse_y, a_angular_velocity_x);
			a_angular_velocity_y = madd(c[VA_Y], friction_impulse_y, a_angular_velocity_y);
			a_angular_velocity_z = madd(c[VA_Z], friction_impulse_y, a_angular_velocity_z);
			b_velocity_x = madd(linear_impulse_x, b_mass_inverse, b_velocity_x);
			b_velocity_y = madd(linear_impulse_y, b_mass_inverse, b_velocity_y);
			b_velocity_z = madd(linear_impulse_z, b_mass_inverse, b_velocity_z);
			b_angular_velocity_x = madd(c[UB_X], friction_impulse_x, b_angular_velocity_x);
			b_angular_velocity_y = madd(c[UB_Y], friction_impulse_x, b_angular_velocity_y);
			b_angular_velocity_z = madd(c[UB_Z], friction_impulse_x, b_angular_velocity_z);
			b_angular_velocity_x = madd(c[VB_X], friction_impulse_y, b_angular_velocity_x);
			b_angular_velocity_y = madd(c[VB_Y], friction_impulse_y, b_angular_velocity_y);
			b_angular_velocity_z = madd(c[VB_Z], friction_impulse_y, b_angular_velocity_z);
			ma[l].velocity[0] = a_velocity_x; ma[l].velocity[1] = a_velocity_y; ma[l].velocity[2] = a_velocity_z;
			ma[l].angular_velocity[0] = a_angular_velocity_x; ma[l].angular_velocity[1] = a_angular_velocity_y; ma[l].angular_velocity[2] = a_angular_velocity_z;
			ma[l].unused1 = 0.0f;  // nudge.cpp:4823
			mb[l].velocity[0] = b_velocity_x; mb[l].velocity[1] = b_velocity_y; mb[l].velocity[2] = b_velocity_z;
			mb[l].angular_velocity[0] = b_angular_velocity_x; mb[l].angular_velocity[1] = b_angular_velocity_y; mb[l].angular_velocity[2] = b_angular_velocity_z;
			mb[l].unused1 = 0.0f;  // nudge.cpp:4849
		}
		for (u32 l = 0; l < 8; ++l) bodies.momentum[data->a[i*8 + l]] = ma[l];
		for (u32 l = 0; l < 8; ++l) bodies.momentum[data->b[i*8 + l]] = mb[l];
	}
}

void nbo_update_cached_impulses(nbo_constraint_data* data, nbo_impulse_data* contact_impulses) {  // nudge.cpp:4857-4884
	for (u32 i = 0; i < data->batches*8; ++i) {
		u32 contact = data->constraint_to_contact[i];
		const float* c = &data->rows[(size_t)i*ROW_FLOATS];
		const float* st = &data->states[(size_t)i*3];
		float* impulse = contact_impulses->data[contact].impulse;
		impulse[0] = st[0]*c[N_X] + st[1]*c[U_X] + st[2]*c[V_X];
		impulse[1] = st[0]*c[N_Y] + st[1]*c[U_Y] + st[2]*c[V_Y];
		impulse[2] = st[0]*c[N_Z] + st[1]*c[U_Z] + st[2]*c[V_Z];
	}
}

void nbo_advance(const nbo_active_bodies* active_bodies, const nbo_body_data* bodies_p, float time_step) {  // nudge.cpp:4886-4926
	const nbo_body_data& bodies = *bodies_p;
	float half_time_step = 0.5f * time_step;
	for (u32 n = 0; n < active_bodies->count; ++n) {
		u32 i = active_bodies->indices[n];
		float3 velocity = f3(bodies.momentum[i].velocity);
		float3 angular_velocity = f3(bodies.momentum[i].angular_velocity);
		if (dot(velocity, velocity) < 1e-2f && dot(angular_velocity, angular_velocity) < 1e-1f) {
			if (bodies.idle_counters[i] < 0xff) ++bodies.idle_counters[i];
		}
		else bodies.idle_counters[i] = 0;
		Rot dr = { angular_velocity, 0.0f };
		dr = rotmul(dr, rot(bodies.transforms[i].rotation));
		dr.v = mul(dr.v, half_time_step); dr.s *= half_time_step;
		nbo_transform& t = bodies.transforms[i];
		t.position[0] += velocity.x * time_step; t.position[1] += velocity.y * time_step; t.position[2] += velocity.z * time_step;
		t.rotation[0] += dr.v.x; t.rotation[1] += dr.v.y; t.rotation[2] += dr.v.z; t.rotation[3] += dr.s;
		float f = 1.0f / sqrtf(t.rotation[3]*t.rotation[3] + t.rotation[0]*t.rotation[0] + t.rotation[1]*t.rotation[1] + t.rotation[2]*t.rotation[2]);  // nudge.cpp:1128-1133
		t.rotation[0] *= f; t.rotation[1] *= f; t.rotation[2] *= f; t.rotation[3] *= f;
	}
}

void nbo_free_impulses(nbo_impulse_data* p) { delete p; }
void nbo_free_constraints(nbo_constraint_data* p) { delete p; }

void nbo_impulses_get(nbo_impulse_data* m, u32 contact_count, u32* sorted_contacts, float* data, u32* culled_count, u64* culled_tags, u32* culled_features, float* culled_data, u32 culled_capacity) {
	if (sorted_contacts) memcpy(sorted_contacts, m->sorted_contacts.data(), sizeof(u32)*contact_count);
	if (data) memcpy(data, m->data.data(), sizeof(nbo_impulse)*contact_count);
	u32 cc = (u32)m->culled_tags.size();
	if (culled_count) *culled_count = cc;
	u32 n = cc < culled_capacity ? cc : culled_capacity;
	if (culled_tags) memcpy(culled_tags, m->culled_tags.data(), sizeof(u64)*n);
	if (culled_features) memcpy(culled_features, m->culled_features.data(), sizeof(u32)*n);
	if (culled_data) memcpy(culled_data, m->culled_data.data(), sizeof(nbo_impulse)*n);
}

uint32_t nbo_constraints_batches(nbo_constraint_data* d) { return d->batches; }

void nbo_constraints_get(nbo_constraint_data* d, u32* c2c, u32* a, u32* b, float* rows, float* states) {
	size_t lanes = (size_t)d->batches*8;
	if (c2c) memcpy(c2c, d->constraint_to_contact.data(), 4*lanes);
	if (a) memcpy(a, d->a.data(), 4*lanes);
	if (b) memcpy(b, d->b.data(), 4*lanes);
	if (rows) memcpy(rows, d->rows.data(), 4*lanes*ROW_FLOATS);
	if (states) memcpy(states, d->states.data(), 4*lanes*3);
}

}
