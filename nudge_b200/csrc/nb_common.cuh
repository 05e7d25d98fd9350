// nudge_b200 — shared device helpers: exact-arithmetic primitives and small vector math.
//
// Arithmetic contract: this translation unit is compiled with -fmad=false so that
// `a*b + c` is two IEEE roundings exactly as the reference's `*`/`+` on __m128/__m256 are under
// -ffp-contract=off; FMAs appear only where the reference writes madd/msub (nudge.cpp:270-284, 578-592) and
// are spelled nb_madd/nb_msub here.  Division and sqrt are IEEE (-prec-div=true -prec-sqrt=true, -ftz=false).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/nudge_b200.h"

typedef uint32_t u32;
typedef uint64_t u64;

#define NB_DEV __device__ __forceinline__
#define NB_SIGN 0x80000000u

// rcpps / rsqrtps tables sampled from the host CPU at nb_create (nudge.cpp:296-302, 604-610):
// rcp depends on sign/exponent + top 11 mantissa bits, rsqrt on exponent parity + top 10 mantissa bits.
__constant__ u32 c_rcp_lut[2048];
__constant__ u32 c_rsqrt_lut[2048];
// copies in global memory; kernels that hit the tables with divergent indices stage these into shared memory
__device__ u32 g_rcp_lut[2048];
__device__ u32 g_rsqrt_lut[2048];

NB_DEV u32 asu(float x) { return __float_as_uint(x); }
NB_DEV float asf(u32 x) { return __uint_as_float(x); }
NB_DEV float nb_min(float x, float y) { return (y < x) ? y : x; }   // first operand on NaN: nudge.cpp:286-289, 594-597
NB_DEV float nb_max(float x, float y) { return (y > x) ? y : x; }   // nudge.cpp:291-294, 599-602
NB_DEV float nb_madd(float x, float y, float z) { return __fmaf_rn(x, y, z); }
NB_DEV float nb_msub(float x, float y, float z) { return __fmaf_rn(x, y, -z); }
NB_DEV float nb_abs(float x) { return asf(asu(x) & 0x7fffffffu); }
NB_DEV float nb_xor(float x, u32 m) { return asf(asu(x) ^ m); }
NB_DEV float nb_neg(float x) { return asf(asu(x) ^ NB_SIGN); }
NB_DEV int nb_toint(float x) {  // cvttps2dq: truncation; out of range / NaN -> 0x80000000 (nudge.cpp:336-338)
	if (!(x > -2147483904.0f && x < 2147483648.0f)) return (int)0x80000000;
	return __float2int_rz(x);
}

// rcpps emulation: table of rcp(1.m) for the top 11 mantissa bits, exponent arithmetic done exactly.
// Denormal inputs behave as zero, results below 2^-126 flush to zero (measured, see tests/test_lut_model.py).
template<class LUT>
NB_DEV float nb_rcp_t(float x, const LUT lut) {
	u32 u = asu(x), s = u & NB_SIGN, e = (u >> 23) & 0xff, m = u & 0x7fffff;
	if (e == 0) return asf(s | 0x7f800000u);
	if (e == 255) return m ? asf(u | 0x00400000u) : asf(s);
	u32 t = lut[m >> 12];
	int re = (int)((t >> 23) & 0xff) - ((int)e - 127);
	if (re <= 0) return asf(s);
	if (re >= 255) return asf(s | 0x7f800000u);
	return asf(s | ((u32)re << 23) | (t & 0x7fffff));
}
template<class LUT>
NB_DEV float nb_rsqrt_t(float x, const LUT lut) {
	u32 u = asu(x), s = u & NB_SIGN, e = (u >> 23) & 0xff, m = u & 0x7fffff;
	if (e == 255 && m) return asf(u | 0x00400000u);
	if (e == 0) return asf(s | 0x7f800000u);
	if (s) return asf(0xffc00000u);
	if (e == 255) return 0.0f;
	int eu = (int)e - 127;
	u32 p = (u32)eu & 1u;
	u32 t = lut[(p << 10) | (m >> 13)];
	int re = (int)((t >> 23) & 0xff) - ((eu - (int)p) >> 1);
	return asf(((u32)re << 23) | (t & 0x7fffff));
}
NB_DEV float nb_rcp(float x) { return nb_rcp_t(x, c_rcp_lut); }
NB_DEV float nb_rsqrt(float x) { return nb_rsqrt_t(x, c_rsqrt_lut); }

struct f3 { float x, y, z; };
NB_DEV f3 mk3(float x, float y, float z) { f3 r = { x, y, z }; return r; }
NB_DEV f3 ld3(const float* p) { f3 r = { p[0], p[1], p[2] }; return r; }
NB_DEV f3 add3(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
NB_DEV f3 sub3(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
NB_DEV f3 mul3(float a, f3 b) { return mk3(a * b.x, a * b.y, a * b.z); }
NB_DEV f3 mul3(f3 a, float b) { return mk3(a.x * b, a.y * b, a.z * b); }
NB_DEV float dot3(f3 a, f3 b) { return a.x*b.x + a.y*b.y + a.z*b.z; }                                  // nudge.cpp:1104-1106
NB_DEV f3 cross3(f3 a, f3 b) { return mk3(a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x); }  // nudge.cpp:1112-1115
struct quat { f3 v; float s; };
NB_DEV quat mkq(float4 q) { quat r = { { q.x, q.y, q.z }, q.w }; return r; }
NB_DEV f3 qrot(quat l, f3 r) {  // nudge.cpp:1117-1120
	f3 t = mul3(2.0f, cross3(l.v, r));
	return add3(add3(r, mul3(l.s, t)), cross3(l.v, t));
}
NB_DEV quat qmul(quat l, quat r) {  // nudge.cpp:1122-1126
	quat q;
	q.v = add3(add3(mul3(r.v, l.s), mul3(l.v, r.s)), cross3(l.v, r.v));
	q.s = l.s*r.s - dot3(l.v, r.v);
	return q;
}
struct mat3 { f3 c0, c1, c2; };
NB_DEV mat3 qmatrix(quat q) {  // nudge.cpp:1142-1163
	float kx = q.v.x + q.v.x, ky = q.v.y + q.v.y, kz = q.v.z + q.v.z;
	float xx = kx*q.v.x, yy = ky*q.v.y, zz = kz*q.v.z;
	float xy = kx*q.v.y, xz = kx*q.v.z, yz = ky*q.v.z;
	float sx = kx*q.s, sy = ky*q.s, sz = kz*q.s;
	mat3 m;
	m.c0 = mk3(1.0f - yy - zz, xy + sz, xz - sy);
	m.c1 = mk3(xy - sz, 1.0f - xx - zz, yz + sx);
	m.c2 = mk3(xz + sy, yz - sx, 1.0f - xx - yy);
	return m;
}

// A Transform (nudge.h:34-38) is two float4 rows in HBM: (position, body-as-bits) and rotation.
struct xform { float4 p; float4 q; };
NB_DEV xform ld_xform(const nb_transform* t, u32 i) {
	const float4* p = reinterpret_cast<const float4*>(t + i);
	xform x; x.p = p[0]; x.q = p[1];
	return x;
}
NB_DEV void st_xform(nb_transform* t, u32 i, xform x) {
	float4* p = reinterpret_cast<float4*>(t + i);
	p[0] = x.p; p[1] = x.q;
}

// order-preserving float <-> uint map for atomicMin/atomicMax on floats
NB_DEV u32 f2ord(float f) { u32 u = asu(f); return (u & NB_SIGN) ? ~u : (u | NB_SIGN); }
NB_DEV float ord2f(u32 o) { return asf((o & NB_SIGN) ? (o & 0x7fffffffu) : ~o); }
