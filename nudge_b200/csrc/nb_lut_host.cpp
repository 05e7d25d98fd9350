// nudge_b200 — host-side calibration of the rcpps / rsqrtps tables.
// The reference calls _mm_rcp_ps / _mm_rsqrt_ps (nudge.cpp:296-302, 604-610), whose results are defined by the
// CPU, not by IEEE 754.  To reproduce them on the GPU we sample the host CPU once per context: rcp for every
// value of the top 11 mantissa bits in [1,2), rsqrt for both exponent parities and the top 10 mantissa bits in
// [1,4).  nb_host_check_lut_model() then verifies on 2^20 probes (plus specials) that "table entry + exact
// exponent arithmetic" reproduces the instruction bit for bit on this CPU; nb_lut_model_exact() reports it.
// This is calibration of constants, not a CPU compute path: no simulation data ever passes through here.
#include <immintrin.h>
#include <stdint.h>
#include <string.h>

typedef uint32_t u32;

static inline u32 asu(float x) { u32 u; memcpy(&u, &x, 4); return u; }
static inline float asf(u32 u) { float x; memcpy(&x, &u, 4); return x; }
static inline float hw_rcp(float x) { return _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(x))); }
static inline float hw_rsqrt(float x) { return _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x))); }

static float model_rcp(float x, const u32* lut) {
	u32 u = asu(x), s = u & 0x80000000u, e = (u >> 23) & 0xff, m = u & 0x7fffff;
	if (e == 0) return asf(s | 0x7f800000u);
	if (e == 255) return m ? asf(u | 0x00400000u) : asf(s);
	u32 t = lut[m >> 12];
	int re = (int)((t >> 23) & 0xff) - ((int)e - 127);
	if (re <= 0) return asf(s);
	if (re >= 255) return asf(s | 0x7f800000u);
	return asf(s | ((u32)re << 23) | (t & 0x7fffff));
}

static float model_rsqrt(float x, const u32* lut) {
	u32 u = asu(x), s = u & 0x80000000u, e = (u >> 23) & 0xff, m = u & 0x7fffff;
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

extern "C" void nb_host_sample_luts(u32* rcp_lut, u32* rsqrt_lut) {
	unsigned csr = _mm_getcsr();
	_mm_setcsr(csr & ~0x8040u);  // FTZ/DAZ off while sampling
	for (u32 i = 0; i < 2048; ++i) rcp_lut[i] = asu(hw_rcp(asf(0x3f800000u | (i << 12))));
	for (u32 p = 0; p < 2; ++p)
		for (u32 i = 0; i < 1024; ++i) rsqrt_lut[(p << 10) | i] = asu(hw_rsqrt(asf(((127u + p) << 23) | (i << 13))));
	_mm_setcsr(csr);
}

extern "C" int nb_host_check_lut_model(const u32* rcp_lut, const u32* rsqrt_lut) {
	unsigned csr = _mm_getcsr();
	_mm_setcsr(csr & ~0x8040u);
	int ok = 1;
	u32 state = 0x12345678u;
	for (u32 i = 0; i < (1u << 20) && ok; ++i) {
		state = state * 1664525u + 1013904223u;
		u32 bits = state ^ (state >> 15);
		float x = asf(bits);
		float a = hw_rcp(x), b = model_rcp(x, rcp_lut);
		bool nan_a = a != a, nan_b = b != b;
		if (nan_a != nan_b || (!nan_a && asu(a) != asu(b))) ok = 0;
		a = hw_rsqrt(x); b = model_rsqrt(x, rsqrt_lut);
		nan_a = a != a; nan_b = b != b;
		if (nan_a != nan_b || (!nan_a && asu(a) != asu(b))) ok = 0;
	}
	// every mantissa in [1,2) / [1,4) for a dense check of the "top bits only" hypothesis
	for (u32 m = 0; m < (1u << 23) && ok; m += 7) {
		float x = asf(0x3f800000u | m);
		if (asu(hw_rcp(x)) != asu(model_rcp(x, rcp_lut))) ok = 0;
		if (asu(hw_rsqrt(x)) != asu(model_rsqrt(x, rsqrt_lut))) ok = 0;
		x = asf(0x40000000u | m);
		if (asu(hw_rsqrt(x)) != asu(model_rsqrt(x, rsqrt_lut))) ok = 0;
	}
	_mm_setcsr(csr);
	return ok;
}
