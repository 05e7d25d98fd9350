// nudge_b200 — collide(): collider world transforms + AABBs, Morton order, implicit 8-ary AABB tree broadphase,
// island detection, pair partition, box-box / box-sphere / sphere-sphere narrowphase with feature tags,
// active-body list and contact compaction.  Replaces nudge::collide (nudge.cpp:3000-4009) and the three
// narrowphase routines (nudge.cpp:1177-2604).
//
// Design (B200-first): the reference's two-level "8 AABBs per coarse AABB, all coarse pairs" scheme
// (nudge.cpp:3204-3399) is O((K/8)^2); here the same groups-of-8 idea is carried up log8(K) levels over the
// Morton-sorted colliders, one thread per collider walks the tree, and pairs are appended with
// warp-aggregated atomics.  The pair SET is identical (all strictly overlapping AABBs); orientation and order
// come from the Morton rank and a radix sort exactly as nudge.cpp:3493-3498 defines them.  The narrowphase is
// one thread per pair, run twice (count, then emit at scanned offsets) so that contacts land in the
// reference's order: box-box face contacts, box-box edge contacts, box-sphere, sphere-sphere.
#pragma once
#include "nb_prims.cuh"

enum {  // device counters (u32 each)
	CNT_PAIRS, CNT_LIVE0, CNT_LIVE1, CNT_LIVE2, CNT_LIVE3, CNT_SLEEP_COARSE, CNT_LIVE_TOTAL,
	CNT_FACE, CNT_EDGE, CNT_OTHER, CNT_STAGED, CNT_CONTACTS, CNT_SLEEP_FINE, CNT_SLEEPING, CNT_ACTIVE,
	CNT_CACHE, CNT_CULLED, CNT_FULL_BATCHES, CNT_BATCHES, CNT_LEVELS, CNT_ENTRIES, CNT_OVERFLOW, CNT_LVCH0, CNT_LVCH1, CNT_LVCH2,
	CNT_BMIN0, CNT_BMIN1, CNT_BMIN2, CNT_BMIN3, CNT_BMAX0, CNT_BMAX1, CNT_BMAX2, CNT_BMAX3,
	CNT_BAR0, CNT_BAR1, CNT_SCRATCH0, CNT_SCRATCH1, CNT_EXT_SUM, CNT_GRID_LEVEL, CNT_LARGE, CNT_SURV, CNT_TMP,
	CNT_EXT_HIST = 64 /* 18 bins */, CNT__COUNT = 96
};
enum { OVF_PAIRS = 1, OVF_CONTACTS = 2, OVF_SCHED = 4, OVF_LEVELS = 8 };

// ---------------- K1: collider world transform + AABB (nudge.cpp:3021-3079) ----------------
__global__ void __launch_bounds__(NB_BLOCK) k_collider_world(u32 nboxes, u32 nspheres, const nb_transform* body_xf,
		const nb_transform* box_xf, const nb_box_collider* box_data, const u32* box_tags,
		const nb_transform* sph_xf, const nb_sphere_collider* sph_data, const u32* sph_tags,
		nb_transform* world_xf, float4* aabb_min, float4* aabb_max, u32* col_tag, u32* col_body, u32* counts) {
	u32 K = nboxes + nspheres;
	float mn[3] = { INFINITY, INFINITY, INFINITY }, mx[3] = { -INFINITY, -INFINITY, -INFINITY };
	float ext_sum = 0.0f;
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < K; i += gridDim.x * blockDim.x) {
		bool is_box = i < nboxes;
		xform l = is_box ? ld_xform(box_xf, i) : ld_xform(sph_xf, i - nboxes);
		u32 body = asu(l.p.w);
		xform b = ld_xform(body_xf, body);
		// Transform * Transform: nudge.cpp:1165-1175
		quat bq = mkq(b.q);
		f3 p = add3(qrot(bq, mk3(l.p.x, l.p.y, l.p.z)), mk3(b.p.x, b.p.y, b.p.z));
		quat q = qmul(bq, mkq(l.q));
		xform w; w.p = make_float4(p.x, p.y, p.z, asf(body)); w.q = make_float4(q.v.x, q.v.y, q.v.z, q.s);
		f3 ext;
		if (is_box) {
			mat3 m = qmatrix(q);
			float4 s = reinterpret_cast<const float4*>(box_data)[i];
			m.c0 = mul3(m.c0, s.x); m.c1 = mul3(m.c1, s.y); m.c2 = mul3(m.c2, s.z);
			ext = mk3(fabsf(m.c0.x) + fabsf(m.c1.x) + fabsf(m.c2.x), fabsf(m.c0.y) + fabsf(m.c1.y) + fabsf(m.c2.y), fabsf(m.c0.z) + fabsf(m.c1.z) + fabsf(m.c2.z));
		}
		else {
			float r = sph_data[i - nboxes].radius;
			ext = mk3(r, r, r);
		}
		float4 lo = make_float4(p.x - ext.x, p.y - ext.y, p.z - ext.z, 0.0f);
		float4 hi = make_float4(p.x + ext.x, p.y + ext.y, p.z + ext.z, 0.0f);
		st_xform(world_xf, i, w);
		aabb_min[i] = lo; aabb_max[i] = hi;
		col_tag[i] = is_box ? box_tags[i] : sph_tags[i - nboxes];
		col_body[i] = body;
		mn[0] = fminf(mn[0], lo.x); mn[1] = fminf(mn[1], lo.y); mn[2] = fminf(mn[2], lo.z);
		mx[0] = fmaxf(mx[0], lo.x); mx[1] = fmaxf(mx[1], lo.y); mx[2] = fmaxf(mx[2], lo.z);
		ext_sum += 2.0f * fmaxf(ext.x, fmaxf(ext.y, ext.z));
	}
	// mean AABB extent (only steers the grid cell size of the broadphase, never a result): warp sum, one float atomic per warp
	#pragma unroll
	for (int d = 16; d; d >>= 1) ext_sum += __shfl_xor_sync(0xffffffffu, ext_sum, d);
	if ((threadIdx.x & 31) == 0 && ext_sum > 0.0f) atomicAdd(reinterpret_cast<float*>(&counts[CNT_EXT_SUM]), ext_sum);
	// scene bounds over AABB *mins* (nudge.cpp:3087-3094): min/max are exact, so any reduction order gives the same bits
	#pragma unroll
	for (int k = 0; k < 3; ++k) {
		#pragma unroll
		for (int d = 16; d; d >>= 1) {
			mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], d));
			mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], d));
		}
	}
	if ((threadIdx.x & 31) == 0) {
		#pragma unroll
		for (int k = 0; k < 3; ++k) {
			atomicMin(&counts[CNT_BMIN0 + k], f2ord(mn[k]));
			atomicMax(&counts[CNT_BMAX0 + k], f2ord(mx[k]));
		}
	}
}

// ---------------- K2: Morton codes (nudge.cpp:3096-3163, 2606-2645) ----------------
NB_DEV void dilate3(u32 x, u32 offset, u32& lo32, u32& hi32) {
	u32 lo24 = x & 0xff, hi24 = (x >> 8) & 0xff;
	lo24 = (lo24 | (lo24 << 8)) & 0x0f00f00fu; hi24 = (hi24 | (hi24 << 8)) & 0x0f00f00fu;
	lo24 = (lo24 | (lo24 << 4)) & 0xc30c30c3u; hi24 = (hi24 | (hi24 << 4)) & 0xc30c30c3u;
	lo24 = (lo24 | (lo24 << 2)) & 0x49249249u; hi24 = (hi24 | (hi24 << 2)) & 0x49249249u;
	lo32 = (lo24 << offset) | (hi24 << (24 + offset));
	hi32 = hi24 >> (8 - offset);
}

struct MortonFrame { float L[4]; float ms[3]; };
NB_DEV MortonFrame morton_frame(const u32* counts) {
	MortonFrame f;
	float smin[4], smax[4], sc[4];
	#pragma unroll
	for (int k = 0; k < 3; ++k) { smin[k] = ord2f(counts[CNT_BMIN0 + k]); smax[k] = ord2f(counts[CNT_BMAX0 + k]); }
	smin[3] = 0.0f; smax[3] = 0.0f;
	#pragma unroll
	for (int k = 0; k < 4; ++k) sc[k] = 65535.0f * nb_rcp(smax[k] - smin[k]);
	float A[4] = { nb_min(sc[0], sc[2]), nb_min(sc[1], sc[2]), nb_min(sc[2], sc[0]), nb_min(sc[2], sc[1]) };
	f.L[0] = nb_min(A[0], A[1]); f.L[1] = nb_min(A[1], A[0]); f.L[2] = nb_min(A[2], A[3]); f.L[3] = nb_min(A[3], A[2]);
	f.ms[0] = smin[0] * f.L[0]; f.ms[1] = smin[1] * f.L[1]; f.ms[2] = smin[2] * f.L[2];
	return f;
}
// quantised min corner of collider i exactly as k_morton computes it (i = ORIGINAL collider index: the lane-dependent scale)
NB_DEV void morton_quantise(const MortonFrame& f, float4 p, u32 i, u32& x, u32& y, u32& z) {
	float s = f.L[i & 3];
	x = (u32)nb_toint(nb_msub(p.x, s, f.ms[0])); y = (u32)nb_toint(nb_msub(p.y, s, f.ms[1])); z = (u32)nb_toint(nb_msub(p.z, s, f.ms[2]));
}

// Also bins the colliders by the grid level they need (smallest l with quantised extent + 2 <= 2^l; 17 = never) for k_grid_setup.
NB_DEV u32 grid_level_needed(float4 lo, float4 hi, float scale) {
	float ext = fmaxf(hi.x - lo.x, fmaxf(hi.y - lo.y, hi.z - lo.z)) * scale + 2.0f;
	u32 l = 1;
	while (l <= 16 && !(ext <= (float)(1u << l))) ++l;  // NaN never fits
	return l;
}
__global__ void __launch_bounds__(NB_BLOCK) k_morton(u32 K, const float4* aabb_min, const float4* aabb_max, u32* counts, u64* keys, u32* vals, u64* keybits /* OR, AND of all codes: steers the sort's bucket digit */) {
	__shared__ u32 s_hist[18];
	u64 k_or = 0, k_and = ~(u64)0;
	if (threadIdx.x < 18) s_hist[threadIdx.x] = 0;
	__syncthreads();
	float smin[4], smax[4], sc[4];
	#pragma unroll
	for (int k = 0; k < 3; ++k) { smin[k] = ord2f(counts[CNT_BMIN0 + k]); smax[k] = ord2f(counts[CNT_BMAX0 + k]); }
	smin[3] = 0.0f; smax[3] = 0.0f;  // the unused0 lane of the AABB (nudge.cpp:879-884)
	#pragma unroll
	for (int k = 0; k < 4; ++k) sc[k] = 65535.0f * nb_rcp(smax[k] - smin[k]);
	float A[4] = { nb_min(sc[0], sc[2]), nb_min(sc[1], sc[2]), nb_min(sc[2], sc[0]), nb_min(sc[2], sc[1]) };  // nudge.cpp:3098
	float L[4] = { nb_min(A[0], A[1]), nb_min(A[1], A[0]), nb_min(A[2], A[3]), nb_min(A[3], A[2]) };          // nudge.cpp:3099
	float ms[3] = { smin[0] * L[0], smin[1] * L[1], smin[2] * L[2] };                                         // nudge.cpp:3100
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < K; i += gridDim.x * blockDim.x) {
		float4 p = aabb_min[i];
		float s = L[i & 3];  // SoA lane j is scaled by AoS lane j&3 (nudge.cpp:3143-3145); all four hold the same min
		u32 x = (u32)nb_toint(nb_msub(p.x, s, ms[0]));
		u32 y = (u32)nb_toint(nb_msub(p.y, s, ms[1]));
		u32 z = (u32)nb_toint(nb_msub(p.z, s, ms[2]));
		u32 lx, hx, ly, hy, lz, hz;
		dilate3(x, 2, lx, hx); dilate3(y, 1, ly, hy); dilate3(z, 0, lz, hz);
		const u64 code = (u64)(lx | ly | lz) | ((u64)(hx | hy | hz) << 32);
		keys[i] = code;
		vals[i] = i;
		k_or |= code; k_and &= code;
		atomicAdd(&s_hist[grid_level_needed(p, aabb_max[i], L[0])], 1u);
	}
	#pragma unroll
	for (int d = 16; d; d >>= 1) { k_or |= __shfl_xor_sync(0xffffffffu, k_or, d); k_and &= __shfl_xor_sync(0xffffffffu, k_and, d); }
	if ((threadIdx.x & 31) == 0) { atomicOr((unsigned long long*)&keybits[0], (unsigned long long)k_or); atomicAnd((unsigned long long*)&keybits[1], (unsigned long long)k_and); }
	__syncthreads();
	if (threadIdx.x < 18 && s_hist[threadIdx.x]) atomicAdd(&counts[CNT_EXT_HIST + threadIdx.x], s_hist[threadIdx.x]);
}

// ---------------- K3/K4: Morton-ordered leaves and the implicit 8-ary AABB tree ----------------
__global__ void __launch_bounds__(NB_BLOCK) k_leaves(u32 K, const u32* sorted_vals, const u64* sorted_keys, const float4* aabb_min, const float4* aabb_max,
													 u32* order, u32* rank, float4* leaf_min, float4* leaf_max, u64* mkeys) {
	for (u32 pos = blockIdx.x * blockDim.x + threadIdx.x; pos < K; pos += gridDim.x * blockDim.x) {
		u32 i = sorted_vals[pos];
		order[pos] = i; rank[i] = pos; mkeys[pos] = sorted_keys[pos];
		float4 lo = aabb_min[i], hi = aabb_max[i];
		float vol = (hi.x - lo.x) * (hi.y - lo.y) * (hi.z - lo.z);
		lo.w = vol; hi.w = vol;
		leaf_min[pos] = lo; leaf_max[pos] = hi;
	}
}

__global__ void __launch_bounds__(NB_BLOCK) k_build_level(const float4* cmin, const float4* cmax, u32 n_child, float4* pmin, float4* pmax, u32 n_parent) {
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n_parent; i += gridDim.x * blockDim.x) {
		u32 b = i * 8, e = min(b + 8, n_child);
		float4 lo = make_float4(INFINITY, INFINITY, INFINITY, 0.0f), hi = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
		for (u32 c = b; c < e; ++c) {
			float4 a = cmin[c], d = cmax[c];
			lo.x = fminf(lo.x, a.x); lo.y = fminf(lo.y, a.y); lo.z = fminf(lo.z, a.z);
			hi.x = fmaxf(hi.x, d.x); hi.y = fmaxf(hi.y, d.y); hi.z = fmaxf(hi.z, d.z);
			hi.w = fmaxf(hi.w, d.w);  // largest leaf volume below this node
		}
		pmin[i] = lo; pmax[i] = hi;
	}
}

#define NB_MAX_LEVELS 9
struct Tree { const float4* mn[NB_MAX_LEVELS]; const float4* mx[NB_MAX_LEVELS]; u32 n[NB_MAX_LEVELS]; int levels; };

// ---------------- K5: all strictly overlapping AABB pairs (the set nudge.cpp:3275-3489 produces) ----------------
// Each unordered pair is reported by exactly one side: the leaf with the smaller (volume, position).  Subtrees whose
// largest leaf volume is below every member's are skipped, so a giant collider (the ground) costs nothing.
// Warp-cooperative: a warp owns 32 Morton-adjacent leaves ("members", staged in shared memory) and walks the tree once with
// a warp-uniform stack.  One step expands a node: lane = (child c = lane & 7, member subset lane >> 3), every lane loads its
// child's box (one coalesced 256-byte read for the 8 children) and tests it against its 8 members, so the walk pays one
// dependent memory round trip per NODE instead of per child.  Hits are staged in shared memory and flushed 32 at a time.
struct PairStage { float lo[3][32]; float hi[3][32]; float vol[32]; u32 orig[32]; u64 buf[32 + 256]; };  // one block of 8 leaves can yield 8 x 32 pairs

__global__ void __launch_bounds__(NB_BLOCK) k_find_pairs(Tree T, u32 K, const u32* order, u32 kbits, u64* pair_keys, u32 max_pairs, u32* counts) {
	__shared__ PairStage stage[NB_WARPS];
	const u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	PairStage& S = stage[wid];
	const u32 c = lane & 7, sub = lane >> 3;
	const u32 warps_total = gridDim.x * NB_WARPS;
	const u32 groups = (K + 31) / 32;
	const int top = T.levels - 1;
	for (u32 g = blockIdx.x * NB_WARPS + wid; g < groups; g += warps_total) {
		{
			const u32 p = g * 32 + lane;
			float4 qlo = make_float4(INFINITY, INFINITY, INFINITY, 0.0f), qhi = make_float4(-INFINITY, -INFINITY, -INFINITY, INFINITY);  // padding: overlaps nothing, prunes everything
			u32 o = 0;
			if (p < K) { qlo = T.mn[0][p]; qhi = T.mx[0][p]; o = order[p]; }
			S.lo[0][lane] = qlo.x; S.lo[1][lane] = qlo.y; S.lo[2][lane] = qlo.z;
			S.hi[0][lane] = qhi.x; S.hi[1][lane] = qhi.y; S.hi[2][lane] = qhi.z;
			S.vol[lane] = qhi.w; S.orig[lane] = o;
		}
		__syncwarp();
		u32 cnt = 0;
		// warp-uniform stack: per level the node whose children are being visited and the mask of children still to expand
		u32 parent[NB_MAX_LEVELS], todo[NB_MAX_LEVELS];
		int level = top;  // "level" = level of the children being tested
		parent[top] = 0; todo[top] = 0;
		bool first = true;
		while (true) {
			u32 base_child, nchild;
			if (first) { base_child = 0; nchild = T.n[top]; first = false; }  // the top level has at most 8 nodes: treat it as one child block
			else {
				// next node to expand: lowest set bit of todo at the current level, else pop
				while (level <= top && todo[level] == 0) ++level;
				if (level > top) break;
				u32 bit = __ffs(todo[level]) - 1;
				todo[level] &= todo[level] - 1;
				u32 node = parent[level] * 8 + bit;  // a node of `level`; its children live one level down
				--level;
				parent[level] = node; todo[level] = 0;
				base_child = node * 8; nchild = min(8u, T.n[level] - base_child);
			}
			// ---- test the children of this block against all 32 members ----
			const u32 child = base_child + c;
			const bool cvalid = c < nchild;
			float4 lo = make_float4(0, 0, 0, 0), hi = lo;
			if (cvalid) { lo = T.mn[level][child]; hi = T.mx[level][child]; }
			u32 hits = 0;  // bit k: member sub*8+k hits this child
			#pragma unroll
			for (u32 k = 0; k < 8; ++k) {
				const u32 mbr = sub * 8 + k;
				// strict interval overlap on three axes: nudge.cpp:3306-3310 / 3386-3390
				bool hit = cvalid && hi.x > S.lo[0][mbr] && S.hi[0][mbr] > lo.x && hi.y > S.lo[1][mbr] && S.hi[1][mbr] > lo.y && hi.z > S.lo[2][mbr] && S.hi[2][mbr] > lo.z;
				const float qvol = S.vol[mbr];
				if (level) hit = hit && !(hi.w < qvol);
				else { const u32 p = g * 32 + mbr; hit = hit && child != p && (hi.w > qvol || (hi.w == qvol && child > p)); }
				hits |= (hit ? 1u : 0u) << k;
			}
			if (level) {
				const u32 m = __ballot_sync(0xffffffffu, hits != 0);
				todo[level] = (m | (m >> 8) | (m >> 16) | (m >> 24)) & 0xffu;  // child accepted if any member subset wants it
				if (level == top) parent[level] = 0;
			}
			else {
				// leaves: every set bit of `hits` is a pair (member sub*8+k, leaf child)
				const u32 mine = __popc(hits);
				u32 incl = mine;
				#pragma unroll
				for (int d = 1; d < 32; d <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (u32)d) incl += t; }
				const u32 total = __shfl_sync(0xffffffffu, incl, 31);
				if (total) {
					u32 at = cnt + incl - mine;
					if (hits) {
						const u32 on = order[child];
						u32 h = hits;
						while (h) {
							const u32 k = __ffs(h) - 1; h &= h - 1;
							const u32 mbr = sub * 8 + k, p = g * 32 + mbr;
							const u32 po = S.orig[mbr];
							// Morton positions: the earlier one goes to the high half, the later one to the low half (nudge.cpp:3495)
							S.buf[at++] = p < child ? (((u64)po << kbits) | on) : (((u64)on << kbits) | po);
						}
					}
					cnt += total;  // at most 31 + 256
					__syncwarp();
					const u32 full = cnt & ~31u;
					for (u32 off = 0; off < full; off += 32) {
						u32 base = 0;
						if (lane == 0) base = atomicAdd(&counts[CNT_PAIRS], 32u);
						base = __shfl_sync(0xffffffffu, base, 0);
						if (base + lane < max_pairs) pair_keys[base + lane] = S.buf[off + lane];
						else atomicOr(&counts[CNT_OVERFLOW], OVF_PAIRS);
					}
					if (full) {
						const u64 rest = S.buf[full + lane];  // only the first cnt - full entries are meaningful
						__syncwarp();
						S.buf[lane] = rest;
						cnt -= full;
						__syncwarp();
					}
				}
				++level;  // back to the level whose todo mask we were consuming
			}
		}
		if (cnt) {
			u32 base = 0;
			if (lane == 0) base = atomicAdd(&counts[CNT_PAIRS], cnt);
			base = __shfl_sync(0xffffffffu, base, 0);
			if (lane < cnt) {
				if (base + lane < max_pairs) pair_keys[base + lane] = S.buf[lane];
				else atomicOr(&counts[CNT_OVERFLOW], OVF_PAIRS);
			}
		}
		__syncwarp();
	}
}

// ---------------- K5b: grid broadphase on top of the Morton order (default) ----------------
// The Morton-sorted colliders are already sorted by octree cell at every level: the colliders of the level-j cell with integer
// coordinates (cx,cy,cz) are the contiguous range of sorted positions whose 48-bit code has the prefix morton(cx,cy,cz) >> 3j.
// Level j is chosen per step so that a cell is at least twice the mean AABB extent.  A collider whose extent (in quantised
// units, +2 for rounding) fits one cell is "small": two overlapping small colliders have min-corner cells that differ by at
// most 1 per axis, so a small collider only has to look into the 27 cells around its own (found through a hash table
// cell -> sorted range).  The few "large" colliders (the ground; also anything whose quantised corner wrapped past 65535,
// nudge.cpp:2613-2616 masks it) are tested against everybody by brute force.  The union is exactly the set of strictly
// overlapping pairs, reported once, oriented by Morton position like nudge.cpp:3495.
#define NB_GRID_EMPTY (~(u64)0)
NB_DEV u64 morton48_of(u32 x, u32 y, u32 z) {
	u32 lx, hx, ly, hy, lz, hz;
	dilate3(x, 2, lx, hx); dilate3(y, 1, ly, hy); dilate3(z, 0, lz, hz);
	return (u64)(lx | ly | lz) | ((u64)(hx | hy | hz) << 32);
}
NB_DEV u32 grid_hash(u64 prefix, u32 mask) { return (u32)((prefix * 0x9E3779B97F4A7C15ull) >> 32) & mask; }

// Cell edge 2^j: the finest level that leaves at most NB_GRID_MAX_LARGE colliders too big for a cell (those go through the
// brute-force kernel, K tests each), read off the histogram k_morton made.
#define NB_GRID_MAX_LARGE 64u
__global__ void k_grid_setup(u32 K, u32* counts) {
	u32 j = 16, above = counts[CNT_EXT_HIST + 17];
	for (u32 l = 16; l >= 1; --l) {  // above = colliders needing a level > l
		if (above <= NB_GRID_MAX_LARGE) j = l;
		above += counts[CNT_EXT_HIST + l];
	}
	counts[CNT_GRID_LEVEL] = j; counts[CNT_LARGE] = 0;
}

__global__ void __launch_bounds__(NB_BLOCK) k_grid_build(u32 K, const u32* order, float4* leaf_min, const float4* leaf_max, const u64* mkeys,
		uint8_t* smallf, u32* large_list, u64* table_keys, u64* table_vals, u32 table_mask, u32* counts) {
	const MortonFrame f = morton_frame(counts);
	const u32 j = counts[CNT_GRID_LEVEL];
	for (u32 p = blockIdx.x * blockDim.x + threadIdx.x; p < K; p += gridDim.x * blockDim.x) {
		float4 lo = leaf_min[p], hi = leaf_max[p];
		u32 qx, qy, qz; morton_quantise(f, lo, order[p], qx, qy, qz);
		bool small = grid_level_needed(lo, hi, f.L[0]) <= j && qx < 65536u && qy < 65536u && qz < 65536u;  // NaN extents are "large" too
		smallf[p] = small ? 1 : 0;
		lo.w = small ? -INFINITY : INFINITY; leaf_min[p] = lo;  // k_grid_pairs reads the class with the box: 0 > lo.w <=> small
		if (!small) large_list[atomicAdd(&counts[CNT_LARGE], 1u)] = p;
		const u64 prefix = mkeys[p] >> (3 * j);
		if (p == 0 || (mkeys[p - 1] >> (3 * j)) != prefix) {  // first collider of its cell: publish the cell's range
			u32 e = p + 1;
			while (e < K && (mkeys[e] >> (3 * j)) == prefix) ++e;
			u32 slot = grid_hash(prefix, table_mask);
			while (true) {
				u64 old = atomicCAS((unsigned long long*)&table_keys[slot], (unsigned long long)NB_GRID_EMPTY, (unsigned long long)prefix);
				if (old == NB_GRID_EMPTY) break;
				slot = (slot + 1) & table_mask;
			}
			table_vals[slot] = (u64)p | ((u64)e << 32);
		}
	}
}

NB_DEV bool boxes_overlap(float4 alo, float4 ahi, float4 blo, float4 bhi) {  // strict, nudge.cpp:3306-3310 / 3386-3390
	return bhi.x > alo.x && ahi.x > blo.x && bhi.y > alo.y && ahi.y > blo.y && bhi.z > alo.z && ahi.z > blo.z;
}
NB_DEV void emit_pair(u32 p, u32 q, const u32* order, u32 kbits, u64* pair_keys, u32 max_pairs, u32* counts) {  // p < q: Morton positions
	u32 slot = warp_append_slot(&counts[CNT_PAIRS]);
	if (slot < max_pairs) pair_keys[slot] = ((u64)order[p] << kbits) | (u64)order[q];  // hi = earlier, lo = later (nudge.cpp:3495)
	else atomicOr(&counts[CNT_OVERFLOW], OVF_PAIRS);
}

// One warp per small collider.  Lanes 0..26 each look up one neighbouring cell (cells whose Morton prefix is below the
// collider's own hold only earlier positions and are skipped); the candidate ranges are then flattened with a warp scan and
// dealt round-robin to all 32 lanes, so the walk takes ceil(candidates / 32) steps however unevenly the cells are filled.
// Hits go to a per-warp shared buffer (shared-memory atomic for the slot) that is flushed to the pair list with one global
// atomic per ~64 pairs: one global atomic per pair would serialise on the counter.
#define NB_GP_BUF 128
__global__ void __launch_bounds__(NB_BLOCK) k_grid_pairs(u32 K, const u32* order, const float4* leaf_min, const float4* leaf_max, const uint8_t* smallf, const u64* mkeys,
		const u64* table_keys, const u64* table_vals, u32 table_mask, u32 kbits, u64* pair_keys, u32 max_pairs, u32* counts) {
	__shared__ u64 buf[NB_WARPS][NB_GP_BUF];
	__shared__ u32 fill[NB_WARPS];
	const MortonFrame f = morton_frame(counts);
	const u32 j = counts[CNT_GRID_LEVEL];
	const int ncell = 1 << (16 - j);
	const u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
	const int dx = (int)(lane % 3) - 1, dy = (int)((lane / 3) % 3) - 1, dz = (int)(lane / 9) - 1;
	if (lane == 0) fill[wid] = 0;
	__syncwarp();
	auto flush = [&]() {  // warp-converged
		u32 cnt = min(fill[wid], (u32)NB_GP_BUF), base = 0;
		if (lane == 0) base = atomicAdd(&counts[CNT_PAIRS], cnt);
		base = __shfl_sync(0xffffffffu, base, 0);
		for (u32 i = lane; i < cnt; i += 32) {
			if (base + i < max_pairs) pair_keys[base + i] = buf[wid][i];
			else atomicOr(&counts[CNT_OVERFLOW], OVF_PAIRS);
		}
		__syncwarp();
		if (lane == 0) fill[wid] = 0;
		__syncwarp();
	};
	for (u32 p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < K; p += nwarps) {  // p is warp-uniform
		if (!smallf[p]) continue;
		const float4 lo = leaf_min[p], hi = leaf_max[p];
		u32 qb = 0, len = 0;
		if (lane < 27) {
			u32 qx, qy, qz; morton_quantise(f, lo, order[p], qx, qy, qz);
			// a cell on the + side only matters if this box reaches into it (max corner quantised like the min corner, +1 for rounding)
			u32 hx, hy, hz; morton_quantise(f, hi, order[p], hx, hy, hz);
			const int cx = (int)(qx >> j), cy = (int)(qy >> j), cz = (int)(qz >> j);
			const bool reach = (dx <= 0 || ((hx + 1) >> j) > (u32)cx) && (dy <= 0 || ((hy + 1) >> j) > (u32)cy) && (dz <= 0 || ((hz + 1) >> j) > (u32)cz);
			const int nx = cx + dx, ny = cy + dy, nz = cz + dz;
			if (reach && nx >= 0 && ny >= 0 && nz >= 0 && nx < ncell && ny < ncell && nz < ncell) {
				const u64 prefix = morton48_of((u32)nx << j, (u32)ny << j, (u32)nz << j) >> (3 * j);
				if (prefix >= (mkeys[p] >> (3 * j))) {
					u32 slot = grid_hash(prefix, table_mask);
					u64 k;
					while ((k = table_keys[slot]) != prefix && k != NB_GRID_EMPTY) slot = (slot + 1) & table_mask;
					if (k != NB_GRID_EMPTY) {
						const u64 range = table_vals[slot];
						qb = max((u32)range, p + 1);  // each pair once: from its earlier Morton position
						const u32 qe = (u32)(range >> 32);
						len = qe > qb ? qe - qb : 0;
					}
				}
			}
		}
		const u32 incl = warp_incl_scan(len), off = incl - len, total = __shfl_sync(0xffffffffu, incl, 31);
		const u64 hi_key = (u64)order[p] << kbits;  // hi = earlier, lo = later (nudge.cpp:3495)
		for (u32 t0 = 0; t0 < total; t0 += 32) {
			const u32 t = t0 + lane;
			// cell of candidate t: the last lane whose offset is <= t (offsets are non-decreasing)
			u32 c = 0;
			#pragma unroll
			for (int step = 16; step; step >>= 1) {
				u32 probe = __shfl_sync(0xffffffffu, off, (c + step) & 31);
				if (c + step < 32 && probe <= t) c += step;
			}
			const u32 q = __shfl_sync(0xffffffffu, qb, c) + (t - __shfl_sync(0xffffffffu, off, c));
			bool hit = false;
			if (t < total) {
				const float4 a = __ldg(leaf_min + q), bq = __ldg(leaf_max + q);  // two 128-bit loads; a.w carries the class
				hit = (bq.x > lo.x) & (hi.x > a.x) & (bq.y > lo.y) & (hi.y > a.y) & (bq.z > lo.z) & (hi.z > a.z) & (0.0f > a.w);  // strict, nudge.cpp:3306-3310
			}
			if (hit) {
				u32 slot = atomicAdd(&fill[wid], 1u);
				if (slot < NB_GP_BUF) buf[wid][slot] = hi_key | (u64)order[q];
				else {  // more than a buffer of hits from one step: straight to the list
					u32 g = atomicAdd(&counts[CNT_PAIRS], 1u);
					if (g < max_pairs) pair_keys[g] = hi_key | (u64)order[q];
					else atomicOr(&counts[CNT_OVERFLOW], OVF_PAIRS);
				}
			}
			__syncwarp();
			if (fill[wid] > NB_GP_BUF - 32) flush();
		}
	}
	if (fill[wid]) flush();
}

__global__ void __launch_bounds__(NB_BLOCK) k_large_pairs(u32 K, const u32* order, const float4* leaf_min, const float4* leaf_max, const uint8_t* smallf,
		const u32* large_list, u32 kbits, u64* pair_keys, u32 max_pairs, u32* counts) {
	const u32 nlarge = counts[CNT_LARGE];
	for (u32 l = 0; l < nlarge; ++l) {
		const u32 pl = large_list[l];
		const float4 lo = leaf_min[pl], hi = leaf_max[pl];
		for (u32 q = blockIdx.x * blockDim.x + threadIdx.x; q < K; q += gridDim.x * blockDim.x) {
			if (q == pl || (!smallf[q] && q < pl)) continue;  // a pair of two large colliders is reported from the earlier one
			if (boxes_overlap(lo, hi, leaf_min[q], leaf_max[q])) emit_pair(min(pl, q), max(pl, q), order, kbits, pair_keys, max_pairs, counts);
		}
	}
}

__global__ void k_clamp_count(u32* counts, int which, u32 cap) {
	if (counts[which] > cap) counts[which] = cap;
}

// ---------------- islands: lock-free union-find, smaller index becomes the root (nudge.cpp:3500-3703, 3788-3971) ----------------
// Invariant: parent[x] <= x, and only a current root is ever hooked (by CAS) under a smaller index.  Finds therefore
// terminate and stay inside x's set even when they read stale L1 lines, so they use ordinary cacheable loads: the root
// of the one giant pile component is read by every thread and would otherwise serialise in a single L2 slice.
NB_DEV u32 uf_find(u32* parent, u32 x) {  // with path halving: every write points a node at one of its ancestors
	while (true) {
		u32 p = parent[x];
		if (p == x) return x;
		u32 gp = parent[p];
		if (gp == p) return p;
		parent[x] = gp;
		x = gp;
	}
}
NB_DEV u32 uf_find_fresh(u32* parent, u32 x) {  // L2-coherent reads, for the final flatten
	while (true) { u32 p = __ldcg(parent + x); if (p == x) return x; x = p; }
}
NB_DEV void uf_unite(u32* parent, u32 a, u32 b) {
	if (!a || !b) return;  // body 0 is the static world and is ignored (nudge.cpp:3517-3519, 3583-3585)
	while (true) {
		a = uf_find(parent, a); b = uf_find(parent, b);
		if (a == b) return;
		if (a < b) { u32 t = a; a = b; b = t; }
		u32 old = atomicCAS(&parent[a], a, b);
		if (old == a) return;
		a = old;  // a was not a root any more: continue from its real parent
	}
}
// Islands and sleeping (nudge.cpp:3500-3703, 3788-3971): a set is active iff ANY member is awake (idle counter != 0xff), so a set
// can only be inactive if ALL its bodies are asleep.  Hence only edges between two asleep bodies need a union; an edge between an
// asleep and an awake body "taints" the asleep one (its whole asleep component is then active), and edges between awake bodies
// carry no information.  The resulting active/inactive decision per body is exactly the reference's, at almost no cost while the
// scene is moving.  active[] is indexed by the root of a body's asleep-component (an awake body is its own root).
NB_DEV void island_edge(u32* parent, u32* taint, const uint8_t* idle, u32 a, u32 b) {
	if (!a || !b || a == b) return;  // body 0 is the static world and is ignored (nudge.cpp:3517-3519, 3583-3585)
	const bool sa = idle[a] == 0xff, sb = idle[b] == 0xff;
	if (sa && sb) uf_unite(parent, a, b);
	else if (sa) taint[a] = 1;
	else if (sb) taint[b] = 1;
}
__global__ void __launch_bounds__(NB_BLOCK) k_uf_init(u32* parent, u32* active, u32* taint, u32 B) {
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) { parent[i] = i; active[i] = 0; taint[i] = 0; }
}
__global__ void __launch_bounds__(NB_BLOCK) k_uf_union_conn(u32* parent, u32* taint, const uint8_t* idle, const nb_body_pair* conn, u32 n) {
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) island_edge(parent, taint, idle, conn[i].a, conn[i].b);
}
__global__ void __launch_bounds__(NB_BLOCK) k_uf_union_pairs(u32* parent, u32* taint, const uint8_t* idle, const u64* pair_keys, u32 kbits, const u32* col_body, const u32* counts) {
	u32 n = counts[CNT_PAIRS];
	u64 mask = ((u64)1 << kbits) - 1;
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		u64 k = pair_keys[i];
		island_edge(parent, taint, idle, col_body[(u32)(k & mask)], col_body[(u32)(k >> kbits)]);
	}
}
__global__ void __launch_bounds__(NB_BLOCK) k_uf_union_contacts(u32* parent, u32* taint, const uint8_t* idle, const uint2* bodies, const u32* counts) {
	u32 n = counts[CNT_STAGED];
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { uint2 ab = bodies[i]; island_edge(parent, taint, idle, ab.x, ab.y); }
}
__global__ void __launch_bounds__(NB_BLOCK) k_uf_flatten_active(u32* parent, u32* active, const u32* taint, const uint8_t* idle, u32 B) {
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
		u32 r = uf_find_fresh(parent, i);
		parent[i] = r;  // roots never change here, so concurrent flattening is benign
		if (i >= 1 && (idle[i] != 0xff || taint[i])) active[r] = 1;  // nudge.cpp:3669-3672
	}
}

// ---------------- coarse island filter + partition by shape type (nudge.cpp:3674-3751) ----------------
__global__ void __launch_bounds__(NB_BLOCK) k_pair_flags(const u64* pair_keys, u32 kbits, u32 nboxes, const u32* col_body, const u32* parent, const u32* active,
														 u32* flags /*[5][stride]*/, u32 stride, const u32* counts) {
	u32 n = counts[CNT_PAIRS];
	u64 mask = ((u64)1 << kbits) - 1;
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		u64 k = pair_keys[i];
		u32 lo = (u32)(k & mask), hi = (u32)(k >> kbits);
		u32 a = col_body[lo], b = col_body[hi];
		int cls = -1;
		if (a != b) {
			u32 set = a ? parent[a] : parent[b];  // sets[0] = 0 and both bodies share a set (nudge.cpp:3663, 3688)
			cls = active[set] ? (int)((lo >= nboxes ? 1u : 0u) | (hi >= nboxes ? 2u : 0u)) : 4;
		}
		#pragma unroll
		for (int c = 0; c < 5; ++c) flags[c*stride + i] = (c == cls) ? 1u : 0u;
	}
}

__global__ void __launch_bounds__(NB_BLOCK) k_partition(const u64* pair_keys, u32 kbits, const u32* flags, const u32* offs, u32 stride,
														const u32* col_tag, uint2* live, u64* sleeping_pairs, u32* counts) {
	u32 n = counts[CNT_PAIRS];
	u64 mask = ((u64)1 << kbits) - 1;
	u32 base[4] = { 0, counts[CNT_LIVE0], counts[CNT_LIVE0] + counts[CNT_LIVE1], counts[CNT_LIVE0] + counts[CNT_LIVE1] + counts[CNT_LIVE2] };
	if (blockIdx.x == 0 && threadIdx.x == 0) counts[CNT_LIVE_TOTAL] = base[3] + counts[CNT_LIVE3];
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		u64 k = pair_keys[i];
		u32 lo = (u32)(k & mask), hi = (u32)(k >> kbits);
		#pragma unroll
		for (int c = 0; c < 4; ++c)
			if (flags[c*stride + i]) live[base[c] + offs[c*stride + i]] = (c == 2) ? make_uint2(hi, lo) : make_uint2(lo, hi);  // bucket 2 halves swapped (nudge.cpp:3746-3751)
		if (flags[4*stride + i]) {
			u64 ta = col_tag[lo], tb = col_tag[hi];
			sleeping_pairs[offs[4*stride + i]] = ta > tb ? ta | (tb << 32) : tb | (ta << 32);  // nudge.cpp:3697
		}
	}
}

// ---------------- narrowphase: box-box (nudge.cpp:1177-2487), one thread per pair ----------------
struct BoxIn { xform t; float3 s; u32 tag; };

NB_DEV void rel_rotation(float4 qa, float4 qb, float* m) {  // nudge.cpp:1228-1268 / 1465-1505
	f3 t = cross3(mk3(qb.x, qb.y, qb.z), mk3(qa.x, qa.y, qa.z));
	float rx = qa.x*qb.w - qb.x*qa.w - t.x;
	float ry = qa.y*qb.w - qb.y*qa.w - t.y;
	float rz = qa.z*qb.w - qb.z*qa.w - t.z;
	float rs = qa.x*qb.x + qa.y*qb.y + qa.z*qb.z + qa.w*qb.w;
	float kx = rx + rx, ky = ry + ry, kz = rz + rz;
	float xx = kx*rx, yy = ky*ry, zz = kz*rz, xy = kx*ry, xz = kx*rz, yz = ky*rz, sx = kx*rs, sy = ky*rs, sz = kz*rs;
	m[0] = 1.0f - yy - zz; m[1] = xy + sz; m[2] = xz - sy;
	m[3] = xy - sz; m[4] = 1.0f - xx - zz; m[5] = yz + sx;
	m[6] = xz + sy; m[7] = yz - sx; m[8] = 1.0f - xx - yy;
}

NB_DEV float sel3(float3 v, u32 k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

// Pass 1: most separating face, nudge.cpp:1195-1410.  Returns false if separated; may swap a/b.
NB_DEV bool bb_faces(const BoxIn& A, const BoxIn& B, float& pen, u32& face, bool& swapped) {
	float m[9]; rel_rotation(A.t.q, B.t.q, m);
	float vx_x = nb_abs(m[0]), vx_y = nb_abs(m[1]), vx_z = nb_abs(m[2]);
	float vy_x = nb_abs(m[3]), vy_y = nb_abs(m[4]), vy_z = nb_abs(m[5]);
	float vz_x = nb_abs(m[6]), vz_y = nb_abs(m[7]), vz_z = nb_abs(m[8]);
	float3 sa = A.s, sb = B.s;
	float pax = sb.x + vx_x*sa.x + vy_x*sa.y + vz_x*sa.z;
	float pay = sb.y + vx_y*sa.x + vy_y*sa.y + vz_y*sa.z;
	float paz = sb.z + vx_z*sa.x + vy_z*sa.y + vz_z*sa.z;
	float pbx = sa.x + vx_x*sb.x + vx_y*sb.y + vx_z*sb.z;
	float pby = sa.y + vy_x*sb.x + vy_y*sb.y + vy_z*sb.z;
	float pbz = sa.z + vz_x*sb.x + vz_y*sb.y + vz_z*sb.z;
	f3 delta = mk3(A.t.p.x - B.t.p.x, A.t.p.y - B.t.p.y, A.t.p.z - B.t.p.z);
	f3 qa = mk3(A.t.q.x, A.t.q.y, A.t.q.z), qb = mk3(B.t.q.x, B.t.q.y, B.t.q.z);
	f3 t = cross3(qb, delta); t = add3(t, t);
	f3 u = cross3(qb, t);
	pax -= nb_abs(u.x + delta.x - B.t.q.w*t.x); pay -= nb_abs(u.y + delta.y - B.t.q.w*t.y); paz -= nb_abs(u.z + delta.z - B.t.q.w*t.z);
	t = cross3(delta, qa); t = add3(t, t);
	u = cross3(qa, t);
	pbx -= nb_abs(u.x - delta.x - A.t.q.w*t.x); pby -= nb_abs(u.y - delta.y - A.t.q.w*t.y); pbz -= nb_abs(u.z - delta.z - A.t.q.w*t.z);
	float payz = nb_min(pay, paz), pbyz = nb_min(pby, pbz);
	float pa = nb_min(pax, payz), pb = nb_min(pbx, pbyz);
	float p = nb_min(pa, pb);
	u32 aface = (payz == pa ? 1u : 0u) + (paz == pa ? 1u : 0u);
	u32 bface = (pbyz == pb ? 1u : 0u) + (pbz == pb ? 1u : 0u);
	swapped = pa == p;  // nudge.cpp:1381-1387
	face = swapped ? aface : bface;
	pen = p;
	return p > 0.0f;
}

struct ContactOut { float4* data; uint2* bodies; u64* tags; u32* features; };

NB_DEV void put_contact(const ContactOut& o, u32 at, float px, float py, float pz, float pen, float nx, float ny, float nz, u32 a, u32 b, u64 tag, u32 feature) {
	o.data[2*at + 0] = make_float4(px, py, pz, pen);
	o.data[2*at + 1] = make_float4(nx, ny, nz, 0.5f);  // friction is fixed: nudge.cpp:2105, 2456, 2515, 2598
	o.bodies[at] = make_uint2(a, b);
	o.tags[at] = tag; o.features[at] = feature;
}

// Pass 2 (nudge.cpp:1432-2136).  A owns the most separating face.  Returns 0 = separated, 1 = face contacts
// (count in n_out; written at `at` when EMIT), 2 = edge candidate (edge_pen / edge_feature / edge_swap set).
// Reserves n (< 32) consecutive slots for each calling lane with one atomic per converged group of lanes.
NB_DEV u32 warp_reserve(u32* counter, u32 n) {
	u32 m = __activemask();
	u32 lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
	u32 prefix = 0, total = 0;
	#pragma unroll
	for (int b = 0; b < 5; ++b) {
		u32 v = __ballot_sync(m, (n >> b) & 1u);
		prefix += __popc(v & lt) << b; total += __popc(v) << b;
	}
	u32 leader = __ffs(m) - 1, base = 0;
	if (lane == leader) base = atomicAdd(counter, total);
	return __shfl_sync(m, base, leader) + prefix;
}

// `reserve` (optional, EMIT only): the contacts go to slots reserved from this counter once their number is known; `at` returns the first slot.
template<bool EMIT>
NB_DEV int bb_face_or_edge(const BoxIn& A, const BoxIn& B, float face_penetration, u32 a_face, const ContactOut& out, u32& at_io, u32 limit,
						   u32& n_out, float& edge_pen, u32& edge_feature, bool& edge_swap, const u32* s_rsqrt, u32* reserve = nullptr) {
	u32 at = at_io;
	float a_to_b[9]; rel_rotation(A.t.q, B.t.q, a_to_b);
	float3 sa = A.s, sb = B.s;
	f3 delta = mk3(A.t.p.x - B.t.p.x, A.t.p.y - B.t.p.y, A.t.p.z - B.t.p.z);
	f3 qa = mk3(A.t.q.x, A.t.q.y, A.t.q.z);
	f3 t = cross3(delta, qa); t = add3(t, t);
	f3 u = cross3(qa, t);
	float b_offset[3] = { u.x - delta.x - A.t.q.w*t.x, u.y - delta.y - A.t.q.w*t.y, u.z - delta.z - A.t.q.w*t.z };

	float epa[9], epb[9];
	#pragma unroll
	for (int i = 0; i < 3; ++i) {  // nudge.cpp:1578-1640
		float acx = a_to_b[0*3 + i], acy = a_to_b[1*3 + i], acz = a_to_b[2*3 + i];
		float bcx = a_to_b[i*3 + 0], bcy = a_to_b[i*3 + 1], bcz = a_to_b[i*3 + 2];
		float ac2x = acx*acx, ac2y = acy*acy, ac2z = acz*acz;
		float bc2x = bcx*bcx, bc2y = bcy*bcy, bc2z = bcz*bcz;
		float aacx = nb_abs(acx), aacy = nb_abs(acy), aacz = nb_abs(acz);
		float abcx = nb_abs(bcx), abcy = nb_abs(bcy), abcz = nb_abs(bcz);
		float ra[3] = { ac2y + ac2z, ac2z + ac2x, ac2x + ac2y };
		float rb[3] = { bc2y + bc2z, bc2z + bc2x, bc2x + bc2y };
		#pragma unroll
		for (int k = 0; k < 3; ++k) {  // rsqrt | cmp_le -> NaN for degenerate axes (nudge.cpp:1611-1619)
			ra[k] = asf(asu(nb_rsqrt_t(ra[k], s_rsqrt)) | (ra[k] <= 1e-3f ? 0xffffffffu : 0u));
			rb[k] = asf(asu(nb_rsqrt_t(rb[k], s_rsqrt)) | (rb[k] <= 1e-3f ? 0xffffffffu : 0u));
		}
		float pa0 = aacy*sa.z + aacz*sa.y, pa1 = aacz*sa.x + aacx*sa.z, pa2 = aacx*sa.y + aacy*sa.x;
		float pb0 = abcy*sb.z + abcz*sb.y, pb1 = abcz*sb.x + abcx*sb.z, pb2 = abcx*sb.y + abcy*sb.x;
		float o0 = nb_abs(acy*b_offset[2] - acz*b_offset[1]);
		float o1 = nb_abs(acz*b_offset[0] - acx*b_offset[2]);
		float o2 = nb_abs(acx*b_offset[1] - acy*b_offset[0]);
		epa[i*3 + 0] = (pa0 - o0) * ra[0]; epa[i*3 + 1] = (pa1 - o1) * ra[1]; epa[i*3 + 2] = (pa2 - o2) * ra[2];
		epb[i*3 + 0] = pb0 * rb[0]; epb[i*3 + 1] = pb1 * rb[1]; epb[i*3 + 2] = pb2 * rb[2];
	}
	u32 a_edge = 0, b_edge = 0;
	float penetration = face_penetration;
	#pragma unroll
	for (int i = 0; i < 3; ++i)
		#pragma unroll
		for (int j = 0; j < 3; ++j) {  // nudge.cpp:1647-1657
			float p = epa[i*3 + j] + epb[j*3 + i];
			bool mk = penetration > p;
			penetration = nb_min(penetration, p);
			if (mk) { a_edge = j; b_edge = i; }
		}
	bool is_edge = face_penetration > penetration + 1e-3f;  // nudge.cpp:1659-1661
	n_out = 0;
	if (!(penetration > 0.0f)) return 0;
	if (is_edge) {  // nudge.cpp:2116-2135
		edge_pen = penetration;
		bool keep = A.tag > B.tag;
		edge_feature = keep ? a_edge | (b_edge << 16) : b_edge | (a_edge << 16);
		edge_swap = !keep;
		return 2;
	}

	// ---- face-face: nudge.cpp:1678-2112 ----
	float dirs0 = nb_abs(a_to_b[a_face*3 + 0]), dirs1 = nb_abs(a_to_b[a_face*3 + 1]), dirs2 = nb_abs(a_to_b[a_face*3 + 2]);
	bool bit1 = dirs1 >= nb_max(dirs2, dirs0), bit2 = dirs2 >= nb_max(dirs1, dirs0);  // nudge.cpp:1721-1723
	float c0[3] = { a_to_b[0] * sb.x, a_to_b[3] * sb.x, a_to_b[6] * sb.x };
	float c1[3] = { a_to_b[1] * sb.y, a_to_b[4] * sb.y, a_to_b[7] * sb.y };
	float c2[3] = { a_to_b[2] * sb.z, a_to_b[5] * sb.z, a_to_b[8] * sb.z };
	u32 b_face = 0;
	float cc[3], dx[3], dy[3];
	if (bit2) { b_face = 2; for (int k = 0; k < 3; ++k) { cc[k] = c2[k]; dx[k] = c0[k]; dy[k] = c1[k]; } }
	else if (bit1) { b_face = 1; for (int k = 0; k < 3; ++k) { cc[k] = c1[k]; dx[k] = c2[k]; dy[k] = c0[k]; } }
	else { for (int k = 0; k < 3; ++k) { cc[k] = c0[k]; dx[k] = c1[k]; dy[k] = c2[k]; } }
	u32 b_positive_face_bit = ((asu(b_offset[a_face]) ^ asu(cc[a_face])) >> 31) << a_face;
	u32 b_offset_neg = (asu(b_offset[a_face]) >> 31) << a_face;
	if (!b_positive_face_bit) for (int k = 0; k < 3; ++k) cc[k] = nb_neg(cc[k]);
	for (int k = 0; k < 3; ++k) cc[k] += b_offset[k];

	u32 X = (a_face + 1) % 3, Y = (a_face + 2) % 3, Z = a_face;
	float sx = sel3(sa, X), sy = sel3(sa, Y), cx = cc[X], cy = cc[Y];
	float d0 = dx[X], d1 = dx[Y], d2 = dy[X], d3 = dy[Y];

	float support_x[16], support_y[16];
	u32 mask = 0;
	u32 edge_axis_near = 0, edge_axis_far = 0;
	{
		const u32 npnp[4] = { NB_SIGN, 0, NB_SIGN, 0 }, pnpn[4] = { 0, NB_SIGN, 0, NB_SIGN }, nnpp[4] = { NB_SIGN, NB_SIGN, 0, 0 };
		bool mask0[4], mask1[4];
		float k0 = cx*d3 - cy*d2, k1 = cx*d1 - cy*d0, k2 = d0*d3 - d1*d2;  // nudge.cpp:1814-1815
		float ox = k0, oy = k1, delta_max = nb_abs(k2);
		float sd0 = d0*sy, sd1 = d1*sx, sd2 = d2*sy, sd3 = d3*sx;           // nudge.cpp:1821
		#pragma unroll
		for (int l = 0; l < 4; ++l) {
			float corner1x = cx + nb_xor(d0, npnp[l]) + nb_xor(d2, nnpp[l]);
			float corner1y = cy + nb_xor(d1, npnp[l]) + nb_xor(d3, nnpp[l]);
			float delta_x = ox + nb_xor(sd2, nnpp[l]) + nb_xor(sd3, npnp[l]);
			float delta_y = oy + nb_xor(sd0, nnpp[l]) + nb_xor(sd1, npnp[l]);
			mask0[l] = nb_max(nb_abs(delta_x), nb_abs(delta_y)) <= delta_max;
			mask1[l] = (nb_abs(corner1x) <= sx) && (nb_abs(corner1y) <= sy);
			support_x[l] = nb_xor(sx, pnpn[l]); support_y[l] = nb_xor(sy, nnpp[l]);
			support_x[4 + l] = corner1x; support_y[4 + l] = corner1y;
		}
		// Don't allow edge intersections if both vertices are inside: nudge.cpp:1834-1836
		bool pre[8] = { mask0[3] && mask0[1], mask0[2] && mask0[0], mask0[0] && mask0[1], mask0[2] && mask0[3],
						mask1[1] && mask1[0], mask1[3] && mask1[2], mask1[2] && mask1[0], mask1[3] && mask1[1] };
		float dxy[4] = { d0, d1, d2, d3 };
		float rdxy[4] = { 1.0f/d0, 1.0f/d1, 1.0f/d2, 1.0f/d3 };  // nudge.cpp:1849
		#pragma unroll
		for (int l = 0; l < 4; ++l) {
			const int i02 = (l < 2) ? 0 : 2, i13 = (l < 2) ? 1 : 3, i20 = (l < 2) ? 2 : 0, i31 = (l < 2) ? 3 : 1;
			float offset_x = dxy[i02], offset_y = dxy[i13];
			float pivot_x = cx + nb_xor(dxy[i20], npnp[l]);
			float pivot_y = cy + nb_xor(dxy[i31], npnp[l]);
			float pos_x = asf((asu(offset_x) & NB_SIGN) | asu(sx));  // copy sign: nudge.cpp:1858-1859
			float pos_y = asf((asu(offset_y) & NB_SIGN) | asu(sy));
			float rx = rdxy[i02], ry = rdxy[i13];
			float near_x = (pos_x + pivot_x) * rx, far_x = (pos_x - pivot_x) * rx;
			float near_y = (pos_y + pivot_y) * ry, far_y = (pos_y - pivot_y) * ry;
			float ea = nb_min(1.0f, near_x), eb = nb_min(1.0f, far_x);
			if (ea > near_y) edge_axis_near |= 1u << l;
			if (eb > far_y) edge_axis_far |= 1u << l;
			ea = nb_min(ea, near_y); eb = nb_min(eb, far_y);
			bool mm = (ea + eb) > 0.0f;
			bool mask_a = !(ea == 1.0f) && mm;  // _mm_cmpneq_ps is unordered: true on NaN (nudge.cpp:328-330, 1886)
			bool mask_b = !(eb == 1.0f) && mm;
			support_x[8 + l] = pivot_x - offset_x * ea; support_y[8 + l] = pivot_y - offset_y * ea;
			support_x[12 + l] = pivot_x + offset_x * eb; support_y[12 + l] = pivot_y + offset_y * eb;
			mask |= (mask0[l] ? 1u : 0u) << l;
			mask |= (mask1[l] ? 1u : 0u) << (4 + l);
			mask |= ((!pre[l] && mask_a) ? 1u : 0u) << (8 + l);
			mask |= ((!pre[4 + l] && mask_b) ? 1u : 0u) << (12 + l);
		}
	}

	// z-plane through face b: nudge.cpp:1973-2019
	float plane0, plane1, plane2;
	{
		float dxt[3] = { dx[X], dx[Y], dx[Z] }, dyt[3] = { dy[X], dy[Y], dy[Z] }, ct[3] = { cc[X], cc[Y], cc[Z] };
		float zn0 = dxt[1]*dyt[2] - dxt[2]*dyt[1];
		float zn1 = dxt[2]*dyt[0] - dxt[0]*dyt[2];
		float zn2 = dxt[0]*dyt[1] - dxt[1]*dyt[0];
		float dt = ct[0]*zn0 + ct[1]*zn1 + ct[2]*zn2;
		float inv = 1.0f / zn2;
		plane0 = nb_neg(zn0) * inv; plane1 = nb_neg(zn1) * inv; plane2 = dt * inv;
	}
	u32 z_sign = b_offset_neg ? NB_SIGN : 0;
	float penetration_offset = sel3(sa, Z);
	float pens[16], support_z[16];
	u32 penetration_mask = 0;
	#pragma unroll
	for (int i = 0; i < 16; ++i) {
		float x = support_x[i], y = support_y[i];
		float z = x*plane0 + y*plane1 + plane2;
		float pen = penetration_offset - nb_xor(z, z_sign);
		z += pen * nb_xor(0.5f, z_sign);
		if (pen > 0.0f) penetration_mask |= 1u << i;
		pens[i] = pen; support_z[i] = z;
	}
	mask &= penetration_mask;
	n_out = __popc(mask);
	if (!EMIT) return 1;
	if (reserve) { at = warp_reserve(reserve, n_out); at_io = at; }

	// labels: nudge.cpp:1902-1970
	u32 a_sign_face_bit = b_offset_neg ? (1u << a_face) : 0;
	u32 b_sign_face_bit = b_positive_face_bit ? 0 : (1u << b_face);
	u32 a_vertices = 0x12003624u >> (3 - a_face);
	u32 b_vertices = 0x00122436u >> (3 - b_face);
	u32 a_face_bits = 0xffff0000u | a_sign_face_bit;
	u32 b_face_bits = 0x0000ffffu | (b_sign_face_bit << 16);
	u32 winding = (asu(d0) >> 31) | ((asu(d1) >> 31) << 1) | ((asu(d2) >> 31) << 2) | ((asu(d3) >> 31) << 3);
	u64 a_edge_map = 0x1200362424003612llu >> (3 - a_face);
	u64 b_edge_map = 0x2400361212003624llu >> (3 - b_face);
	u32 face_bits = a_sign_face_bit | (a_sign_face_bit << 8) | (b_sign_face_bit << 16) | (b_sign_face_bit << 24);

	// a to world: nudge.cpp:2028-2056 (the diagonal is built as -((p + q) - 1))
	float w0[3], w1[3], w2[3];
	{
		float qx = A.t.q.x, qy = A.t.q.y, qz = A.t.q.z, qs = A.t.q.w;
		float kx = qx + qx, ky = qy + qy, kz = qz + qz, ks = nb_neg(qs + qs);
		w0[0] = nb_neg((ky*qy + kz*qz) - 1.0f); w0[1] = kx*qy + kz*qs; w0[2] = kx*qz + ks*qy;
		w1[0] = kx*qy + ks*qz; w1[1] = nb_neg((kz*qz + kx*qx) - 1.0f); w1[2] = ky*qz + kx*qs;
		w2[0] = kx*qz + ky*qs; w2[1] = ky*qz + ks*qx; w2[2] = nb_neg((kx*qx + ky*qy) - 1.0f);
	}
	float wn[3];
	for (int k = 0; k < 3; ++k) wn[k] = a_face == 0 ? w0[k] : (a_face == 1 ? w1[k] : w2[k]);
	if (b_offset_neg) for (int k = 0; k < 3; ++k) wn[k] = nb_neg(wn[k]);
	u32 a_body = asu(A.t.p.w), b_body = asu(B.t.p.w);
	u32 a_tag = A.tag, b_tag = B.tag;
	bool tag_swap = false;
	if (b_tag > a_tag) {  // nudge.cpp:2074-2087
		u32 tt = a_tag; a_tag = b_tag; b_tag = tt;
		tt = a_body; a_body = b_body; b_body = tt;
		tag_swap = true;
		for (int k = 0; k < 3; ++k) wn[k] = nb_neg(wn[k]);
	}
	u64 high_tag = (u64)a_tag | ((u64)b_tag << 32);
	u32 afi = (a_face ^ 1) ^ (a_face >> 1);  // nudge.cpp:2022
	u32 iX = (afi + 1) % 3, iY = (afi + 2) % 3, iZ = afi;
	while (mask) {
		u32 index = __ffs(mask) - 1;
		mask &= mask - 1;
		// support tag of point `index` (nudge.cpp:1914-1970), computed on demand
		u32 st;
		if (index < 4) st = ((a_vertices >> (8 * index)) & 0x7) | a_face_bits;
		else if (index < 8) {
			u32 l = index - 4;
			u32 bv = l == 0 ? (b_vertices << 16) : (l == 1 ? (b_vertices << 8) : (l == 2 ? b_vertices : (b_vertices >> 8)));
			st = (bv & 0x70000) | b_face_bits;
		}
		else {
			u32 l = index & 3;
			bool is_far = index >= 12;
			u32 w = is_far ? (winding ^ 0xf) : winding;
			u32 yb = ((is_far ? edge_axis_far : edge_axis_near) >> l) & 1;
			u32 e = yb*2 + ((w >> ((l < 2 ? 0 : 2) + yb)) & 1);
			u32 b_edge_l = ((u32)((b_edge_map >> (l << 4)) & 0x0707) << 16) | face_bits;
			st = (u32)((a_edge_map >> (e << 4)) & 0x0707) | b_edge_l;
		}
		float sp[3] = { support_x[index], support_y[index], support_z[index] };
		float lx = sp[iX], ly = sp[iY], lz = sp[iZ];
		float wx = w0[0]*lx + w1[0]*ly + w2[0]*lz + A.t.p.x;
		float wy = w0[1]*lx + w1[1]*ly + w2[1]*lz + A.t.p.y;
		float wz = w0[2]*lx + w1[2]*ly + w2[2]*lz + A.t.p.z;
		u32 feature = tag_swap ? ((st >> 16) | (st << 16)) : st;  // nudge.cpp:2108
		if (at < limit) put_contact(out, at, wx, wy, wz, pens[index], wn[0], wn[1], wn[2], a_body, b_body, high_tag, feature);
		++at;
	}
	return 1;
}

// Pass 3: edge-edge closest points, nudge.cpp:2157-2479.
NB_DEV void bb_edge(const BoxIn& A, const BoxIn& B, float pen, u32 edge, const ContactOut& out, u32 at, const u32* s_rsqrt) {
	float ab[3][3], bb[3][3];
	#pragma unroll
	for (int w = 0; w < 2; ++w) {
		float4 q = w ? B.t.q : A.t.q;
		float kx = q.x + q.x, ky = q.y + q.y, kz = q.z + q.z;
		float xx = kx*q.x, yy = ky*q.y, zz = kz*q.z, xy = kx*q.y, xz = kx*q.z, yz = ky*q.z, sx = kx*q.w, sy = ky*q.w, sz = kz*q.w;
		float (*m)[3] = w ? bb : ab;
		m[0][0] = 1.0f - yy - zz; m[0][1] = xy + sz; m[0][2] = xz - sy;
		m[1][0] = xy - sz; m[1][1] = 1.0f - xx - zz; m[1][2] = yz + sx;
		m[2][0] = xz + sy; m[2][1] = yz - sx; m[2][2] = 1.0f - xx - yy;
	}
	u32 ua = (edge & 2) ? 2 : ((edge & 1) ? 1 : 0);                       // blendv on shifted bits, nudge.cpp:2257-2278
	u32 ub = (edge & (2u << 16)) ? 2 : ((edge & (1u << 16)) ? 1 : 0);
	f3 u = mk3(ab[ua][0], ab[ua][1], ab[ua][2]);
	f3 v = mk3(bb[ub][0], bb[ub][1], bb[ub][2]);
	f3 n = cross3(u, v);
	f3 delta = mk3(B.t.p.x - A.t.p.x, B.t.p.y - A.t.p.y, B.t.p.z - A.t.p.z);
	u32 flip = asu(n.x*delta.x + n.y*delta.y + n.z*delta.z) & NB_SIGN;
	n.x = nb_xor(n.x, flip); n.y = nb_xor(n.y, flip); n.z = nb_xor(n.z, flip);
	float sa[3] = { A.s.x, A.s.y, A.s.z }, sb[3] = { B.s.x, B.s.y, B.s.z };
	u32 asg[3], bsg[3];
	#pragma unroll
	for (int k = 0; k < 3; ++k) {
		asg[k] = asu(ab[k][0]*n.x + ab[k][1]*n.y + ab[k][2]*n.z) & NB_SIGN;
		bsg[k] = asu(bb[k][0]*n.x + bb[k][1]*n.y + bb[k][2]*n.z) & NB_SIGN;
	}
	u32 edge_x = (asg[0] >> 31) | ((bsg[0] ^ NB_SIGN) >> 15);
	u32 edge_y = (asg[1] >> 30) | ((bsg[1] ^ NB_SIGN) >> 14);
	u32 edge_z = (asg[2] >> 29) | ((bsg[2] ^ NB_SIGN) >> 13);
	u32 elo = edge & 0xffff, ehi = edge >> 16;  // per 16-bit lane (e + 1) + (e >> 1) == 1 << e for e in 0..2 (nudge.cpp:2381)
	u32 edge_w = (((elo + 1) + (elo >> 1)) & 0xffff) | ((((ehi + 1) + (ehi >> 1)) & 0xffff) << 16);
	u32 tag_hi = edge_x | edge_y | edge_z | edge_w;
	u32 tag_lo = tag_hi & ~edge_w;
	u32 tag = tag_lo | (tag_hi << 8);
	#pragma unroll
	for (int k = 0; k < 3; ++k) { sa[k] = nb_xor(sa[k], asg[k]); sb[k] = nb_xor(sb[k], bsg[k]); }
	#pragma unroll
	for (int k = 0; k < 3; ++k)
		#pragma unroll
		for (int j = 0; j < 3; ++j) { ab[k][j] *= sa[k]; bb[k][j] *= sb[k]; }
	float apos[3] = { A.t.p.x, A.t.p.y, A.t.p.z }, bpos[3] = { B.t.p.x, B.t.p.y, B.t.p.z };
	float ca[3], cb[3], o[3];
	#pragma unroll
	for (int j = 0; j < 3; ++j) {
		ca[j] = ab[0][j] + ab[1][j] + ab[2][j] + apos[j];
		cb[j] = bb[0][j] + bb[1][j] + bb[2][j] - bpos[j];  // negated on purpose (nudge.cpp:2428)
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
	float px = (ca[0] - cb[0])*0.5f + u.x*s_a + v.x*s_b;
	float py = (ca[1] - cb[1])*0.5f + u.y*s_a + v.y*s_b;
	float pz = (ca[2] - cb[2])*0.5f + u.z*s_a + v.z*s_b;
	float fn = nb_rsqrt_t(n.x*n.x + n.y*n.y + n.z*n.z, s_rsqrt);  // nudge.cpp:842-847, 2453
	put_contact(out, at, px, py, pz, pen, n.x*fn, n.y*fn, n.z*fn, asu(A.t.p.w), asu(B.t.p.w), (u64)A.tag | ((u64)B.tag << 32), tag);
}

NB_DEV bool sphere_sphere(float ra, float rb, xform ta, xform tb, float* o /* p[3], pen, n[3] */) {  // nudge.cpp:2489-2521
	float r = ra + rb;
	f3 pa = mk3(ta.p.x, ta.p.y, ta.p.z);
	f3 dp = sub3(mk3(tb.p.x, tb.p.y, tb.p.z), pa);
	float l2 = dot3(dp, dp);
	if (l2 > r*r) return false;
	f3 n;
	float l = sqrtf(l2);
	if (l2 > 1e-4f) n = mul3(dp, 1.0f / l);
	else n = mk3(1.0f, 0.0f, 0.0f);
	f3 p = add3(pa, mul3(n, l - rb));
	o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = r - l; o[4] = n.x; o[5] = n.y; o[6] = n.z;
	return true;
}

NB_DEV bool box_sphere(float3 size, float radius, xform ta, xform tb, float* o) {  // nudge.cpp:2523-2604
	quat a_to_world = mkq(ta.q);
	quat world_to_a = a_to_world; world_to_a.v.x = -world_to_a.v.x; world_to_a.v.y = -world_to_a.v.y; world_to_a.v.z = -world_to_a.v.z;
	f3 apos = mk3(ta.p.x, ta.p.y, ta.p.z);
	f3 offset_b = qrot(world_to_a, sub3(mk3(tb.p.x, tb.p.y, tb.p.z), apos));
	float dx = fabsf(offset_b.x), dy = fabsf(offset_b.y), dz = fabsf(offset_b.z);
	float w = size.x + radius, h = size.y + radius, d = size.z + radius;
	if (dx >= w || dy >= h || dz >= d) return false;
	f3 n; float penetration; float r = radius;
	u32 outside_x = dx > size.x, outside_y = dy > size.y, outside_z = dz > size.z;
	if (outside_x + outside_y + outside_z >= 2) {
		f3 corner = mk3(outside_x ? (offset_b.x > 0.0f ? size.x : -size.x) : offset_b.x,
						outside_y ? (offset_b.y > 0.0f ? size.y : -size.y) : offset_b.y,
						outside_z ? (offset_b.z > 0.0f ? size.z : -size.z) : offset_b.z);
		f3 dp = sub3(offset_b, corner);
		float l2 = dot3(dp, dp);
		if (l2 > r*r) return false;
		float l = sqrtf(l2);
		float m = 1.0f / l;
		n = mul3(dp, m);
		penetration = r - l;
	}
	else if (w - dx < h - dy && w - dx < d - dz) { n = mk3(offset_b.x > 0.0f ? 1.0f : -1.0f, 0.0f, 0.0f); penetration = w - dx; }
	else if (h - dy < d - dz) { n = mk3(0.0f, offset_b.y > 0.0f ? 1.0f : -1.0f, 0.0f); penetration = h - dy; }
	else { n = mk3(0.0f, 0.0f, offset_b.z > 0.0f ? 1.0f : -1.0f); penetration = d - dz; }
	f3 p = sub3(offset_b, mul3(n, r));
	p = add3(qrot(a_to_world, p), apos);
	n = qrot(a_to_world, n);
	o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = penetration; o[4] = n.x; o[5] = n.y; o[6] = n.z;
	return true;
}

NB_DEV BoxIn load_box(const nb_transform* world_xf, const nb_box_collider* box_data, const u32* col_tag, u32 i) {
	BoxIn b;
	b.t = ld_xform(world_xf, i);
	float4 s = reinterpret_cast<const float4*>(box_data)[i];
	b.s = make_float3(s.x, s.y, s.z);
	b.tag = col_tag[i];
	return b;
}

// The narrowphase runs in three thread-per-pair passes so that the expensive face clipping only runs on dense warps:
//   k_np_faces  every live pair: box-box -> pass 1 (face SAT, cheap, rejects ~85 %); sphere pairs -> the whole (cheap) test
//   k_np_clip   surviving box-box pairs (compacted): pass 2 (+ pass 3 for edge contacts) -> contacts in a scratch buffer + counts
//   k_np_emit   moves the scratch contacts (and computes the sphere contacts) to the scanned offsets, in the reference's contact order
//               (box-box face contacts, box-box edge contacts, box-sphere, sphere-sphere: nudge.cpp:3753-3786)

__global__ void __launch_bounds__(NB_BLOCK) k_np_faces(const uint2* live, u32 nboxes, const nb_transform* world_xf, const nb_box_collider* box_data,
		const nb_sphere_collider* sph_data, const u32* col_tag, u32* cnt /*[4][stride]: face, edge, other, survivor*/, u32 stride, float* np_pen, u32* np_info, u32* counts) {
	u32 n = counts[CNT_LIVE_TOTAL];
	u32 n_bb = counts[CNT_LIVE0], n_bs_end = n_bb + counts[CNT_LIVE1] + counts[CNT_LIVE2];
	for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
		uint2 pr = live[j];  // x = low half, y = high half of the reference's pair word
		u32 surv = 0, other = 0;
		if (j < n_bb) {
			BoxIn A = load_box(world_xf, box_data, col_tag, pr.x), B = load_box(world_xf, box_data, col_tag, pr.y);  // a = low, b = high (nudge.cpp:1202-1203)
			float pen; u32 face; bool swapped;
			if (bb_faces(A, B, pen, face, swapped)) { surv = 1; np_pen[j] = pen; np_info[j] = face | (swapped ? 4u : 0u); }
		}
		else {
			float o[7];
			u32 a = pr.y, b = pr.x;  // a = high half, b = low half (nudge.cpp:3759-3760, 3775-3776)
			xform ta = ld_xform(world_xf, a), tb = ld_xform(world_xf, b);
			bool hit;
			if (j < n_bs_end) { float4 sz = reinterpret_cast<const float4*>(box_data)[a]; hit = box_sphere(make_float3(sz.x, sz.y, sz.z), sph_data[b - nboxes].radius, ta, tb, o); }
			else hit = sphere_sphere(sph_data[a - nboxes].radius, sph_data[b - nboxes].radius, ta, tb, o);
			other = hit ? 1u : 0u;
		}
		cnt[0*stride + j] = 0; cnt[1*stride + j] = 0; cnt[2*stride + j] = other; cnt[3*stride + j] = surv;
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) counts[CNT_TMP] = 0;
}

__global__ void __launch_bounds__(NB_BLOCK) k_np_list(const u32* cnt, const u32* offs, u32 stride, u32* np_list, const u32* counts) {
	u32 n = counts[CNT_LIVE0];
	for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
		if (cnt[3*stride + j]) np_list[offs[3*stride + j]] = j;
}

// Clips the surviving box-box pairs ONCE: the contacts go to a scratch contact buffer in arrival order (np_start[j] = first slot),
// and k_np_emit moves them to their final, pair-ordered place after the scan.
__global__ void __launch_bounds__(NB_BLOCK) k_np_clip(const uint2* live, const u32* np_list, const float* np_pen, const u32* np_info, const nb_transform* world_xf,
		const nb_box_collider* box_data, const u32* col_tag, u32* cnt, u32 stride, ContactOut tmp, u32* np_start, u32 max_contacts, u32* counts) {
	// the rsqrtps table goes to shared memory: 32 different indices per warp would serialise on the constant cache
	__shared__ u32 s_rsqrt[2048];
	for (u32 i = threadIdx.x; i < 2048; i += blockDim.x) s_rsqrt[i] = g_rsqrt_lut[i];
	__syncthreads();
	u32 n = counts[CNT_SURV];
	for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
		u32 j = np_list[k];
		uint2 pr = live[j];
		u32 info = np_info[j];
		BoxIn A = load_box(world_xf, box_data, col_tag, (info & 4) ? pr.y : pr.x), B = load_box(world_xf, box_data, col_tag, (info & 4) ? pr.x : pr.y);  // a owns the best face (nudge.cpp:1381-1387)
		float epen; u32 efeat; bool eswap; u32 nface = 0, at = 0;
		int kind = bb_face_or_edge<true>(A, B, np_pen[j], info & 3, tmp, at, max_contacts, nface, epen, efeat, eswap, s_rsqrt, &counts[CNT_TMP]);
		if (kind == 2) {
			at = warp_reserve(&counts[CNT_TMP], 1);
			if (eswap) { BoxIn t = A; A = B; B = t; }
			if (at < max_contacts) bb_edge(A, B, epen, efeat, tmp, at, s_rsqrt);
		}
		np_start[j] = at;
		cnt[0*stride + j] = kind == 1 ? nface : 0;
		cnt[1*stride + j] = kind == 2 ? 1u : 0u;
	}
}

NB_DEV void move_contact(const ContactOut& from, u32 src, const ContactOut& to, u32 dst) {
	to.data[2*dst + 0] = from.data[2*src + 0]; to.data[2*dst + 1] = from.data[2*src + 1];
	to.bodies[dst] = from.bodies[src]; to.tags[dst] = from.tags[src]; to.features[dst] = from.features[src];
}

__global__ void __launch_bounds__(NB_BLOCK) k_np_emit(const uint2* live, const u32* np_list, const u32* np_start, ContactOut tmp, u32 nboxes, const nb_transform* world_xf,
		const nb_box_collider* box_data, const nb_sphere_collider* sph_data, const u32* col_tag, const u32* cnt, const u32* offs, u32 stride,
		ContactOut out, u32 max_contacts, u32* counts) {
	const u32 n_surv = counts[CNT_SURV], n_bb = counts[CNT_LIVE0], n_all = counts[CNT_LIVE_TOTAL];
	const u32 n_bs_end = n_bb + counts[CNT_LIVE1] + counts[CNT_LIVE2];
	const u32 edge_base = counts[CNT_FACE], other_base = edge_base + counts[CNT_EDGE];
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		u32 staged = other_base + counts[CNT_OTHER];
		if (staged > max_contacts || counts[CNT_TMP] > max_contacts) { atomicOr(&counts[CNT_OVERFLOW], OVF_CONTACTS); staged = min(staged, max_contacts); }
		counts[CNT_STAGED] = staged;
	}
	const u32 total = n_surv + (n_all - n_bb);
	for (u32 t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
		if (t < n_surv) {
			u32 j = np_list[t];
			u32 nf = cnt[0*stride + j], ne = cnt[1*stride + j];
			if (!(nf | ne)) continue;
			u32 src = np_start[j], dst = nf ? offs[0*stride + j] : edge_base + offs[1*stride + j];
			for (u32 i = 0; i < nf + ne; ++i)
				if (src + i < max_contacts && dst + i < max_contacts) move_contact(tmp, src + i, out, dst + i);
		}
		else {
			u32 j = n_bb + (t - n_surv);
			if (!cnt[2*stride + j]) continue;
			uint2 pr = live[j];
			float o[7];
			u32 a = pr.y, b = pr.x;
			xform ta = ld_xform(world_xf, a), tb = ld_xform(world_xf, b);
			if (j < n_bs_end) { float4 sz = reinterpret_cast<const float4*>(box_data)[a]; box_sphere(make_float3(sz.x, sz.y, sz.z), sph_data[b - nboxes].radius, ta, tb, o); }
			else sphere_sphere(sph_data[a - nboxes].radius, sph_data[b - nboxes].radius, ta, tb, o);
			u32 at = other_base + offs[2*stride + j];
			if (at < max_contacts)
				put_contact(out, at, o[0], o[1], o[2], o[3], o[4], o[5], o[6], asu(ta.p.w), asu(tb.p.w), (u64)col_tag[a] | ((u64)col_tag[b] << 32), 0);  // nudge.cpp:3767, 3784
		}
	}
}

// ---------------- fine islands: active bodies + contact compaction (nudge.cpp:3965-4006) ----------------
__global__ void __launch_bounds__(NB_BLOCK) k_body_flags(const u32* parent, const u32* active, u32* flags, u32 B) {
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x)
		flags[i] = (i >= 1 && active[parent[i]]) ? 1u : 0u;
}
__global__ void __launch_bounds__(NB_BLOCK) k_active_scatter(const u32* flags, const u32* offs, u32* active_idx, u32 B) {
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x)
		if (flags[i]) active_idx[offs[i]] = i;
}
__global__ void __launch_bounds__(NB_BLOCK) k_contact_flags(const uint2* bodies, const u64* tags, const u32* parent, const u32* active, u32* flags, u32 stride, const u32* counts) {
	u32 n = counts[CNT_STAGED];
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		uint2 ab = bodies[i];
		// The reference decides per span of equal pair tags from the span's first contact (nudge.cpp:3976-3999);
		// every contact of a span has the same bodies, so the per-contact decision is the same.
		u32 set = ab.x ? parent[ab.x] : parent[ab.y];
		bool keep = active[set] != 0;
		flags[i] = keep ? 1u : 0u;
		flags[stride + i] = (!keep && (i == 0 || tags[i - 1] != tags[i])) ? 1u : 0u;
	}
}
__global__ void __launch_bounds__(NB_BLOCK) k_contact_compact(ContactOut st, ContactOut fin, const u32* flags, const u32* offs, u32 stride, u64* sleeping_pairs, u32* counts) {
	u32 n = counts[CNT_STAGED];
	u32 sleep_base = counts[CNT_SLEEP_COARSE];
	if (blockIdx.x == 0 && threadIdx.x == 0) counts[CNT_SLEEPING] = sleep_base + counts[CNT_SLEEP_FINE];
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		if (flags[i]) {
			u32 d = offs[i];
			fin.data[2*d] = st.data[2*i]; fin.data[2*d + 1] = st.data[2*i + 1];
			fin.bodies[d] = st.bodies[i]; fin.tags[d] = st.tags[i]; fin.features[d] = st.features[i];
		}
		if (flags[stride + i]) sleeping_pairs[sleep_base + offs[stride + i]] = st.tags[i];
	}
}
