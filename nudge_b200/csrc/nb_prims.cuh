// nudge_b200 — device-wide primitives: multi-counter exclusive scan, stable LSD radix sort, grid barrier.
//
// All element counts are read from device memory (the step never synchronises with the host), so every
// kernel runs on a fixed grid sized for the SM count and walks its share with a stride or a contiguous tile.
// These replace the reference's sequential radix sorts (nudge.cpp:2647-2836) and its many serial
// "count, then append" loops.
#pragma once
#include "nb_common.cuh"

#define NB_BLOCK 256
#define NB_WARPS (NB_BLOCK/32)

struct Launch { cudaStream_t stream; unsigned long long* counter; int sms; };

static inline unsigned nb_grid_for(unsigned n_cap, int sms, int per_sm = 8) {
	unsigned g = (n_cap + NB_BLOCK - 1) / NB_BLOCK;
	unsigned cap = (unsigned)(sms * per_sm);
	if (g > cap) g = cap;
	return g ? g : 1;
}

// ---------------- block-level exclusive scan (256 threads) ----------------
NB_DEV u32 warp_incl_scan(u32 v) {
	u32 lane = threadIdx.x & 31;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		u32 t = __shfl_up_sync(0xffffffffu, v, d);
		if (lane >= (u32)d) v += t;
	}
	return v;
}

// returns exclusive prefix of v within the block; *total = block sum.  smem: NB_WARPS+1 words.  Ends with a barrier.
NB_DEV u32 block_excl_scan(u32 v, u32* total, u32* smem) {
	u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	u32 incl = warp_incl_scan(v);
	if (lane == 31) smem[wid] = incl;
	__syncthreads();
	if (wid == 0) {
		u32 w = lane < NB_WARPS ? smem[lane] : 0;
		u32 wi = warp_incl_scan(w);
		if (lane < NB_WARPS) smem[lane] = wi - w;
		if (lane == NB_WARPS - 1) smem[NB_WARPS] = wi;
	}
	__syncthreads();
	u32 r = incl - v + smem[wid];
	*total = smem[NB_WARPS];
	__syncthreads();
	return r;
}

// ---------------- device-wide exclusive scan over N interleaved counters ----------------
// in/out: N arrays laid out as arr[c*stride + i]; count read from *n_ptr (or n_host if n_ptr is null).
// totals[c] receives the sum of counter c.  Three kernels: tile sums, spine, downsweep.
#define NB_SCAN_GRID 592
#define NB_SCAN_ITEMS 4

NB_DEV void scan_tile_range(u32 n, u32& begin, u32& end) {
	u32 chunk = NB_BLOCK * NB_SCAN_ITEMS;
	u32 tiles = (n + chunk - 1) / chunk;
	u32 per = (tiles + gridDim.x - 1) / gridDim.x;
	begin = min(n, blockIdx.x * per * chunk);
	end = min(n, begin + per * chunk);
}

template<int N>
__global__ void __launch_bounds__(NB_BLOCK) k_scan_reduce(const u32* in, u32 stride, const u32* n_ptr, u32 n_host, u32* block_sums) {
	__shared__ u32 sm[NB_WARPS + 1];
	u32 n = n_ptr ? *n_ptr : n_host;
	u32 begin, end; scan_tile_range(n, begin, end);
	u32 acc[N];
	#pragma unroll
	for (int c = 0; c < N; ++c) acc[c] = 0;
	for (u32 i = begin + threadIdx.x; i < end; i += NB_BLOCK)
		#pragma unroll
		for (int c = 0; c < N; ++c) acc[c] += in[c*stride + i];
	#pragma unroll
	for (int c = 0; c < N; ++c) {
		u32 total; block_excl_scan(acc[c], &total, sm);
		if (threadIdx.x == 0) block_sums[c*NB_SCAN_GRID + blockIdx.x] = total;
	}
}

template<int N>
__global__ void __launch_bounds__(1024) k_scan_spine(u32* block_sums, u32* totals) {
	// one block; NB_SCAN_GRID <= 1024 entries per counter
	__shared__ u32 sm[33];
	for (int c = 0; c < N; ++c) {
		u32 v = threadIdx.x < NB_SCAN_GRID ? block_sums[c*NB_SCAN_GRID + threadIdx.x] : 0;
		u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
		u32 incl = warp_incl_scan(v);
		if (lane == 31) sm[wid] = incl;
		__syncthreads();
		if (wid == 0) {
			u32 w = sm[lane];
			u32 wi = warp_incl_scan(w);
			sm[lane] = wi - w;
			if (lane == 31) sm[32] = wi;
		}
		__syncthreads();
		if (threadIdx.x < NB_SCAN_GRID) block_sums[c*NB_SCAN_GRID + threadIdx.x] = incl - v + sm[wid];
		if (threadIdx.x == 0 && totals) totals[c] = sm[32];
		__syncthreads();
	}
}

template<int N>
__global__ void __launch_bounds__(NB_BLOCK) k_scan_down(const u32* in, u32* out, u32 stride, const u32* n_ptr, u32 n_host, const u32* block_sums) {
	__shared__ u32 sm[NB_WARPS + 1];
	u32 n = n_ptr ? *n_ptr : n_host;
	u32 begin, end; scan_tile_range(n, begin, end);
	u32 run[N];
	#pragma unroll
	for (int c = 0; c < N; ++c) run[c] = block_sums[c*NB_SCAN_GRID + blockIdx.x];
	for (u32 base = begin; base < end; base += NB_BLOCK * NB_SCAN_ITEMS) {
		u32 i0 = base + threadIdx.x * NB_SCAN_ITEMS;
		#pragma unroll
		for (int c = 0; c < N; ++c) {
			u32 v[NB_SCAN_ITEMS]; u32 s = 0;
			#pragma unroll
			for (int k = 0; k < NB_SCAN_ITEMS; ++k) { v[k] = (i0 + k < end) ? in[c*stride + i0 + k] : 0; s += v[k]; }
			u32 total; u32 ex = block_excl_scan(s, &total, sm) + run[c];
			#pragma unroll
			for (int k = 0; k < NB_SCAN_ITEMS; ++k) { if (i0 + k < end) out[c*stride + i0 + k] = ex; ex += v[k]; }
			run[c] += total;
		}
	}
}

template<int N>
static void nb_scan(const Launch& L, const u32* in, u32* out, u32 stride, const u32* n_ptr, u32 n_host, u32* block_sums /*N*NB_SCAN_GRID*/, u32* totals) {
	k_scan_reduce<N><<<NB_SCAN_GRID, NB_BLOCK, 0, L.stream>>>(in, stride, n_ptr, n_host, block_sums);
	k_scan_spine<N><<<1, 1024, 0, L.stream>>>(block_sums, totals);
	k_scan_down<N><<<NB_SCAN_GRID, NB_BLOCK, 0, L.stream>>>(in, out, stride, n_ptr, n_host, block_sums);
	*L.counter += 3;
}

// ---------------- stable LSD radix sort, 8 bits per pass, u64 keys + u32 payload ----------------
#define NB_SORT_GRID 592

NB_DEV void sort_tile_range(u32 n, u32& begin, u32& end) {
	u32 chunks = (n + NB_BLOCK - 1) / NB_BLOCK;
	u32 per = (chunks + gridDim.x - 1) / gridDim.x;
	begin = min(n, blockIdx.x * per * NB_BLOCK);
	end = min(n, begin + per * NB_BLOCK);
}

__global__ void __launch_bounds__(NB_BLOCK) k_sort_hist(const u64* keys, const u32* n_ptr, u32 shift, u32* hist /*[256][NB_SORT_GRID]*/) {
	__shared__ u32 h[256];
	h[threadIdx.x] = 0;
	__syncthreads();
	u32 n = *n_ptr;
	u32 begin, end; sort_tile_range(n, begin, end);
	for (u32 i = begin + threadIdx.x; i < end; i += NB_BLOCK)
		atomicAdd(&h[(u32)(keys[i] >> shift) & 0xff], 1u);
	__syncthreads();
	hist[threadIdx.x * NB_SORT_GRID + blockIdx.x] = h[threadIdx.x];
}

template<bool HAS_VALS>
__global__ void __launch_bounds__(NB_BLOCK) k_sort_scatter(const u64* keys_in, u64* keys_out, const u32* vals_in, u32* vals_out,
															const u32* n_ptr, u32 shift, const u32* hist_scanned, const u32* digit_base) {
	__shared__ u32 running[256];
	__shared__ u32 chunk_base[256];
	__shared__ u32 wc[NB_WARPS][256];
	u32 n = *n_ptr;
	u32 begin, end; sort_tile_range(n, begin, end);
	running[threadIdx.x] = digit_base[threadIdx.x] + hist_scanned[threadIdx.x * NB_SORT_GRID + blockIdx.x];
	u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	for (u32 base = begin; base < end; base += NB_BLOCK) {
		u32 i = base + threadIdx.x;
		bool valid = i < end;
		u64 key = valid ? keys_in[i] : 0;
		u32 val = (HAS_VALS && valid) ? vals_in[i] : 0;
		u32 d = valid ? ((u32)(key >> shift) & 0xff) : 0xffffffffu;
		u32 peers = __match_any_sync(0xffffffffu, d);
		u32 rank = __popc(peers & ((1u << lane) - 1u));
		#pragma unroll
		for (int w = 0; w < NB_WARPS; ++w) wc[w][threadIdx.x] = 0;
		__syncthreads();
		if (valid && rank == 0) wc[wid][d] = __popc(peers);
		__syncthreads();
		{
			u32 sum = 0;
			#pragma unroll
			for (int w = 0; w < NB_WARPS; ++w) { u32 c = wc[w][threadIdx.x]; wc[w][threadIdx.x] = sum; sum += c; }
			u32 b = running[threadIdx.x];
			chunk_base[threadIdx.x] = b;
			running[threadIdx.x] = b + sum;
		}
		__syncthreads();
		if (valid) {
			u32 pos = chunk_base[d] + wc[wid][d] + rank;
			keys_out[pos] = key;
			if (HAS_VALS) vals_out[pos] = val;
		}
		__syncthreads();
	}
}

// Turns the block histogram hist[digit][block] into scatter offsets in ONE launch: block d scans row d in place (exclusive over
// the sort blocks) and publishes the digit total; the last block to finish (completion counter) prefixes the 256 totals into
// digit_base[].  A key's final position is digit_base[d] + hist[d][block] + its rank inside the block.
__global__ void __launch_bounds__(1024) k_sort_offsets(u32* hist /*[256][NB_SORT_GRID]*/, u32* digit_base /*[256] + [256] totals + [1] counter*/) {
	__shared__ u32 sm[33];
	__shared__ bool last;
	const u32 d = blockIdx.x;
	u32* totals = digit_base + 256;
	u32* counter = digit_base + 512;
	u32 v = threadIdx.x < NB_SORT_GRID ? hist[d * NB_SORT_GRID + threadIdx.x] : 0;
	u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	u32 incl = warp_incl_scan(v);
	if (lane == 31) sm[wid] = incl;
	__syncthreads();
	if (wid == 0) { u32 w = sm[lane]; u32 wi = warp_incl_scan(w); sm[lane] = wi - w; if (lane == 31) sm[32] = wi; }
	__syncthreads();
	if (threadIdx.x < NB_SORT_GRID) hist[d * NB_SORT_GRID + threadIdx.x] = incl - v + sm[wid];
	if (threadIdx.x == 0) {
		totals[d] = sm[32];
		__threadfence();
		last = atomicAdd(counter, 1u) == gridDim.x - 1;
	}
	__syncthreads();
	if (last) {
		__threadfence();
		u32 t = threadIdx.x < 256 ? ((volatile u32*)totals)[threadIdx.x] : 0;
		u32 inc2 = warp_incl_scan(t);
		__syncthreads();
		if (lane == 31) sm[wid] = inc2;
		__syncthreads();
		if (wid == 0) { u32 w = sm[lane]; u32 wi = warp_incl_scan(w); sm[lane] = wi - w; }
		__syncthreads();
		if (threadIdx.x < 256) digit_base[threadIdx.x] = inc2 - t + sm[wid];
		if (threadIdx.x == 0) *counter = 0;  // ready for the next pass
	}
}

// ---------------- grid-wide barrier for cooperative (co-resident) launches ----------------
NB_DEV u32 ld_acquire_u32(const u32* p) {
	u32 v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

NB_DEV void grid_barrier(u32* bar /* [0]=arrivals, [1]=generation */, u32 nblocks) {
	__syncthreads();
	if (threadIdx.x == 0) {
		u32 gen = ld_acquire_u32(bar + 1);
		__threadfence();
		if (atomicAdd(bar, 1u) == nblocks - 1) {
			bar[0] = 0;
			__threadfence();
			atomicAdd(bar + 1, 1u);
		}
		else {
			while (ld_acquire_u32(bar + 1) == gen) { }
		}
		__threadfence();
	}
	__syncthreads();
}

// ---------------- single-launch radix sort (cooperative) ----------------
// All passes of one sort in ONE cooperative launch: one block of 1024 threads per SM, two grid barriers per pass.  A block owns a
// contiguous range of 4096-key tiles.  Per pass: block digit histogram -> [block][digit] matrix (L2 resident) | barrier | every
// block derives its own scatter offsets from the matrix, ranks its tiles (match_any per warp, 4 keys per thread) and scatters |
// barrier.  When every block owns at most one tile (n <= 4096 * blocks) the keys stay in registers between the histogram and
// the scatter.  Small inputs (n <= NB_CS_SMALL) are sorted by block 0 alone with block-level barriers; n == 0 costs one empty
// launch.  Same stable LSD order as nb_radix_sort.
#define NB_CS_THREADS 1024
#define NB_CS_WARPS 32
#define NB_CS_ITEMS 4
#define NB_CS_TILE (NB_CS_THREADS * NB_CS_ITEMS)
#define NB_CS_SMALL (4 * NB_CS_TILE)
struct SortPasses { int n; int shift[12]; };

template<bool HAS_VALS>
__global__ void __launch_bounds__(NB_CS_THREADS) k_sort_coop(u64* k0, u64* k1, u32* v0, u32* v1, const u32* n_ptr, u32* hist /*[gridDim][256]*/, u32* bar, SortPasses P) {
	__shared__ u32 h[256];        // digit totals of this block's range, then of the whole input
	__shared__ u32 tot[256];      // digit totals of the current tile
	__shared__ u32 running[256];  // next free output slot per digit for this block
	__shared__ u32 below[4][256], total[4][256];
	__shared__ u32 wc[NB_CS_WARPS][256];
	__shared__ u32 sm[8];
	const u32 n = *n_ptr;
	if (n == 0) return;
	const bool small = n <= NB_CS_SMALL;
	if (small && blockIdx.x != 0) return;
	const u32 G = small ? 1u : gridDim.x, b = blockIdx.x;
	const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, lt = (1u << lane) - 1u;
	const u32 tiles = (n + NB_CS_TILE - 1) / NB_CS_TILE, per = (tiles + G - 1) / G;
	const u32 begin = min(n, b * per * NB_CS_TILE), end = min(n, begin + per * NB_CS_TILE);
	const bool fused = per == 1;
	u64* kin = k0; u64* kout = k1; u32* vin = v0; u32* vout = v1;
	u64 key[NB_CS_ITEMS]; u32 val[NB_CS_ITEMS], dg[NB_CS_ITEMS], rk[NB_CS_ITEMS];

	for (int p = 0; p < P.n; ++p) {
		const u32 shift = (u32)P.shift[p];
		// ranks one tile: key/val/dg/rk in registers, wc[w][d] = keys of digit d in warps before w, tot[d] = tile total
		auto rank_tile = [&](u32 base) {
			for (u32 w = tid; w < NB_CS_WARPS * 256; w += NB_CS_THREADS) (&wc[0][0])[w] = 0;
			__syncthreads();
			#pragma unroll
			for (int r = 0; r < NB_CS_ITEMS; ++r) {  // all loads first: the warp barriers below would serialise them
				u32 i = base + wid * (32 * NB_CS_ITEMS) + r * 32 + lane;
				key[r] = i < end ? __ldcg(kin + i) : 0;
				if (HAS_VALS) val[r] = i < end ? __ldcg(vin + i) : 0;
			}
			#pragma unroll
			for (int r = 0; r < NB_CS_ITEMS; ++r) {
				u32 i = base + wid * (32 * NB_CS_ITEMS) + r * 32 + lane;
				bool valid = i < end;
				u32 d = valid ? ((u32)(key[r] >> shift) & 0xff) : 0xffffffffu;
				u32 peers = __match_any_sync(0xffffffffu, d);
				u32 leader = __ffs(peers) - 1, old = 0;
				if (valid && lane == leader) { old = wc[wid][d]; wc[wid][d] = old + __popc(peers); }
				old = __shfl_sync(0xffffffffu, old, leader);
				dg[r] = d; rk[r] = old + __popc(peers & lt);
				__syncwarp();
			}
			__syncthreads();
			if (tid < 256) {
				u32 sum = 0;
				#pragma unroll 8
				for (int w = 0; w < NB_CS_WARPS; ++w) { u32 c = wc[w][tid]; wc[w][tid] = sum; sum += c; }
				tot[tid] = sum;
			}
			__syncthreads();
		};
		auto scatter_tile = [&]() {
			#pragma unroll
			for (int r = 0; r < NB_CS_ITEMS; ++r)
				if (dg[r] != 0xffffffffu) {
					u32 pos = running[dg[r]] + wc[wid][dg[r]] + rk[r];
					kout[pos] = key[r];
					if (HAS_VALS) vout[pos] = val[r];
				}
		};

		if (fused) {
			rank_tile(begin);
			if (tid < 256) h[tid] = tot[tid];
		}
		else {
			if (tid < 256) h[tid] = 0;
			__syncthreads();
			for (u32 i = begin + tid; i < end; i += NB_CS_THREADS) atomicAdd(&h[(u32)(__ldcg(kin + i) >> shift) & 0xff], 1u);
		}
		__syncthreads();
		u32 mine = 0;  // keys of digit tid in blocks before this one
		if (!small) {
			if (tid < 256) hist[b * 256 + tid] = h[tid];
			grid_barrier(bar, G);
			u32 d = tid & 255, q = tid >> 8, bl = 0, tt = 0;
			for (u32 r0 = q; r0 < G; r0 += 32) {  // eight independent loads in flight per round trip
				u32 v[8];
				#pragma unroll
				for (int u = 0; u < 8; ++u) { u32 r = r0 + 4 * u; v[u] = r < G ? __ldcg(&hist[r * 256 + d]) : 0; }
				#pragma unroll
				for (int u = 0; u < 8; ++u) { tt += v[u]; if (r0 + 4 * u < b) bl += v[u]; }
			}
			below[q][d] = bl; total[q][d] = tt;
			__syncthreads();
			if (tid < 256) { h[tid] = total[0][tid] + total[1][tid] + total[2][tid] + total[3][tid]; mine = below[0][tid] + below[1][tid] + below[2][tid] + below[3][tid]; }
		}
		{	// exclusive scan of the 256 digit totals (threads 0..255 = warps 0..7)
			u32 t = tid < 256 ? h[tid] : 0;
			u32 incl = warp_incl_scan(t);
			if (tid < 256 && lane == 31) sm[wid] = incl;
			__syncthreads();
			if (wid == 0) { u32 w = lane < 8 ? sm[lane] : 0; u32 wi = warp_incl_scan(w); if (lane < 8) sm[lane] = wi - w; }
			__syncthreads();
			if (tid < 256) running[tid] = incl - t + sm[wid] + mine;
			__syncthreads();
		}
		if (fused) scatter_tile();
		else
			for (u32 base = begin; base < end; base += NB_CS_TILE) {
				rank_tile(base);
				scatter_tile();
				__syncthreads();
				if (tid < 256) running[tid] += tot[tid];
			}
		if (!small) grid_barrier(bar, G);
		else { __threadfence(); __syncthreads(); }
		{ u64* t = kin; kin = kout; kout = t; u32* tv = vin; vin = vout; vout = tv; }
	}
}

struct SortBuffers { u64* keys[2]; u32* vals[2]; u32* hist; u32* block_sums; u32* bar; int coop_blocks; /* 0 = three launches per pass */ };

// Sorts bits [begin_bit, end_bit) of keys[cur] (+vals[cur]); returns which buffer (0/1) holds the result.
static int nb_radix_sort(const Launch& L, const SortBuffers& B, const u32* n_ptr, int begin_bit, int end_bit, bool has_vals, int cur, int begin_bit2 = 0, int end_bit2 = 0) {
	if (B.coop_blocks) {
		SortPasses P; P.n = 0;
		for (int shift = begin_bit; shift < end_bit; shift += 8) P.shift[P.n++] = shift;
		for (int shift = begin_bit2; shift < end_bit2; shift += 8) P.shift[P.n++] = shift;
		u64* k0 = B.keys[cur]; u64* k1 = B.keys[cur ^ 1]; u32* v0 = B.vals[cur]; u32* v1 = B.vals[cur ^ 1]; u32* hist = B.hist; u32* bar = B.bar;
		void* args[] = { &k0, &k1, &v0, &v1, &n_ptr, &hist, &bar, &P };
		cudaLaunchCooperativeKernel(has_vals ? (void*)k_sort_coop<true> : (void*)k_sort_coop<false>, dim3(B.coop_blocks), dim3(NB_CS_THREADS), args, 0, L.stream);
		*L.counter += 1;
		return cur ^ (P.n & 1);
	}
	if (end_bit2 > begin_bit2) { cur = nb_radix_sort(L, B, n_ptr, begin_bit, end_bit, has_vals, cur); begin_bit = begin_bit2; end_bit = end_bit2; }
	for (int shift = begin_bit; shift < end_bit; shift += 8) {
		k_sort_hist<<<NB_SORT_GRID, NB_BLOCK, 0, L.stream>>>(B.keys[cur], n_ptr, (u32)shift, B.hist);
		*L.counter += 1;
		k_sort_offsets<<<256, 1024, 0, L.stream>>>(B.hist, B.block_sums);
		if (has_vals) k_sort_scatter<true><<<NB_SORT_GRID, NB_BLOCK, 0, L.stream>>>(B.keys[cur], B.keys[cur ^ 1], B.vals[cur], B.vals[cur ^ 1], n_ptr, (u32)shift, B.hist, B.block_sums);
		else k_sort_scatter<false><<<NB_SORT_GRID, NB_BLOCK, 0, L.stream>>>(B.keys[cur], B.keys[cur ^ 1], nullptr, nullptr, n_ptr, (u32)shift, B.hist, B.block_sums);
		*L.counter += 2;
		cur ^= 1;
	}
	return cur;
}

// ---------------- warp-aggregated append ----------------
NB_DEV u32 warp_append_slot(u32* counter) {
	u32 active = __activemask();
	u32 lane = threadIdx.x & 31;
	u32 leader = __ffs(active) - 1;
	u32 base = 0;
	if (lane == leader) base = atomicAdd(counter, (u32)__popc(active));
	base = __shfl_sync(active, base, leader);
	return base + __popc(active & ((1u << lane) - 1u));
}
