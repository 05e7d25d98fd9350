// Microbenchmark for DESIGN.md §2.8: what does ONE hand-off of the solver's dataflow cost at best?
// Two threads on different SMs bounce a self-validating 128-bit word (payload + token) through L2 exactly the way k_solve does
// (st.relaxed.gpu.b128 / polling ld.relaxed.gpu.b128); a third variant bounces it between two warps of one block through shared
// memory (what block-local chains would pay).  Prints nanoseconds per one-way hand-off.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o handoff_latency scripts/handoff_latency.cu && ./handoff_latency
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ float4 ld128(const float4* p) {
	float4 v;
	asm volatile("{\n .reg .b128 q;\n ld.relaxed.gpu.global.b128 q, [%4];\n mov.b128 {%0,%1,%2,%3}, q;\n}" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st128(float4* p, float4 v) {
	asm volatile("{\n .reg .b128 q;\n mov.b128 q, {%1,%2,%3,%4};\n st.relaxed.gpu.global.b128 [%0], q;\n}" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// block 0 and block `partner` (placed on another SM by giving every block a full SM's worth of shared memory)
__global__ void pingpong_l2(float4* word, int rounds, int partner, long long* cycles) {
	extern __shared__ char pad[];
	if (threadIdx.x != 0 || (blockIdx.x != 0 && blockIdx.x != partner)) return;
	const int me = blockIdx.x == 0 ? 0 : 1;
	long long t0 = clock64();
	for (int r = 0; r < rounds; ++r) {
		unsigned expect = 2u * r + me;          // tokens 0,1,2,...: even ones are mine to answer if me == 0
		float4 v;
		do { v = ld128(word); } while (__float_as_uint(v.w) != expect);
		v.x += 1.0f; v.w = __uint_as_float(expect + 1);
		st128(word, v);
	}
	if (me == 0) *cycles = clock64() - t0;
}

__global__ void pingpong_shared(int rounds, long long* cycles) {
	__shared__ float4 word;
	if (threadIdx.x == 0) word = make_float4(0, 0, 0, __uint_as_float(0u));
	__syncthreads();
	const int lane = threadIdx.x & 31, me = threadIdx.x >> 5;
	if (lane != 0 || me > 1) return;
	volatile float4* w = &word;
	long long t0 = clock64();
	for (int r = 0; r < rounds; ++r) {
		unsigned expect = 2u * r + me;
		float x, tok;
		do { x = w->x; tok = w->w; } while (__float_as_uint(tok) != expect);
		w->x = x + 1.0f; __threadfence_block(); w->w = __uint_as_float(expect + 1);
	}
	if (me == 0) *cycles = clock64() - t0;
}

// two CTAs of one thread-block cluster bounce the word through distributed shared memory: CTA 1 polls and writes CTA 0's word
__global__ void __cluster_dims__(2, 1, 1) pingpong_dsmem(int rounds, long long* cycles) {
	__shared__ float4 word;
	cg::cluster_group cluster = cg::this_cluster();
	if (threadIdx.x == 0) word = make_float4(0, 0, 0, __uint_as_float(0u));
	cluster.sync();
	const int me = cluster.block_rank();
	volatile float4* w = (volatile float4*)cluster.map_shared_rank(&word, 0);   // both CTAs use CTA 0's copy
	if (threadIdx.x == 0) {
		long long t0 = clock64();
		for (int r = 0; r < rounds; ++r) {
			unsigned expect = 2u * r + me;
			float x, tok;
			do { x = w->x; tok = w->w; } while (__float_as_uint(tok) != expect);
			w->x = x + 1.0f; __threadfence_block(); w->w = __uint_as_float(expect + 1);
		}
		if (me == 0) *cycles = clock64() - t0;
	}
	cluster.sync();
}

int main() {
	cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
	float4* word; long long* cyc; cudaMalloc(&word, 16); cudaMalloc(&cyc, 8);
	const int rounds = 20000;
	int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	const size_t smem = 200 * 1024;  // one block per SM
	cudaFuncSetAttribute(pingpong_l2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	const int partners[] = { 1, 2, prop.multiProcessorCount / 2, prop.multiProcessorCount - 1 };
	for (int p : partners) {
		cudaMemset(word, 0, 16);
		pingpong_l2<<<prop.multiProcessorCount, 32, smem>>>(word, rounds, p, cyc);
		long long c = 0; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
		printf("L2 ping-pong block 0 <-> block %3d: %.0f cycles = %.0f ns per one-way hand-off (%s)\n", p, c / (2.0 * rounds), c / (2.0 * rounds) * 1e6 / khz, cudaGetErrorString(cudaGetLastError()));
	}
	pingpong_shared<<<1, 64>>>(rounds, cyc);
	long long c = 0; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
	printf("shared-memory ping-pong, two warps of one block: %.0f cycles = %.0f ns per hand-off\n", c / (2.0 * rounds), c / (2.0 * rounds) * 1e6 / khz);
	pingpong_dsmem<<<2, 32>>>(rounds, cyc);
	c = 0; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
	printf("DSMEM ping-pong, two CTAs of one cluster (word in CTA 0): %.0f cycles = %.0f ns per hand-off (%s)\n", c / (2.0 * rounds), c / (2.0 * rounds) * 1e6 / khz, cudaGetErrorString(cudaGetLastError()));
	return 0;
}
