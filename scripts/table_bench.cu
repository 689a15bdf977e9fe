// Micro-benchmark: what the memory system gives for random 32-byte-slot table operations.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/table_bench scripts/table_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

struct __align__(32) slot { unsigned long long key; double re, im; unsigned long long rep; };

__host__ __device__ inline uint64_t mix64(uint64_t x) {
	x ^= x >> 32; x *= 0xd6e8feb86659fd93ull; x ^= x >> 32; x *= 0xd6e8feb86659fd93ull; x ^= x >> 32; return x;
}

// mode 0: load key only; 1: load + 2 RED f64; 2: 2 RED only; 3: load + 1 RED; 4: CAS only; 5: load+2RED with ILP 4; 6: plain RMW store (no atomics)
template <int MODE, int ILP>
__global__ void __launch_bounds__(256) bench(slot *t, uint64_t cap, uint64_t n, uint64_t distinct, unsigned long long *sink, uint64_t window) {
	uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	unsigned long long acc = 0;
	for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += stride * ILP) {
		uint64_t idx[ILP];
		unsigned long long seen[ILP];
#pragma unroll
		for (int j = 0; j < ILP; ++j) {
			uint64_t i = i0 + j * stride;
			// WINDOW > 0: ops are grouped in runs of WINDOW*256 consecutive ops (what one warp of the real kernel
			// does for one family): the run revisits the same 256 keys WINDOW times
			uint64_t key = window ? (mix64((i / (window * 256)) * 0x9e3779b97f4a7c15ull) + (i % 256) * 0x632be59bd9b4e019ull) % distinct
			                      : mix64(i * 0x9e3779b97f4a7c15ull) % distinct; // which distinct object
			idx[j] = __umul64hi(mix64(key + 1), cap);
			if (MODE == 0 || MODE == 1 || MODE == 3 || MODE == 5) seen[j] = __ldcg(&t[idx[j]].key);
			if (MODE == 4) seen[j] = atomicCAS(&t[idx[j]].key, 0ull, key + 1);
			if (MODE == 6) seen[j] = __ldcg(&t[idx[j]].key);
		}
#pragma unroll
		for (int j = 0; j < ILP; ++j) {
			if (i0 + j * stride >= n) continue;
			if (MODE == 0 || MODE == 4) acc += seen[j];
			if (MODE == 1 || MODE == 5) { if (seen[j] != 12345) { atomicAdd(&t[idx[j]].re, 1.0); atomicAdd(&t[idx[j]].im, 1.0); } }
			if (MODE == 2) { atomicAdd(&t[idx[j]].re, 1.0); atomicAdd(&t[idx[j]].im, 1.0); }
			if (MODE == 3) { if (seen[j] != 12345) atomicAdd(&t[idx[j]].re, 1.0); }
			if (MODE == 6) { t[idx[j]].re = (double)seen[j] + 1.0; }
		}
	}
	if (acc == 0xdeadbeef) *sink = acc;
}

template <int MODE, int ILP>
void run(const char *name, slot *t, uint64_t cap, uint64_t n, uint64_t distinct, unsigned long long *sink, int grid, uint64_t window = 0) {
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	float best = 1e30f;
	for (int rep = 0; rep < 3; ++rep) {
		cudaMemset(t, 0, cap * sizeof(slot));
		cudaEventRecord(a);
		bench<MODE, ILP><<<grid, 256>>>(t, cap, n, distinct, sink, window);
		cudaEventRecord(b); cudaEventSynchronize(b);
		float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
	}
	printf("  %-34s %8.3f ms  %7.2f G ops/s  (%.1f ps/op)\n", name, best, n / best / 1e6, best * 1e9 / n);
}

int main(int argc, char **argv) {
	uint64_t n = 130000000ull;
	int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	unsigned long long *sink; cudaMalloc(&sink, 8);
	for (int gran : {32}) {
		cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
		size_t g; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
		for (uint64_t distinct : {1000000ull, 12400000ull, 130000000ull}) {
			uint64_t cap = distinct * 3;
			slot *t; cudaMalloc(&t, cap * sizeof(slot));
			printf("L2 fetch granularity %zu, %llu distinct keys, table %.2f GB, %llu ops\n", g, (unsigned long long)distinct, cap * 32 / 1e9, (unsigned long long)n);
			int grid = sms * 8;
			run<0, 1>("load key", t, cap, n, distinct, sink, grid);
			run<0, 4>("load key, 4 in flight", t, cap, n, distinct, sink, grid);
			run<2, 1>("2 RED f64", t, cap, n, distinct, sink, grid);
			run<2, 4>("2 RED f64, 4 in flight", t, cap, n, distinct, sink, grid);
			run<3, 4>("load + 1 RED, 4 in flight", t, cap, n, distinct, sink, grid);
			run<1, 1>("load + 2 RED", t, cap, n, distinct, sink, grid);
			run<5, 4>("load + 2 RED, 4 in flight", t, cap, n, distinct, sink, grid);
			run<4, 4>("CAS only, 4 in flight", t, cap, n, distinct, sink, grid);
			run<6, 4>("load + plain store, 4 in flight", t, cap, n, distinct, sink, grid);
			run<5, 4>("load + 2 RED, 4 in flight, window 10", t, cap, n, distinct, sink, grid, 10);
			run<5, 4>("load + 2 RED, 4 in flight, window 100", t, cap, n, distinct, sink, grid, 100);
			cudaFree(t);
		}
	}
	return 0;
}
