// standalone check of sort.cuh (stable LSD radix sort) -- built and run by hand on the GPU box
#include <algorithm>
#include <cstdio>
#include <numeric>
#include <vector>
#include "../quids_b200/csrc/sort.cuh"
using namespace qb;
struct widen { const uint32_t *v; __device__ uint64_t operator()(uint64_t j) const { return v[j]; } };
int main() {
	for (uint64_t n : {1000ull, 4096ull, 100000ull, 3000000ull}) {
		std::vector<uint32_t> keys(n); std::vector<uint64_t> vals(n);
		uint64_t x = 88172645463325252ull;
		for (uint64_t i = 0; i < n; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; keys[i] = (uint32_t)(x >> (n < 5000 ? 56 : 20)); vals[i] = i; }
		uint64_t tiles = (n + SORT_TILE - 1) / SORT_TILE;
		uint32_t *k[2], *hist; uint64_t *v[2], *base, *ws;
		cudaMalloc(&k[0], 4 * n); cudaMalloc(&k[1], 4 * n); cudaMalloc(&v[0], 8 * n); cudaMalloc(&v[1], 8 * n);
		cudaMalloc(&hist, 4 * SORT_BINS * tiles); cudaMalloc(&base, 8 * (SORT_BINS * tiles + 1));
		uint64_t stiles = (SORT_BINS * tiles + SCAN_TILE - 1) / SCAN_TILE; cudaMalloc(&ws, 8 * (stiles + 2));
		cudaMemcpy(k[0], keys.data(), 4 * n, cudaMemcpyHostToDevice); cudaMemcpy(v[0], vals.data(), 8 * n, cudaMemcpyHostToDevice);
		int src = 0;
		cudaFuncSetAttribute((const void *)radix_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SORT_STAGE_BYTES);
		for (int shift = 0; shift < 32; shift += 8, src ^= 1) {
			radix_histogram_kernel<<<tiles, SORT_THREADS>>>(k[src], n, shift, hist, tiles);
			cudaMemset(ws, 0, 8 * (stiles + 2));
			scan_state st{ws + 1, (unsigned int *)ws};
			exclusive_scan_kernel<<<stiles, SCAN_THREADS>>>(widen{hist}, base, (uint64_t)SORT_BINS * tiles, st);
			radix_scatter_kernel<<<tiles, SORT_THREADS, SORT_STAGE_BYTES>>>(k[src], v[src], n, shift, base, tiles, k[src ^ 1], v[src ^ 1]);
		}
		std::vector<uint32_t> ok(n); std::vector<uint64_t> ov(n);
		cudaMemcpy(ok.data(), k[src], 4 * n, cudaMemcpyDeviceToHost); cudaMemcpy(ov.data(), v[src], 8 * n, cudaMemcpyDeviceToHost);
		printf("n=%llu err=%s ", (unsigned long long)n, cudaGetErrorString(cudaGetLastError()));
		std::vector<uint64_t> ref(n); std::iota(ref.begin(), ref.end(), 0);
		std::stable_sort(ref.begin(), ref.end(), [&](uint64_t a, uint64_t b) { return keys[a] < keys[b]; });
		uint64_t bad = 0, unsorted = 0;
		for (uint64_t i = 0; i < n; ++i) { bad += ov[i] != ref[i]; if (i && ok[i] < ok[i - 1]) ++unsorted; }
		printf("mismatches vs stable_sort: %llu, inversions: %llu\n", (unsigned long long)bad, (unsigned long long)unsorted);
	}
}
