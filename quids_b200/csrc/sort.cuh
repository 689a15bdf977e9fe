// sort.cuh -- stable LSD radix sort of (32-bit key, 64-bit value) pairs, 8 bits per pass.
//
// Used to order the parents of a rule iteration by the rule's LOCALITY KEY before children are
// generated: parents with equal keys produce the same children, so after the sort all inserts for
// one object reach the interference table within a short time window and are served by L2 instead
// of DRAM (measured on B200: a random read-modify-write of a 32-byte slot costs ~60-70 ps when the
// table lives in DRAM, ~21 ps when the touched slots stay L2 resident; scripts/table_bench.cu).
// The reference has no counterpart: its per-bucket hash maps are cache resident by construction
// (quids.hpp:740-809 partitions the children by hash first, which on the GPU would cost a second
// pass over all children).
//
// Per pass: tile histograms -> one exclusive scan (decoupled look-back, scan.cuh) over the
// digit-major histogram matrix -> stable scatter.  Element order inside a tile: warp w owns the
// contiguous elements [w * 32 * ROUNDS, (w + 1) * 32 * ROUNDS), visited in rows of 32.
#pragma once

#include <quids/device/common.cuh>
#include "scan.cuh"

namespace qb {

constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ROUNDS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ROUNDS;
constexpr int SORT_BINS = 256;

__device__ __forceinline__ uint64_t sort_element(uint64_t tile, unsigned warp, unsigned round, unsigned lane) {
	return tile * SORT_TILE + (uint64_t)warp * 32 * SORT_ROUNDS + round * 32 + lane;
}

// hist[digit * tiles + tile] = number of elements of the tile with that digit
static __global__ void __launch_bounds__(SORT_THREADS) radix_histogram_kernel(const uint32_t *keys, uint64_t n, int shift, uint32_t *hist, uint64_t tiles) {
	__shared__ unsigned int s_hist[SORT_BINS];
	s_hist[threadIdx.x] = 0;
	__syncthreads();
	const unsigned warp = threadIdx.x >> 5, lane = lane_id();
	for (unsigned r = 0; r < SORT_ROUNDS; ++r) {
		const uint64_t i = sort_element(blockIdx.x, warp, r, lane);
		if (i < n)
			atomicAdd(&s_hist[(keys[i] >> shift) & (SORT_BINS - 1)], 1u);
	}
	__syncthreads();
	hist[(uint64_t)threadIdx.x * tiles + blockIdx.x] = s_hist[threadIdx.x];
}

// base[digit * tiles + tile] = exclusive scan of hist (position of the tile's first element with that digit).
// The tile is first sorted by digit in SHARED memory (rank = digit start inside the tile + the warp's base + the rank among
// the warp's elements), then written out in that order: consecutive threads hold consecutive elements of one digit, i.e.
// consecutive destinations -- runs of ~16 elements (128-byte value runs) instead of one scattered 8-byte store per lane.
struct sort_stage {
	uint64_t vals[SORT_TILE];
	uint64_t base[SORT_BINS];  // base[digit * tiles + tile] - digit start inside the tile
	uint32_t keys[SORT_TILE];
	unsigned int count[SORT_WARPS][SORT_BINS]; // per warp: elements seen so far per digit, then the warp's base inside the tile
	unsigned int start[SORT_BINS];             // first local position of every digit
};
constexpr size_t SORT_STAGE_BYTES = sizeof(sort_stage);

static __global__ void __launch_bounds__(SORT_THREADS) radix_scatter_kernel(const uint32_t *keys_in, const uint64_t *vals_in, uint64_t n, int shift,
                                                                            const uint64_t *base, uint64_t tiles, uint32_t *keys_out, uint64_t *vals_out) {
	extern __shared__ __align__(16) uint8_t s_sort_raw[];
	sort_stage &st = *reinterpret_cast<sort_stage *>(s_sort_raw);
	for (int i = threadIdx.x; i < SORT_WARPS * SORT_BINS; i += SORT_THREADS)
		(&st.count[0][0])[i] = 0;
	__syncthreads();
	const unsigned warp = threadIdx.x >> 5, lane = lane_id();
	const unsigned lt = (1u << lane) - 1;
	uint32_t key[SORT_ROUNDS];
	uint32_t offset[SORT_ROUNDS]; // rank among the warp's elements with the same digit
#pragma unroll
	for (unsigned r = 0; r < SORT_ROUNDS; ++r) {
		const uint64_t i = sort_element(blockIdx.x, warp, r, lane);
		const bool valid = i < n;
		key[r] = valid ? keys_in[i] : 0xffffffffu;
		const unsigned digit = (key[r] >> shift) & (SORT_BINS - 1);
		const unsigned active = __ballot_sync(0xffffffffu, valid);
		offset[r] = 0;
		if (valid) {
			const unsigned peers = __match_any_sync(active, digit);
			const unsigned before = st.count[warp][digit]; // rows are processed in order by the same warp
			offset[r] = before + __popc(peers & lt);
			__syncwarp(active);
			if ((peers & lt) == 0)
				st.count[warp][digit] = before + __popc(peers);
		}
		__syncwarp();
	}
	__syncthreads();
	// per digit (thread d handles digit d): exclusive prefix over the warps, then over the digits
	{
		unsigned run = 0;
		for (int w = 0; w < SORT_WARPS; ++w) {
			const unsigned c = st.count[w][threadIdx.x];
			st.count[w][threadIdx.x] = run;
			run += c;
		}
		uint64_t tile_total;
		const uint32_t begin = (uint32_t)block_exclusive_sum((uint64_t)run, tile_total);
		st.start[threadIdx.x] = begin;
		st.base[threadIdx.x] = base[(uint64_t)threadIdx.x * tiles + blockIdx.x] - begin;
	}
	__syncthreads();
#pragma unroll
	for (unsigned r = 0; r < SORT_ROUNDS; ++r) {
		const uint64_t i = sort_element(blockIdx.x, warp, r, lane);
		if (i < n) {
			const unsigned digit = (key[r] >> shift) & (SORT_BINS - 1);
			const uint32_t q = st.start[digit] + st.count[warp][digit] + offset[r];
			st.keys[q] = key[r];
			st.vals[q] = vals_in[i];
		}
	}
	__syncthreads();
	const uint64_t tile_begin = (uint64_t)blockIdx.x * SORT_TILE;
	const uint32_t in_tile = (uint32_t)min((uint64_t)SORT_TILE, n - tile_begin);
	for (uint32_t q = threadIdx.x; q < in_tile; q += SORT_THREADS) {
		const uint32_t k = st.keys[q];
		const uint64_t dst = st.base[(k >> shift) & (SORT_BINS - 1)] + q;
		keys_out[dst] = k;
		vals_out[dst] = st.vals[q];
	}
}

} // namespace qb
