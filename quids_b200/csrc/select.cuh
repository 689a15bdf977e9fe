// select.cuh -- radix-select of the k-th largest 64-bit key (top-k truncation by probability).
//
// Replaces __gnu_parallel::nth_element on norm(magnitude) of iteration::truncate (quids.hpp:613-642)
// and symbolic_iteration::truncate (quids.hpp:866-900), simple-truncation mode.  Keys are the IEEE
// bit patterns of re^2+im^2 (non-negative doubles order like their unsigned bit patterns).
// The whole selection runs on the device: a histogram pass per digit, then a one-block kernel that
// picks the digit and narrows the prefix; the host never reads an intermediate value.
//
// Result (select_state): `prefix` = key of the k-th largest element (the threshold), `count_gt` =
// number of keys strictly above it, `k` = how many keys EQUAL to the threshold must still be taken
// (ties at the threshold are arbitrary in the reference; here the first ones in storage order win).
#pragma once

#include <quids/device/common.cuh>
#include "scan.cuh"

namespace qb {

constexpr int SELECT_MAX_BITS = 12;
constexpr int SELECT_BINS = 1 << SELECT_MAX_BITS;

struct select_state {
	uint64_t prefix;   // digits decided so far (in place)
	uint64_t mask;     // which bits of prefix are decided
	uint64_t k;        // rank still wanted among the keys matching prefix (1-based, from the top)
	uint64_t count_gt; // keys known to be strictly above the final threshold
	// CANDIDATES: after the first two digits the keys still matching the prefix (typically n / 4096 of them) are copied
	// to a dense buffer by select_filter_kernel, and the remaining digits only read that buffer.  When they do not fit
	// (cand_count > cand_capacity: e.g. all keys equal) the later passes keep reading all the keys.
	unsigned long long cand_count;
	uint64_t cand_capacity;
	unsigned long long hist[SELECT_BINS];
};

constexpr int SELECT_ITEMS = 4; // keys one thread has in flight per loop iteration

__device__ __forceinline__ bool select_uses_candidates(const select_state *st, const uint64_t *cand) {
	return cand != nullptr && st->cand_count <= st->cand_capacity;
}

template <class KeyFn>
__global__ void __launch_bounds__(256) select_histogram_kernel(KeyFn key_of, uint64_t n, select_state *st, int shift, int bits, const uint64_t *cand) {
	__shared__ unsigned int s_hist[SELECT_BINS];
	const int bins = 1 << bits;
	for (int i = threadIdx.x; i < bins; i += blockDim.x)
		s_hist[i] = 0;
	__syncthreads();
	const uint64_t prefix = st->prefix, mask = st->mask;
	const bool from_cand = select_uses_candidates(st, cand);
	if (from_cand)
		n = st->cand_count;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * SELECT_ITEMS;
	// block-uniform trip count so that the warp votes below always see all 32 lanes
	for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x * SELECT_ITEMS; base < n; base += stride) {
		uint64_t key[SELECT_ITEMS];
#pragma unroll
		for (int q = 0; q < SELECT_ITEMS; ++q) { // all loads first
			const uint64_t i = base + (uint64_t)q * blockDim.x + threadIdx.x;
			key[q] = i < n ? (from_cand ? cand[i] : key_of(i)) : 0;
		}
#pragma unroll
		for (int q = 0; q < SELECT_ITEMS; ++q) {
			const uint64_t i = base + (uint64_t)q * blockDim.x + threadIdx.x;
			const bool in = i < n && (key[q] & mask) == prefix;
			const unsigned digit = (unsigned)(key[q] >> shift) & (bins - 1);
			// probabilities come in large groups of equal values: when neighbouring lanes agree on the digit, equal digits
			// are merged inside the warp before touching shared memory; otherwise plain shared-memory atomics
			const unsigned active = __ballot_sync(0xffffffffu, in);
			const unsigned other = __shfl_xor_sync(0xffffffffu, digit, 1);
			const bool clustered = __any_sync(0xffffffffu, in && other == digit);
			if (in) {
				if (clustered) {
					const unsigned peers = __match_any_sync(active, digit);
					if ((__ffs(peers) - 1) == (int)lane_id())
						atomicAdd(&s_hist[digit], (unsigned)__popc(peers));
				} else {
					atomicAdd(&s_hist[digit], 1u);
				}
			}
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < bins; i += blockDim.x)
		if (s_hist[i])
			atomicAdd(&st->hist[i], (unsigned long long)s_hist[i]);
}

// copies the keys that still match the decided digits to `cand` (any order: only the threshold is wanted from them).
// One atomic on the cursor per warp and iteration; keys beyond the capacity are dropped (the count keeps growing, which
// is how the later passes know the buffer is incomplete).
template <class KeyFn>
__global__ void __launch_bounds__(256) select_filter_kernel(KeyFn key_of, uint64_t n, select_state *st, uint64_t *cand) {
	const uint64_t prefix = st->prefix, mask = st->mask, capacity = st->cand_capacity;
	const unsigned lane = lane_id();
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * SELECT_ITEMS;
	for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x * SELECT_ITEMS; base < n; base += stride) {
		uint64_t key[SELECT_ITEMS];
#pragma unroll
		for (int q = 0; q < SELECT_ITEMS; ++q) {
			const uint64_t i = base + (uint64_t)q * blockDim.x + threadIdx.x;
			key[q] = i < n ? key_of(i) : 0;
		}
#pragma unroll
		for (int q = 0; q < SELECT_ITEMS; ++q) {
			const uint64_t i = base + (uint64_t)q * blockDim.x + threadIdx.x;
			const bool in = i < n && (key[q] & mask) == prefix;
			const unsigned votes = __ballot_sync(0xffffffffu, in);
			if (votes == 0)
				continue;
			unsigned long long at = 0;
			if (lane == (unsigned)(__ffs(votes) - 1))
				at = atomicAdd(&st->cand_count, (unsigned long long)__popc(votes));
			at = __shfl_sync(0xffffffffu, at, __ffs(votes) - 1) + __popc(votes & ((1u << lane) - 1));
			if (in && at < capacity)
				cand[at] = key[q];
		}
	}
}

// one block: a fresh state for the selection of the k-th largest key
__global__ void __launch_bounds__(SCAN_THREADS) select_init_kernel(select_state *st, uint64_t k, uint64_t cand_capacity) {
	if (threadIdx.x == 0) {
		st->prefix = st->mask = st->count_gt = 0;
		st->k = k;
		st->cand_count = 0;
		st->cand_capacity = cand_capacity;
	}
	for (int i = threadIdx.x; i < SELECT_BINS; i += blockDim.x)
		st->hist[i] = 0;
}

// one block: find the digit holding the k-th largest key, narrow the prefix, clear the histogram
__global__ void __launch_bounds__(SCAN_THREADS) select_pick_kernel(select_state *st, int shift, int bits) {
	constexpr int PER = SELECT_BINS / SCAN_THREADS;
	const int bins = 1 << bits;
	// thread t owns PER consecutive bins counted from the TOP: position p <-> bin bins-1-p
	uint64_t c[PER];
	uint64_t sum = 0;
#pragma unroll
	for (int j = 0; j < PER; ++j) {
		int p = threadIdx.x * PER + j;
		c[j] = p < bins ? st->hist[bins - 1 - p] : 0;
		sum += c[j];
	}
	uint64_t total;
	uint64_t above = block_exclusive_sum(sum, total);
	const uint64_t k = st->k;
	__syncthreads();
#pragma unroll
	for (int j = 0; j < PER; ++j) {
		int p = threadIdx.x * PER + j;
		if (p < bins && above < k && k <= above + c[j]) { // exactly one (thread, j) satisfies this when k <= total
			st->prefix |= (uint64_t)(bins - 1 - p) << shift;
			st->mask |= (uint64_t)(bins - 1) << shift;
			st->k = k - above;
			st->count_gt += above;
		}
		above += c[j];
	}
	__syncthreads();
	for (int i = threadIdx.x; i < SELECT_BINS; i += blockDim.x)
		st->hist[i] = 0;
}

} // namespace qb
