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

#include "common.cuh"
#include "scan.cuh"

namespace qb {

constexpr int SELECT_MAX_BITS = 12;
constexpr int SELECT_BINS = 1 << SELECT_MAX_BITS;

struct select_state {
	uint64_t prefix;   // digits decided so far (in place)
	uint64_t mask;     // which bits of prefix are decided
	uint64_t k;        // rank still wanted among the keys matching prefix (1-based, from the top)
	uint64_t count_gt; // keys known to be strictly above the final threshold
	unsigned long long hist[SELECT_BINS];
};

template <class KeyFn>
__global__ void __launch_bounds__(256) select_histogram_kernel(KeyFn key_of, uint64_t n, select_state *st, int shift, int bits) {
	__shared__ unsigned int s_hist[SELECT_BINS];
	const int bins = 1 << bits;
	for (int i = threadIdx.x; i < bins; i += blockDim.x)
		s_hist[i] = 0;
	__syncthreads();
	const uint64_t prefix = st->prefix, mask = st->mask;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	// block-uniform trip count so that the warp votes below always see all 32 lanes
	for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < n; base += stride) {
		const uint64_t i = base + threadIdx.x;
		uint64_t key = i < n ? key_of(i) : 0;
		bool in = i < n && (key & mask) == prefix;
		unsigned digit = (unsigned)(key >> shift) & (bins - 1);
		// probabilities come in large groups of equal values: merge equal digits inside the warp
		// before touching shared memory
		unsigned active = __ballot_sync(0xffffffffu, in);
		if (in) {
			unsigned peers = __match_any_sync(active, digit);
			if ((__ffs(peers) - 1) == (int)lane_id())
				atomicAdd(&s_hist[digit], (unsigned)__popc(peers));
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < bins; i += blockDim.x)
		if (s_hist[i])
			atomicAdd(&st->hist[i], (unsigned long long)s_hist[i]);
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
