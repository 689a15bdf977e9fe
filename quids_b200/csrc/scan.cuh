// scan.cuh -- single-pass prefix sums with decoupled look-back (one kernel, one read of the input).
//
// Replaces the three scans of the reference: __gnu_parallel::partial_sum over num_childs
// (quids.hpp:568), the serial exclusive scan over the kept parents (quids.hpp:666-671) and the scan
// of padded child sizes (quids.hpp:944-946); the same machinery ranks the elements kept by the
// stream compactions (quids.hpp:819-823 partition by tolerance, survivor selection).
//
// Every tile publishes one 64-bit word: 2 status bits + 62 value bits, so that a status and its
// value are read atomically with one volatile load and no fence is needed.
#pragma once

#include <quids/device/common.cuh>

namespace qb {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;

constexpr uint64_t TILE_INVALID = 0;
constexpr uint64_t TILE_AGGREGATE = 1ull << 62;
constexpr uint64_t TILE_PREFIX = 2ull << 62;
constexpr uint64_t TILE_VALUE_MASK = (1ull << 62) - 1;

// workspace of one scan launch: status words (one per tile) followed by the ticket counter;
// must be zeroed before the launch
struct scan_state {
	uint64_t *status;
	unsigned int *ticket;
};

// tiles are numbered in the order CTAs START (atomic ticket), so a tile only ever waits for tiles
// that are already running
__device__ __forceinline__ unsigned int scan_take_ticket(scan_state st) {
	__shared__ unsigned int s_ticket;
	if (threadIdx.x == 0)
		s_ticket = atomicAdd(st.ticket, 1u);
	__syncthreads();
	return s_ticket;
}

// all threads call; returns the sum of the aggregates of tiles [0, tile)
__device__ __forceinline__ uint64_t scan_lookback(scan_state st, unsigned int tile, uint64_t aggregate) {
	__shared__ uint64_t s_exclusive;
	volatile uint64_t *status = st.status;
	if (threadIdx.x < 32) {
		const unsigned lane = threadIdx.x;
		if (tile == 0) {
			if (lane == 0) {
				status[0] = TILE_PREFIX | aggregate;
				s_exclusive = 0;
			}
		} else {
			if (lane == 0)
				status[tile] = TILE_AGGREGATE | aggregate;
			uint64_t running = 0;
			long long look = (long long)tile - 1;
			while (true) {
				long long idx = look - lane;
				uint64_t w = TILE_PREFIX; // tiles before 0: prefix 0
				if (idx >= 0) {
					w = status[idx];
					while ((w >> 62) == 0)
						w = status[idx];
				}
				unsigned have_prefix = __ballot_sync(0xffffffffu, (w >> 62) == 2);
				unsigned first = have_prefix ? (__ffs(have_prefix) - 1) : 31;
				uint64_t v = (lane <= first) ? (w & TILE_VALUE_MASK) : 0;
				running += warp_sum(v);
				if (have_prefix)
					break;
				look -= 32;
			}
			if (lane == 0) {
				status[tile] = TILE_PREFIX | ((running + aggregate) & TILE_VALUE_MASK);
				s_exclusive = running;
			}
		}
	}
	__syncthreads();
	return s_exclusive;
}

// block-wide exclusive sum of one value per thread; returns the exclusive prefix, total in `total`
__device__ __forceinline__ uint64_t block_exclusive_sum(uint64_t v, uint64_t &total) {
	__shared__ uint64_t s_warp[SCAN_WARPS];
	__shared__ uint64_t s_total;
	const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
	uint64_t inc = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		uint64_t up = __shfl_up_sync(0xffffffffu, inc, o);
		if (lane >= (unsigned)o)
			inc += up;
	}
	if (lane == 31)
		s_warp[warp] = inc;
	__syncthreads();
	if (warp == 0) {
		uint64_t w = lane < SCAN_WARPS ? s_warp[lane] : 0;
		uint64_t winc = w;
#pragma unroll
		for (int o = 1; o < SCAN_WARPS; o <<= 1) {
			uint64_t up = __shfl_up_sync(0xffffffffu, winc, o);
			if (lane >= (unsigned)o)
				winc += up;
		}
		if (lane < SCAN_WARPS)
			s_warp[lane] = winc - w;
		if (lane == SCAN_WARPS - 1)
			s_total = winc;
	}
	__syncthreads();
	total = s_total;
	uint64_t r = s_warp[warp] + inc - v;
	__syncthreads();
	return r;
}

// ---- exclusive scan: out[i] = sum_{j<i} f(j), out[n] = total ------------------------------------
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <class F>
__global__ void __launch_bounds__(SCAN_THREADS) exclusive_scan_kernel(F f, uint64_t *out, uint64_t n, scan_state st) {
	const unsigned int tile = scan_take_ticket(st);
	const uint64_t base = (uint64_t)tile * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
	uint64_t v[SCAN_ITEMS];
	uint64_t sum = 0;
#pragma unroll
	for (int j = 0; j < SCAN_ITEMS; ++j) {
		v[j] = (base + j < n) ? f(base + j) : 0;
		sum += v[j];
	}
	uint64_t tile_total;
	uint64_t within = block_exclusive_sum(sum, tile_total);
	uint64_t before = scan_lookback(st, tile, tile_total);
	uint64_t run = before + within;
#pragma unroll
	for (int j = 0; j < SCAN_ITEMS; ++j) {
		if (base + j < n)
			out[base + j] = run;
		run += v[j];
	}
	// the thread holding the last element also writes the grand total
	if (base < n && base + SCAN_ITEMS >= n)
		out[n] = run;
}

// ---- ranks for a stream compaction with STRIPED items (item j of thread t = tile_base + j*THREADS + t),
// which keeps the global loads of the big streaming kernels coalesced.  Returns, for each kept
// item, its exclusive rank inside the tile in element order; total kept in `total`.
template <int ITEMS>
__device__ __forceinline__ void block_rank_striped(const bool (&keep)[ITEMS], uint32_t (&rank)[ITEMS], uint32_t &total) {
	static_assert(ITEMS * SCAN_WARPS <= 32, "one warp scans the (item, warp) counts");
	__shared__ uint32_t s_cnt[ITEMS * SCAN_WARPS];
	__shared__ uint32_t s_total;
	const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
	const unsigned lt = (1u << lane) - 1;
#pragma unroll
	for (int j = 0; j < ITEMS; ++j) {
		unsigned b = __ballot_sync(0xffffffffu, keep[j]);
		rank[j] = __popc(b & lt);
		if (lane == 0)
			s_cnt[j * SCAN_WARPS + warp] = __popc(b);
	}
	__syncthreads();
	if (warp == 0) {
		uint32_t c = lane < ITEMS * SCAN_WARPS ? s_cnt[lane] : 0;
		uint32_t inc = c;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
			if (lane >= (unsigned)o)
				inc += up;
		}
		if (lane < ITEMS * SCAN_WARPS)
			s_cnt[lane] = inc - c;
		if (lane == 31)
			s_total = inc;
	}
	__syncthreads();
#pragma unroll
	for (int j = 0; j < ITEMS; ++j)
		rank[j] += s_cnt[j * SCAN_WARPS + warp];
	total = s_total;
	__syncthreads();
}

// ---- ranks for a stream compaction with WARP-STRIPED items: warp w of the CTA owns the 32 * ITEMS consecutive
// elements starting at tile_base + w * 32 * ITEMS, item j of lane l = that + j * 32 + l (every load of a warp is one
// coalesced row), so a tile can be as long as the registers allow (ITEMS is not bounded by the warp width).  Two
// predicates are ranked at once (`a` and `b`, exclusive of each other or not); ranks are in element order.
template <int ITEMS>
__device__ __forceinline__ uint64_t warp_striped_index(uint64_t tile_base, int j) {
	return tile_base + (uint64_t)(threadIdx.x >> 5) * (32 * ITEMS) + (uint64_t)j * 32 + (threadIdx.x & 31);
}

template <int ITEMS, bool TWO>
__device__ __forceinline__ void block_rank_warp_striped(const bool (&a)[ITEMS], const bool (&b)[ITEMS], uint32_t (&rank_a)[ITEMS], uint32_t (&rank_b)[ITEMS],
                                                        uint32_t &total_a, uint32_t &total_b) {
	__shared__ uint32_t s_a[SCAN_WARPS], s_b[SCAN_WARPS];
	__shared__ uint32_t s_total_a, s_total_b;
	const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
	const unsigned lt = (1u << lane) - 1;
	uint32_t run_a = 0, run_b = 0;
#pragma unroll
	for (int j = 0; j < ITEMS; ++j) {
		const unsigned ba = __ballot_sync(0xffffffffu, a[j]);
		rank_a[j] = run_a + __popc(ba & lt);
		run_a += __popc(ba);
		if (TWO) {
			const unsigned bb = __ballot_sync(0xffffffffu, b[j]);
			rank_b[j] = run_b + __popc(bb & lt);
			run_b += __popc(bb);
		}
	}
	if (lane == 0) {
		s_a[warp] = run_a;
		s_b[warp] = run_b;
	}
	__syncthreads();
	if (warp == 0) {
		const uint32_t ca = lane < SCAN_WARPS ? s_a[lane] : 0, cb = lane < SCAN_WARPS ? s_b[lane] : 0;
		uint32_t ia = ca, ib = cb;
#pragma unroll
		for (int o = 1; o < SCAN_WARPS; o <<= 1) {
			const uint32_t ua = __shfl_up_sync(0xffffffffu, ia, o), ub = __shfl_up_sync(0xffffffffu, ib, o);
			if (lane >= (unsigned)o) {
				ia += ua;
				ib += ub;
			}
		}
		if (lane < SCAN_WARPS) {
			s_a[lane] = ia - ca;
			s_b[lane] = ib - cb;
		}
		if (lane == SCAN_WARPS - 1) {
			s_total_a = ia;
			s_total_b = ib;
		}
	}
	__syncthreads();
	const uint32_t off_a = s_a[warp], off_b = s_b[warp];
#pragma unroll
	for (int j = 0; j < ITEMS; ++j) {
		rank_a[j] += off_a;
		if (TWO)
			rank_b[j] += off_b;
	}
	total_a = s_total_a;
	total_b = s_total_b;
	__syncthreads();
}

} // namespace qb
