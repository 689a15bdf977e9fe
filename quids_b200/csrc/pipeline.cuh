// pipeline.cuh -- the rule-INDEPENDENT kernels of one rule iteration: table compaction, survivor
// selection, finalisation metadata, normalisation.  (Rule-dependent kernels: engine.cuh.)
#pragma once

#include <quids/device/rule_api.cuh>
#include "scan.cuh"
#include "select.cuh"
#include <quids/device/table.cuh>

namespace qb {

constexpr int COMPACT_ITEMS = 8; // warp-striped (scan.cuh): 2048 elements per tile
constexpr int COMPACT_TILE = SCAN_THREADS * COMPACT_ITEMS;

__device__ __forceinline__ uint64_t key_of_norm(double norm) { return (uint64_t)__double_as_longlong(norm); }

// ---- interference result -> dense list of unique children above the tolerance ------------------------
// One streaming pass over the table (two 16-byte loads per 32-byte slot, coalesced).  Replaces the partition by
// `norm(mag) > tolerance` of quids.hpp:819-823; the strict > and norm = re*re + im*im (separately rounded) are kept.
// The kept entries of a tile go to out[base .. base + total) with base = ONE atomicAdd on the output cursor per tile: the
// order of the list is the order in which the tiles asked, i.e. arbitrary -- as arbitrary as the order of the table itself
// (which child created which slot) and of the reference's own list (quids.hpp:819, an unstable parallel partition).
// Round 1 ranked the tiles with a decoupled look-back instead (stable order): the look-back chain advances about 8e7
// tiles/s, and a table of 1.3e9 slots has 6.5e5 tiles of 2048 -- 8 ms of pure serialisation out of 16 ms for the kernel.
// `n` = slots to scan (a hashed table: capacity + 1 with the dedicated slot of the hash 0; regions: the slots handed out).
//
// FILTER (simple truncation to k survivors with k far below the number of slots): only the entries whose key is at
// least `*floor` are listed -- a lower bound of the k-th largest key, taken from a random sample of the slots
// (sample_norm_keys_kernel) with 6 sigma of margin -- so that the list holds little more than k entries instead of every
// unique child, and the selection after it reads megabytes instead of gigabytes.  `count_kept` still counts every child
// above the tolerance (N_u); the host checks that at least min(k, N_u) entries were listed and redoes the pass unfiltered
// otherwise (never seen: 1e-9 per call).
template <bool FILTER>
__global__ void __launch_bounds__(SCAN_THREADS) table_compact_kernel(table_view t, uint64_t n, double tolerance, uint64_t *ukey, uint32_t *uslot,
                                                                     unsigned long long *count, const uint64_t *floor, unsigned long long *count_kept) {
	__shared__ unsigned long long s_base;
	[[maybe_unused]] const uint64_t floor_key = FILTER ? *floor : 0;
	for (uint64_t base = (uint64_t)blockIdx.x * COMPACT_TILE; base < n; base += (uint64_t)gridDim.x * COMPACT_TILE) {
		bool keep[COMPACT_ITEMS], above[COMPACT_ITEMS];
		uint64_t key[COMPACT_ITEMS];
		ulonglong2 lo[COMPACT_ITEMS], hi[COMPACT_ITEMS];
#pragma unroll
		for (int j = 0; j < COMPACT_ITEMS; ++j) { // all loads first: two 16-byte halves of every slot
			const uint64_t i = warp_striped_index<COMPACT_ITEMS>(base, j);
			lo[j] = hi[j] = make_ulonglong2(0, 0);
			if (i < n) {
				lo[j] = __ldcs(reinterpret_cast<const ulonglong2 *>(t.slots + i));     // key, re
				hi[j] = __ldcs(reinterpret_cast<const ulonglong2 *>(t.slots + i) + 1); // im, rep
			}
		}
#pragma unroll
		for (int j = 0; j < COMPACT_ITEMS; ++j) {
			// a slot is occupied once it has a representative (set by whoever created it; never 0): true for hashed slots,
			// for the dedicated slot of the hash 0, and for region slots (whose object may hash to 0)
			const bool occupied = hi[j].y != 0;
			const double norm = cnorm(cplx{__longlong_as_double((long long)lo[j].y), __longlong_as_double((long long)hi[j].x)});
			keep[j] = occupied && norm > tolerance;
			key[j] = key_of_norm(norm);
			above[j] = keep[j];
			if constexpr (FILTER)
				keep[j] = keep[j] && key[j] >= floor_key;
		}
		uint32_t rank[COMPACT_ITEMS], unused_rank[COMPACT_ITEMS], total, total_above;
		block_rank_warp_striped<COMPACT_ITEMS, FILTER>(keep, above, rank, unused_rank, total, total_above);
		if (threadIdx.x == 0) {
			s_base = total ? atomicAdd(count, (unsigned long long)total) : 0;
			if (FILTER && total_above)
				atomicAdd(count_kept, (unsigned long long)total_above);
		}
		__syncthreads();
		const uint64_t before = s_base;
#pragma unroll
		for (int j = 0; j < COMPACT_ITEMS; ++j)
			if (keep[j]) {
				const uint64_t dst = before + rank[j];
				ukey[dst] = key[j];
				uslot[dst] = (uint32_t)warp_striped_index<COMPACT_ITEMS>(base, j);
			}
		__syncthreads(); // s_base is reused by the next tile
	}
}

// norm keys of `samples` slots drawn at random (with replacement) from the first n slots of the table; empty slots and
// children below the tolerance give key 0
__global__ void __launch_bounds__(256) sample_norm_keys_kernel(table_view t, uint64_t n, double tolerance, uint64_t *keys, uint32_t samples, uint64_t seed) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= samples)
		return;
	const uint64_t at = __umul64hi(mix64(seed + (uint64_t)i * 0x9e3779b97f4a7c15ull), n);
	const ulonglong2 lo = __ldcs(reinterpret_cast<const ulonglong2 *>(t.slots + at));
	const ulonglong2 hi = __ldcs(reinterpret_cast<const ulonglong2 *>(t.slots + at) + 1);
	const double norm = cnorm(cplx{__longlong_as_double((long long)lo.y), __longlong_as_double((long long)hi.x)});
	keys[i] = hi.y != 0 && norm > tolerance ? key_of_norm(norm) : 0;
}

// ---- keep the elements selected by a finished radix select --------------------------------------------
//   key >  threshold                      -> out[rank among those]
//   key == threshold, first `k` of them   -> out[count_gt + rank among those]
// The look-back of a single-pass compaction advances ~32 tiles per L2 round trip (about 8e7 tiles/s measured on B200), so
// a tile must carry >= 100 KB of input to stream at HBM speed: a CTA owns 16384 keys.  Pass A reads the keys ONCE, keeps
// two bits per key in registers and counts; one look-back per CTA; pass B ranks the kept keys row by row from the
// saved bits (no memory traffic but the output; a first version that ranked sub-tile by sub-tile with block-wide
// scans needed 128 registers and 2 instructions per key: 9 ms for 1.3e9 keys, ncu profiles/select_compact_r1i).
// MODE 2: both kinds in one launch; the look-back carries the two running counts packed in one word (31 bits each),
// so this mode is for n < 2^31.  MODE 0 / 1: one kind per launch (any n).
constexpr int SELECT_ROWS = 64; // rows of 32 keys one warp owns (one bit per row and lane in a 64-bit register)
constexpr uint64_t SELECT_TILE = (uint64_t)SCAN_WARPS * 32 * SELECT_ROWS;

// Warp w of the CTA owns the SELECT_ROWS * 32 consecutive keys starting at tile_base + w * 32 * SELECT_ROWS, row r of lane
// l = that + r * 32 + l.  Element order = (warp, row, lane), so a kept key's rank is: CTA prefix (look-back) + warp prefix
// (one block scan) + kept keys of earlier rows (running count of ballots) + earlier lanes of its row.
template <int MODE, class KeyFn, class OutFn>
__global__ void __launch_bounds__(SCAN_THREADS, 4) select_compact_kernel(KeyFn key_of, uint64_t n, const select_state *sel, OutFn out, scan_state st) {
	const unsigned int tile = scan_take_ticket(st);
	const uint64_t threshold = sel->prefix, need = sel->k, count_gt = sel->count_gt;
	const unsigned lane = lane_id();
	const uint64_t first = (uint64_t)tile * SELECT_TILE + (uint64_t)(threadIdx.x >> 5) * (32 * SELECT_ROWS) + lane;
	uint64_t gt_bits = 0, eq_bits = 0;
	constexpr int IN_FLIGHT = 8;
#pragma unroll 1
	for (int r0 = 0; r0 < SELECT_ROWS; r0 += IN_FLIGHT) {
		uint64_t key[IN_FLIGHT];
#pragma unroll
		for (int j = 0; j < IN_FLIGHT; ++j) {
			const uint64_t i = first + (uint64_t)(r0 + j) * 32;
			key[j] = i < n ? key_of(i) : 0;
		}
#pragma unroll
		for (int j = 0; j < IN_FLIGHT; ++j) {
			const uint64_t i = first + (uint64_t)(r0 + j) * 32;
			const bool gt = i < n && MODE != 1 && key[j] > threshold;
			const bool eq = i < n && MODE != 0 && key[j] == threshold;
			gt_bits |= (uint64_t)gt << (r0 + j);
			eq_bits |= (uint64_t)eq << (r0 + j);
		}
	}
	// CTA totals -> one look-back; the exclusive prefix of a warp's first thread is the warp's prefix
	uint64_t cta_total;
	const uint64_t mine = ((uint64_t)__popcll(eq_bits) << 32) | (uint64_t)__popcll(gt_bits);
	const uint64_t warp_prefix = __shfl_sync(0xffffffffu, block_exclusive_sum(mine, cta_total), 0);
	const uint32_t total_gt = (uint32_t)cta_total, total_eq = (uint32_t)(cta_total >> 32);
	const uint64_t aggregate = MODE == 2 ? ((uint64_t)total_eq << 31) | total_gt : MODE == 0 ? total_gt : total_eq;
	const uint64_t before = scan_lookback(st, tile, aggregate);
	uint64_t run_gt = (MODE == 2 ? before & 0x7fffffffull : before) + (uint32_t)warp_prefix;
	uint64_t run_eq = (MODE == 2 ? before >> 31 : before) + (warp_prefix >> 32);
	// rows with nothing kept cost two votes
	unsigned rows_left = __ballot_sync(0xffffffffu, (gt_bits | eq_bits) != 0) ? SELECT_ROWS : 0;
	for (int r = 0; r < (int)rows_left; ++r) {
		const bool gt = (gt_bits >> r) & 1, eq = (eq_bits >> r) & 1;
		const unsigned vg = __ballot_sync(0xffffffffu, gt), ve = __ballot_sync(0xffffffffu, eq);
		if (vg | ve) {
			const unsigned lt = (1u << lane) - 1;
			const uint64_t i = first + (uint64_t)r * 32;
			if (gt)
				out(run_gt + __popc(vg & lt), i);
			if (eq && run_eq + __popc(ve & lt) < need)
				out(count_gt + run_eq + __popc(ve & lt), i);
			run_gt += __popc(vg);
			run_eq += __popc(ve);
		}
	}
}

// ---- finalisation metadata (quids.hpp:933-942) ---------------------------------------------------------
// For every survivor: read its table slot, recover (parent, child_id) from the representative's
// child index, write size / padded size / magnitude of the new object, and reduce sum |mag|^2.
struct finalize_args {
	table_view table;
	const uint32_t *survivor_slot;
	uint64_t n_survivors;
	const uint64_t *child_begin;
	const uint64_t *kept;
	uint64_t n_parents;
	uint32_t uniform_fanout; // != 0: every kept parent has this many children (parent of child c = c / fan-out)
	uint32_t align;
	uint32_t *next_size;
	uint32_t *next_padded;
	cplx *next_mag;
	uint64_t *survivor_parent;
	uint32_t *survivor_child;
	double *partial_norm; // one per CTA
};

__device__ __forceinline__ uint32_t padded_size(uint32_t size, uint32_t align) { // quids.hpp:93-102
	if (align <= 1)
		return size;
	const uint32_t rem = size % align;
	return rem ? size + align - rem : size;
}

__global__ void __launch_bounds__(SCAN_THREADS) finalize_meta_kernel(finalize_args a) {
	__shared__ double s_part[SCAN_WARPS];
	double local = 0;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < a.n_survivors; s += stride) {
		const table_slot *slot = a.table.slots + a.survivor_slot[s];
		const ulonglong2 lo = reinterpret_cast<const ulonglong2 *>(slot)[0];
		const ulonglong2 hi = reinterpret_cast<const ulonglong2 *>(slot)[1];
		const cplx mag{__longlong_as_double((long long)lo.y), __longlong_as_double((long long)hi.x)};
		const uint64_t index = rep_index(hi.y);
		const uint32_t size = rep_size(hi.y);
		const uint64_t p = a.uniform_fanout ? index / a.uniform_fanout : upper_bound_u64(a.child_begin, a.n_parents + 1, index) - 1;
		a.survivor_parent[s] = a.kept ? a.kept[p] : p;
		a.survivor_child[s] = (uint32_t)(index - a.child_begin[p]);
		a.next_size[s] = size;
		a.next_padded[s] = padded_size(size, a.align);
		a.next_mag[s] = mag;
		local += cnorm(mag);
	}
	local = warp_sum(local);
	if (lane_id() == 0)
		s_part[threadIdx.x >> 5] = local;
	__syncthreads();
	if (threadIdx.x == 0) {
		double sum = 0;
		for (int w = 0; w < SCAN_WARPS; ++w)
			sum += s_part[w];
		a.partial_norm[blockIdx.x] = sum;
	}
}

// distributed path: a survivor as it comes back from the rank that owns its hash
struct __align__(8) survivor_record {
	unsigned long long rep; // (child index + 1) << 24 | size, in the symbolic order of the receiving rank
	double re, im;          // merged magnitude
};

struct finalize_record_args {
	const survivor_record *records;
	uint64_t n_survivors;
	const uint64_t *child_begin;
	const uint64_t *kept;
	uint64_t n_parents;
	uint32_t uniform_fanout; // != 0: every kept parent has this many children (parent of child c = c / fan-out)
	uint32_t align;
	uint32_t *next_size;
	uint32_t *next_padded;
	cplx *next_mag;
	uint64_t *survivor_parent;
	uint32_t *survivor_child;
	double *partial_norm;
};

__global__ void __launch_bounds__(SCAN_THREADS) finalize_meta_records_kernel(finalize_record_args a) {
	__shared__ double s_part[SCAN_WARPS];
	double local = 0;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < a.n_survivors; s += stride) {
		const survivor_record r = a.records[s];
		const cplx mag{r.re, r.im};
		const uint64_t index = rep_index(r.rep);
		const uint32_t size = rep_size(r.rep);
		const uint64_t p = a.uniform_fanout ? index / a.uniform_fanout : upper_bound_u64(a.child_begin, a.n_parents + 1, index) - 1;
		a.survivor_parent[s] = a.kept ? a.kept[p] : p;
		a.survivor_child[s] = (uint32_t)(index - a.child_begin[p]);
		a.next_size[s] = size;
		a.next_padded[s] = padded_size(size, a.align);
		a.next_mag[s] = mag;
		local += cnorm(mag);
	}
	local = warp_sum(local);
	if (lane_id() == 0)
		s_part[threadIdx.x >> 5] = local;
	__syncthreads();
	if (threadIdx.x == 0) {
		double sum = 0;
		for (int w = 0; w < SCAN_WARPS; ++w)
			sum += s_part[w];
		a.partial_norm[blockIdx.x] = sum;
	}
}

// number of keys above / equal to the threshold of a finished select (distributed tie sharing)
template <class KeyFn>
__global__ void __launch_bounds__(256) select_count_kernel(KeyFn key_of, uint64_t n, const select_state *sel, unsigned long long *gt_eq) {
	const uint64_t threshold = sel->prefix;
	unsigned long long gt = 0, eq = 0;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const uint64_t key = key_of(i);
		gt += key > threshold;
		eq += key == threshold;
	}
	gt = warp_sum((uint64_t)gt);
	eq = warp_sum((uint64_t)eq);
	if (lane_id() == 0) {
		if (gt) atomicAdd(&gt_eq[0], gt);
		if (eq) atomicAdd(&gt_eq[1], eq);
	}
}

__global__ void select_patch_kernel(select_state *sel, uint64_t need, uint64_t count_gt) {
	sel->k = need;
	sel->count_gt = count_gt;
}

// sum |mag|^2 of a state (normalize() on its own, quids.hpp:1004-1006)
__global__ void __launch_bounds__(SCAN_THREADS) norm_partial_kernel(const cplx *mag, uint64_t n, double *partial) {
	__shared__ double s_part[SCAN_WARPS];
	double local = 0;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
		local += cnorm(mag[i]);
	local = warp_sum(local);
	if (lane_id() == 0)
		s_part[threadIdx.x >> 5] = local;
	__syncthreads();
	if (threadIdx.x == 0) {
		double sum = 0;
		for (int w = 0; w < SCAN_WARPS; ++w)
			sum += s_part[w];
		partial[blockIdx.x] = sum;
	}
}

// fixed-order second stage: the total does not depend on scheduling
__global__ void __launch_bounds__(SCAN_THREADS) norm_total_kernel(const double *partial, int n, double *total) {
	__shared__ double s_part[SCAN_THREADS];
	double local = 0;
	for (int i = threadIdx.x; i < n; i += SCAN_THREADS)
		local += partial[i];
	s_part[threadIdx.x] = local;
	__syncthreads();
	for (int o = SCAN_THREADS / 2; o > 0; o >>= 1) {
		if (threadIdx.x < o)
			s_part[threadIdx.x] += s_part[threadIdx.x + o];
		__syncthreads();
	}
	if (threadIdx.x == 0)
		*total = s_part[0];
}

// mag /= sqrt(total_proba)  (quids.hpp:1008-1013: a real division of both parts)
__global__ void __launch_bounds__(SCAN_THREADS) scale_kernel(cplx *mag, uint64_t n, double factor) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		cplx m = mag[i];
		m.re = __ddiv_rn(m.re, factor);
		m.im = __ddiv_rn(m.im, factor);
		mag[i] = m;
	}
}

} // namespace qb
