// pipeline.cuh -- the rule-INDEPENDENT kernels of one rule iteration: table compaction, survivor
// selection, finalisation metadata, normalisation.  (Rule-dependent kernels: engine.cuh.)
#pragma once

#include "rule_api.cuh"
#include "scan.cuh"
#include "select.cuh"
#include "table.cuh"

namespace qb {

constexpr int COMPACT_ITEMS = 4;
constexpr int COMPACT_TILE = SCAN_THREADS * COMPACT_ITEMS;

__device__ __forceinline__ uint64_t key_of_norm(double norm) { return (uint64_t)__double_as_longlong(norm); }

// ---- interference result -> dense list of unique children above the tolerance ------------------------
// One streaming pass over the table (two 16-byte loads per 32-byte slot, coalesced), stable
// compaction ranked by decoupled look-back.  Replaces the partition by `norm(mag) > tolerance` of
// quids.hpp:819-823; the strict > and norm = re*re + im*im (separately rounded) are kept.
__global__ void __launch_bounds__(SCAN_THREADS) table_compact_kernel(table_view t, double tolerance, uint64_t *ukey, uint32_t *uslot,
                                                                     unsigned long long *count, scan_state st) {
	const unsigned int tile = scan_take_ticket(st);
	const uint64_t n = t.capacity + 1;
	const uint64_t base = (uint64_t)tile * COMPACT_TILE;
	bool keep[COMPACT_ITEMS];
	uint64_t key[COMPACT_ITEMS];
#pragma unroll
	for (int j = 0; j < COMPACT_ITEMS; ++j) {
		const uint64_t i = base + (uint64_t)j * SCAN_THREADS + threadIdx.x;
		keep[j] = false;
		key[j] = 0;
		if (i < n) {
			const ulonglong2 lo = reinterpret_cast<const ulonglong2 *>(t.slots + i)[0]; // key, re
			const ulonglong2 hi = reinterpret_cast<const ulonglong2 *>(t.slots + i)[1]; // im, rep
			// a slot is occupied once it has a representative (set by whoever created it; never 0): true for hashed slots,
			// for the dedicated slot of the hash 0, and for region slots (whose object may hash to 0)
			const bool occupied = hi.y != 0;
			const double norm = cnorm(cplx{__longlong_as_double((long long)lo.y), __longlong_as_double((long long)hi.x)});
			keep[j] = occupied && norm > tolerance;
			key[j] = key_of_norm(norm);
		}
	}
	uint32_t rank[COMPACT_ITEMS], total;
	block_rank_striped<COMPACT_ITEMS>(keep, rank, total);
	const uint64_t before = scan_lookback(st, tile, total);
#pragma unroll
	for (int j = 0; j < COMPACT_ITEMS; ++j)
		if (keep[j]) {
			const uint64_t dst = before + rank[j];
			ukey[dst] = key[j];
			uslot[dst] = (uint32_t)(base + (uint64_t)j * SCAN_THREADS + threadIdx.x);
		}
	if (base + COMPACT_TILE >= n && threadIdx.x == 0)
		*count = before + total;
}

// ---- keep the elements selected by a finished radix select --------------------------------------------
// pass 0: key >  threshold                      -> out[rank]
// pass 1: key == threshold, first `k` of them   -> out[count_gt + rank]
template <class KeyFn, class OutFn>
__global__ void __launch_bounds__(SCAN_THREADS) select_compact_kernel(KeyFn key_of, uint64_t n, const select_state *sel, int pass, OutFn out, scan_state st) {
	const unsigned int tile = scan_take_ticket(st);
	const uint64_t threshold = sel->prefix, need = sel->k, count_gt = sel->count_gt;
	const uint64_t base = (uint64_t)tile * COMPACT_TILE;
	bool keep[COMPACT_ITEMS];
#pragma unroll
	for (int j = 0; j < COMPACT_ITEMS; ++j) {
		const uint64_t i = base + (uint64_t)j * SCAN_THREADS + threadIdx.x;
		keep[j] = false;
		if (i < n) {
			const uint64_t key = key_of(i);
			keep[j] = pass == 0 ? key > threshold : key == threshold;
		}
	}
	uint32_t rank[COMPACT_ITEMS], total;
	block_rank_striped<COMPACT_ITEMS>(keep, rank, total);
	const uint64_t before = scan_lookback(st, tile, total);
#pragma unroll
	for (int j = 0; j < COMPACT_ITEMS; ++j)
		if (keep[j]) {
			const uint64_t r = before + rank[j];
			const uint64_t i = base + (uint64_t)j * SCAN_THREADS + threadIdx.x;
			if (pass == 0)
				out(r, i);
			else if (r < need)
				out(count_gt + r, i);
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
		const uint64_t p = upper_bound_u64(a.child_begin, a.n_parents + 1, index) - 1;
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
		const uint64_t p = upper_bound_u64(a.child_begin, a.n_parents + 1, index) - 1;
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
