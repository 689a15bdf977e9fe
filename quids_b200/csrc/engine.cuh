// engine.cuh -- the rule-dependent kernels of one rule iteration, templated on the device rule type,
// and the glue that exposes them to the rule-independent pipeline of capi.cu through rule_ops.
//
//   num_child_kernel   iteration::compute_num_child        quids.hpp:548-569
//   symbolic_kernel    generate_symbolic_iteration + the   quids.hpp:647-721
//                      insert half of compute_collisions   quids.hpp:785-809  (fused: a child's
//                      (hash, magnitude) goes straight from registers into the interference table,
//                      the 60 B/child symbolic arrays of the reference are never materialised)
//   populate_kernel    symbolic_iteration::finalize        quids.hpp:958-967
//   hash_kernel        rule->hasher over a state (parity tooling, qb_hash_objects)
//   modifier_kernel    iteration::apply_modifier           quids.hpp:973-980
#pragma once

#include "rule_api.cuh"
#include "table.cuh"

namespace qb {

struct engine_launch {
	cudaStream_t stream;
	int sm_count;
	uint64_t *launch_counter;

	iter_view it; // the parent state

	// num_child
	uint32_t *num_childs;
	unsigned int *max_child_size;

	// symbolic
	const uint64_t *child_begin; // exclusive scan of num_childs over the kept parents, n_parents + 1 entries
	const uint64_t *kept;        // object ids of the kept parents, or nullptr = all parents in order
	uint64_t n_parents;
	uint64_t n_children;
	table_view table;
	uint8_t *scratch;        // needs_scratch rules: scratch_stride bytes per resident thread
	uint32_t scratch_stride;

	// populate
	uint64_t n_survivors;
	const uint64_t *survivor_parent; // object id of the parent of each survivor
	const uint32_t *survivor_child;  // its child_id
	uint8_t *next_objects;
	const uint64_t *next_begin;
	const uint32_t *next_size;

	// hash
	uint64_t *hashes;
};

constexpr int ENGINE_THREADS = 256;
constexpr int SYMBOLIC_CHUNK = 2048;      // children handled by one CTA per loop iteration
constexpr int SYMBOLIC_GROUP = ENGINE_THREADS; // parents whose contexts are resident at once

inline int resident_grid(const void *kernel, int threads, int sm_count) {
	int per_sm = 0;
	QB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0));
	if (per_sm < 1)
		per_sm = 1;
	return per_sm * sm_count;
}

template <class Rule>
__global__ void __launch_bounds__(ENGINE_THREADS) num_child_kernel(const Rule rule, iter_view it, uint32_t *num_childs, unsigned int *max_child_size) {
	__shared__ unsigned int s_max;
	if (threadIdx.x == 0)
		s_max = 0;
	__syncthreads();
	unsigned int local_max = 0;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < it.n; i += stride) {
		uint32_t count, bound;
		rule.get_num_child(it.objects + it.begin[i], it.size[i], count, bound);
		num_childs[i] = count;
		local_max = max(local_max, bound);
	}
	atomicMax(&s_max, local_max);
	__syncthreads();
	if (threadIdx.x == 0 && s_max)
		atomicMax(max_child_size, s_max);
}

// One CTA takes SYMBOLIC_CHUNK consecutive children at a time (balanced whatever the fan-out), finds
// the parents they belong to, prepares those parents' contexts in shared memory (one thread per
// parent), then every thread produces children: symbolic() -> table_insert().
template <class Rule>
__global__ void __launch_bounds__(ENGINE_THREADS) symbolic_kernel(const Rule rule, const engine_launch L) {
	typedef typename Rule::ctx_t ctx_t;
	__shared__ ctx_t s_ctx[SYMBOLIC_GROUP];
	__shared__ uint64_t s_child_begin[SYMBOLIC_GROUP + 1];
	__shared__ uint64_t s_object[SYMBOLIC_GROUP]; // byte offset of the parent
	__shared__ uint32_t s_size[SYMBOLIC_GROUP];
	__shared__ cplx s_mag[SYMBOLIC_GROUP];
	__shared__ uint64_t s_parent_range[2];

	uint8_t *scratch = Rule::needs_scratch ? L.scratch + ((size_t)blockIdx.x * ENGINE_THREADS + threadIdx.x) * L.scratch_stride : nullptr;

	const uint64_t num_chunks = div_up<uint64_t>(L.n_children, SYMBOLIC_CHUNK);
	for (uint64_t chunk = blockIdx.x; chunk < num_chunks; chunk += gridDim.x) {
		const uint64_t c0 = chunk * SYMBOLIC_CHUNK;
		const uint64_t c1 = min(c0 + (uint64_t)SYMBOLIC_CHUNK, L.n_children);
		if (threadIdx.x < 2) // parent of child c = last p with child_begin[p] <= c
			s_parent_range[threadIdx.x] = upper_bound_u64(L.child_begin, L.n_parents + 1, threadIdx.x == 0 ? c0 : c1 - 1) - 1;
		__syncthreads();
		const uint64_t p_lo = s_parent_range[0], p_hi = s_parent_range[1];

		for (uint64_t g0 = p_lo; g0 <= p_hi; g0 += SYMBOLIC_GROUP) {
			const uint32_t count = (uint32_t)min((uint64_t)SYMBOLIC_GROUP, p_hi + 1 - g0);
			if (threadIdx.x < count) {
				const uint64_t p = g0 + threadIdx.x;
				const uint64_t oid = L.kept ? L.kept[p] : p;
				const uint64_t off = L.it.begin[oid];
				const uint32_t sz = L.it.size[oid];
				s_child_begin[threadIdx.x] = L.child_begin[p];
				s_object[threadIdx.x] = off;
				s_size[threadIdx.x] = sz;
				s_mag[threadIdx.x] = L.it.mag[oid];
				if (L.child_begin[p + 1] > L.child_begin[p]) // childless parents need no context
					rule.prepare(L.it.objects + off, sz, s_ctx[threadIdx.x]);
			}
			if (threadIdx.x == 0)
				s_child_begin[count] = L.child_begin[g0 + count];
			__syncthreads();

			const uint64_t lo = max(c0, s_child_begin[0]), hi = min(c1, s_child_begin[count]);
			for (uint64_t c = lo + threadIdx.x; c < hi; c += ENGINE_THREADS) {
				const uint32_t j = (uint32_t)upper_bound_u64(s_child_begin, count + 1, c) - 1;
				const uint32_t child_id = (uint32_t)(c - s_child_begin[j]);
				cplx mag = s_mag[j];
				uint32_t size;
				const uint64_t hash = rule.symbolic(L.it.objects + s_object[j], s_size[j], s_ctx[j], child_id, scratch, size, mag);
				table_insert(L.table, hash, mag, rep_pack(c, size));
			}
			__syncthreads();
		}
	}
}

// v1 finalisation: one thread rebuilds one surviving child in place (populate_child_simple) and
// zeroes its alignment padding
template <class Rule>
__global__ void __launch_bounds__(ENGINE_THREADS) populate_kernel(const Rule rule, const engine_launch L) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < L.n_survivors; s += stride) {
		const uint64_t oid = L.survivor_parent[s];
		uint8_t *child = L.next_objects + L.next_begin[s];
		rule.populate_child_simple(L.it.objects + L.it.begin[oid], L.it.size[oid], child, L.survivor_child[s]);
		for (uint64_t b = L.next_begin[s] + L.next_size[s]; b < L.next_begin[s + 1]; ++b)
			L.next_objects[b] = 0;
	}
}

template <class Rule>
__global__ void __launch_bounds__(ENGINE_THREADS) hash_kernel(const Rule rule, iter_view it, uint64_t *hashes) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < it.n; i += stride)
		hashes[i] = rule.hasher(it.objects + it.begin[i], it.size[i]);
}

template <class Modifier>
__global__ void __launch_bounds__(ENGINE_THREADS) modifier_kernel(const Modifier modifier, iter_view it) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < it.n; i += stride) {
		cplx mag = it.mag[i];
		const cplx before = mag;
		modifier(it.objects + it.begin[i], it.size[i], mag);
		if (mag.re != before.re || mag.im != before.im)
			it.mag[i] = mag;
	}
}

// ---- glue -------------------------------------------------------------------------------------------
inline int grid_for(uint64_t n, int threads, int cap) {
	uint64_t g = div_up<uint64_t>(n, threads);
	if (g < 1) g = 1;
	return (int)(g < (uint64_t)cap ? g : (uint64_t)cap);
}

template <class Rule>
struct rule_glue {
	static_assert(sizeof(Rule) <= RULE_STORAGE_BYTES, "device rules are passed by value");
	static_assert(std::is_trivially_copyable<Rule>::value, "device rules are passed by value");

	static void num_child(const void *rule, const engine_launch &L) {
		int grid = grid_for(L.it.n, ENGINE_THREADS, resident_grid((const void *)num_child_kernel<Rule>, ENGINE_THREADS, L.sm_count));
		num_child_kernel<Rule><<<grid, ENGINE_THREADS, 0, L.stream>>>(*static_cast<const Rule *>(rule), L.it, L.num_childs, L.max_child_size);
		++*L.launch_counter;
	}
	static int symbolic_grid(int sm_count) { return resident_grid((const void *)symbolic_kernel<Rule>, ENGINE_THREADS, sm_count); }
	static void symbolic(const void *rule, const engine_launch &L) {
		int grid = grid_for(div_up<uint64_t>(L.n_children, SYMBOLIC_CHUNK) * ENGINE_THREADS, ENGINE_THREADS, symbolic_grid(L.sm_count));
		symbolic_kernel<Rule><<<grid, ENGINE_THREADS, 0, L.stream>>>(*static_cast<const Rule *>(rule), L);
		++*L.launch_counter;
	}
	static void populate(const void *rule, const engine_launch &L) {
		int grid = grid_for(L.n_survivors, ENGINE_THREADS, resident_grid((const void *)populate_kernel<Rule>, ENGINE_THREADS, L.sm_count));
		populate_kernel<Rule><<<grid, ENGINE_THREADS, 0, L.stream>>>(*static_cast<const Rule *>(rule), L);
		++*L.launch_counter;
	}
	static void hash(const void *rule, const engine_launch &L) {
		int grid = grid_for(L.it.n, ENGINE_THREADS, resident_grid((const void *)hash_kernel<Rule>, ENGINE_THREADS, L.sm_count));
		hash_kernel<Rule><<<grid, ENGINE_THREADS, 0, L.stream>>>(*static_cast<const Rule *>(rule), L.it, L.hashes);
		++*L.launch_counter;
	}
	static rule_ops ops(const char *name, int (*make)(const double *, uint32_t, void *)) {
		rule_ops o;
		o.name = name;
		o.make = make;
		o.launch_num_child = num_child;
		o.launch_symbolic = symbolic;
		o.launch_populate = populate;
		o.launch_hash = hash;
		o.needs_scratch = Rule::needs_scratch;
		o.symbolic_grid = symbolic_grid;
		return o;
	}
};

template <class Modifier>
struct modifier_glue {
	static_assert(sizeof(Modifier) <= RULE_STORAGE_BYTES, "device modifiers are passed by value");
	static void launch(const void *modifier, const iter_view &it, cudaStream_t stream, int sm_count) {
		int grid = grid_for(it.n, ENGINE_THREADS, resident_grid((const void *)modifier_kernel<Modifier>, ENGINE_THREADS, sm_count));
		modifier_kernel<Modifier><<<grid, ENGINE_THREADS, 0, stream>>>(*static_cast<const Modifier *>(modifier), it);
	}
};

#define QB_REGISTER_RULE(NAME, TYPE, MAKE) static const int qb_rule_registered_##NAME = ::qb::register_rule(::qb::rule_glue<TYPE>::ops(#NAME, MAKE))
#define QB_REGISTER_MODIFIER(NAME, TYPE, MAKE) \
	static const int qb_modifier_registered_##NAME = ::qb::register_modifier(::qb::modifier_ops{#NAME, MAKE, ::qb::modifier_glue<TYPE>::launch})

} // namespace qb
