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
	uint32_t *num_groups; // only for rules with warp_groups
	unsigned int *max_child_size;
	unsigned int *child_count_range; // [0] = largest child count, [1] = ~smallest

	// symbolic
	const uint64_t *child_begin; // exclusive scan of num_childs over the kept parents, n_parents + 1 entries
	const uint64_t *kept;        // object ids of the kept parents, or nullptr = all parents in order
	const uint64_t *group_begin; // same over the group counts (== child_begin for rules without groups)
	const uint64_t *chunk_parent; // parent holding the first group of every chunk of the symbolic kernel (+ one sentinel)
	uint32_t *item_keys;          // sorted order: key and (parent position << 24 | group) of every group
	uint64_t *item_vals;
	const uint64_t *items;        // item_vals after the sort
	void *parent_ctx;             // sorted order: Rule::ctx_t of every kept parent, prepared once
	uint64_t n_parents;
	uint64_t n_children;
	uint64_t n_groups;
	table_view table;
	bin_view bins;           // records != nullptr: one-child-per-lane rules send their children to the bins of table.cuh instead of the table
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
constexpr int MAX_DEVICES = 64; // per-device launch state (one context per GPU; several contexts may live in one process)
constexpr int SYMBOLIC_CHUNK = 128; // groups of children handled by one warp per loop iteration

inline int resident_grid(const void *kernel, int threads, int sm_count) {
	int per_sm = 0;
	QB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0));
	if (per_sm < 1)
		per_sm = 1;
	return per_sm * sm_count;
}

template <class Rule>
__global__ void __launch_bounds__(ENGINE_THREADS) num_child_kernel(const Rule rule, iter_view it, uint32_t *num_childs, uint32_t *num_groups,
                                                                 unsigned int *max_child_size, unsigned int *child_count_range) {
	__shared__ unsigned int s_max, s_count_max, s_count_min_inv;
	if (threadIdx.x == 0)
		s_max = s_count_max = s_count_min_inv = 0;
	__syncthreads();
	unsigned int local_max = 0, count_max = 0, count_min_inv = 0;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < it.n; i += stride) {
		uint32_t count, bound;
		const uint8_t *parent = it.objects + it.begin[i];
		const uint32_t size = it.size[i];
		rule.get_num_child(parent, size, count, bound);
		num_childs[i] = count;
		if (Rule::warp_groups || Rule::lane_groups)
			num_groups[i] = rule.get_num_group(parent, size, count);
		local_max = max(local_max, bound);
		count_max = max(count_max, count);
		count_min_inv = max(count_min_inv, ~count);
	}
	atomicMax(&s_max, local_max);
	atomicMax(&s_count_max, count_max);
	atomicMax(&s_count_min_inv, count_min_inv);
	__syncthreads();
	if (threadIdx.x == 0) {
		if (s_max)
			atomicMax(max_child_size, s_max);
		// largest and (inverted) smallest child count: equal = every parent has the same fan-out, and the parent of child c
		// is c / fan-out, no search (finalize_meta_kernel)
		atomicMax(child_count_range, s_count_max);
		atomicMax(child_count_range + 1, s_count_min_inv);
	}
}

// children -> interference table, straight from registers
struct table_emitter {
	const table_view &table;
	uint64_t first_child; // index of child 0 of this parent in the symbolic order
	uint32_t created = 0;
	uint32_t regions = 0; // regions of the table created through this emitter (region mode)
	__device__ table_emitter(const table_view &t, uint64_t first) : table(t), first_child(first) {}
	__device__ void operator()(uint32_t child_id, uint64_t hash, uint32_t size, cplx mag) {
		created += table_insert(table, hash, mag, rep_pack(first_child + child_id, size));
	}
	__device__ uint64_t rep(uint32_t child_id, uint32_t size) const { return rep_pack(first_child + child_id, size); }
	// objects whose representative is already known (a rule flushing what it accumulated)
	template <int N, class MagOf, class RepOf>
	__device__ void batch_raw(int count, const uint64_t (&hash)[N], MagOf mag_of, RepOf rep_of) {
		created += table_insert_batch<N>(table, count, hash, mag_of, rep_of);
	}
	// batch<N>(count, hash[N], size, child_id_of(i), mag_of(i)): up to N children of the same size at once
	template <int N, class ChildOf, class MagOf>
	__device__ void batch(int count, const uint64_t (&hash)[N], uint32_t size, ChildOf child_of, MagOf mag_of) {
		const uint64_t first = first_child;
		created += table_insert_batch<N>(table, count, hash, mag_of, [=](int i) { return rep_pack(first + child_of(i), size); });
	}
};

constexpr int SYMBOLIC_THREADS = 128; // per-warp shared-memory slices: 4 warps keep the CTA under the 48 KB static limit
constexpr int ENGINE_WARPS = SYMBOLIC_THREADS / 32;

// Every WARP works on its own: it takes SYMBOLIC_CHUNK consecutive groups of children at a time
// (balanced whatever the fan-out), finds the parents they belong to, prepares those parents' contexts
// in its slice of shared memory (one lane per parent), then produces the groups -- one child per lane
// (symbolic()), or, for rules with warp_groups, one group at a time with all lanes (symbolic_warp()).
// No CTA-wide barrier: an insert is a DRAM round trip of very variable length, and a barrier would
// make every warp wait for the slowest lane of the CTA.
template <class Rule>
__global__ void __launch_bounds__(SYMBOLIC_THREADS, 5) symbolic_kernel(const Rule rule, const engine_launch L) {
	typedef typename Rule::ctx_t ctx_t;
	constexpr int CHUNK = Rule::warp_groups ? 32 : SYMBOLIC_CHUNK;
	constexpr int BATCH = Rule::parents_per_batch; // parents whose contexts one warp holds at a time
	static_assert(BATCH >= 1 && BATCH <= 32, "one lane per staged parent");
	struct warp_slice {
		ctx_t ctx[BATCH];
		uint64_t group_begin[33];
		uint64_t child_begin[32];
		uint64_t object[32]; // byte offset of the parent
		cplx mag[32];
		uint32_t size[32];
		typename Rule::group_ctx_t group_ctx[32];
	};
	__shared__ warp_slice s_slices[ENGINE_WARPS];
	__shared__ typename Rule::workspace_t s_workspace[ENGINE_WARPS];
	warp_slice &s = s_slices[threadIdx.x >> 5];
	const unsigned lane = lane_id();
	// stage of the batch's parent bytes (rules with warp_prepare and prepare_stage_bytes)
	constexpr uint32_t PREPARE_STAGE = Rule::warp_prepare ? Rule::prepare_stage_bytes : 0;
	struct __align__(16) prepare_stage_t {
		uint8_t bytes[PREPARE_STAGE ? PREPARE_STAGE + 32 : 16];
		unsigned long long mbar, pad_;
	};
	__shared__ prepare_stage_t s_prepare_stage[PREPARE_STAGE ? ENGINE_WARPS : 1];
	uint32_t stage_phase = 0;
	if constexpr (PREPARE_STAGE > 0) {
		if (lane == 0) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&s_prepare_stage[threadIdx.x >> 5].mbar)));
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		__syncwarp();
	}

	uint8_t *scratch = Rule::needs_scratch ? L.scratch + ((size_t)blockIdx.x * SYMBOLIC_THREADS + threadIdx.x) * L.scratch_stride : nullptr;
	uint32_t created = 0;
	if constexpr (Rule::warp_groups) {
		rule.init_warp(s_workspace[threadIdx.x >> 5]);
		__syncwarp();
	}

	const uint64_t num_chunks = div_up<uint64_t>(L.n_groups, CHUNK);
	const uint64_t warp_stride = (uint64_t)gridDim.x * ENGINE_WARPS;
	for (uint64_t chunk = (uint64_t)blockIdx.x * ENGINE_WARPS + (threadIdx.x >> 5); chunk < num_chunks; chunk += warp_stride) {
		if (table_overflowed(L.table)) // the table was sized too small: the host redoes the step, stop feeding it
			break;
		const uint64_t c0 = chunk * CHUNK;
		const uint64_t c1 = min(c0 + (uint64_t)CHUNK, L.n_groups);
		// parents of this chunk: from the one holding group c0 to the one holding group c1 - 1
		const uint64_t p_lo = L.chunk_parent[chunk];
		uint64_t p_hi = L.chunk_parent[chunk + 1];
		if (p_hi > p_lo && L.group_begin[p_hi] == c1)
			--p_hi;

		for (uint64_t g0 = p_lo; g0 <= p_hi; g0 += BATCH) {
			const uint32_t count = (uint32_t)min((uint64_t)BATCH, p_hi + 1 - g0);
			uint64_t off = 0;
			uint32_t sz = 0;
			if (lane < count) {
				const uint64_t p = g0 + lane;
				const uint64_t oid = L.kept ? L.kept[p] : p;
				off = L.it.begin[oid];
				sz = L.it.size[oid];
				const uint64_t gb = L.group_begin[p];
				s.group_begin[lane] = gb;
				s.child_begin[lane] = L.child_begin[p];
				s.object[lane] = off;
				s.size[lane] = sz;
				s.mag[lane] = L.it.mag[oid];
				if (L.group_begin[p + 1] > gb) // childless parents need no context
					rule.prepare(L.it.objects + off, sz, s.ctx[lane]);
			}
			if (lane == 0)
				s.group_begin[count] = L.group_begin[g0 + count];
			__syncwarp();
			if constexpr (Rule::warp_prepare) { // contexts built by the whole warp, one parent after the other
				const uint8_t *staged = nullptr; // shared-memory copy of the batch's bytes, which start at offset `staged_from` of the state
				uint64_t staged_from = 0;
				if constexpr (PREPARE_STAGE > 0) {
					if (!L.kept) { // storage order: the batch is the byte range [begin of the first, end of the last]
						staged_from = __shfl_sync(0xffffffffu, off, 0);
						const uint64_t end = __shfl_sync(0xffffffffu, off + sz, count - 1);
						if (end - staged_from <= PREPARE_STAGE) {
							prepare_stage_t &st = s_prepare_stage[threadIdx.x >> 5];
							staged = stage_range_to(st.bytes, &st.mbar, L.it.objects + staged_from, (uint32_t)(end - staged_from), stage_phase);
						}
					}
				}
				auto bytes_of = [&](uint32_t j) { return staged ? staged + (s.object[j] - staged_from) : L.it.objects + s.object[j]; };
				if constexpr (Rule::warp_prepare_pairs) {
					// two parents at a time, one per half warp, when both are small enough for 16 lanes
					for (uint32_t j0 = 0; j0 < count; j0 += 2) {
						const uint32_t j = min(j0 + (lane >> 4), count - 1);
						const bool active = j0 + (lane >> 4) < count && s.group_begin[j + 1] > s.group_begin[j];
						const uint8_t *bytes = bytes_of(j);
						if (__all_sync(0xffffffffu, !active || Rule::fits_half_warp(bytes))) {
							rule.prepare_half_warp(bytes, s.ctx[j], active);
						} else {
							for (uint32_t k = j0; k < min(j0 + 2, count); ++k)
								if (s.group_begin[k + 1] > s.group_begin[k])
									rule.prepare_warp(bytes_of(k), s.size[k], s.ctx[k]);
						}
					}
				} else {
					for (uint32_t j = 0; j < count; ++j)
						if (s.group_begin[j + 1] > s.group_begin[j])
							rule.prepare_warp(bytes_of(j), s.size[j], s.ctx[j]);
				}
				__syncwarp();
			}

			const uint64_t lo = max(c0, s.group_begin[0]), hi = min(c1, s.group_begin[count]);
			if constexpr (Rule::warp_groups) {
				// at most 32 groups here: one lane per group prepares what the whole group shares ...
				if (lo + lane < hi) {
					const uint32_t j = (uint32_t)upper_bound_u64(s.group_begin, count + 1, lo + lane) - 1;
					rule.prepare_group(s.ctx[j], (uint32_t)(lo + lane - s.group_begin[j]), s.mag[j], s.group_ctx[lane]);
				}
				__syncwarp();
				// ... then all lanes together produce one group after the other
				uint32_t j = 0;
				for (uint64_t c = lo; c < hi; ++c) {
					while (s.group_begin[j + 1] <= c)
						++j;
					table_emitter emit(L.table, s.child_begin[j]);
					rule.template symbolic_warp<false>(L.it.objects + s.object[j], s.size[j], s.ctx[j], (uint32_t)(c - s.group_begin[j]), s.group_ctx[c - lo],
					                   s_workspace[threadIdx.x >> 5], emit);
					created += emit.created;
				}
			} else if constexpr (Rule::lane_groups) {
				// one FAN per lane: a few children of one parent that share most of their work (rule_api.cuh)
				for (uint64_t c = lo + lane; c < hi; c += 32) {
					const uint32_t j = (uint32_t)upper_bound_u64(s.group_begin, count + 1, c) - 1;
					const uint64_t first_child = s.child_begin[j];
					rule.symbolic_fan(L.it.objects + s.object[j], s.size[j], s.ctx[j], (uint32_t)(c - s.group_begin[j]), s.mag[j], scratch,
					                  [&](uint32_t child_id, uint64_t hash, uint32_t size, cplx mag) {
						                  if (L.bins.records)
							                  created += bin_emit(L.bins, L.table, hash, mag, rep_pack(first_child + child_id, size));
						                  else
							                  created += table_insert(L.table, hash, mag, rep_pack(first_child + child_id, size));
					                  });
				}
			} else {
				// one child per lane: the loop body of quids.hpp:705-719.  (Collecting a lane's 4 children and inserting them as one
				// batch was measured 2.7x SLOWER on split_merge: the child walks then no longer overlap the inserts' round trips.)
				for (uint64_t c = lo + lane; c < hi; c += 32) {
					const uint32_t j = (uint32_t)upper_bound_u64(s.group_begin, count + 1, c) - 1;
					cplx mag = s.mag[j];
					uint32_t size;
					const uint32_t child_id = (uint32_t)(c - s.group_begin[j]);
					const uint64_t hash = rule.symbolic(L.it.objects + s.object[j], s.size[j], s.ctx[j], child_id, scratch, size, mag);
					if (L.bins.records) // a large table: the child goes to the bin of its table region (table.cuh), bin_insert_kernel does the rest
						created += bin_emit(L.bins, L.table, hash, mag, rep_pack(s.child_begin[j] + child_id, size));
					else
						created += table_insert(L.table, hash, mag, rep_pack(s.child_begin[j] + child_id, size));
				}
			}
			__syncwarp();
		}
	}
	// slots created -> one global atomic per warp (sizes the next call's table)
	created = (uint32_t)warp_sum((uint64_t)created);
	if (lane == 0 && created)
		atomicAdd(L.table.used, (unsigned long long)created);
}

// ---- sorted order: work items = (parent, group), ordered by the rule's group key -------------------------
constexpr int ITEM_GROUP_BITS = 24;
constexpr int POPULATE_COPIES = 4;  // parent copies one warp keeps in flight in the finalisation
constexpr int STAGED_THREADS = 96; // kernels whose warps stage their parents in shared memory: 3 stages of 11 KB per CTA
constexpr int ITEMS_BLOCKS_PER_SM = 5; // occupancy target of the sorted-order kernel (latency bound: ncu shows 29 % issue utilisation at 5)
constexpr int BATCH_BLOCKS_PER_SM = 6; // the same for the batch kernel (80 registers)
constexpr int ITEM_CHUNK = 256; // items one warp takes at a time

// (Measured and dropped: building the contexts in the kernel that counts the children, so that the parents are read once
// instead of twice.  The counting kernel only touches the particle bytes of an object, 0.8 of the 2.6 GB of a 1e7-parent
// state; with the contexts it reads everything and the step got slower, 7.71 -> 7.88 ms at 1e7 parents, 39.7 -> 41.4 ms
// at 1e8.)
// one lane per kept parent writes the keys and values of its groups.  A warp takes 32 consecutive parents; when they
// are consecutive in storage too (no parent truncation) and fit the stage, their bytes come to shared memory with one
// bulk copy and the lanes walk their object there: 32 lanes chasing 32 different objects in global memory cost one
// L1 wavefront per lane and load, the copy costs none.
// what a work item needs from its parent, written ONCE per kept parent by group_items_kernel: the symbolic kernel then does
// one gather per item (this record) instead of a chain item -> kept -> begin / size / magnitude / child_begin / context
// (ncu r1: long-scoreboard bound, 4.4x the compulsory traffic)
template <class Rule>
struct __align__(16) item_parent {
	typename Rule::ctx_t ctx;
	cplx mag;
	uint64_t child_begin;
	uint64_t object; // byte offset of the parent in the state
	uint32_t size;
	uint32_t pad_;
};

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <class Rule>
__global__ void __launch_bounds__(STAGED_THREADS) group_items_kernel(const Rule rule, const engine_launch L) {
	__shared__ warp_stage s_stage[STAGED_THREADS / 32];
	warp_stage &stage = s_stage[threadIdx.x >> 5];
	stage_init(stage);
	uint32_t phase = 0;
	const unsigned lane = lane_id();
	const uint64_t batches = div_up<uint64_t>(L.n_parents, 32);
	const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	for (uint64_t batch = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; batch < batches; batch += warps) {
		const uint64_t p0 = batch * 32, p = p0 + lane;
		const uint32_t in_batch = (uint32_t)min((uint64_t)32, L.n_parents - p0);
		const bool valid = lane < in_batch;
		const uint64_t oid = valid ? (L.kept ? L.kept[p] : p) : 0;
		const uint64_t off = valid ? L.it.begin[oid] : 0;
		const uint32_t size = valid ? L.it.size[oid] : 0;
		const uint8_t *object = L.it.objects + off;
		if (!L.kept) { // storage order: the batch is the byte range [begin of the first, end of the last]
			const uint64_t lo = __shfl_sync(0xffffffffu, off, 0);
			const uint64_t hi = __shfl_sync(0xffffffffu, off + size, in_batch - 1);
			if (hi - lo <= STAGE_BYTES) {
				const uint8_t *staged = stage_range(stage, L.it.objects + lo, (uint32_t)(hi - lo), phase);
				object = staged + (off - lo);
			}
		}
		if (!valid)
			continue;
		const uint64_t first = L.group_begin[p];
		const uint32_t count = (uint32_t)(L.group_begin[p + 1] - first);
		if (count == 0)
			continue;
		// the parent's context is prepared here once (this lane already walks the object) instead of once
		// per work item in the symbolic kernel, where 32 lanes would each chase a different object
		item_parent<Rule> record;
		rule.prepare(object, size, record.ctx);
		record.mag = L.it.mag[oid];
		record.child_begin = L.child_begin[p];
		record.object = off;
		record.size = size;
		record.pad_ = 0;
		static_cast<item_parent<Rule> *>(L.parent_ctx)[p] = record;
		const typename Rule::ctx_t &ctx = record.ctx;
		bool keyed = false;
		if constexpr (Rule::has_group_keys_from_ctx)
			keyed = rule.group_keys_from_ctx(ctx, count, L.item_keys + first);
		if (!keyed)
			rule.group_keys(object, size, count, L.item_keys + first);
		for (uint32_t g = 0; g < count; ++g)
			L.item_vals[first + g] = (p << ITEM_GROUP_BITS) | g;
	}
}

// every warp takes ITEM_CHUNK consecutive items of the sorted order; one lane per item prepares its
// parent's context and the group's root, then all lanes produce one group after the other with
// ACCUMULATE = true; what the rule still holds at the end of the chunk is flushed
template <class Rule, int BLOCKS_PER_SM>
__global__ void __launch_bounds__(SYMBOLIC_THREADS, BLOCKS_PER_SM) symbolic_items_kernel(const Rule rule, const engine_launch L) {
	typedef typename Rule::ctx_t ctx_t;
	struct warp_slice {
		ctx_t ctx[32];
		typename Rule::group_ctx_t group_ctx[32];
		uint64_t child_begin[32];
		uint64_t object[32];
		uint32_t size[32];
		uint32_t group[32];
	};
	__shared__ warp_slice s_slices[ENGINE_WARPS];
	__shared__ typename Rule::items_workspace_t s_workspace[ENGINE_WARPS];
	warp_slice &s = s_slices[threadIdx.x >> 5];
	typename Rule::items_workspace_t &ws = s_workspace[threadIdx.x >> 5];
	const unsigned lane = lane_id();
	uint32_t created = 0, regions = 0;
	rule.init_warp(ws);
	__syncwarp();

	const uint64_t num_chunks = div_up<uint64_t>(L.n_groups, ITEM_CHUNK);
	const uint64_t warp_stride = (uint64_t)gridDim.x * ENGINE_WARPS;
	for (uint64_t chunk = (uint64_t)blockIdx.x * ENGINE_WARPS + (threadIdx.x >> 5); chunk < num_chunks; chunk += warp_stride) {
		if (table_overflowed(L.table))
			break;
		const uint64_t c0 = chunk * ITEM_CHUNK, c1 = min(c0 + (uint64_t)ITEM_CHUNK, L.n_groups);
		const item_parent<Rule> *parents = static_cast<const item_parent<Rule> *>(L.parent_ctx);
		uint64_t next_item = c0 + lane < c1 ? L.items[c0 + lane] : 0;
		for (uint64_t b = c0; b < c1; b += 32) {
			const uint32_t count = (uint32_t)min((uint64_t)32, c1 - b);
			// the next batch's item is requested before this batch is processed, and its parent record is pulled into L2:
			// the gather below depends on the item, and dependent round trips to DRAM were the kernel's top stall
			const uint64_t item = next_item;
			next_item = b + 32 + lane < c1 ? L.items[b + 32 + lane] : 0;
			if (b + 32 + lane < c1) {
				const char *ahead = reinterpret_cast<const char *>(parents + (next_item >> ITEM_GROUP_BITS));
				prefetch_l2(ahead);
				prefetch_l2(ahead + sizeof(item_parent<Rule>) - 1);
			}
			if (lane < count) {
				const uint64_t p = item >> ITEM_GROUP_BITS;
				const uint32_t group = (uint32_t)(item & ((1u << ITEM_GROUP_BITS) - 1));
				const item_parent<Rule> &record = parents[p];
				s.ctx[lane] = record.ctx;
				s.child_begin[lane] = record.child_begin;
				s.object[lane] = record.object;
				s.size[lane] = record.size;
				s.group[lane] = group;
				rule.prepare_group(s.ctx[lane], group, record.mag, s.group_ctx[lane]);
			}
			__syncwarp();
			if constexpr (Rule::has_run_identity) {
				// stretches of items with the same identity: the head goes through the warp-wide path (it opens a run or
				// continues the one that is open), the others join it one per lane
				typename Rule::run_id_t id{};
				if (lane < count)
					id = rule.run_identity(s.ctx[lane], s.group[lane]);
				const typename Rule::run_id_t before = id.shuffle_up();
				const bool follows = lane > 0 && lane < count && id == before;
				unsigned heads = __ballot_sync(0xffffffffu, lane < count && !follows);
				while (heads) {
					const uint32_t h = __ffs(heads) - 1;
					heads &= heads - 1;
					const uint32_t e = heads ? __ffs(heads) - 1 : count;
					table_emitter emit(L.table, s.child_begin[h]);
					rule.template symbolic_warp<true>(L.it.objects + s.object[h], s.size[h], s.ctx[h], s.group[h], s.group_ctx[h], ws, emit);
					created += emit.created;
					regions += emit.regions;
					if (lane > h && lane < e)
						rule.continue_run(s.ctx[lane], s.group_ctx[lane], ws);
					__syncwarp();
				}
			} else {
				for (uint32_t i = 0; i < count; ++i) {
					table_emitter emit(L.table, s.child_begin[i]);
					rule.template symbolic_warp<true>(L.it.objects + s.object[i], s.size[i], s.ctx[i], s.group[i], s.group_ctx[i], ws, emit);
					created += emit.created;
					regions += emit.regions;
				}
			}
			__syncwarp();
		}
		table_emitter emit(L.table, 0);
		rule.flush_warp(ws, emit);
		created += emit.created;
		regions += emit.regions;
	}
	{
		table_emitter emit(L.table, 0);
		rule.finish_warp(ws, emit); // whatever the rule keeps per warp across chunks (the unused part of its range of table slots)
	}
	created = (uint32_t)warp_sum((uint64_t)created);
	regions = (uint32_t)warp_sum((uint64_t)regions);
	if (lane == 0 && created)
		atomicAdd(L.table.used, (unsigned long long)created);
	if (lane == 0 && regions)
		atomicAdd(L.table.regions, (unsigned long long)regions);
}

// BATCH mode of the sorted order (rules with has_region_batch, region mode, short runs): a warp takes 32 items at a time, one
// lane per item gathers its parent record and the root magnitude of its group, then the rule handles the whole batch
// (rule.region_batch: all directory probes of the batch in flight together, one slot allocation, one publication).
template <class Rule, int BLOCKS_PER_SM>
__global__ void __launch_bounds__(SYMBOLIC_THREADS, BLOCKS_PER_SM) symbolic_items_batch_kernel(const Rule rule, const engine_launch L) {
	if constexpr (Rule::has_region_batch) {
		typedef typename Rule::ctx_t ctx_t;
		struct warp_slice {
			ctx_t ctx[32];
			cplx root[32];
			uint64_t child_begin[32];
			uint32_t size[32];
			uint32_t group[32];
		};
		__shared__ warp_slice s_slices[ENGINE_WARPS];
		__shared__ typename Rule::items_workspace_t s_workspace[ENGINE_WARPS];
		warp_slice &s = s_slices[threadIdx.x >> 5];
		typename Rule::items_workspace_t &ws = s_workspace[threadIdx.x >> 5];
		const unsigned lane = lane_id();
		uint32_t created = 0, regions = 0;
		rule.init_warp(ws);
		__syncwarp();

		const item_parent<Rule> *parents = static_cast<const item_parent<Rule> *>(L.parent_ctx);
		const uint64_t num_chunks = div_up<uint64_t>(L.n_groups, ITEM_CHUNK);
		const uint64_t warp_stride = (uint64_t)gridDim.x * ENGINE_WARPS;
		for (uint64_t chunk = (uint64_t)blockIdx.x * ENGINE_WARPS + (threadIdx.x >> 5); chunk < num_chunks; chunk += warp_stride) {
			if (table_overflowed(L.table))
				break;
			const uint64_t c0 = chunk * ITEM_CHUNK, c1 = min(c0 + (uint64_t)ITEM_CHUNK, L.n_groups);
			uint64_t next_item = c0 + lane < c1 ? L.items[c0 + lane] : 0;
			for (uint64_t b = c0; b < c1; b += 32) {
				const uint32_t count = (uint32_t)min((uint64_t)32, c1 - b);
				// the next batch's item is requested before this batch is processed, and its parent record is pulled into L2
				const uint64_t item = next_item;
				next_item = b + 32 + lane < c1 ? L.items[b + 32 + lane] : 0;
				if (b + 32 + lane < c1) {
					const char *ahead = reinterpret_cast<const char *>(parents + (next_item >> ITEM_GROUP_BITS));
					prefetch_l2(ahead);
					prefetch_l2(ahead + sizeof(item_parent<Rule>) - 1);
				}
				if (lane < count) {
					const uint64_t p = item >> ITEM_GROUP_BITS;
					const uint32_t group = (uint32_t)(item & ((1u << ITEM_GROUP_BITS) - 1));
					const item_parent<Rule> &record = parents[p];
					s.ctx[lane] = record.ctx;
					s.child_begin[lane] = record.child_begin;
					s.size[lane] = record.size;
					s.group[lane] = group;
					s.root[lane] = rule.root_magnitude(s.ctx[lane], group, record.mag);
				}
				__syncwarp();
				rule.region_batch(s.ctx, s.root, s.child_begin, s.size, s.group, count, ws, L.table, created, regions);
				__syncwarp();
			}
		}
		created = (uint32_t)warp_sum((uint64_t)created);
		regions = (uint32_t)warp_sum((uint64_t)regions);
		if (lane == 0 && created)
			atomicAdd(L.table.used, (unsigned long long)created);
		if (lane == 0 && regions)
			atomicAdd(L.table.regions, (unsigned long long)regions);
	}
}

// parent holding the first group of every chunk (one binary search per chunk, all in parallel, instead of
// a chain of dependent loads at the head of every chunk of the symbolic kernel)
static __global__ void __launch_bounds__(ENGINE_THREADS) chunk_parent_kernel(const uint64_t *group_begin, uint64_t n_parents, uint64_t n_groups, uint32_t chunk,
                                                                    uint64_t num_chunks, uint64_t *chunk_parent) {
	const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (c < num_chunks)
		chunk_parent[c] = upper_bound_u64(group_begin, n_parents + 1, c * chunk) - 1;
	else if (c == num_chunks)
		chunk_parent[c] = n_parents - 1;
}

// finalisation (quids.hpp:958-967): every surviving child is rebuilt in place from its parent.
// Rules with edit_child: a warp takes 32 survivors at a time -- one lane per survivor fetches its
// metadata (coalesced), the warp copies the 32 parents with 8-byte words, four copies in flight, then
// one lane per survivor applies the rule's edits and zeroes the alignment padding.
// Other rules: one thread per child runs populate_child_simple.
template <class Rule>
__global__ void __launch_bounds__(ENGINE_THREADS) populate_kernel(const Rule rule, const engine_launch L) {
	if constexpr (Rule::has_edit_child) {
		const unsigned lane = lane_id();
		const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
		const uint64_t batches = div_up<uint64_t>(L.n_survivors, 32);
		for (uint64_t batch = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; batch < batches; batch += warps) {
			const uint64_t mine = batch * 32 + lane;
			const bool valid = mine < L.n_survivors;
			const uint8_t *parent = nullptr;
			uint8_t *child = nullptr;
			uint32_t size = 0, child_id = 0, padded = 0;
			if (valid) {
				const uint64_t oid = L.survivor_parent[mine];
				const uint64_t begin = L.next_begin[mine];
				parent = L.it.objects + L.it.begin[oid];
				child = L.next_objects + begin;
				size = L.it.size[oid];
				child_id = L.survivor_child[mine];
				padded = (uint32_t)(L.next_begin[mine + 1] - begin);
			}
			const unsigned count = __popc(__ballot_sync(0xffffffffu, valid));
			// small objects of one size (a qubit register: 24 bytes): one object would keep 3 of the 32 lanes busy, so the
			// words of the whole batch are spread over the lanes instead -- word k of the batch = word k % w of object k / w
			const uint32_t size0 = __shfl_sync(0xffffffffu, size, 0);
			const bool regular = !valid || (size == size0 && padded == size0 && ((reinterpret_cast<uintptr_t>(parent) | reinterpret_cast<uintptr_t>(child)) & 7) == 0);
			if (size0 > 0 && size0 <= 128 && (size0 & 7) == 0 && __all_sync(0xffffffffu, regular)) {
				const uint32_t w = size0 / 8, total = count * w;
				const uint32_t inv = (65536u + w - 1) / w; // k / w for k < 512
				uint2 *out = reinterpret_cast<uint2 *>(reinterpret_cast<uint8_t *>(__shfl_sync(0xffffffffu, reinterpret_cast<uintptr_t>(child), 0)));
				for (uint32_t k0 = 0; k0 < total; k0 += 32 * POPULATE_COPIES) {
					uint2 word[POPULATE_COPIES];
#pragma unroll
					for (int q = 0; q < POPULATE_COPIES; ++q) {
						const uint32_t k = k0 + q * 32 + lane;
						const uint32_t obj = min((k * inv) >> 16, count - 1);
						const uint2 *from = reinterpret_cast<const uint2 *>(__shfl_sync(0xffffffffu, reinterpret_cast<uintptr_t>(parent), obj));
						if (k < total)
							word[q] = from[k - obj * w];
					}
#pragma unroll
					for (int q = 0; q < POPULATE_COPIES; ++q) {
						const uint32_t k = k0 + q * 32 + lane;
						if (k < total)
							out[k] = word[q];
					}
				}
				__syncwarp();
				if (valid)
					rule.edit_child(parent, size, child, child_id);
				continue;
			}
			for (unsigned j = 0; j < count; j += POPULATE_COPIES) {
				// several parents at a time: all their loads are issued before the first store
				uint2 word[POPULATE_COPIES];
				const uint8_t *src[POPULATE_COPIES];
				uint8_t *dst[POPULATE_COPIES];
				uint32_t bytes[POPULATE_COPIES];
				bool fast[POPULATE_COPIES];
#pragma unroll
				for (int q = 0; q < POPULATE_COPIES; ++q) {
					const unsigned from = min(j + q, count - 1);
					src[q] = reinterpret_cast<const uint8_t *>(__shfl_sync(0xffffffffu, reinterpret_cast<uintptr_t>(parent), from));
					dst[q] = reinterpret_cast<uint8_t *>(__shfl_sync(0xffffffffu, reinterpret_cast<uintptr_t>(child), from));
					bytes[q] = j + q < count ? __shfl_sync(0xffffffffu, size, from) : 0;
					fast[q] = ((reinterpret_cast<uintptr_t>(src[q]) | reinterpret_cast<uintptr_t>(dst[q])) & 7) == 0 && bytes[q] <= 256;
					if (fast[q] && lane < bytes[q] / 8)
						word[q] = reinterpret_cast<const uint2 *>(src[q])[lane];
				}
#pragma unroll
				for (int q = 0; q < POPULATE_COPIES; ++q) {
					if (fast[q]) { // 8-byte words, then the (size % 8) trailing bytes
						if (lane < bytes[q] / 8)
							reinterpret_cast<uint2 *>(dst[q])[lane] = word[q];
						const uint32_t tail = bytes[q] & ~7u;
						if (tail + lane < bytes[q])
							dst[q][tail + lane] = src[q][tail + lane];
					} else { // any size, any alignment
						for (uint32_t b = lane; b < bytes[q]; b += 32)
							dst[q][b] = src[q][b];
					}
				}
			}
			__syncwarp();
			if (valid) {
				rule.edit_child(parent, size, child, child_id);
				for (uint32_t b = size; b < padded; ++b)
					child[b] = 0;
			}
		}
	} else {
		const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
		for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < L.n_survivors; s += stride) {
			const uint64_t oid = L.survivor_parent[s];
			uint8_t *child = L.next_objects + L.next_begin[s];
			rule.populate_child_simple(L.it.objects + L.it.begin[oid], L.it.size[oid], child, L.survivor_child[s]);
			for (uint64_t b = L.next_begin[s] + L.next_size[s]; b < L.next_begin[s + 1]; ++b)
				L.next_objects[b] = 0;
		}
	}
}

// ---- finalisation of rules WITHOUT edit_child (children of any size, e.g. split_merge): staged ------------------------
// One lane per child running populate_child_simple straight on global memory costs one L1 wavefront per lane for
// every 1/2/4-byte access of the walk (32 different parents, 32 different children per warp instruction).  Here a
// warp takes 32 consecutive survivors and, for as many of them as fit its two stages,
//   1. fetches their parents into shared memory, one bulk copy (cp.async.bulk, mbarrier) per parent;
//   2. builds the children in shared memory, one lane per child, padding zeroed (same 16-byte phase as their place in
//      the next state, so the rule's alignment assumptions hold);
//   3. writes the children -- consecutive in the next state, hence ONE contiguous byte range -- with a single bulk
//      store (cp.async.bulk shared -> global; the < 16-byte head and tail by the lanes).
// A child or parent too large for a stage is built in place by one lane.
// Stage sizes: a round handles the survivors whose parents AND children fit; 32 grown 12-node graphs (300-330 bytes each, slots
// of 336) need 10.75 KB on either side -- with 8.5 KB stages every batch of 32 took two rounds (27 + 5 lanes busy).  Three
// warps per CTA: 3 x 21.5 KB = 64.6 KB, three CTAs per SM.
constexpr uint32_t POPULATE_PARENT_STAGE = 10752;
constexpr uint32_t POPULATE_CHILD_STAGE = 10752;
constexpr int POPULATE_THREADS = 96;
struct __align__(16) populate_stage {
	uint8_t parents[POPULATE_PARENT_STAGE];
	uint8_t children[POPULATE_CHILD_STAGE];
	unsigned long long mbar;
	unsigned long long pad_;
};

__device__ __forceinline__ uint32_t warp_inclusive_sum(uint32_t v) {
	const unsigned lane = lane_id();
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t up = __shfl_up_sync(0xffffffffu, v, o);
		if (lane >= (unsigned)o)
			v += up;
	}
	return v;
}

template <class Rule>
__global__ void __launch_bounds__(POPULATE_THREADS) populate_staged_kernel(const Rule rule, const engine_launch L) {
	extern __shared__ __align__(16) uint8_t s_populate[];
	populate_stage &st = reinterpret_cast<populate_stage *>(s_populate)[threadIdx.x >> 5];
	const unsigned lane = lane_id();
	const uint32_t bar = smem_addr(&st.mbar);
	if (lane == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncwarp();
	uint32_t phase = 0;
	const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	const uint64_t batches = div_up<uint64_t>(L.n_survivors, 32);
	// what a lane needs to know about its survivor: fetched one batch AHEAD (survivor -> parent -> offset and size of the
	// parent is a chain of dependent round trips to DRAM; requested while the previous batch is built, it costs nothing)
	struct survivor_meta {
		uint64_t parent_begin, dst, dst_end;
		uint32_t psize, csize, child_id;
	};
	auto fetch = [&](uint64_t batch) {
		survivor_meta m{0, 0, 0, 0, 0, 0};
		const uint64_t mine = batch * 32 + lane;
		if (batch < batches && mine < L.n_survivors) {
			const uint64_t oid = L.survivor_parent[mine];
			m.parent_begin = L.it.begin[oid];
			m.psize = L.it.size[oid];
			m.dst = L.next_begin[mine];
			m.dst_end = L.next_begin[mine + 1];
			m.csize = L.next_size[mine];
			m.child_id = L.survivor_child[mine];
		}
		return m;
	};
	bool store_in_flight = false; // a bulk store of the children stage may still be reading it
	const uint64_t first_batch = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	survivor_meta ahead = fetch(first_batch);
	for (uint64_t batch = first_batch; batch < batches; batch += warps) {
		const uint64_t mine = batch * 32 + lane;
		const bool valid = mine < L.n_survivors;
		const survivor_meta meta = ahead;
		ahead = fetch(batch + warps);
		const uint8_t *parent = L.it.objects;
		uint64_t dst = 0;
		uint32_t psize = 0, csize = 0, cbytes = 0, child_id = 0, lead = 0, pbytes = 0, pslot = 0;
		if (valid) {
			parent = L.it.objects + meta.parent_begin;
			psize = meta.psize;
			dst = meta.dst;
			cbytes = (uint32_t)(meta.dst_end - dst);
			csize = meta.csize;
			child_id = meta.child_id;
			lead = (uint32_t)(reinterpret_cast<uintptr_t>(parent) & 15);
			pbytes = (lead + psize + 15u) & ~15u;
			// the lanes walk their parents in step: equal strides that are a multiple of 128 bytes (fresh 12-node graphs:
			// 256) would put every lane on the same bank (ncu: 7e8 conflicts); an odd multiple of 16 spreads them over 8
			pslot = (pbytes & 16u) ? pbytes : pbytes + 16;
		}
		const uint32_t count = __popc(__ballot_sync(0xffffffffu, valid));
		const uint32_t p_incl = warp_inclusive_sum(pslot), c_incl = warp_inclusive_sum(cbytes);
		uint32_t start = 0;
		while (start < count) {
			const uint32_t p0 = __shfl_sync(0xffffffffu, p_incl - pslot, start), c0 = __shfl_sync(0xffffffffu, c_incl - cbytes, start);
			const uint64_t dst0 = __shfl_sync(0xffffffffu, dst, start);
			uint8_t *out = L.next_objects + dst0;
			const uint32_t clead = (uint32_t)(reinterpret_cast<uintptr_t>(out) & 15);
			const bool fits = lane >= start && lane < count && p_incl - p0 <= POPULATE_PARENT_STAGE && clead + (c_incl - c0) <= POPULATE_CHILD_STAGE;
			const uint32_t m = __popc(__ballot_sync(0xffffffffu, fits)); // both sums grow with the lane: the lanes that fit are start .. start + m - 1
			if (m == 0) { // too large for the stages: built in place
				if (lane == start) {
					uint8_t *child = L.next_objects + dst;
					rule.populate_child_simple(parent, psize, child, child_id);
					for (uint32_t b = csize; b < cbytes; ++b)
						child[b] = 0;
				}
				__syncwarp();
				++start;
				continue;
			}
			const uint32_t end = start + m;
			const bool in = lane >= start && lane < end;
			// 1. parents -> shared memory
			const uint32_t total = (uint32_t)warp_sum((uint64_t)(in ? pbytes : 0u));
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the lanes' earlier reads of the stage come first
			__syncwarp();
			if (lane == start)
				asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(total) : "memory");
			__syncwarp();
			uint8_t *staged = st.parents + (p_incl - pslot - p0);
			if (in)
				asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(staged)), "l"(parent - lead),
				             "r"(pbytes), "r"(bar)
				             : "memory");
			uint32_t done = 0;
			while (!done)
				asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(phase) : "memory");
			phase ^= 1;
			// 2. children built in shared memory (once the previous round's bulk store has read them out of the stage)
			if (store_in_flight) {
				if (lane == 0)
					asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
				__syncwarp();
				store_in_flight = false;
			}
			if (in) {
				uint8_t *child = st.children + clead + (c_incl - cbytes - c0);
				rule.populate_child_simple(staged + lead, psize, child, child_id);
				for (uint32_t b = csize; b < cbytes; ++b)
					child[b] = 0;
			}
			// 3. one contiguous range of the next state
			const uint32_t len = __shfl_sync(0xffffffffu, c_incl, end - 1) - c0;
			const uint8_t *from = st.children + clead;
			const uint32_t head = min(len, (16u - clead) & 15u), body = (len - head) & ~15u, tail = len - head - body;
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the children become visible to the copy engine
			__syncwarp();
			if (lane == 0 && body) {
				asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + head), "r"(smem_addr(from + head)), "r"(body) : "memory");
				asm volatile("cp.async.bulk.commit_group;" ::: "memory");
			}
			if (lane < head)
				out[lane] = from[lane];
			if (lane < tail)
				out[head + body + lane] = from[head + body + lane];
			store_in_flight = body != 0; // waited for before the stage is written again: the store overlaps the next round's parent fetch
			__syncwarp();
			start = end;
		}
	}
	if (lane == 0)
		asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <class Rule>
__global__ void __launch_bounds__(ENGINE_THREADS) hash_kernel(const Rule rule, iter_view it, uint64_t *hashes) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < it.n; i += stride)
		hashes[i] = rule.hasher(it.objects + it.begin[i], it.size[i]);
}

// family keys of all objects (distributed path, route.inc.cuh).  A warp takes 32 consecutive objects: their bytes come to shared
// memory with one bulk copy and every lane walks its object there (as in group_items_kernel: 32 lanes chasing 32 different
// objects in global memory pay one L1 wavefront per lane and load)
template <class Rule>
__global__ void __launch_bounds__(STAGED_THREADS) family_kernel(const Rule rule, iter_view it, uint64_t *family) {
	__shared__ warp_stage s_stage[STAGED_THREADS / 32];
	warp_stage &stage = s_stage[threadIdx.x >> 5];
	stage_init(stage);
	uint32_t phase = 0;
	const unsigned lane = lane_id();
	const uint64_t batches = div_up<uint64_t>(it.n, 32);
	const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	for (uint64_t batch = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; batch < batches; batch += warps) {
		const uint64_t p0 = batch * 32, p = p0 + lane;
		const uint32_t in_batch = (uint32_t)min((uint64_t)32, it.n - p0);
		const bool valid = lane < in_batch;
		const uint64_t off = valid ? it.begin[p] : 0;
		const uint32_t size = valid ? it.size[p] : 0;
		const uint8_t *object = it.objects + off;
		const uint64_t lo = __shfl_sync(0xffffffffu, off, 0);
		const uint64_t hi = __shfl_sync(0xffffffffu, off + size, in_batch - 1);
		if (hi - lo <= STAGE_BYTES) {
			const uint8_t *staged = stage_range(stage, it.objects + lo, (uint32_t)(hi - lo), phase);
			object = staged + (off - lo);
		}
		if (valid)
			family[p] = rule.family_key(object, size);
	}
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

// sum over the objects of observable(object) * |mag|^2 (quids.hpp:208-234), K values per object in one pass over the
// state; per-CTA partial sums (fixed order inside a CTA), summed in CTA order by observable_total_kernel
template <class Observable>
__global__ void __launch_bounds__(ENGINE_THREADS) observable_kernel(const Observable observable, iter_view it, double *partial) {
	constexpr int K = Observable::values;
	__shared__ double s_part[ENGINE_THREADS / 32][K];
	double local[K];
#pragma unroll
	for (int k = 0; k < K; ++k)
		local[k] = 0;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < it.n; i += stride) {
		double value[K];
		observable(it.objects + it.begin[i], it.size[i], value);
		const double weight = cnorm(it.mag[i]);
#pragma unroll
		for (int k = 0; k < K; ++k)
			local[k] += value[k] * weight;
	}
#pragma unroll
	for (int k = 0; k < K; ++k) {
		local[k] = warp_sum(local[k]);
		if (lane_id() == 0)
			s_part[threadIdx.x >> 5][k] = local[k];
	}
	__syncthreads();
	if (threadIdx.x < K) {
		double sum = 0;
		for (int w = 0; w < ENGINE_THREADS / 32; ++w)
			sum += s_part[w][threadIdx.x];
		partial[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = sum;
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
		num_child_kernel<Rule><<<grid, ENGINE_THREADS, 0, L.stream>>>(*static_cast<const Rule *>(rule), L.it, L.num_childs, L.num_groups, L.max_child_size, L.child_count_range);
		++*L.launch_counter;
	}
	static uint64_t symbolic_chunks(uint64_t n_groups) { return div_up<uint64_t>(n_groups, Rule::warp_groups ? 32 : SYMBOLIC_CHUNK); }
	static int symbolic_grid(int sm_count) { return resident_grid((const void *)symbolic_kernel<Rule>, SYMBOLIC_THREADS, sm_count); }
	static void symbolic(const void *rule, const engine_launch &L) {
		constexpr int chunk = Rule::warp_groups ? 32 : SYMBOLIC_CHUNK;
		const uint64_t warps = div_up<uint64_t>(L.n_groups, chunk);
		Rule::prepare_device(L.stream);
		chunk_parent_kernel<<<(unsigned)div_up<uint64_t>(warps + 1, ENGINE_THREADS), ENGINE_THREADS, 0, L.stream>>>(
		    L.group_begin, L.n_parents, L.n_groups, chunk, warps, const_cast<uint64_t *>(L.chunk_parent));
		++*L.launch_counter;
		int grid = grid_for(warps * 32, SYMBOLIC_THREADS, symbolic_grid(L.sm_count));
		symbolic_kernel<Rule><<<grid, SYMBOLIC_THREADS, 0, L.stream>>>(*static_cast<const Rule *>(rule), L);
		++*L.launch_counter;
	}
	static void group_items(const void *rule, const engine_launch &L) {
		if constexpr (Rule::has_group_key) {
			int grid = grid_for(L.n_parents, STAGED_THREADS, resident_grid((const void *)group_items_kernel<Rule>, STAGED_THREADS, L.sm_count));
			group_items_kernel<Rule><<<grid, STAGED_THREADS, 0, L.stream>>>(*static_cast<const Rule *>(rule), L);
			++*L.launch_counter;
		}
	}
	static void symbolic_items(const void *rule, const engine_launch &L) {
		if constexpr (Rule::has_group_key) {
			const uint64_t warps = div_up<uint64_t>(L.n_groups, ITEM_CHUNK);
			Rule::prepare_device(L.stream);
			// occupancy target: the register budget follows from it (5 CTAs of 4 warps: 96 registers).  QB_ITEMS_BLOCKS is a developer
			// knob for A/B runs (4: 128 registers, no spills; 6: 80 registers)
			static const int blocks = getenv("QB_ITEMS_BLOCKS") ? atoi(getenv("QB_ITEMS_BLOCKS")) : ITEMS_BLOCKS_PER_SM;
			auto launch = [&](auto kernel) {
				int grid = grid_for(warps * 32, SYMBOLIC_THREADS, resident_grid((const void *)kernel, SYMBOLIC_THREADS, L.sm_count));
				kernel<<<grid, SYMBOLIC_THREADS, 0, L.stream>>>(*static_cast<const Rule *>(rule), L);
			};
			if (blocks == 4)
				launch(symbolic_items_kernel<Rule, 4>);
			else if (blocks == 6)
				launch(symbolic_items_kernel<Rule, 6>);
			else
				launch(symbolic_items_kernel<Rule, ITEMS_BLOCKS_PER_SM>);
			++*L.launch_counter;
		}
	}
	static void symbolic_items_batch(const void *rule, const engine_launch &L) {
		if constexpr (Rule::has_group_key && Rule::has_region_batch) {
			const uint64_t warps = div_up<uint64_t>(L.n_groups, ITEM_CHUNK);
			Rule::prepare_device(L.stream);
			// CTAs per SM (the register budget follows): measured on the loop state at 1e7 parents, 4 -> 15.5 ms, 5 -> 14.8, 6 -> 14.0
			static const int blocks = getenv("QB_ITEMS_BLOCKS") ? atoi(getenv("QB_ITEMS_BLOCKS")) : BATCH_BLOCKS_PER_SM; // (developer knob, as above)
			auto launch = [&](auto kernel) {
				int grid = grid_for(warps * 32, SYMBOLIC_THREADS, resident_grid((const void *)kernel, SYMBOLIC_THREADS, L.sm_count));
				kernel<<<grid, SYMBOLIC_THREADS, 0, L.stream>>>(*static_cast<const Rule *>(rule), L);
			};
			if (blocks == 5)
				launch(symbolic_items_batch_kernel<Rule, 5>);
			else if (blocks == 7)
				launch(symbolic_items_batch_kernel<Rule, 7>);
			else if (blocks == 8)
				launch(symbolic_items_batch_kernel<Rule, 8>);
			else
				launch(symbolic_items_batch_kernel<Rule, BATCH_BLOCKS_PER_SM>);
			++*L.launch_counter;
		}
	}
	static void populate(const void *rule, const engine_launch &L) {
		if constexpr (Rule::has_edit_child) {
			int grid = grid_for(L.n_survivors, ENGINE_THREADS, resident_grid((const void *)populate_kernel<Rule>, ENGINE_THREADS, L.sm_count));
			populate_kernel<Rule><<<grid, ENGINE_THREADS, 0, L.stream>>>(*static_cast<const Rule *>(rule), L);
		} else {
			constexpr size_t smem = sizeof(populate_stage) * (POPULATE_THREADS / 32);
			// CTAs of this kernel one SM holds (shared-memory bound); the opt-in to more than 48 KB is a per-DEVICE attribute
			static int per_sm_of_device[MAX_DEVICES] = {};
			int device = 0;
			QB_CUDA(cudaGetDevice(&device));
			int &per_sm = per_sm_of_device[device % MAX_DEVICES];
			if (per_sm == 0) {
				QB_CUDA(cudaFuncSetAttribute((const void *)populate_staged_kernel<Rule>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
				QB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)populate_staged_kernel<Rule>, POPULATE_THREADS, smem));
				if (per_sm < 1)
					per_sm = 1;
			}
			int grid = grid_for(L.n_survivors, POPULATE_THREADS, per_sm * L.sm_count);
			populate_staged_kernel<Rule><<<grid, POPULATE_THREADS, smem, L.stream>>>(*static_cast<const Rule *>(rule), L);
		}
		++*L.launch_counter;
	}
	static void hash(const void *rule, const engine_launch &L) {
		int grid = grid_for(L.it.n, ENGINE_THREADS, resident_grid((const void *)hash_kernel<Rule>, ENGINE_THREADS, L.sm_count));
		hash_kernel<Rule><<<grid, ENGINE_THREADS, 0, L.stream>>>(*static_cast<const Rule *>(rule), L.it, L.hashes);
		++*L.launch_counter;
	}
	static void family(const void *rule, const engine_launch &L) {
		if constexpr (Rule::has_family) {
			int grid = grid_for(L.it.n, STAGED_THREADS, resident_grid((const void *)family_kernel<Rule>, STAGED_THREADS, L.sm_count));
			family_kernel<Rule><<<grid, STAGED_THREADS, 0, L.stream>>>(*static_cast<const Rule *>(rule), L.it, L.hashes);
			++*L.launch_counter;
		}
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
		o.warp_groups = Rule::warp_groups;
		o.has_groups = Rule::warp_groups || Rule::lane_groups;
		o.has_group_key = Rule::has_group_key;
		o.region_size_limit = Rule::region_size_limit;
		o.ctx_bytes = sizeof(item_parent<Rule>); // per kept parent in sorted order
		o.group_capacity = Rule::group_capacity;
		o.launch_group_items = group_items;
		o.launch_symbolic_items = symbolic_items;
		o.has_region_batch = Rule::has_group_key && Rule::has_region_batch;
		o.launch_symbolic_items_batch = symbolic_items_batch;
		o.has_family = Rule::has_family;
		o.launch_family = family;
		o.symbolic_grid = symbolic_grid;
		o.symbolic_chunks = symbolic_chunks;
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

template <class Observable>
struct observable_glue {
	static_assert(sizeof(Observable) <= RULE_STORAGE_BYTES, "device observables are passed by value");
	static_assert(Observable::values >= 1 && Observable::values <= OBSERVABLE_MAX_VALUES, "values per object");
	static int launch(const void *observable, const iter_view &it, double *partial, cudaStream_t stream, int sm_count) {
		int grid = grid_for(it.n, ENGINE_THREADS, resident_grid((const void *)observable_kernel<Observable>, ENGINE_THREADS, sm_count));
		observable_kernel<Observable><<<grid, ENGINE_THREADS, 0, stream>>>(*static_cast<const Observable *>(observable), it, partial);
		return grid;
	}
};

#define QB_REGISTER_OBSERVABLE(NAME, TYPE, MAKE) \
	static const int qb_observable_registered_##NAME = \
	    ::qb::register_observable(::qb::observable_ops{#NAME, TYPE::values, MAKE, ::qb::observable_glue<TYPE>::launch})
#define QB_REGISTER_RULE(NAME, TYPE, MAKE) static const int qb_rule_registered_##NAME = ::qb::register_rule(::qb::rule_glue<TYPE>::ops(#NAME, MAKE))
#define QB_REGISTER_MODIFIER(NAME, TYPE, MAKE) \
	static const int qb_modifier_registered_##NAME = ::qb::register_modifier(::qb::modifier_ops{#NAME, MAKE, ::qb::modifier_glue<TYPE>::launch})

} // namespace qb
