// rule_api.cuh -- what a rule is on the device, and how rules reach the engine.
//
// The reference's plug-in interface is the abstract class quids::rule (quids.hpp:105-146) with four
// virtual methods called from OpenMP loops.  Host objects with vtables cannot be called from a
// kernel, so here a rule is a trivially copyable struct passed BY VALUE to the kernels, carrying
// its parameters and implementing the same four methods as __device__ members (static dispatch:
// the engine's kernels are templates instantiated per rule type):
//
//     get_num_child(parent, parent_size, num_child&, max_child_size&)        quids.hpp:115
//     populate_child(parent, parent_size, child, child_id, size&, mag&)      quids.hpp:124
//     populate_child_simple(parent, parent_size, child, child_id)            quids.hpp:131-136 (optional)
//     hasher(object, size)                                                   quids.hpp:143-145 (optional, default = libstdc++ murmur)
//
// plus two OPTIONAL hooks that only exist on the GPU path:
//     prepare(parent, parent_size, ctx&)   per-parent precomputation shared by all its children
//     symbolic(parent, parent_size, ctx, child_id, scratch, size&, mag&) -> hash
//                                          size, magnitude and hash of a child WITHOUT writing its
//                                          bytes; the default materialises the child into `scratch`
//                                          with populate_child and calls hasher on it, exactly like
//                                          the symbolic loop of the reference (quids.hpp:705-719)
//     get_num_group(parent, parent_size, num_child) / symbolic_warp(..., group, ..., workspace, emit)
//                                          children of one parent may be produced in GROUPS by a whole
//                                          warp that shares work between siblings (e.g. expands the
//                                          binary tree of their choices level by level in shared memory,
//                                          reusing the common prefix of hash and magnitude); a group
//                                          must emit each of its children exactly once.  Default: one
//                                          child per group, one lane per child.
//
// QB_REGISTER_RULE(name, type, make) instantiates the engine kernels for `type` in the translation
// unit where it appears and adds the rule to the registry looked up by qb_rule_id().
#pragma once

#include "common.cuh"

namespace qb {

// ---- default hasher: libstdc++ std::hash<std::string_view> = _Hash_bytes (64-bit murmur variant,
// gcc libsupc++/hash_bytes.cc, seed 0xc70f6907); the byte source is a functor so that rules can hash
// a child that only exists as "parent with one byte changed"
template <class ByteAt>
__device__ __forceinline__ uint64_t murmur_bytes(ByteAt at, uint32_t len) {
	uint64_t h = 0xc70f6907ull ^ ((uint64_t)len * MURMUR_MUL);
	const uint32_t whole = len & ~7u;
	for (uint32_t i = 0; i < whole; i += 8) {
		uint64_t w = 0;
#pragma unroll
		for (int b = 7; b >= 0; --b)
			w = (w << 8) | (uint64_t)at(i + b);
		h ^= shift_mix(w * MURMUR_MUL) * MURMUR_MUL;
		h *= MURMUR_MUL;
	}
	if (len & 7u) {
		uint64_t w = 0;
		for (uint32_t b = len & 7u; b-- > 0;)
			w = (w << 8) + (uint64_t)at(whole + b);
		h ^= w;
		h *= MURMUR_MUL;
	}
	h = shift_mix(h) * MURMUR_MUL;
	return shift_mix(h);
}

__device__ __forceinline__ uint64_t murmur_bytes(const uint8_t *p, uint32_t len) {
	return murmur_bytes([p](uint32_t i) { return p[i]; }, len);
}

struct no_ctx {};

template <class Derived>
struct rule_base {
	typedef no_ctx ctx_t;
	// true: `symbolic` needs a scratch buffer of max_child_size bytes per thread
	static constexpr bool needs_scratch = true;

	__device__ const Derived &self() const { return *static_cast<const Derived *>(this); }

	__device__ void prepare(const uint8_t *, uint32_t, no_ctx &) const {}
	// optional: the context of a parent is built by a whole warp -- prepare_warp(parent, parent_size, ctx&) is called by all
	// 32 lanes with the same arguments after prepare() -- and, for large contexts, fewer parents are staged at a time
	static constexpr bool warp_prepare = false;
	static constexpr int parents_per_batch = 32;
	// with warp_prepare: the rule can also build TWO contexts at a time, one per half warp --
	//     static bool fits_half_warp(parent);  prepare_half_warp(parent, ctx&, active)   (all 32 lanes, arguments of the lane's half)
	static constexpr bool warp_prepare_pairs = false;
	// with warp_prepare: the parents of a batch (consecutive in the state) are first copied to shared memory in one bulk copy
	// when they fit in this many bytes, and prepare_warp reads them there: its chains of dependent loads (node count -> name
	// offsets -> atoms) then cost shared-memory latency instead of one DRAM round trip per link and parent.  0 = no staging
	static constexpr uint32_t prepare_stage_bytes = 0;

	__device__ uint64_t hasher(const uint8_t *object, uint32_t size) const { return murmur_bytes(object, size); }

	__device__ void populate_child_simple(const uint8_t *parent, uint32_t parent_size, uint8_t *child, uint32_t child_id) const {
		uint32_t size;
		cplx mag{1, 0};
		self().populate_child(parent, parent_size, child, child_id, size, mag);
	}

	template <class Ctx>
	__device__ uint64_t symbolic(const uint8_t *parent, uint32_t parent_size, const Ctx &, uint32_t child_id, uint8_t *scratch,
	                             uint32_t &size, cplx &mag) const {
		self().populate_child(parent, parent_size, scratch, child_id, size, mag);
		return self().hasher(scratch, size);
	}

	// false: a group is a single child handled by one lane (the engine calls symbolic() itself and
	// skips the second prefix sum).  true: the rule provides
	//     workspace_t                                  per-warp shared memory it needs
	//     get_num_group(parent, parent_size, num_child) -> number of groups
	//     group_ctx_t, prepare_group(ctx, group, parent_mag, group_ctx&)   one lane per group
	//     symbolic_warp<ACCUMULATE>(parent, parent_size, ctx, group, group_ctx, workspace&, emit)
	// symbolic_warp is called by all 32 lanes of a warp with identical arguments; every child of the
	// group must be emitted exactly once, by any lane: emit(child_id, hash, size, mag) or
	// emit.batch<N>(count, hash[N], size, child_id_of(i), mag_of(i)).
	static constexpr bool warp_groups = false;
	// LANE groups ("fans"): a group is a few children of one parent produced by ONE lane that shares work between them (e.g.
	// siblings that differ in their last choices share the walk over the parent up to there).  The rule provides
	//     get_num_group(parent, parent_size, num_child) -> number of fans
	//     symbolic_fan(parent, parent_size, ctx, fan, parent_mag, scratch, emit)     emit(child_id, hash, size, mag) once per child
	static constexpr bool lane_groups = false;
	struct workspace_t {};
	typedef workspace_t items_workspace_t; // the sorted order may need less (or other) per-warp memory than the unsorted one
	// per-group precomputation done by ONE lane per group, 32 groups at a time (whatever is the same
	// for all children of the group and would otherwise be recomputed identically by all 32 lanes)
	struct group_ctx_t {};
	template <class Ctx>
	__device__ void prepare_group(const Ctx &, uint32_t, cplx, group_ctx_t &) const {}

	__device__ uint32_t get_num_group(const uint8_t *, uint32_t, uint32_t num_child) const { return num_child; }

	// optional, for rules with warp_groups: ORDERING of the groups.  group_keys() gives every group of a
	// parent a 32-bit key such that groups producing the SAME set of objects (from different parents)
	// have equal keys.  The engine then sorts all (parent, group) work items by key and hands them to
	// the warps in that order with symbolic_warp<true>: the rule may keep the magnitudes of a run of
	// equal-target groups in its per-warp workspace and emit each object once per run (flush_warp)
	// instead of once per child -- the global interference table then sees ~N_u inserts, not N_c.
	// Purely a performance device: any keys give correct results.
	//     init_warp(workspace&)                       once per warp, before the first group
	//     symbolic_warp<ACCUMULATE>(...)              ACCUMULATE = true only in sorted order
	//     flush_warp(workspace&, emit)                emit whatever the workspace still holds
	// optional, for rules whose children have the parent's size and differ from it in a few bytes:
	//     edit_child(parent, parent_size, child, child_id)   child already holds a copy of the parent
	// The finalisation then copies parents to children itself, a warp at a time with coalesced wide
	// loads and several copies in flight, and calls edit_child from one lane per child.
	static constexpr bool has_edit_child = false;
	__device__ void edit_child(const uint8_t *, uint32_t, uint8_t *, uint32_t) const {}

	// optional, with has_group_key: items that CONTINUE the run of the item before them are handed to the rule
	// one per lane instead of one per warp-wide call:
	//     run_id_t run_identity(ctx, group)            equal identities = same objects; compared with ==
	//     continue_run(ctx, group_ctx, workspace&)     this lane's group joins the open run (shared-memory atomics)
	// The head of every stretch of equal identities still goes through symbolic_warp<true>.
	static constexpr bool has_run_identity = false;
	// optional: bool group_keys_from_ctx(ctx, num_groups, keys) -- the keys from the prepared context alone; false =
	// not for this parent, group_keys() is called with the object
	static constexpr bool has_group_keys_from_ctx = false;
	// optional, with has_run_identity: the rule can send a run to a REGION of the table (table.cuh: region_acquire) when
	// every object of the state is smaller than this many bytes (0 = the rule does not use regions)
	static constexpr uint32_t region_size_limit = 0;
	// optional, with regions: BATCH mode for states whose runs are short (the host picks it from the run lengths it sees).  The
	// rule handles 32 consecutive items of the sorted order at once, one lane per item for the bookkeeping:
	//     cplx root_magnitude(ctx, group, parent_mag)                 one lane per item
	//     region_batch(ctx[32], root[32], child_begin[32], size[32], group[32], count, workspace&, table, created&, regions&)   all lanes
	static constexpr bool has_region_batch = false;

	// optional, distributed path: a FAMILY is a set of objects closed under the rule -- every child of a member is a member --
	// so that objects of different families never interfere.  family_key(parent, size) gives equal keys to the members of a
	// family (a hash of what the rule leaves untouched).  quids::mpi::simulate then moves every PARENT to the rank that owns
	// its family (parents are a few hundred bytes, their children thousands of records) and interference needs no exchange at
	// all (route.inc.cuh).  Only used when every object of the state is smaller than region_size_limit.
	static constexpr bool has_family = false;
	__device__ uint64_t family_key(const uint8_t *, uint32_t) const { return 0; }

	static constexpr bool has_group_key = false;
	static constexpr uint32_t group_capacity = 1; // most children one group can hold (bounds what a run can send to the table)
	__device__ void group_keys(const uint8_t *, uint32_t, uint32_t, uint32_t *) const {}
	__device__ void init_warp(workspace_t &) const {}
	template <class Emit>
	__device__ void flush_warp(workspace_t &, Emit &) const {}
	// optional: called once per warp when the sorted-order kernel ends (after the last flush_warp)
	template <class WS, class Emit>
	__device__ void finish_warp(WS &, Emit &) const {}
	// optional, HOST side: called before the symbolic kernels of this rule type are launched on `stream` (the device is
	// current): device-resident tables the rule's kernels read, built once per device, ordered before the launch by the stream
	static void prepare_device(cudaStream_t) {}
};

// ---- modifiers: f(begin, end, mag&) in place (quids.hpp:86,973-980) as a device functor
//     __device__ void operator()(uint8_t *object, uint32_t size, cplx &mag) const

// ---- views handed to the kernels ----------------------------------------------------------------
struct iter_view {
	uint8_t *objects;
	const uint64_t *begin;
	const uint32_t *size;
	cplx *mag;
	uint64_t n;
};

struct table_view;
struct engine_launch; // defined in engine.cuh

// what the engine needs from one rule type (filled by the template in engine.cuh)
struct rule_ops {
	const char *name;
	// builds the device rule (<= 256 bytes, trivially copyable) from the ABI parameter doubles
	int (*make)(const double *params, uint32_t num_params, void *rule_storage);
	void (*launch_num_child)(const void *rule, const engine_launch &L);
	void (*launch_symbolic)(const void *rule, const engine_launch &L);
	void (*launch_populate)(const void *rule, const engine_launch &L);
	void (*launch_hash)(const void *rule, const engine_launch &L);
	bool needs_scratch;
	bool warp_groups;
	bool has_groups; // warp_groups or lane_groups: the groups are a second index space (num_groups, group_begin)
	bool has_group_key;
	uint32_t region_size_limit;
	uint32_t group_capacity;
	size_t ctx_bytes;
	void (*launch_group_items)(const void *rule, const engine_launch &L);
	void (*launch_symbolic_items)(const void *rule, const engine_launch &L);
	bool has_region_batch;
	void (*launch_symbolic_items_batch)(const void *rule, const engine_launch &L); // region mode, short runs
	bool has_family;
	void (*launch_family)(const void *rule, const engine_launch &L); // L.hashes[i] = family_key of object i
	int (*symbolic_grid)(int sm_count); // CTAs the symbolic kernel is launched with at most (sizes the scratch)
	uint64_t (*symbolic_chunks)(uint64_t n_groups); // work chunks of the symbolic kernel (sizes chunk_parent)
};

struct modifier_ops {
	const char *name;
	int (*make)(const double *params, uint32_t num_params, void *storage);
	void (*launch)(const void *modifier, const iter_view &it, cudaStream_t stream, int sm_count);
};

// ---- observables: f(begin, end) -> value, averaged with weights |mag|^2 (iteration::average_value, quids.hpp:208-234)
// as a device functor producing `values` numbers per object in one pass:
//     static constexpr int values;   __device__ void operator()(const uint8_t *object, uint32_t size, double *out) const
constexpr int OBSERVABLE_MAX_VALUES = 4;
struct observable_ops {
	const char *name;
	int values;
	int (*make)(const double *params, uint32_t num_params, void *storage);
	// partial[k * grid + block] = this CTA's share of value k; returns the grid used
	int (*launch)(const void *observable, const iter_view &it, double *partial, cudaStream_t stream, int sm_count);
};

constexpr size_t RULE_STORAGE_BYTES = 256;

int register_rule(const rule_ops &ops);
int register_modifier(const modifier_ops &ops);
int register_observable(const observable_ops &ops);
const observable_ops *find_observable(int id);
const rule_ops *find_rule(int id);
const modifier_ops *find_modifier(int id);

} // namespace qb
