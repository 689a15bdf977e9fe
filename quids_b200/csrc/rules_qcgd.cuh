// rules_qcgd.cuh -- device implementations of the Quantum Causal Graph Dynamics rules
// (reference: src/rules/qcgd.hpp).
//
// Object layout (qcgd.hpp:63-112), n = number of nodes:
//     u16 n | u8 left[n] | u8 right[n] | u16 name_begin[n+1] | atom names[name_begin[n]]
//     atom (the reference's sub_node, 16 bytes) = { i16 hmlz, i16 kind, 4 bytes padding, u64 hash }
//     kind: -3 = ".l" wrapper, -2 = ".r" wrapper, -1 = element, >= 0 = pair, value = offset to the
//     right subtree; a node's name is the prefix-order tree names[name_begin[i] .. name_begin[i+1])
// Atoms start at byte 4+4n, so their u64 hash is only 4-byte aligned: it is read as two u32.
// Objects must start 4-byte aligned (true for any align_byte_length that is a multiple of 4, and
// for align 0/1 as long as the state only holds QCGD objects, whose sizes are multiples of 4).
//
// Unlike the reference, which edits a copy of the parent through pointer accessors, the rules here
// are written as a WALK over the parent that announces the child's nodes in order to an emitter:
// one emitter folds the child's hash and size without writing anything (symbolic phase), another
// writes the child's bytes (finalisation).
#pragma once

#include <quids/device/rule_api.cuh>

namespace qb {
namespace qcgd {

enum : int { DOT_L = -3, DOT_R = -2, ELEMENT = -1 };

// qcgd.hpp:11-25
__host__ __device__ __forceinline__ uint64_t hash_combine(uint64_t seed, uint64_t v) {
	seed *= MURMUR_MUL;
	seed ^= v >> 47;
	seed *= MURMUR_MUL;
	seed ^= v;
	seed *= MURMUR_MUL;
	return seed + 0xe6546b64ull;
}

// same function for a value below 2^47 (a node index): `seed ^= v >> 47` is then a no-op and the
// first two multiplications merge into one by MURMUR_MUL^2 (mod 2^64)
__host__ __device__ __forceinline__ uint64_t hash_combine_index(uint64_t seed, uint32_t index) {
	constexpr uint64_t MUL2 = MURMUR_MUL * MURMUR_MUL;
	return ((seed * MUL2) ^ (uint64_t)index) * MURMUR_MUL + 0xe6546b64ull;
}

// FOLD TABLE.  hash_graph's two particle hashes are folds of hash_combine over the INDICES of the nodes that hold a particle,
// in increasing order: a function of the particle mask alone.  For graphs of at most FOLD_TABLE_BITS nodes the fold of every
// mask is tabulated once per device (512 KB, L2 resident): the hash of a child whose particle masks are known costs two
// loads and the two final folds instead of one dependent chain of 64-bit multiplications per node.
constexpr uint32_t FOLD_TABLE_BITS = 16;
static __device__ uint64_t g_particle_fold[1u << FOLD_TABLE_BITS];

static __global__ void __launch_bounds__(256) particle_fold_init_kernel() {
	const uint32_t mask = blockIdx.x * blockDim.x + threadIdx.x;
	uint64_t h = 0;
	for (uint32_t i = 0; i < FOLD_TABLE_BITS; ++i)
		if ((mask >> i) & 1)
			h = hash_combine_index(h, i);
	g_particle_fold[mask] = h;
}

// host: the table of the current device is filled by a kernel on `stream` the first time a rule that reads it is launched there
inline void particle_fold_prepare(cudaStream_t stream) {
	static bool ready[64] = {};
	int device = 0;
	QB_CUDA(cudaGetDevice(&device));
	if (ready[device & 63])
		return;
	particle_fold_init_kernel<<<(1u << FOLD_TABLE_BITS) / 256, 256, 0, stream>>>();
	QB_CUDA(cudaGetLastError());
	ready[device & 63] = true;
}

__device__ __forceinline__ uint64_t particle_fold(uint32_t mask) { return __ldg(&g_particle_fold[mask]); }
// any mask of up to 64 nodes: the nodes the table covers are looked up, the others folded one by one
__device__ __forceinline__ uint64_t particle_fold_wide(uint64_t mask) {
	uint64_t h = particle_fold((uint32_t)mask & ((1u << FOLD_TABLE_BITS) - 1));
	for (uint64_t m = mask >> FOLD_TABLE_BITS; m; m &= m - 1)
		h = hash_combine_index(h, FOLD_TABLE_BITS + (uint32_t)__ffsll((long long)m) - 1);
	return h;
}

struct atom {
	int hmlz; // "has most-left zero" flag / element + 1   (qcgd.hpp:36,40-42)
	int kind;
	uint64_t hash;
};

// read-only view of one object
struct graph {
	const uint8_t *p;
	uint32_t n;

	__device__ explicit graph(const uint8_t *p_) : p(p_), n(*reinterpret_cast<const uint16_t *>(p_)) {}
	__device__ bool left(uint32_t i) const { return p[2 + i]; }
	__device__ bool right(uint32_t i) const { return p[2 + n + i]; }
	__device__ uint32_t name_begin(uint32_t i) const { return reinterpret_cast<const uint16_t *>(p + 2 + 2 * n)[i]; }
	__device__ const uint32_t *atom_words(uint32_t k) const { return reinterpret_cast<const uint32_t *>(p + 4 + 4 * n) + 4 * (size_t)k; }
	__device__ atom get(uint32_t k) const {
		const uint32_t *w = atom_words(k);
		atom a;
		a.hmlz = (int16_t)(w[0] & 0xffff);
		a.kind = (int16_t)(w[0] >> 16);
		a.hash = (uint64_t)w[2] | ((uint64_t)w[3] << 32);
		return a;
	}
	__device__ uint64_t atom_hash(uint32_t k) const {
		const uint32_t *w = atom_words(k);
		return (uint64_t)w[2] | ((uint64_t)w[3] << 32);
	}
	// split site / merge site at node i (qcgd.hpp:174-179)
	__device__ void site(uint32_t i, bool &split, bool &merge) const {
		split = left(i) && right(i);
		merge = !split && i + 1 < n && left(i) && right(i + 1) && !left(i + 1);
	}
};

// the 2 n particle bytes (left[n] then right[n], each 0 or 1) of a graph of at most 32 nodes as one 64-bit mask: bit i =
// left[i], bit n + i = right[i].  Aligned 32-bit loads + funnel shifts instead of one byte load per particle, then four
// bytes become four bits with one multiplication ((b0 | b1<<8 | b2<<16 | b3<<24) * 0x00204081 has b0..b3 at bits 21..24).
__device__ __forceinline__ uint64_t particle_mask(const uint8_t *object, uint32_t n) {
	const uintptr_t first = reinterpret_cast<uintptr_t>(object) + 2;
	const uint32_t *w = reinterpret_cast<const uint32_t *>(first & ~(uintptr_t)3);
	const uint32_t shift = (uint32_t)(first & 3) * 8;
	const uint32_t words = (2 * n + 3) / 4;
	uint64_t mask = 0;
	uint32_t lo = w[0];
	for (uint32_t k = 0; k < words; ++k) {
		const uint32_t hi = w[k + 1]; // at most 3 bytes past the particles: still inside the object (name_begin follows)
		const uint32_t x = __funnelshift_r(lo, hi, shift);
		mask |= (uint64_t)(((x & 0x01010101u) * 0x00204081u >> 21) & 0xfu) << (4 * k);
		lo = hi;
	}
	return 2 * n < 64 ? mask & ((1ull << (2 * n)) - 1) : mask;
}

// qcgd.hpp:122-146
__device__ inline uint64_t hash_graph(const uint8_t *object) {
	graph g(object);
	uint64_t hl = 0, hr = 0, hn = 0;
	for (uint32_t i = 0; i < g.n; ++i) {
		if (g.left(i))
			hl = hash_combine(hl, i);
		if (g.right(i))
			hr = hash_combine(hr, i);
		hn = hash_combine(hn, g.atom_hash(g.name_begin(i)));
	}
	return hash_combine(hash_combine(hn, hl), hr);
}

// the four amplitudes of a binary choice (qcgd.hpp:466-471, used :501-506, :575-580, :686-696):
//   index = taken * 2 + conjugated:  stay, -conj(stay), go, conj(go)
struct amplitudes {
	cplx f[4];
	__host__ void set(double theta, double phi, double xi) {
		std::complex<double> go = std::polar(std::sin(theta), phi), stay = std::polar(std::cos(theta), xi);
		f[0] = cplx{stay.real(), stay.imag()};
		f[1] = cplx{-stay.real(), stay.imag()};
		f[2] = cplx{go.real(), go.imag()};
		f[3] = cplx{go.real(), -go.imag()};
	}
	// selects instead of an indexed load: the rule lives in the kernel parameters (constant bank), a
	// dynamic index would force a local-memory copy of the table
	__device__ cplx get(bool taken, bool conjugated) const {
		const cplx a = conjugated ? f[1] : f[0], b = conjugated ? f[3] : f[2];
		return taken ? b : a;
	}
};

// ===================================================================================================
// erase_create (qcgd.hpp:459-532) and coin (qcgd.hpp:534-605): every ELIGIBLE node consumes one bit
// of child_id, a set bit toggles both particles of the node.  erase_create: eligible = left == right;
// coin: eligible = left != right.  Size unchanged, names untouched.  In both rules the amplitude is
// conjugated exactly when the left particle is present.
// ===================================================================================================
constexpr int FLIP_LEVELS = 7;               // tree levels one warp expands in shared memory
constexpr int FLIP_BLOCK = 1 << FLIP_LEVELS; // children per group (at most)

struct flip_ctx {
	uint64_t left, right; // particle masks (only when n <= 64)
	uint64_t names_hash;  // fold of the first-atom hashes: the same for every child
	uint32_t n;
	uint32_t prefix_bits;         // left-particle bits of the first k - levels eligible nodes (the ones the group index decides)
	uint8_t eligible;             // number of eligible nodes k: the parent has 2^k children
	uint8_t levels;               // min(k, FLIP_LEVELS): tree levels of one group
	uint8_t tree_bits;            // left-particle bits of the tree nodes (bit l = level l)
	uint8_t pos[FLIP_LEVELS + 1]; // node index of the last `levels` eligible nodes, then n
};

struct flip_root { // state before the first tree level of one group
	uint64_t hl, hr;
	cplx mag;
};

// per-warp shared memory.  Tree states: after level l, state i (i < 2^l) = fold of hash_graph's left /
// right hashes and the magnitude product over every node before the next tree node, for choice bits i.
// Accumulators: the objects of the current run of equal-target groups, indexed by the TARGET's particle
// bits on the tree nodes (leaf index xor the parent's own bits), so that every member of a family adds
// into the same slots.
template <bool WITH_MAG_TREE>
struct flip_workspace_t {
	uint64_t hl[FLIP_BLOCK], hr[FLIP_BLOCK]; // tree states; once a run is open: hash and representative of its objects
	double re[WITH_MAG_TREE ? FLIP_BLOCK : 1], im[WITH_MAG_TREE ? FLIP_BLOCK : 1]; // magnitude tree (unsorted order only)
	double acc_re[FLIP_BLOCK], acc_im[FLIP_BLOCK];
	// what the accumulators hold: family (eligible nodes, particles elsewhere, names) and target of the group
	uint64_t run_eligible, run_fixed, run_names;
	uint32_t run_n, run_target, run_leaves, run_valid;
	uint32_t run_patterns[FLIP_BLOCK / 32]; // which parent patterns (accumulator indices) have received a root in this run
	cplx amp[4]; // the rule's four amplitudes, index = taken * 2 + conjugated: one 16-byte shared load per factor
	// region mode: the group that opened the run, kept so that the run can still write the objects' hashes and
	// representatives if it turns out to be the one that creates their region
	flip_ctx open_ctx;
	flip_root open_root;
	uint64_t open_first_child;
	uint32_t open_group, open_size;
	region_chunk chunk; // this warp's private range of table slots
};

template <bool WANT_EQUAL>
struct flip_rule : rule_base<flip_rule<WANT_EQUAL>> {
	amplitudes amp;

	__device__ uint64_t hasher(const uint8_t *object, uint32_t) const { return hash_graph(object); }

	__device__ void get_num_child(const uint8_t *parent, uint32_t parent_size, uint32_t &num_child, uint32_t &max_child_size) const {
		graph g(parent);
		max_child_size = parent_size;
		uint32_t eligible = 0;
		if (g.n >= 1 && g.n <= 32 && (reinterpret_cast<uintptr_t>(parent) & 1) == 0) { // a few word loads instead of one byte load per particle
			const uint64_t both = particle_mask(parent, g.n), all = (1ull << g.n) - 1;
			const uint64_t l = both & all, r = both >> g.n;
			eligible = __popcll((WANT_EQUAL ? ~(l ^ r) : (l ^ r)) & all);
		} else {
			for (uint32_t i = 0; i < g.n; ++i)
				eligible += (g.left(i) == g.right(i)) == WANT_EQUAL;
		}
		num_child = 1u << eligible;
	}

	__device__ void populate_child(const uint8_t *parent, uint32_t parent_size, uint8_t *child, uint32_t child_id, uint32_t &size, cplx &mag) const {
		graph g(parent);
		size = parent_size;
		const uint32_t *src = reinterpret_cast<const uint32_t *>(parent);
		uint32_t *dst = reinterpret_cast<uint32_t *>(child);
		for (uint32_t w = 0; w < parent_size / 4; ++w)
			dst[w] = src[w];
		for (uint32_t i = 0; i < g.n; ++i) {
			bool l = g.left(i), r = g.right(i);
			if ((l == r) != WANT_EQUAL)
				continue;
			bool taken = child_id & 1;
			child_id >>= 1;
			mag = cmul(mag, amp.get(taken, l));
			if (taken) {
				child[2 + i] = !l;
				child[2 + g.n + i] = !r;
			}
		}
	}
};

template <bool WANT_EQUAL>
struct flip_rule_fused : flip_rule<WANT_EQUAL> {
	typedef flip_ctx ctx_t;
	typedef flip_workspace_t<true> workspace_t;        // unsorted order: hash and magnitude trees
	typedef flip_workspace_t<false> items_workspace_t; // sorted order: hash tree and accumulators
	typedef flip_root group_ctx_t;
	static constexpr bool needs_scratch = false;
	static constexpr bool warp_groups = true;
	static constexpr bool has_group_key = true;
	static constexpr uint32_t group_capacity = FLIP_BLOCK;
	static constexpr bool has_edit_child = true;
	static constexpr bool has_run_identity = true;
	// sorted order with table regions needs every parent to go through the mask-based path (n <= 64 nodes); an object
	// of fewer bytes than the smallest 65-node graph (4 + 4 n + 16 per name atom) cannot have more nodes
	static constexpr uint32_t region_size_limit = 4 + 20 * 65;
	static void prepare_device(cudaStream_t stream) { particle_fold_prepare(stream); }

	// what makes two groups produce the same objects: family (eligible nodes, particles elsewhere, names, size) and target
	struct run_id_t {
		uint64_t eligible, fixed, names;
		uint32_t n, target;
		__device__ bool operator==(const run_id_t &o) const {
			return eligible == o.eligible && fixed == o.fixed && names == o.names && n == o.n && target == o.target;
		}
		__device__ run_id_t shuffle_up() const {
			run_id_t r;
			r.eligible = __shfl_up_sync(0xffffffffu, eligible, 1);
			r.fixed = __shfl_up_sync(0xffffffffu, fixed, 1);
			r.names = __shfl_up_sync(0xffffffffu, names, 1);
			r.n = __shfl_up_sync(0xffffffffu, n, 1);
			r.target = __shfl_up_sync(0xffffffffu, target, 1);
			return r;
		}
	};
	__device__ run_id_t run_identity(const flip_ctx &ctx, uint32_t group) const {
		run_id_t id;
		const uint64_t all = ctx.n >= 64 ? ~0ull : ((1ull << ctx.n) - 1);
		id.eligible = (WANT_EQUAL ? ~(ctx.left ^ ctx.right) : (ctx.left ^ ctx.right)) & all;
		id.fixed = ctx.left & ~id.eligible & all;
		id.names = ctx.names_hash;
		id.n = ctx.n;
		id.target = group ^ ctx.prefix_bits;
		if (ctx.n > 64) // wide graphs have no runs: make every item its own stretch
			id.target = group, id.eligible = ~0ull, id.fixed = (uint64_t)lane_id();
		return id;
	}
	// FAMILY (rule_api.cuh): a child keeps its parent's node count, eligible nodes, particles on the other nodes and names, so
	// all of that is an invariant of the rule: parents that differ in any of it have no child in common.  (Graphs of more
	// than 64 nodes have no masks: the caller only routes by family when every object is below region_size_limit.)
	static constexpr bool has_family = true;
	__device__ uint64_t family_key(const uint8_t *parent, uint32_t parent_size) const {
		flip_ctx ctx;
		prepare(parent, parent_size, ctx);
		const uint64_t all = ctx.n >= 64 ? ~0ull : ((1ull << ctx.n) - 1);
		const uint64_t eligible = (WANT_EQUAL ? ~(ctx.left ^ ctx.right) : (ctx.left ^ ctx.right)) & all;
		const uint64_t fixed = ctx.left & ~eligible & all;
		return mix64(eligible ^ mix64(fixed + 0x9e3779b97f4a7c15ull * (ctx.n + 1ull)) ^ mix64(ctx.names_hash ^ 0xc2b2ae3d27d4eb4full));
	}

	// this lane's group has the identity of the run that is open: its root joins the sum of its parent pattern
	template <class WS>
	__device__ void continue_run(const flip_ctx &ctx, const flip_root &root, WS &ws) const {
		atomicOr(&ws.run_patterns[ctx.tree_bits >> 5], 1u << (ctx.tree_bits & 31));
		atomicAdd(&ws.acc_re[ctx.tree_bits], root.mag.re);
		atomicAdd(&ws.acc_im[ctx.tree_bits], root.mag.im);
	}

	// a child is its parent with some eligible nodes toggled
	__device__ void edit_child(const uint8_t *parent, uint32_t, uint8_t *child, uint32_t child_id) const {
		graph g(parent);
		for (uint32_t i = 0; i < g.n; ++i) {
			const bool l = g.left(i), r = g.right(i);
			if ((l == r) != WANT_EQUAL)
				continue;
			if (child_id & 1) {
				child[2 + i] = !l;
				child[2 + g.n + i] = !r;
			}
			child_id >>= 1;
		}
	}

	// The 2^k children of a parent are the leaves of a binary tree over its k eligible nodes, and
	// hash_graph folds the nodes in index order: two children share the fold (and the magnitude
	// product) up to the first eligible node where their choices differ.  A warp therefore produces
	// children in groups of up to 2^FLIP_LEVELS: the group index fixes the choices of the FIRST
	// k - levels eligible nodes (the low bits of child_id); the tree over the last `levels` ones is
	// expanded level by level in shared memory, one lane per tree node.  ~2 fold steps per child
	// instead of n, and all lanes run the same control flow because they share the parent.
	//
	// FAMILIES.  A child keeps its parent's eligible nodes, its particles on the other nodes and its
	// names, so two parents that agree on those produce the SAME 2^k objects.  A group's objects are
	// those whose particles on the first k - levels eligible nodes equal  target = group xor the
	// parent's own bits there.  group_keys() hashes (family, target); in sorted order the warp meets
	// all groups with the same objects in a row, adds their magnitudes in shared memory (only the
	// magnitude tree is expanded after the first group: the hashes are known), and sends each object
	// to the global interference table once per run.
	// Graphs wider than the 64-bit masks: groups of 32 children, one per lane, walking the bytes.
	__device__ uint32_t get_num_group(const uint8_t *parent, uint32_t, uint32_t num_child) const {
		const uint32_t n = *reinterpret_cast<const uint16_t *>(parent);
		if (n > 64)
			return (num_child + 31) / 32;
		return num_child > (uint32_t)FLIP_BLOCK ? num_child >> FLIP_LEVELS : 1;
	}

	// the same keys from a prepared context (n <= 64: the masks say everything group_keys reads from the object)
	static constexpr bool has_group_keys_from_ctx = true;
	__device__ bool group_keys_from_ctx(const flip_ctx &ctx, uint32_t num_groups, uint32_t *keys) const {
		if (ctx.n > 64)
			return false;
		const uint64_t all = ctx.n == 64 ? ~0ull : ((1ull << ctx.n) - 1);
		const uint64_t eligible = (WANT_EQUAL ? ~(ctx.left ^ ctx.right) : (ctx.left ^ ctx.right)) & all;
		const uint64_t family = mix64(mix64(eligible + 0x9e3779b97f4a7c15ull * ctx.n) ^ (ctx.left & ~eligible));
		for (uint32_t group = 0; group < num_groups; ++group)
			keys[group] = (uint32_t)(mix64(family + 0x9e3779b97f4a7c15ull * ((group ^ ctx.prefix_bits) + 1)) >> 32);
		return true;
	}

	__device__ void group_keys(const uint8_t *parent, uint32_t, uint32_t num_groups, uint32_t *keys) const {
		graph g(parent);
		uint64_t family = g.n;
		uint32_t eligible = 0, prefix_bits = 0;
		for (uint32_t i = 0; i < g.n; ++i) { // pass 1: the family, and k
			const bool l = g.left(i), r = g.right(i);
			const bool is_eligible = (l == r) == WANT_EQUAL;
			eligible += is_eligible;
			family = family * 3 + (is_eligible ? 2 : (l ? 1 : 0));
			if ((i & 31) == 31)
				family = mix64(family);
		}
		family = mix64(family);
		if (g.n <= 64) { // the parent's own bits on the nodes the group index decides
			const uint32_t prefix = eligible > (uint32_t)FLIP_LEVELS ? eligible - FLIP_LEVELS : 0;
			uint32_t seen = 0;
			for (uint32_t i = 0; i < g.n && seen < prefix; ++i)
				if ((g.left(i) == g.right(i)) == WANT_EQUAL) {
					prefix_bits |= (uint32_t)g.left(i) << seen;
					++seen;
				}
		}
		for (uint32_t group = 0; group < num_groups; ++group)
			keys[group] = (uint32_t)(mix64(family + 0x9e3779b97f4a7c15ull * ((group ^ prefix_bits) + 1)) >> 32);
	}

	__device__ void prepare(const uint8_t *parent, uint32_t, flip_ctx &ctx) const {
		graph g(parent);
		ctx.n = g.n;
		if (g.n >= 1 && g.n <= 32) { // masks with a few word loads, everything else with bit operations
			const uint64_t both = particle_mask(parent, g.n);
			const uint64_t all = (1ull << g.n) - 1;
			const uint64_t l = both & all, r = both >> g.n;
			uint64_t hn = 0;
			for (uint32_t i = 0; i < g.n; ++i)
				hn = hash_combine(hn, g.atom_hash(g.name_begin(i)));
			uint64_t elig = (WANT_EQUAL ? ~(l ^ r) : (l ^ r)) & all;
			const uint32_t eligible = __popcll(elig), levels = min(eligible, (uint32_t)FLIP_LEVELS);
			ctx.left = l;
			ctx.right = r;
			ctx.names_hash = hn;
			ctx.eligible = (uint8_t)eligible;
			ctx.levels = (uint8_t)levels;
			uint32_t prefix_bits = 0, tree_bits = 0;
			for (uint32_t seen = 0; elig; ++seen) { // eligible nodes in index order: the first eligible - levels are the prefix
				const uint32_t i = __ffsll((long long)elig) - 1;
				elig &= elig - 1;
				const uint32_t bit = (uint32_t)((l >> i) & 1);
				if (seen + levels >= eligible) {
					ctx.pos[seen + levels - eligible] = (uint8_t)i;
					tree_bits |= bit << (seen + levels - eligible);
				} else {
					prefix_bits |= bit << seen;
				}
			}
			ctx.pos[levels] = (uint8_t)g.n;
			ctx.prefix_bits = prefix_bits;
			ctx.tree_bits = (uint8_t)tree_bits;
			return;
		}
		uint64_t l = 0, r = 0, hn = 0;
		uint32_t eligible = 0;
		for (uint32_t i = 0; i < g.n; ++i) {
			const bool li = g.left(i), ri = g.right(i);
			if (i < 64) {
				l |= (uint64_t)li << i;
				r |= (uint64_t)ri << i;
			}
			eligible += (li == ri) == WANT_EQUAL;
			hn = hash_combine(hn, g.atom_hash(g.name_begin(i)));
		}
		ctx.left = l;
		ctx.right = r;
		ctx.names_hash = hn;
		ctx.eligible = (uint8_t)eligible;
		const uint32_t levels = min(eligible, (uint32_t)FLIP_LEVELS);
		ctx.levels = (uint8_t)levels;
		uint32_t prefix_bits = 0, tree_bits = 0;
		if (g.n <= 64) {
			uint32_t seen = 0;
			for (uint32_t i = 0; i < g.n; ++i)
				if ((((l ^ r) >> i) & 1) != (uint64_t)WANT_EQUAL) {
					const uint32_t bit = (uint32_t)((l >> i) & 1);
					if (seen + levels >= eligible) {
						ctx.pos[seen + levels - eligible] = (uint8_t)i;
						tree_bits |= bit << (seen + levels - eligible);
					} else {
						prefix_bits |= bit << seen;
					}
					++seen;
				}
			ctx.pos[levels] = (uint8_t)g.n;
		}
		ctx.prefix_bits = prefix_bits;
		ctx.tree_bits = (uint8_t)tree_bits;
	}

	// one child on its own (also the symbolic hook of the one-child-per-lane path)
	__device__ uint64_t symbolic(const uint8_t *parent, uint32_t parent_size, const flip_ctx &, uint32_t child_id, uint8_t *, uint32_t &size,
	                             cplx &mag) const {
		size = parent_size;
		graph g(parent);
		uint64_t hl = 0, hr = 0, hn = 0;
		for (uint32_t i = 0; i < g.n; ++i) {
			bool l = g.left(i), r = g.right(i);
			if ((l == r) == WANT_EQUAL) {
				const bool taken = child_id & 1;
				child_id >>= 1;
				mag = cmul(mag, this->amp.get(taken, l));
				l ^= taken;
				r ^= taken;
			}
			if (l) hl = hash_combine_index(hl, i);
			if (r) hr = hash_combine_index(hr, i);
			hn = hash_combine(hn, g.atom_hash(g.name_begin(i)));
		}
		return hash_combine(hash_combine(hn, hl), hr);
	}

	// root of a group: everything before the first tree node, choices taken from the group index
	__device__ void prepare_group(const flip_ctx &ctx, uint32_t group, cplx mag, flip_root &root) const {
		uint64_t hl = 0, hr = 0;
		if (ctx.n <= 64) {
			const uint64_t left = ctx.left, right = ctx.right;
			const uint64_t eligible = WANT_EQUAL ? ~(left ^ right) : (left ^ right);
			uint32_t bits = group;
			const uint32_t first = ctx.pos[0];
			for (uint32_t i = 0; i < first; ++i) {
				bool l = (left >> i) & 1, r = (right >> i) & 1;
				if ((eligible >> i) & 1) {
					const bool taken = bits & 1;
					bits >>= 1;
					mag = cmul(mag, this->amp.get(taken, l));
					l ^= taken;
					r ^= taken;
				}
				if (l) hl = hash_combine_index(hl, i);
				if (r) hr = hash_combine_index(hr, i);
			}
		}
		root.hl = hl;
		root.hr = hr;
		root.mag = mag;
	}

	template <class WS>
	__device__ void init_warp(WS &ws) const {
		if (lane_id() == 0) {
			ws.run_valid = 0;
			ws.chunk.next = ws.chunk.end = 0;
		}
		if (lane_id() < 4)
			ws.amp[lane_id()] = this->amp.f[lane_id()];
	}

	// A run's accumulators hold, per pattern t of the parents' own particles on the tree nodes, the SUM of the roots of
	// the groups with that pattern (a continuing group costs one addition).  The magnitude a group gives to object s is
	// root * prod_l amp[taken_l][t_l] with taken_l = s_l xor t_l, so the objects' magnitudes are
	//     acc[s] = sum_t R[t] * prod_l A[s_l xor t_l][t_l]
	// -- a Kronecker product of one 2x2 matrix per tree level, applied in place level by level (64 butterflies each),
	// instead of one product chain per child.  Same value as summing the children one by one up to rounding
	// (the interference table adds them in no particular order either).
	template <class WS>
	__device__ __forceinline__ void spread_run(WS &ws) const {
		const uint32_t lane = lane_id();
		const uint32_t leaves = ws.run_leaves;
		const cplx a00 = ws.amp[0], a01 = ws.amp[1], a10 = ws.amp[2], a11 = ws.amp[3]; // index = taken * 2 + parent's bit
		// ONE parent pattern in the run (the usual case once a state has grown: about one group per region): the objects'
		// magnitudes are root * prod_l amp[s_l xor t_l][t_l], a binary tree of 2^levels - 1 complex products instead of the
		// levels * 2^(levels - 1) butterflies below (127 against 448 x 4 products for a full group).  Same values: a
		// butterfly with one zero input is the same rounded product plus an exact zero.
		uint32_t patterns = 0, t0 = 0;
#pragma unroll
		for (int w = 0; w < FLIP_BLOCK / 32; ++w) {
			const uint32_t m = ws.run_patterns[w];
			patterns += __popc(m);
			if (m)
				t0 = w * 32 + (__ffs(m) - 1);
		}
		if (patterns == 1) {
			if (lane == 0 && t0 != 0) {
				ws.acc_re[0] = ws.acc_re[t0];
				ws.acc_im[0] = ws.acc_im[t0];
				ws.acc_re[t0] = 0;
				ws.acc_im[t0] = 0;
			}
			__syncwarp();
			for (uint32_t bit = 1; bit < leaves; bit <<= 1) {
				const bool parent_bit = (t0 & bit) != 0;
				// object bit 0 / 1 on this node: taken = object bit xor parent's bit
				const cplx f0 = parent_bit ? a11 : a00, f1 = parent_bit ? a01 : a10;
				for (uint32_t i = lane; i < bit; i += 32) {
					const cplx m{ws.acc_re[i], ws.acc_im[i]};
					const cplx y0 = cmul(m, f0), y1 = cmul(m, f1);
					ws.acc_re[i] = y0.re;
					ws.acc_im[i] = y0.im;
					ws.acc_re[i + bit] = y1.re;
					ws.acc_im[i + bit] = y1.im;
				}
				__syncwarp();
			}
			return;
		}
		for (uint32_t bit = 1; bit < leaves; bit <<= 1) {
			for (uint32_t b = lane; b < leaves / 2; b += 32) {
				const uint32_t i0 = ((b & ~(bit - 1)) << 1) | (b & (bit - 1)), i1 = i0 | bit;
				const cplx x0{ws.acc_re[i0], ws.acc_im[i0]}, x1{ws.acc_re[i1], ws.acc_im[i1]};
				// products rounded like the reference's, then added: two contributions that cancel exactly there cancel exactly here
				const cplx y0 = cadd(cmul(x0, a00), cmul(x1, a11)); // object bit 0: pattern 0 stays, pattern 1 goes
				const cplx y1 = cadd(cmul(x0, a10), cmul(x1, a01)); // object bit 1: pattern 0 goes, pattern 1 stays
				ws.acc_re[i0] = y0.re;
				ws.acc_im[i0] = y0.im;
				ws.acc_re[i1] = y1.re;
				ws.acc_im[i1] = y1.im;
			}
			__syncwarp();
		}
	}

	// identity of a run's objects -> key of their region in the directory
	__device__ static uint64_t region_key(uint64_t eligible, uint64_t fixed, uint32_t target, uint64_t names, uint32_t n) {
		return mix64(eligible ^ mix64(fixed + 0x9e3779b97f4a7c15ull * (target + 1ull)) ^ mix64(names ^ (0xc2b2ae3d27d4eb4full * n)));
	}

	// ---- BATCH mode of the sorted order (engine.cuh, symbolic_items_batch_kernel): states whose runs are short (about one
	// group per region: a grown state).  The accumulating path above walks the runs one after the other, and every run pays
	// its own directory probe, its own publication and a dozen single-lane bookkeeping steps while 31 lanes wait.  Here a warp
	// takes 32 items at once:
	//   1. one lane per item: run identity; the HEAD of every stretch of equal identities probes the directory (all probes of
	//      the batch in flight together) and learns whether its stretch creates the region or adds to it;
	//   2. the regions created by the batch get consecutive slots with ONE addition to the table's cursor (exact: no chunks,
	//      no unused tails to zero);
	//   3. stretch by stretch, all lanes: lane l owns objects l, l + 32, ... of the region.  Magnitudes = one product chain per
	//      item of the stretch (same factors in the same order as the reference's child by child products), summed; a stretch
	//      of many items goes through the butterflies of spread_run instead.  Hashes of a created region = particle masks of the
	//      head's parent with the group's and the object's toggles -> fold table -> the two final folds.  Created regions are
	//      written whole (one full sector per object), the others receive RED.F64 additions;
	//   4. the creators publish their regions together (one release per lane, after all the slots of the batch).
	// Creators never wait for anybody, and a stretch that adds is handled after the batch's own creators are published: a
	// region created and extended inside one batch (equal 32-bit sort keys interleaved) cannot deadlock.
	// A run that crosses a batch boundary is simply two stretches: the second one adds to the region the first one created.
	static constexpr bool has_region_batch = true;
	template <class WS>
	__device__ __noinline__ void spread_run_out_of_line(WS &ws) const { spread_run(ws); }
	static constexpr uint32_t BATCH_MAX_CHAINS = 6; // stretches of more items use the butterflies

	// magnitude of a group's root: the parent's magnitude times the factors of the eligible nodes the group index decides
	__device__ __forceinline__ cplx root_magnitude(const flip_ctx &ctx, uint32_t group, cplx mag) const {
		const uint32_t prefix = ctx.eligible - ctx.levels;
		for (uint32_t b = 0; b < prefix; ++b)
			mag = cmul(mag, this->amp.get((group >> b) & 1, (ctx.prefix_bits >> b) & 1));
		return mag;
	}

	template <class WS>
	__device__ __forceinline__ void region_batch(const flip_ctx *ctx, const cplx *root, const uint64_t *child_begin, const uint32_t *size, const uint32_t *group,
	                                             uint32_t count, WS &ws, const table_view &table, uint32_t &created, uint32_t &regions) const {
		constexpr int PER_LANE = FLIP_BLOCK / 32;
		const uint32_t lane = lane_id();
		const bool valid = lane < count;
		// 1. runs inside the batch and their regions.  The items are sorted by 24 bits of a key: the items of one run are
		// close to each other but not always neighbours (another run with the same 24 bits may sit in between), so the members
		// of a run are found by matching the region keys of all 32 items, not by comparing neighbours
		uint64_t key = 0;
		uint32_t my_levels = 0;
		if (valid) {
			const run_id_t id = run_identity(ctx[lane], group[lane]);
			key = region_key(id.eligible, id.fixed, id.target, id.names, id.n);
			if (key == 0)
				key = 1;
			my_levels = ctx[lane].levels;
		}
		const unsigned in_batch = __ballot_sync(0xffffffffu, valid);
		const unsigned members = valid ? __match_any_sync(in_batch, (unsigned long long)key) : 0u; // the items of this lane's run
		const bool head = valid && (uint32_t)__ffs(members) - 1 == lane;
		region_entry *entry = nullptr;
		bool made = false, failed = false;
		if (head) {
			uint64_t i = __umul64hi(mix64(key), table.dir_capacity);
			unsigned long long seen = atomicCAS(&table.dir[i].key, 0ull, (unsigned long long)key);
			for (uint32_t probes = 0;; ++probes) {
				if (seen == 0) {
					made = true;
					break;
				}
				if (seen == key)
					break;
				if (probes > TABLE_MAX_PROBES || ((probes & 63) == 63 && table_overflowed_lane(table))) {
					*table.overflow = 1;
					failed = true;
					break;
				}
				if (++i == table.dir_capacity)
					i = 0;
				seen = __ldcg(&table.dir[i].key);
				if (seen == 0)
					seen = atomicCAS(&table.dir[i].key, 0ull, (unsigned long long)key);
			}
			entry = table.dir + i;
		}
		// 2. slots of the regions this batch creates
		const uint32_t want = head && made ? 1u << my_levels : 0u;
		uint32_t incl = want;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= (uint32_t)o)
				incl += up;
		}
		const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
		unsigned long long base = 0;
		if (total) {
			if (lane == 0)
				base = atomicAdd(table.cursor, (unsigned long long)total);
			base = __shfl_sync(0xffffffffu, base, 0);
			if (base + total > table.capacity) { // the table is too small: nothing of it may be touched, whoever waits for these regions gives up too
				if (lane == 0)
					*table.overflow = 1;
				if (head && made)
					atomicExch(&entry->base, ~0ull);
				failed = true;
			}
		}
		base += incl - want; // (heads that create)

		// 3. one stretch: magnitudes of the lane's objects, then the region's slots
		auto stretch = [&](uint32_t h, unsigned long long first_slot, bool creates) {
			const unsigned run = __shfl_sync(0xffffffffu, members, h); // the items of the run (h is the first)
			const flip_ctx &c = ctx[h];
			const uint32_t levels = c.levels, leaves = 1u << levels;
			const uint32_t tree_bits = c.tree_bits;
			// a new region: the particle masks of its objects are known before their magnitudes -- the lookups in the fold table
			// are issued first and their latency hides behind the product chains
			uint64_t hl[PER_LANE], hr[PER_LANE];
			if (creates) {
				const uint64_t left = c.left, right = c.right;
				const uint32_t g = group[h], shift = c.eligible - levels; // child_id = group | leaf << shift
				uint64_t toggles = 0;
				{ // the group index decides the first `shift` eligible nodes
					const uint64_t all = c.n >= 64 ? ~0ull : ((1ull << c.n) - 1);
					uint64_t el = (WANT_EQUAL ? ~(left ^ right) : (left ^ right)) & all;
					for (uint32_t b = 0; b < shift; ++b) {
						const uint32_t i = (uint32_t)__ffsll((long long)el) - 1;
						el &= el - 1;
						toggles |= (uint64_t)((g >> b) & 1) << i;
					}
				}
				{ // tree levels 0-4: the lane's bits of the object, xor the head parent's own
					const uint32_t leaf_low = (lane ^ tree_bits) & 31, low = levels < 5 ? levels : 5;
					for (uint32_t l = 0; l < low; ++l)
						toggles |= (uint64_t)((leaf_low >> l) & 1) << c.pos[l];
				}
				const bool narrow = c.n <= FOLD_TABLE_BITS;
#pragma unroll
				for (int q = 0; q < PER_LANE; ++q) {
					const uint32_t slot = lane + 32 * q;
					if (slot < leaves) {
						uint64_t t = toggles;
#pragma unroll
						for (int l = 5; l < FLIP_LEVELS; ++l)
							if ((uint32_t)l < levels)
								t |= (uint64_t)((((uint32_t)q >> (l - 5)) ^ (tree_bits >> l)) & 1) << c.pos[l];
						if (narrow) {
							hl[q] = particle_fold((uint32_t)(left ^ t));
							hr[q] = particle_fold((uint32_t)(right ^ t));
						} else {
							hl[q] = particle_fold_wide(left ^ t);
							hr[q] = particle_fold_wide(right ^ t);
						}
					}
				}
			}
			cplx mag[PER_LANE];
#pragma unroll
			for (int q = 0; q < PER_LANE; ++q)
				mag[q] = cplx{0, 0};
			if ((uint32_t)__popc(run) <= BATCH_MAX_CHAINS) {
				const uint32_t low = levels < 5 ? levels : 5;
				for (unsigned todo = run; todo; todo &= todo - 1) {
					const uint32_t j = (uint32_t)__ffs(todo) - 1;
					// amp index = taken * 2 + parent's bit, taken = object's bit xor parent's bit
					const uint32_t t = ctx[j].tree_bits;
					cplx m = root[j];
					for (uint32_t l = 0; l < low; ++l) {
						const uint32_t tb = (t >> l) & 1, sb = (lane >> l) & 1;
						m = cmul(m, ws.amp[((sb ^ tb) << 1) | tb]);
					}
#pragma unroll
					for (int q = 0; q < PER_LANE; ++q) {
						cplx mq = m;
#pragma unroll
						for (int l = 5; l < FLIP_LEVELS; ++l)
							if ((uint32_t)l < levels) {
								const uint32_t tb = (t >> l) & 1, sb = ((uint32_t)q >> (l - 5)) & 1;
								mq = cmul(mq, ws.amp[((sb ^ tb) << 1) | tb]);
							}
						mag[q] = j == h ? mq : cadd(mag[q], mq);
					}
				}
			} else {
				// many items: their roots are summed per parent pattern, then one Kronecker transform for all of them
				__syncwarp();
				for (uint32_t leaf = lane; leaf < leaves; leaf += 32)
					ws.acc_re[leaf] = ws.acc_im[leaf] = 0.0;
				if (lane < FLIP_BLOCK / 32)
					ws.run_patterns[lane] = 0;
				if (lane == 0)
					ws.run_leaves = leaves;
				__syncwarp();
				if ((run >> lane) & 1) {
					const uint32_t t = ctx[lane].tree_bits;
					atomicOr(&ws.run_patterns[t >> 5], 1u << (t & 31));
					atomicAdd(&ws.acc_re[t], root[lane].re);
					atomicAdd(&ws.acc_im[t], root[lane].im);
				}
				__syncwarp();
				spread_run_out_of_line(ws);
#pragma unroll
				for (int q = 0; q < PER_LANE; ++q) {
					const uint32_t slot = lane + 32 * q;
					if (slot < leaves)
						mag[q] = cplx{ws.acc_re[slot], ws.acc_im[slot]};
				}
				__syncwarp();
			}
			table_slot *slots = table.slots + first_slot;
			if (!creates) {
#pragma unroll
				for (int q = 0; q < PER_LANE; ++q) {
					const uint32_t slot = lane + 32 * q;
					if (slot < leaves) {
						atomicAdd(&slots[slot].re, mag[q].re); // results unused -> RED.ADD.F64 on consecutive sectors
						atomicAdd(&slots[slot].im, mag[q].im);
					}
				}
				return;
			}
			// the region is new: hash, summed magnitude and representative of every object, one full 32-byte sector each
			const uint32_t g = group[h], shift = c.eligible - levels;
			const uint64_t names_hash = c.names_hash, first_child = child_begin[h];
			const uint32_t bytes = size[h];
#pragma unroll
			for (int q = 0; q < PER_LANE; ++q) {
				const uint32_t slot = lane + 32 * q;
				if (slot < leaves) {
					const uint32_t leaf = slot ^ tree_bits;
					const unsigned long long key = hash_combine(hash_combine(names_hash, hl[q]), hr[q]);
					const unsigned long long rep = rep_pack(first_child + (g | (leaf << shift)), bytes);
					ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(slots + slot);
					dst[0] = make_ulonglong2(key, (unsigned long long)__double_as_longlong(mag[q].re));
					dst[1] = make_ulonglong2((unsigned long long)__double_as_longlong(mag[q].im), rep);
				}
			}
		};

		const bool creates = head && made && !failed;
		for (unsigned todo = __ballot_sync(0xffffffffu, creates); todo; todo &= todo - 1) {
			const uint32_t h = (uint32_t)__ffs(todo) - 1;
			stretch(h, __shfl_sync(0xffffffffu, base, h), true);
		}
		// 4. the batch's regions become visible to the other runs of the same objects
		__syncwarp();
		if (creates) {
			// st.release.gpu: the warp's slot writes (ordered before this lane by __syncwarp) become visible before the base does
			asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(&entry->base), "l"(base + 1) : "memory");
			created += want;
			regions += 1;
		}
		for (unsigned todo = __ballot_sync(0xffffffffu, head && !made && !failed); todo; todo &= todo - 1) {
			const uint32_t h = (uint32_t)__ffs(todo) - 1;
			unsigned long long at = ~0ull;
			if (lane == h) {
				// published (release) once the creator has written the slots; the acquire load orders this stretch's additions after them
				while ((at = load_acquire(&entry->base)) == 0)
					if (table_overflowed_lane(table)) {
						at = ~0ull;
						break;
					}
				if (at != ~0ull)
					at -= 1;
			}
			at = __shfl_sync(0xffffffffu, at, h);
			if (at != ~0ull)
				stretch(h, at, false);
		}
	}

	// the objects of the current run go to the global table, four per lane at a time
	template <class WS, class Emit>
	__device__ void flush_warp(WS &ws, Emit &emit) const {
		__syncwarp();
		if (ws.run_valid && emit.table.dir) {
			// region mode: one directory probe for the whole run, then the run's objects land on consecutive slots
			const uint32_t lane = lane_id(), leaves = ws.run_leaves;
			spread_run(ws);
			region_grant grant{~0ull, nullptr, 0, 0, false};
			if (lane == 0) {
				grant = region_acquire(emit.table, ws.chunk, region_key(ws.run_eligible, ws.run_fixed, ws.run_target, ws.run_names, ws.run_n), leaves);
				emit.regions += grant.created;
			}
			const unsigned long long base = __shfl_sync(0xffffffffu, grant.base, 0);
			const int made = __shfl_sync(0xffffffffu, (int)grant.created, 0);
			const unsigned long long retire_from = __shfl_sync(0xffffffffu, grant.retire_from, 0), retire_count = __shfl_sync(0xffffffffu, grant.retire_count, 0);
			if (retire_count)
				region_retire(emit.table, retire_from, retire_count);
			if (base != ~0ull) {
				table_slot *slots = emit.table.slots + base;
				if (made) {
					// first run of these objects anywhere: it writes the slots WHOLE -- hash (tree of the opening group), summed
					// magnitude, representative -- one full 32-byte sector per object, then publishes the region
					expand_full<false>(ws.open_ctx, ws.open_root, ws);
					const uint32_t levels = ws.open_ctx.levels, tree_bits = ws.open_ctx.tree_bits;
					const uint32_t shift = ws.open_ctx.eligible - levels; // child_id = group | leaf << shift
					const uint64_t names_hash = ws.open_ctx.names_hash;
					for (uint32_t slot = lane; slot < leaves; slot += 32) {
						const uint32_t leaf = slot ^ tree_bits;
						const unsigned long long key = hash_combine(hash_combine(names_hash, ws.hl[leaf]), ws.hr[leaf]);
						const unsigned long long rep = rep_pack(ws.open_first_child + (ws.open_group | (leaf << shift)), ws.open_size);
						ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(slots + slot);
						dst[0] = make_ulonglong2(key, (unsigned long long)__double_as_longlong(ws.acc_re[slot]));
						dst[1] = make_ulonglong2((unsigned long long)__double_as_longlong(ws.acc_im[slot]), rep);
					}
					__syncwarp();
					if (lane == 0) {
						region_publish(grant);
						emit.created += leaves;
					}
				} else {
					for (uint32_t i = lane; i < leaves; i += 32) {
						atomicAdd(&slots[i].re, ws.acc_re[i]); // results unused -> RED.ADD.F64 on consecutive sectors
						atomicAdd(&slots[i].im, ws.acc_im[i]);
					}
				}
			}
		} else if (ws.run_valid) {
			const uint32_t leaves = ws.run_leaves;
			spread_run(ws);
			for (uint32_t base = lane_id(); base < leaves; base += 128) {
				uint64_t hash[4];
				int count = 0;
#pragma unroll
				for (int q = 0; q < 4; ++q)
					if (base + q * 32 < leaves) {
						hash[q] = ws.hl[base + q * 32];
						count = q + 1;
					}
				emit.template batch_raw<4>(
				    count, hash, [&ws, base](int q) { return cplx{ws.acc_re[base + q * 32], ws.acc_im[base + q * 32]}; },
				    [&ws, base](int q) { return ws.hr[base + q * 32]; });
			}
		}
		__syncwarp();
		if (lane_id() == 0)
			ws.run_valid = 0;
		__syncwarp();
	}

	// the warp is done: what it did not use of its last range of table slots must read as empty (table.cuh, region_retire)
	template <class WS, class Emit>
	__device__ void finish_warp(WS &ws, Emit &emit) const {
		__syncwarp();
		if (emit.table.dir && ws.chunk.end > ws.chunk.next)
			region_retire(emit.table, ws.chunk.next, ws.chunk.end - ws.chunk.next);
		__syncwarp();
		if (lane_id() == 0)
			ws.chunk.next = ws.chunk.end = 0;
	}

	// full expansion of one group: tree states (hash folds and magnitudes) of all its leaves in ws
	template <bool WITH_MAG = true, class WS>
	__device__ __forceinline__ void expand_full(const flip_ctx &ctx, const flip_root &root, WS &ws) const {
		const uint32_t lane = lane_id();
		const uint64_t left = ctx.left, right = ctx.right;
		const uint32_t levels = ctx.levels;
		if (lane == 0) {
			ws.hl[0] = root.hl;
			ws.hr[0] = root.hr;
			if (WITH_MAG) {
				ws.re[0] = root.mag.re;
				ws.im[0] = root.mag.im;
			}
		}
		__syncwarp();
		// level l: tree node pos[l] with both choices, then the non-eligible nodes up to the next tree
		// node.  State i keeps choice 0 in place; choice 1 becomes state i + 2^l (bit l of the leaf index).
		for (uint32_t l = 0; l < levels; ++l) {
			const uint32_t e = ctx.pos[l], end = ctx.pos[l + 1];
			const bool pl = (left >> e) & 1, pr = (right >> e) & 1;
			const cplx stay = this->amp.get(false, pl), go = this->amp.get(true, pl);
			const uint32_t width = 1u << l;
			for (uint32_t i = lane; i < width; i += 32) {
				uint64_t hl0 = ws.hl[i], hr0 = ws.hr[i];
				const cplx m = WITH_MAG ? cplx{ws.re[i], ws.im[i]} : cplx{0, 0};
				uint64_t hl1 = hl0, hr1 = hr0;
				// choice 0 keeps the parent's particles, choice 1 toggles both
				if (pl) hl0 = hash_combine_index(hl0, e); else hl1 = hash_combine_index(hl1, e);
				if (pr) hr0 = hash_combine_index(hr0, e); else hr1 = hash_combine_index(hr1, e);
				for (uint32_t j = e + 1; j < end; ++j) {
					if ((left >> j) & 1) {
						hl0 = hash_combine_index(hl0, j);
						hl1 = hash_combine_index(hl1, j);
					}
					if ((right >> j) & 1) {
						hr0 = hash_combine_index(hr0, j);
						hr1 = hash_combine_index(hr1, j);
					}
				}
				ws.hl[i] = hl0;
				ws.hr[i] = hr0;
				ws.hl[i + width] = hl1;
				ws.hr[i + width] = hr1;
				if (WITH_MAG) {
					const cplx m0 = cmul(m, stay), m1 = cmul(m, go);
					ws.re[i] = m0.re;
					ws.im[i] = m0.im;
					ws.re[i + width] = m1.re;
					ws.im[i + width] = m1.im;
				}
			}
			__syncwarp();
		}
	}

	// a new run starts (rare next to the groups that continue one): the previous run goes to the table,
	// this group is expanded in full and opens the accumulators.  Kept out of line so that the hot path
	// of symbolic_warp<true> stays small.
	template <class WS, class Emit>
	__device__ void open_run(uint32_t parent_size, const flip_ctx &ctx, uint32_t group, const flip_root &root, WS &ws,
	                                      Emit &emit, uint64_t eligible, uint64_t fixed, uint32_t target) const {
		const uint32_t lane = lane_id();
		const uint32_t levels = ctx.levels, leaves = 1u << levels, tree_bits = ctx.tree_bits;
		flush_warp(ws, emit);
		if (lane == 0) {
			ws.run_eligible = eligible;
			ws.run_fixed = fixed;
			ws.run_names = ctx.names_hash;
			ws.run_n = ctx.n;
			ws.run_target = target;
			ws.run_leaves = leaves;
			ws.run_valid = 1;
#pragma unroll
			for (int w = 0; w < FLIP_BLOCK / 32; ++w)
				ws.run_patterns[w] = 0;
			ws.run_patterns[tree_bits >> 5] = 1u << (tree_bits & 31);
		}
		if (emit.table.dir) { // region mode: hashes and representatives are only needed if this run creates the region (flush_warp)
			if (lane == 0) {
				region_prefetch(emit.table, region_key(eligible, fixed, target, ctx.names_hash, ctx.n)); // probed when the run is flushed
				ws.open_ctx = ctx;
				ws.open_root = root;
				ws.open_first_child = emit.first_child;
				ws.open_group = group;
				ws.open_size = parent_size;
			}
			for (uint32_t leaf = lane; leaf < leaves; leaf += 32) {
				ws.acc_re[leaf] = leaf == tree_bits ? root.mag.re : 0.0;
				ws.acc_im[leaf] = leaf == tree_bits ? root.mag.im : 0.0;
			}
			__syncwarp();
			return;
		}
		expand_full<false>(ctx, root, ws);
		// hash and representative take the place of the tree states (another slot of the same arrays:
		// read everything first, then write)
		const uint32_t shift = ctx.eligible - levels; // child_id = group | leaf << shift
		const uint64_t names_hash = ctx.names_hash;
		uint64_t hash[FLIP_BLOCK / 32];
#pragma unroll
		for (int q = 0; q < FLIP_BLOCK / 32; ++q) {
			const uint32_t leaf = lane + q * 32;
			if (leaf < leaves)
				hash[q] = hash_combine(hash_combine(names_hash, ws.hl[leaf]), ws.hr[leaf]);
		}
		__syncwarp();
#pragma unroll
		for (int q = 0; q < FLIP_BLOCK / 32; ++q) {
			const uint32_t leaf = lane + q * 32;
			if (leaf < leaves) {
				const uint32_t slot = leaf ^ tree_bits;
				ws.hl[slot] = hash[q];
				ws.hr[slot] = emit.rep(group | (leaf << shift), parent_size);
				ws.acc_re[leaf] = leaf == tree_bits ? root.mag.re : 0.0;
				ws.acc_im[leaf] = leaf == tree_bits ? root.mag.im : 0.0;
			}
		}
		__syncwarp();
	}

	template <bool ACCUMULATE, class WS, class Emit>
	__device__ void symbolic_warp(const uint8_t *parent, uint32_t parent_size, const flip_ctx &ctx, uint32_t group, const flip_root &root,
	                              WS &ws, Emit &emit) const {
		const uint32_t lane = lane_id();
		if (ctx.n > 64) { // wide graph: 32 children per group, one per lane
			const uint32_t child_id = group * 32 + lane;
			if (ctx.eligible < 32 && child_id < (1u << ctx.eligible)) {
				uint32_t size;
				cplx mag = root.mag;
				const uint64_t hash = symbolic(parent, parent_size, ctx, child_id, nullptr, size, mag);
				emit(child_id, hash, size, mag);
			}
			return;
		}
		const uint32_t levels = ctx.levels;
		const uint32_t leaves = 1u << levels;

		if (ACCUMULATE) {
			const uint64_t left = ctx.left, right = ctx.right;
			const uint64_t all = ctx.n == 64 ? ~0ull : ((1ull << ctx.n) - 1);
			const uint64_t eligible = (WANT_EQUAL ? ~(left ^ right) : (left ^ right)) & all;
			const uint64_t fixed = left & ~eligible & all;
			const uint32_t target = group ^ ctx.prefix_bits;
			const bool same = ws.run_valid && ws.run_eligible == eligible && ws.run_fixed == fixed && ws.run_names == ctx.names_hash &&
			                  ws.run_n == ctx.n && ws.run_target == target;
			if (!same) {
				open_run(parent_size, ctx, group, root, ws, emit, eligible, fixed, target);
				return;
			}
			// the objects of this group are already in the accumulators: its root joins the sum of its pattern
			if (lane == 0) {
				ws.run_patterns[ctx.tree_bits >> 5] |= 1u << (ctx.tree_bits & 31);
				ws.acc_re[ctx.tree_bits] += root.mag.re;
				ws.acc_im[ctx.tree_bits] += root.mag.im;
			}
			__syncwarp();
			return;
		}

		// unsorted order: expand, finish the hash and insert, four per lane at a time (four table loads
		// in flight); magnitudes stay in shared memory until their entry is resolved
		expand_full(ctx, root, ws);
		const uint32_t shift = ctx.eligible - levels; // child_id = group | leaf << shift
		const uint64_t names_hash = ctx.names_hash;
		for (uint32_t base = lane; base < leaves; base += 128) {
			uint64_t hash[4];
			int count = 0;
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const uint32_t leaf = base + q * 32;
				if (leaf < leaves) {
					hash[q] = hash_combine(hash_combine(names_hash, ws.hl[leaf]), ws.hr[leaf]);
					count = q + 1;
				}
			}
			emit.template batch<4>(
			    count, hash, parent_size, [=](int q) { return group | ((base + q * 32) << shift); },
			    [&ws, base](int q) { return cplx{ws.re[base + q * 32], ws.im[base + q * 32]}; });
		}
		__syncwarp();
	}
};

// ===================================================================================================
// split_merge (qcgd.hpp:607-1036)
// ===================================================================================================
struct split_merge_plan {
	bool first_split, last_merge;
	uint32_t bits;    // child_id bits left for the general walk
	uint32_t child_n; // number of nodes of the child
};

// decides the wrap-around sites, counts the child's nodes and (optionally) accumulates the magnitude
// (qcgd.hpp:647-700).  A first split that is NOT taken leaves its bit to the general walk, which
// then meets node 0 as an ordinary split site (:655-659).
template <bool WITH_MAG>
__device__ inline split_merge_plan plan_split_merge(const graph &g, uint32_t child_id, const amplitudes &amp, cplx &mag) {
	split_merge_plan pl;
	const uint32_t n = g.n;
	bool fs = g.left(0) && g.right(0);
	bool lm = !fs && n > 1 && g.right(0) && g.left(n - 1) && !g.right(n - 1);
	uint32_t bits = child_id;
	fs = fs && (bits & 1);
	if (fs) {
		if (WITH_MAG) mag = cmul(mag, amp.get(true, false));
		bits >>= 1;
	}
	if (lm) {
		const bool taken = bits & 1;
		if (WITH_MAG) mag = cmul(mag, amp.get(taken, true));
		lm = taken;
		bits >>= 1;
	}
	pl.first_split = fs;
	pl.last_merge = lm;
	pl.bits = bits;
	uint32_t cn = n + fs - lm;
	uint32_t b = bits;
	for (uint32_t i = (uint32_t)fs + lm; i < n - lm; ++i) {
		bool split, merge;
		g.site(i, split, merge);
		if (split || merge) {
			const bool taken = b & 1;
			b >>= 1;
			if (taken)
				cn += (int)split - (int)merge;
			if (WITH_MAG) mag = cmul(mag, amp.get(taken, merge));
		}
	}
	pl.child_n = cn;
	return pl;
}

// announces the child's nodes in order (qcgd.hpp:702-845).  Emit provides
//   copy(i, l, r)        node i of the parent, unchanged
//   left_half(i, l, r)   left  name of the split of node i  (operations::left,  qcgd.hpp:194-200)
//   right_half(i, l, r)  right name of the split of node i  (operations::right, qcgd.hpp:202-208)
//   merged(i, j)         merge of nodes i and j, both particles set (operations::merge, qcgd.hpp:181-192)
template <class Emit>
__device__ inline void walk_split_merge(const graph &g, const split_merge_plan &pl, Emit &out) {
	const uint32_t n = g.n;
	bool overflow = false;
	if (pl.first_split) {
		// when the most-left element of node 0's name is not 0, the left half wraps to the END (:709-749, 838-845)
		const atom a0 = g.get(0);
		const bool most_left_zero = !(a0.kind >= 0 && g.get(1).hmlz > 0);
		if (most_left_zero) {
			out.left_half(0, true, false);
			out.right_half(0, false, true);
		} else {
			overflow = true;
			out.right_half(0, false, true);
		}
	}
	if (pl.last_merge)
		out.merged(n - 1, 0);
	uint32_t bits = pl.bits;
	for (uint32_t i = (uint32_t)pl.first_split + pl.last_merge; i < n - pl.last_merge; ++i) {
		bool split, merge;
		g.site(i, split, merge);
		bool taken = false;
		if (split || merge) {
			taken = bits & 1;
			bits >>= 1;
		}
		if (taken && split) {
			out.left_half(i, true, false);
			out.right_half(i, false, true);
		} else if (taken && merge) {
			out.merged(i, i + 1);
			++i; // node i+1 is consumed (:816)
		} else {
			out.copy(i, g.left(i), g.right(i));
		}
	}
	if (overflow)
		out.left_half(0, true, false);
}

// symbolic emitter: folds hash_graph of the child and counts its atoms, writes nothing
struct hash_emitter {
	const graph &g;
	uint64_t hl = 0, hr = 0, hn = 0;
	uint32_t index = 0, atoms = 0;

	__device__ explicit hash_emitter(const graph &g_) : g(g_) {}
	__device__ void node(bool l, bool r, uint64_t first_hash, uint32_t count) {
		if (l) hl = hash_combine(hl, index);
		if (r) hr = hash_combine(hr, index);
		hn = hash_combine(hn, first_hash);
		atoms += count;
		++index;
	}
	__device__ void copy(uint32_t i, bool l, bool r) {
		const uint32_t b = g.name_begin(i);
		node(l, r, g.atom_hash(b), g.name_begin(i + 1) - b);
	}
	__device__ void left_half(uint32_t i, bool l, bool r) {
		const uint32_t b = g.name_begin(i), len = g.name_begin(i + 1) - b;
		const atom a = g.get(b);
		if (a.kind >= 0)
			node(l, r, g.atom_hash(b + 1), a.kind - 1);
		else
			node(l, r, hash_combine(a.hash, (uint64_t)(int64_t)DOT_L), len + 1);
	}
	__device__ void right_half(uint32_t i, bool l, bool r) {
		const uint32_t b = g.name_begin(i), len = g.name_begin(i + 1) - b;
		const atom a = g.get(b);
		if (a.kind >= 0)
			node(l, r, g.atom_hash(b + a.kind), len - a.kind);
		else
			node(l, r, hash_combine(a.hash, (uint64_t)(int64_t)DOT_R), len + 1);
	}
	__device__ void merged(uint32_t i, uint32_t j) {
		const uint32_t bi = g.name_begin(i), li = g.name_begin(i + 1) - bi;
		const uint32_t bj = g.name_begin(j), lj = g.name_begin(j + 1) - bj;
		const atom a = g.get(bi), c = g.get(bj);
		if (a.kind == DOT_L && c.kind == DOT_R && g.atom_hash(bi + 1) == g.atom_hash(bj + 1)) // X.l v X.r -> X, on the hash only
			node(true, true, g.atom_hash(bi + 1), li - 1);
		else
			node(true, true, hash_combine(a.hash, c.hash), li + lj + 1);
	}
	__device__ uint64_t hash() const { return hash_combine(hash_combine(hn, hl), hr); }
	__device__ uint32_t size() const { return 4 + 4 * index + 16 * atoms; }
};

// finalisation emitter: writes the child's bytes; the 4 padding bytes of created atoms are zero,
// copied atoms keep the parent's
struct byte_emitter {
	const graph &g;
	uint8_t *child;
	uint32_t child_n, index = 0, atoms = 0;

	__device__ byte_emitter(const graph &g_, uint8_t *child_, uint32_t child_n_) : g(g_), child(child_), child_n(child_n_) {
		*reinterpret_cast<uint16_t *>(child) = (uint16_t)child_n;
	}
	__device__ uint16_t *name_begin() { return reinterpret_cast<uint16_t *>(child + 2 + 2 * child_n); }
	__device__ uint32_t *atom_words(uint32_t k) { return reinterpret_cast<uint32_t *>(child + 4 + 4 * child_n) + 4 * (size_t)k; }
	__device__ void open(bool l, bool r) {
		child[2 + index] = l;
		child[2 + child_n + index] = r;
		name_begin()[index] = (uint16_t)atoms;
	}
	__device__ void close() {
		++index;
		name_begin()[index] = (uint16_t)atoms;
	}
	__device__ void put(int hmlz, int kind, uint64_t hash) {
		uint32_t *w = atom_words(atoms++);
		w[0] = (uint32_t)(uint16_t)(int16_t)hmlz | ((uint32_t)(uint16_t)(int16_t)kind << 16);
		w[1] = 0;
		w[2] = (uint32_t)hash;
		w[3] = (uint32_t)(hash >> 32);
	}
	__device__ void put_range(uint32_t first, uint32_t last) {
		for (uint32_t k = first; k < last; ++k) {
			const uint32_t *s = g.atom_words(k);
			uint32_t *w = atom_words(atoms++);
			w[0] = s[0];
			w[1] = s[1];
			w[2] = s[2];
			w[3] = s[3];
		}
	}
	__device__ void copy(uint32_t i, bool l, bool r) {
		open(l, r);
		put_range(g.name_begin(i), g.name_begin(i + 1));
		close();
	}
	__device__ void left_half(uint32_t i, bool l, bool r) {
		const uint32_t b = g.name_begin(i), e = g.name_begin(i + 1);
		const atom a = g.get(b);
		open(l, r);
		if (a.kind >= 0) {
			put_range(b + 1, b + a.kind);
		} else {
			put(a.hmlz < 0 ? -1 : 1, DOT_L, hash_combine(a.hash, (uint64_t)(int64_t)DOT_L)); // qcgd.hpp:43-51
			put_range(b, e);
		}
		close();
	}
	__device__ void right_half(uint32_t i, bool l, bool r) {
		const uint32_t b = g.name_begin(i), e = g.name_begin(i + 1);
		const atom a = g.get(b);
		open(l, r);
		if (a.kind >= 0) {
			put_range(b + a.kind, e);
		} else {
			put(1, DOT_R, hash_combine(a.hash, (uint64_t)(int64_t)DOT_R));
			put_range(b, e);
		}
		close();
	}
	__device__ void merged(uint32_t i, uint32_t j) {
		const uint32_t bi = g.name_begin(i), ei = g.name_begin(i + 1);
		const uint32_t bj = g.name_begin(j), ej = g.name_begin(j + 1);
		const atom a = g.get(bi), c = g.get(bj);
		open(true, true);
		if (a.kind == DOT_L && c.kind == DOT_R && g.atom_hash(bi + 1) == g.atom_hash(bj + 1)) {
			put_range(bi + 1, ei);
		} else {
			put((a.hmlz < 0 || c.hmlz < 0) ? -1 : 1, (int)(ei - bi) + 1, hash_combine(a.hash, c.hash)); // qcgd.hpp:52-60
			put_range(bi, ei);
			put_range(bj, ej);
		}
		close();
	}
	__device__ uint32_t size() const { return 4 + 4 * child_n + 16 * atoms; }
};

struct split_merge : rule_base<split_merge> {
	amplitudes amp;

	__device__ uint64_t hasher(const uint8_t *object, uint32_t) const { return hash_graph(object); }

	__device__ void get_num_child(const uint8_t *parent, uint32_t parent_size, uint32_t &num_child, uint32_t &max_child_size) const {
		graph g(parent);
		max_child_size = 4 * parent_size; // qcgd.hpp:624
		const uint32_t n = g.n;
		if (n >= 1 && n <= 32 && (reinterpret_cast<uintptr_t>(parent) & 1) == 0) { // the same count from the particle masks (a few word loads)
			const uint64_t both = particle_mask(parent, n), all = (1ull << n) - 1;
			const uint32_t l = (uint32_t)(both & all), r = (uint32_t)(both >> n);
			const uint32_t split = l & r;
			const uint32_t merge = ~split & l & (r >> 1) & ~(l >> 1) & (n > 1 ? (0xffffffffu >> (33 - n)) : 0u); // i + 1 < n
			const bool wrap_merge = !(split & 1) && n > 1 && (r & 1) && ((l >> (n - 1)) & 1) && !((r >> (n - 1)) & 1);
			num_child = 1u << (__popc(split | merge) + (wrap_merge ? 1 : 0));
			return;
		}
		const bool fs = g.left(0) && g.right(0);
		const bool lm = !fs && n > 1 && g.right(0) && g.left(n - 1) && !g.right(n - 1);
		uint32_t sites = (fs || lm) ? 1 : 0;
		for (uint32_t i = fs; i < n - lm; ++i) {
			bool split, merge;
			g.site(i, split, merge);
			sites += split || merge;
		}
		num_child = 1u << sites;
	}

	__device__ void populate_child(const uint8_t *parent, uint32_t, uint8_t *child, uint32_t child_id, uint32_t &size, cplx &mag) const {
		graph g(parent);
		split_merge_plan pl = plan_split_merge<true>(g, child_id, amp, mag);
		byte_emitter out(g, child, pl.child_n);
		walk_split_merge(g, pl, out);
		size = out.size();
	}

	__device__ void populate_child_simple(const uint8_t *parent, uint32_t, uint8_t *child, uint32_t child_id) const {
		graph g(parent);
		cplx unused{1, 0};
		split_merge_plan pl = plan_split_merge<false>(g, child_id, amp, unused);
		byte_emitter out(g, child, pl.child_n);
		walk_split_merge(g, pl, out);
	}
};

// What all the children of one parent share in the symbolic phase, computed ONCE per parent by a whole warp (lane i =
// node i) and kept in shared memory: the particle masks, the split / merge site masks, and for every node the (first
// hash, number of atoms) of each name it can contribute to a child -- its own, the two halves of its split, its merge
// with the next node (node n-1: with node 0, the wrap-around merge).  A child is then a walk over these tables:
// no access to the object, no dependent loads, 1-3 hash folds per node.
constexpr uint32_t SPLIT_MERGE_MAX_NODES = 32;
struct split_merge_ctx {
	uint32_t n; // 0: not prepared (more than 32 nodes): the children walk the object itself
	uint32_t left, right, split, merge;
	uint32_t most_left_zero; // qcgd.hpp:709-749: where the left half of a first split goes
	// the sites in the order they consume the bits of child_id (a wrap-around merge first, then node order): their number, and
	// for bit b whether its site is a merge.  The magnitude of child c is the parent's times amp(bit b of c, site b is a merge)
	// for b = 0, 1, ... in this order (qcgd.hpp:647-700) -- it does not depend on the walk
	uint32_t num_sites, site_is_merge;
	// [0] the node's own name; [1] split site: left half, merge site (and node n-1): the merged name; [2] split site: right half
	uint64_t hash[3][SPLIT_MERGE_MAX_NODES]; // first hash of the name
	uint16_t len[3][SPLIT_MERGE_MAX_NODES];  // its number of atoms
};

struct split_merge_fused : split_merge {
	static constexpr bool needs_scratch = false;
	typedef split_merge_ctx ctx_t;
	static constexpr bool warp_prepare = true;
	static constexpr int parents_per_batch = 6; // 1 KB of context per parent
	static constexpr uint32_t prepare_stage_bytes = 3072; // 6 graphs of up to 500 bytes
	static void prepare_device(cudaStream_t stream) { particle_fold_prepare(stream); }

	__device__ void prepare(const uint8_t *, uint32_t, split_merge_ctx &ctx) const { ctx.n = 0; }

	// all 32 lanes, same arguments
	__device__ void prepare_warp(const uint8_t *parent, uint32_t, split_merge_ctx &ctx) const { prepare_lanes<32>(parent, ctx, true); }

	// TWO parents per warp (graphs of at most 16 nodes, the usual case): lanes 0-15 build the context of one parent, lanes 16-31
	// that of the next -- the same instruction stream serves both, half the warp instructions per parent.  Called by all 32
	// lanes; `parent` / `ctx` are those of the lane's half, `active` = the half has a parent
	static constexpr bool warp_prepare_pairs = true;
	__device__ static bool fits_half_warp(const uint8_t *parent) { return *reinterpret_cast<const uint16_t *>(parent) <= 16; }
	__device__ void prepare_half_warp(const uint8_t *parent, split_merge_ctx &ctx, bool active) const { prepare_lanes<16>(parent, ctx, active); }

	template <uint32_t WIDTH>
	__device__ __forceinline__ void prepare_lanes(const uint8_t *parent, split_merge_ctx &ctx, bool active) const {
		const uint32_t i = lane_id() & (WIDTH - 1), shift = lane_id() & ~(WIDTH - 1); // node of this lane, first lane of its group
		const uint32_t n = active ? *reinterpret_cast<const uint16_t *>(parent) : 0;
		const bool ok = n >= 1 && n <= (WIDTH < SPLIT_MERGE_MAX_NODES ? WIDTH : SPLIT_MERGE_MAX_NODES);
		const graph g(ok ? parent : reinterpret_cast<const uint8_t *>(&ctx.n)); // (never dereferenced beyond n when !ok: n is taken as 0 below)
		const bool here = ok && i < n;
		const uint32_t group_mask = WIDTH == 32 ? 0xffffffffu : ((1u << WIDTH) - 1);
		const uint32_t l = (__ballot_sync(0xffffffffu, here && parent[2 + i]) >> shift) & group_mask;
		const uint32_t r = (__ballot_sync(0xffffffffu, here && parent[2 + n + i]) >> shift) & group_mask;
		if (!ok) {
			if (active && i == 0)
				ctx.n = 0;
			return;
		}
		const uint32_t split = l & r;
		const uint32_t merge = ~split & l & (r >> 1) & ~(l >> 1) & (n > 1 ? (0xffffffffu >> (33 - n)) : 0u); // i + 1 < n
		const bool wrap_merge = !(split & 1) && n > 1 && (r & 1) && ((l >> (n - 1)) & 1) && !((r >> (n - 1)) & 1);
		if (i == 0) {
			ctx.n = n;
			ctx.left = l;
			ctx.right = r;
			ctx.split = split;
			ctx.merge = merge;
			ctx.most_left_zero = !(g.get(0).kind >= 0 && g.get(1).hmlz > 0);
			uint32_t count = wrap_merge ? 1 : 0, types = wrap_merge ? 1 : 0;
			for (uint32_t sites = split | merge; sites; sites &= sites - 1) {
				types |= ((merge >> (__ffs(sites) - 1)) & 1) << count;
				++count;
			}
			ctx.num_sites = count;
			ctx.site_is_merge = types;
		}
		if (!here)
			return;
		const uint32_t begin = g.name_begin(i), len = g.name_begin(i + 1) - begin;
		const atom first = g.get(begin);
		ctx.hash[0][i] = first.hash;
		ctx.len[0][i] = (uint16_t)len;
		if ((split >> i) & 1) { // operations::left / right, qcgd.hpp:194-208
			if (first.kind >= 0) {
				ctx.hash[1][i] = g.atom_hash(begin + 1);
				ctx.len[1][i] = (uint16_t)(first.kind - 1);
				ctx.hash[2][i] = g.atom_hash(begin + first.kind);
				ctx.len[2][i] = (uint16_t)(len - first.kind);
			} else {
				ctx.hash[1][i] = hash_combine(first.hash, (uint64_t)(int64_t)DOT_L);
				ctx.hash[2][i] = hash_combine(first.hash, (uint64_t)(int64_t)DOT_R);
				ctx.len[1][i] = ctx.len[2][i] = (uint16_t)(len + 1);
			}
		} else if (((merge >> i) & 1) || (wrap_merge && i == n - 1)) { // operations::merge, qcgd.hpp:181-192
			const uint32_t j = i + 1 < n ? i + 1 : 0;
			const uint32_t other = g.name_begin(j), other_len = g.name_begin(j + 1) - other;
			const atom second = g.get(other);
			if (first.kind == DOT_L && second.kind == DOT_R && g.atom_hash(begin + 1) == g.atom_hash(other + 1)) { // X.l v X.r -> X
				ctx.hash[1][i] = g.atom_hash(begin + 1);
				ctx.len[1][i] = (uint16_t)(len - 1);
			} else {
				ctx.hash[1][i] = hash_combine(first.hash, second.hash);
				ctx.len[1][i] = (uint16_t)(len + other_len + 1);
			}
		}
	}

	__device__ uint64_t symbolic(const uint8_t *parent, uint32_t, const split_merge_ctx &ctx, uint32_t child_id, uint8_t *, uint32_t &size, cplx &mag) const {
		const uint32_t n = ctx.n;
		if (n == 0) { // not prepared: walk the object (same result)
			graph g(parent);
			split_merge_plan pl = plan_split_merge<true>(g, child_id, amp, mag);
			hash_emitter out(g);
			walk_split_merge(g, pl, out);
			size = out.size();
			return out.hash();
		}
		const uint32_t left = ctx.left, right = ctx.right, split = ctx.split, merge = ctx.merge;
		uint64_t hn = 0;
		uint64_t child_left = 0, child_right = 0; // particle masks of the child (at most 2 n <= 64 nodes)
		uint32_t index = 0, atoms = 0, bits = child_id;
		// one node of the child: selects, no branches (the lanes of a warp are different children of the same few parents
		// and would take every path of a branchy walk one after the other: 2560 instructions per child, ncu
		// profiles/split_merge_sym_r1i, against ~800 this way).  The particles only set a bit: their two hashes are folds over
		// the node indices, looked up in the fold table at the end (two 64-bit multiplication chains per node less)
		auto node = [&](bool l, bool r, uint32_t which, uint32_t i) {
			child_left |= (uint64_t)l << index;
			child_right |= (uint64_t)r << index;
			hn = hash_combine(hn, ctx.hash[which][i]);
			atoms += ctx.len[which][i];
			++index;
		};
		// the wrap-around sites first (qcgd.hpp:647-668); a first split that is not taken leaves its bit to the walk
		const bool first_split = (split & 1) && (bits & 1);
		bool last_merge = !(split & 1) && n > 1 && (right & 1) && ((left >> (n - 1)) & 1) && !((right >> (n - 1)) & 1);
		bool overflow = false;
		// magnitude: one factor per site, in bit order
		for (uint32_t b = 0; b < ctx.num_sites; ++b)
			mag = cmul(mag, amp.get((child_id >> b) & 1, (ctx.site_is_merge >> b) & 1));
		if (first_split) {
			bits >>= 1;
			if (ctx.most_left_zero)
				node(true, false, 1, 0);
			else
				overflow = true; // the left half goes to the end
			node(false, true, 2, 0);
		}
		if (last_merge) {
			last_merge = bits & 1;
			bits >>= 1;
			if (last_merge)
				node(true, true, 1, n - 1);
		}
		// the general walk: every iteration emits exactly one node of the child -- the node itself, a merge (consumes two
		// nodes), or one half of a split (the right half is `pending` for the next iteration)
		const uint32_t end = n - last_merge;
		bool pending = false;
		for (uint32_t i = (uint32_t)first_split + last_merge; i < end;) {
			const bool is_split = (split >> i) & 1, is_merge = (merge >> i) & 1;
			const bool site = (is_split || is_merge) && !pending;
			const bool taken = site && (bits & 1);
			bits >>= site ? 1 : 0;
			const bool l = pending ? false : (taken ? true : (bool)((left >> i) & 1));
			const bool r = pending ? true : (taken ? is_merge : (bool)((right >> i) & 1));
			node(l, r, pending ? 2u : (taken ? 1u : 0u), i);
			const bool left_half = taken && is_split;
			i += left_half ? 0u : (taken ? 2u : 1u); // a taken merge swallows node i + 1 (:816)
			pending = left_half;
		}
		if (overflow)
			node(true, false, 1, 0);
		size = 4 + 4 * index + 16 * atoms;
		uint64_t hl = 0, hr = 0;
		if (index <= FOLD_TABLE_BITS) {
			hl = particle_fold((uint32_t)child_left);
			hr = particle_fold((uint32_t)child_right);
		} else {
			for (uint64_t m = child_left; m; m &= m - 1)
				hl = hash_combine_index(hl, (uint32_t)__ffsll((long long)m) - 1);
			for (uint64_t m = child_right; m; m &= m - 1)
				hr = hash_combine_index(hr, (uint32_t)__ffsll((long long)m) - 1);
		}
		return hash_combine(hash_combine(hn, hl), hr);
	}
	// ---- FANS (rule_api.cuh, lane_groups).  The bits of child_id are consumed in walk order, so the children that differ only
	// in their LAST choices share the walk over the parent up to the first of those sites.  One lane produces the 2^F children
	// fan | x << (S - F) of a parent with S sites, F = min(fork_max, S - 1): it walks to the (S - F)-th site once and forks there
	// (and again at the last site when F = 2) -- 1.2 walks for two children of a parent with four sites instead of two.  (Bit 0
	// is never forked: the wrap-around sites keep their place at the head of the walk.)
	static constexpr bool lane_groups = true;
	// children per fan = 2^fork_max at most.  Measured on the loop state (1e7 parents, child generation): 0 -> 15.2 ms, 1 -> 13.2 ms,
	// 2 -> 15.2 ms: four children per lane share more of the walk, but a batch of six parents then fills only 21 lanes
	uint32_t fork_max = 1; // (QB_SPLIT_MERGE_FORK = 0 / 1 / 2: developer knob for A/B runs)
	__device__ uint32_t fork_bits(uint32_t sites) const { return min(fork_max, sites >= 3 ? 2u : (sites == 2 ? 1u : 0u)); }
	__device__ uint32_t get_num_group(const uint8_t *, uint32_t, uint32_t num_child) const {
		const uint32_t sites = 31 - __clz(num_child);
		return num_child >> fork_bits(sites);
	}

	struct walk_state {
		uint64_t hn, left, right; // fold of the names, particle masks of the child so far
		uint32_t index, atoms, i; // nodes and atoms of the child so far, next node of the parent
		bool pending;             // the right half of a split is next
	};

	template <class Emit>
	__device__ void symbolic_fan(const uint8_t *parent, uint32_t parent_size, const split_merge_ctx &ctx, uint32_t fan, cplx parent_mag, uint8_t *scratch,
	                             Emit emit) const {
		const uint32_t n = ctx.n;
		if (n == 0) { // not prepared (more than 32 nodes): every child walks the object
			uint32_t num_child, unused;
			get_num_child(parent, parent_size, num_child, unused);
			const uint32_t sites = 31 - __clz(num_child), forked = fork_bits(sites);
			for (uint32_t x = 0; x < (1u << forked); ++x) {
				const uint32_t child_id = fan | (x << (sites - forked));
				cplx mag = parent_mag;
				uint32_t size;
				const uint64_t hash = symbolic(parent, parent_size, ctx, child_id, scratch, size, mag);
				emit(child_id, hash, size, mag);
			}
			return;
		}
		const uint32_t left = ctx.left, right = ctx.right, split = ctx.split, merge = ctx.merge;
		const uint32_t sites = ctx.num_sites, forked = fork_bits(sites), shared = sites - forked;
		cplx mag = parent_mag; // factors of the shared choices, in bit order
		for (uint32_t b = 0; b < shared; ++b)
			mag = cmul(mag, amp.get((fan >> b) & 1, (ctx.site_is_merge >> b) & 1));

		auto node = [&](walk_state &w, bool l, bool r, uint32_t which, uint32_t i) {
			w.left |= (uint64_t)l << w.index;
			w.right |= (uint64_t)r << w.index;
			w.hn = hash_combine(w.hn, ctx.hash[which][i]);
			w.atoms += ctx.len[which][i];
			++w.index;
		};
		// the wrap-around sites first (qcgd.hpp:647-668); a first split that is not taken leaves its bit to the walk
		walk_state st{0, 0, 0, 0, 0, 0, false};
		uint32_t bits = fan, todo = shared; // choices of the shared part still to be consumed
		const bool first_split = (split & 1) && (bits & 1);
		bool last_merge = !(split & 1) && n > 1 && (right & 1) && ((left >> (n - 1)) & 1) && !((right >> (n - 1)) & 1);
		bool overflow = false;
		if (first_split) {
			bits >>= 1;
			--todo;
			if (ctx.most_left_zero)
				node(st, true, false, 1, 0);
			else
				overflow = true; // the left half goes to the end
			node(st, false, true, 2, 0);
		}
		if (last_merge) {
			last_merge = bits & 1;
			bits >>= 1;
			--todo;
			if (last_merge)
				node(st, true, true, 1, n - 1);
		}
		st.i = (uint32_t)first_split + last_merge;
		const uint32_t end = n - last_merge;
		// the general walk (see symbolic()): every iteration emits one node of the child; stops in front of a site when the
		// choices given to it are used up
		auto run = [&](walk_state &w, uint32_t choice_bits, uint32_t choices) {
			while (w.i < end) {
				const uint32_t i = w.i;
				const bool is_split = (split >> i) & 1, is_merge = (merge >> i) & 1;
				const bool site = (is_split || is_merge) && !w.pending;
				if (site && choices == 0)
					return;
				const bool taken = site && (choice_bits & 1);
				choice_bits >>= site ? 1 : 0;
				choices -= site ? 1 : 0;
				const bool l = w.pending ? false : (taken ? true : (bool)((left >> i) & 1));
				const bool r = w.pending ? true : (taken ? is_merge : (bool)((right >> i) & 1));
				node(w, l, r, w.pending ? 2u : (taken ? 1u : 0u), i);
				const bool left_half = taken && is_split;
				w.i = i + (left_half ? 0u : (taken ? 2u : 1u)); // a taken merge swallows node i + 1 (:816)
				w.pending = left_half;
			}
		};
		auto finish = [&](walk_state &w, cplx child_mag, uint32_t child_id) {
			if (overflow)
				node(w, true, false, 1, 0);
			uint64_t hl = 0, hr = 0;
			if (w.index <= FOLD_TABLE_BITS) {
				hl = particle_fold((uint32_t)w.left);
				hr = particle_fold((uint32_t)w.right);
			} else {
				for (uint64_t m = w.left; m; m &= m - 1)
					hl = hash_combine_index(hl, (uint32_t)__ffsll((long long)m) - 1);
				for (uint64_t m = w.right; m; m &= m - 1)
					hr = hash_combine_index(hr, (uint32_t)__ffsll((long long)m) - 1);
			}
			emit(child_id, hash_combine(hash_combine(w.hn, hl), hr), 4 + 4 * w.index + 16 * w.atoms, child_mag);
		};
		run(st, bits, todo);
		const uint32_t fan1 = forked >= 1 ? 2 : 1, fan2 = forked >= 2 ? 2 : 1;
		for (uint32_t x1 = 0; x1 < fan1; ++x1) {
			walk_state w1 = st;
			cplx m1 = mag;
			if (forked >= 1) {
				m1 = cmul(mag, amp.get(x1, (ctx.site_is_merge >> shared) & 1));
				run(w1, x1, 1);
			}
			for (uint32_t x2 = 0; x2 < fan2; ++x2) {
				walk_state w2 = w1;
				cplx m2 = m1;
				if (forked >= 2) {
					m2 = cmul(m1, amp.get(x2, (ctx.site_is_merge >> (shared + 1)) & 1));
					run(w2, x2, 1);
				}
				finish(w2, m2, fan | (x1 << shared) | (x2 << (shared + 1)));
			}
		}
	}
};

template <class Rule>
inline void apply_developer_knobs(Rule &) {}
inline void apply_developer_knobs(split_merge_fused &r) {
	if (const char *fork = getenv("QB_SPLIT_MERGE_FORK"))
		r.fork_max = (uint32_t)atoi(fork) > 2 ? 2u : (uint32_t)atoi(fork);
}

template <class Rule>
inline int make_qcgd_rule(const double *params, uint32_t num_params, void *storage) {
	if (num_params < 1)
		return QB_ERR_ARG;
	Rule r;
	apply_developer_knobs(r);
	r.amp.set(params[0], num_params > 1 ? params[1] : 0.0, num_params > 2 ? params[2] : 0.0);
	memcpy(storage, &r, sizeof r);
	return QB_OK;
}

// ---- modifiers step / reversed_step (qcgd.hpp:443-457): left particles move one node to the left,
// right particles one node to the right (cyclically); reversed_step undoes it
template <bool REVERSED>
struct step_modifier {
	__device__ void operator()(uint8_t *object, uint32_t, cplx &) const {
		const uint32_t n = *reinterpret_cast<const uint16_t *>(object);
		if (n == 0)
			return;
		if (reinterpret_cast<uintptr_t>(object) & 3) { // (objects of a QCGD state start 4-byte aligned; anything else: byte by byte)
			uint8_t *down = object + 2 + (REVERSED ? n : 0); // array rotated towards index 0
			uint8_t *up = object + 2 + (REVERSED ? 0 : n);   // array rotated towards index n-1
			const uint8_t first = down[0];
			for (uint32_t i = 0; i + 1 < n; ++i)
				down[i] = down[i + 1];
			down[n - 1] = first;
			const uint8_t last = up[n - 1];
			for (uint32_t i = n - 1; i > 0; --i)
				up[i] = up[i - 1];
			up[0] = last;
			return;
		}
		// The 2 n particle bytes sit at [2, 2 + 2n): one half moves one byte down, the other one byte up, both cyclically.  One
		// thread per object with a byte load and a byte store per particle costs 4 n memory instructions whose 32 lanes touch 32
		// different sectors (ncu: 2.2 ms per pass over 1e7 graphs, bound by the load/store unit); here the object's leading words
		// are read and written ONCE each, as 32-bit words, and the bytes are moved in registers: word k of the result is blended
		// from word k itself, the word shifted by one byte either way (funnel shifts over the neighbours) and the two bytes that
		// wrap around.
		uint32_t *w = reinterpret_cast<uint32_t *>(object);
		const uint32_t words = (2 + 2 * n + 3) / 4;
		const uint32_t down_begin = 2 + (REVERSED ? n : 0), up_begin = 2 + (REVERSED ? 0 : n);
		const uint32_t down_wrap = object[down_begin], up_wrap = object[up_begin + n - 1]; // what re-enters at the other end
		uint32_t prev = 0, cur = w[0];
		for (uint32_t k = 0; k < words; ++k) {
			const uint32_t next = k + 1 < words ? w[k + 1] : 0;
			const uint32_t from_above = __funnelshift_r(cur, next, 8); // byte b = old byte 4k + b + 1
			const uint32_t from_below = __funnelshift_l(prev, cur, 8); // byte b = old byte 4k + b - 1
			uint32_t out = cur;
#pragma unroll
			for (uint32_t b = 0; b < 4; ++b) {
				const uint32_t p = 4 * k + b, lane_mask = 0xffu << (8 * b);
				if (p >= down_begin && p < down_begin + n)
					out = (out & ~lane_mask) | ((p + 1 < down_begin + n ? from_above : down_wrap << (8 * b)) & lane_mask);
				else if (p >= up_begin && p < up_begin + n)
					out = (out & ~lane_mask) | ((p > up_begin ? from_below : up_wrap << (8 * b)) & lane_mask);
			}
			if (out != cur)
				w[k] = out;
			prev = cur;
			cur = next;
		}
	}
};

// ---- observables of utils::serialize (qcgd.hpp:309-372) -----------------------------------------------------------------
// number of nodes; its square; density = (number of particles) / (2 n); its square -- the four averages of one
// serialize() call in a single pass over the state
struct stats_observable {
	static constexpr int values = 4;
	__device__ void operator()(const uint8_t *object, uint32_t, double *out) const {
		const uint32_t n = *reinterpret_cast<const uint16_t *>(object);
		const double nodes = (double)n;
		double density = 0; // the reference adds left + right per node in a PROBA_TYPE, then divides by 2 n (:329-334)
		for (uint32_t i = 0; i < n; ++i)
			density += (double)(object[2 + i] + object[2 + n + i]);
		density /= 2 * nodes;
		out[0] = nodes;
		out[1] = nodes * nodes;
		out[2] = density;
		out[3] = density * density;
	}
};
struct size_observable {
	static constexpr int values = 1;
	__device__ void operator()(const uint8_t *object, uint32_t, double *out) const { out[0] = (double)*reinterpret_cast<const uint16_t *>(object); }
};

} // namespace qcgd
} // namespace qb
