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

#include "rule_api.cuh"

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
	__device__ cplx get(bool taken, bool conjugated) const { return f[(taken ? 2 : 0) + (conjugated ? 1 : 0)]; }
};

// ===================================================================================================
// erase_create (qcgd.hpp:459-532) and coin (qcgd.hpp:534-605): every ELIGIBLE node consumes one bit
// of child_id, a set bit toggles both particles of the node.  erase_create: eligible = left == right;
// coin: eligible = left != right.  Size unchanged, names untouched.  In both rules the amplitude is
// conjugated exactly when the left particle is present.
// ===================================================================================================
struct flip_ctx {
	uint64_t left, right; // particle masks (only when n <= 64)
	uint64_t names_hash;  // fold of the first-atom hashes: the same for every child
	uint32_t n;
};

template <bool WANT_EQUAL>
struct flip_rule : rule_base<flip_rule<WANT_EQUAL>> {
	amplitudes amp;

	__device__ uint64_t hasher(const uint8_t *object, uint32_t) const { return hash_graph(object); }

	__device__ void get_num_child(const uint8_t *parent, uint32_t parent_size, uint32_t &num_child, uint32_t &max_child_size) const {
		graph g(parent);
		max_child_size = parent_size;
		uint32_t eligible = 0;
		for (uint32_t i = 0; i < g.n; ++i)
			eligible += (g.left(i) == g.right(i)) == WANT_EQUAL;
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
	static constexpr bool needs_scratch = false;

	__device__ void prepare(const uint8_t *parent, uint32_t, flip_ctx &ctx) const {
		graph g(parent);
		ctx.n = g.n;
		uint64_t l = 0, r = 0, hn = 0;
		for (uint32_t i = 0; i < g.n; ++i) {
			if (i < 64) {
				l |= (uint64_t)g.left(i) << i;
				r |= (uint64_t)g.right(i) << i;
			}
			hn = hash_combine(hn, g.atom_hash(g.name_begin(i)));
		}
		ctx.left = l;
		ctx.right = r;
		ctx.names_hash = hn;
	}

	__device__ uint64_t symbolic(const uint8_t *parent, uint32_t parent_size, const flip_ctx &ctx, uint32_t child_id, uint8_t *, uint32_t &size,
	                             cplx &mag) const {
		size = parent_size;
		uint64_t hl = 0, hr = 0;
		if (ctx.n <= 64) {
			const uint64_t all = ctx.n == 64 ? ~0ull : ((1ull << ctx.n) - 1);
			uint64_t eligible = (WANT_EQUAL ? ~(ctx.left ^ ctx.right) : (ctx.left ^ ctx.right)) & all;
			uint64_t toggled = 0;
			while (eligible) {
				const uint64_t lowest = eligible & (0 - eligible);
				eligible ^= lowest;
				const bool taken = child_id & 1;
				child_id >>= 1;
				mag = cmul(mag, this->amp.get(taken, ctx.left & lowest));
				if (taken)
					toggled |= lowest;
			}
			uint64_t l = ctx.left ^ toggled, r = ctx.right ^ toggled;
			while (l) {
				hl = hash_combine(hl, (uint64_t)(__ffsll((long long)l) - 1));
				l &= l - 1;
			}
			while (r) {
				hr = hash_combine(hr, (uint64_t)(__ffsll((long long)r) - 1));
				r &= r - 1;
			}
		} else { // wide graphs: same walk on the bytes
			graph g(parent);
			for (uint32_t i = 0; i < g.n; ++i) {
				bool l = g.left(i), r = g.right(i);
				if ((l == r) == WANT_EQUAL) {
					const bool taken = child_id & 1;
					child_id >>= 1;
					mag = cmul(mag, this->amp.get(taken, l));
					if (taken) {
						l = !l;
						r = !r;
					}
				}
				if (l)
					hl = hash_combine(hl, i);
				if (r)
					hr = hash_combine(hr, i);
			}
		}
		return hash_combine(hash_combine(ctx.names_hash, hl), hr);
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

struct split_merge_fused : split_merge {
	static constexpr bool needs_scratch = false;

	__device__ uint64_t symbolic(const uint8_t *parent, uint32_t, const no_ctx &, uint32_t child_id, uint8_t *, uint32_t &size, cplx &mag) const {
		graph g(parent);
		split_merge_plan pl = plan_split_merge<true>(g, child_id, amp, mag);
		hash_emitter out(g);
		walk_split_merge(g, pl, out);
		size = out.size();
		return out.hash();
	}
};

template <class Rule>
inline int make_qcgd_rule(const double *params, uint32_t num_params, void *storage) {
	if (num_params < 1)
		return QB_ERR_ARG;
	Rule r;
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
	}
};

} // namespace qcgd
} // namespace qb
