/*
 * oracle.cpp -- CPU restatement of the QuIDS rule-application step.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_api.h): the product path (quids_b200/csrc) never links,
 * loads or calls this file.  It is the checker the CUDA path is compared with.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   - the known answers of SURVEY.md appendix A.4 (taken from the reference run in this image),
 *   - golden vectors produced by the unmodified reference (oracle/_ref, see gen_golden.py),
 *   - the live reference library when /root/reference is present.
 *
 * The code is deliberately written in a different shape from the reference: QCGD objects are
 * DECODED into a small graph value, transformed, and ENCODED back, instead of being edited in
 * place through pointer accessors.  All citations are into /root/reference/src.
 */
#include "oracle_api.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <unordered_map>
#include <vector>

namespace {

typedef std::complex<double> cplx;

/* ------------------------------------------------------------------------------------------
 * state: structure of arrays, unpadded (padding is a storage detail of quids.hpp:174-188 that
 * never reaches a rule: rules see [begin, begin+size) only)
 * ------------------------------------------------------------------------------------------ */
struct state_t {
	std::vector<uint8_t> bytes;
	std::vector<uint64_t> begin{0};
	std::vector<cplx> mag;
	double total_proba = 1; /* quids.hpp:154 */

	size_t n() const { return mag.size(); }
	uint32_t size(size_t i) const { return (uint32_t)(begin[i + 1] - begin[i]); }
	const uint8_t *obj(size_t i) const { return bytes.data() + begin[i]; }
	uint8_t *obj(size_t i) { return bytes.data() + begin[i]; }
	void clear() {
		bytes.clear();
		begin.assign(1, 0);
		mag.clear();
		total_proba = 1;
	}
	void push(const uint8_t *b, uint32_t sz, cplx m) {
		bytes.insert(bytes.end(), b, b + sz);
		begin.push_back(bytes.size());
		mag.push_back(m);
	}
};

/* ------------------------------------------------------------------------------------------
 * default hasher: std::hash<std::string_view> (quids.hpp:143-145) = libstdc++ _Hash_bytes,
 * 64-bit variant (gcc libsupc++/hash_bytes.cc), seed 0xc70f6907.  Restated, not called.
 * ------------------------------------------------------------------------------------------ */
const uint64_t MURMUR_MUL = 0xc6a4a7935bd1e995ull;

inline uint64_t shift_mix(uint64_t v) { return v ^ (v >> 47); }

uint64_t hash_bytes(const uint8_t *p, size_t len) {
	uint64_t h = 0xc70f6907ull ^ (len * MURMUR_MUL);
	size_t whole = len & ~(size_t)7;
	for (size_t i = 0; i < whole; i += 8) {
		uint64_t w;
		memcpy(&w, p + i, 8); /* little endian host */
		h ^= shift_mix(w * MURMUR_MUL) * MURMUR_MUL;
		h *= MURMUR_MUL;
	}
	if (len & 7) {
		uint64_t w = 0;
		for (size_t i = len & 7; i-- > 0;)
			w = (w << 8) + p[whole + i];
		h ^= w;
		h *= MURMUR_MUL;
	}
	h = shift_mix(h) * MURMUR_MUL;
	return shift_mix(h);
}

/* qcgd.hpp:11-25 */
inline uint64_t hash_combine(uint64_t seed, uint64_t v) {
	seed *= MURMUR_MUL;
	seed ^= v >> 47;
	seed *= MURMUR_MUL;
	seed ^= v;
	seed *= MURMUR_MUL;
	return seed + 0xe6546b64ull;
}

/* ------------------------------------------------------------------------------------------
 * QCGD graph value (layout: qcgd.hpp:63-112)
 *   u16 n | u8 left[n] | u8 right[n] | u16 name_begin[n+1] | sub_node names[name_begin[n]]
 *   sub_node = { i16 hmlz_and_element, i16 right_or_type, 4 bytes padding, u64 hash }
 * ------------------------------------------------------------------------------------------ */
enum { DOT_L = -3, DOT_R = -2, ELEMENT = -1 }; /* qcgd.hpp:29-34; >= 0: offset to right subtree */

struct atom_t { /* one sub_node */
	int16_t hmlz;
	int16_t kind;
	uint64_t hash;
};
typedef std::vector<atom_t> name_t; /* prefix-order name tree of one node */

atom_t atom_element(int16_t e) { /* qcgd.hpp:40-42 */
	return atom_t{(int16_t)(e == 0 ? -1 : e + 1), ELEMENT, (uint64_t)e};
}
atom_t atom_wrap(const atom_t &a, int16_t kind) { /* qcgd.hpp:43-51 */
	int16_t hmlz = (kind == DOT_L && a.hmlz < 0) ? -1 : 1;
	return atom_t{hmlz, kind, hash_combine(a.hash, (uint64_t)(int64_t)kind)};
}
atom_t atom_pair(const atom_t &l, const atom_t &r, int16_t right_offset) { /* qcgd.hpp:52-60 */
	int16_t hmlz = (l.hmlz < 0 || r.hmlz < 0) ? -1 : 1;
	return atom_t{hmlz, right_offset, hash_combine(l.hash, r.hash)};
}

struct graph_t {
	std::vector<uint8_t> left, right;
	std::vector<name_t> name;
	size_t n() const { return name.size(); }
};

graph_t decode(const uint8_t *p) {
	graph_t g;
	uint16_t n;
	memcpy(&n, p, 2);
	g.left.assign(p + 2, p + 2 + n);
	g.right.assign(p + 2 + n, p + 2 + 2 * n);
	const uint8_t *nb = p + 2 + 2 * (size_t)n;
	const uint8_t *nodes = p + 4 + 4 * (size_t)n;
	g.name.resize(n);
	for (uint16_t i = 0; i < n; ++i) {
		uint16_t b, e;
		memcpy(&b, nb + 2 * i, 2);
		memcpy(&e, nb + 2 * (i + 1), 2);
		for (uint16_t k = b; k < e; ++k) {
			atom_t a;
			memcpy(&a.hmlz, nodes + 16 * (size_t)k, 2);
			memcpy(&a.kind, nodes + 16 * (size_t)k + 2, 2);
			memcpy(&a.hash, nodes + 16 * (size_t)k + 8, 8);
			g.name[i].push_back(a);
		}
	}
	return g;
}

size_t encoded_size(const graph_t &g) {
	size_t atoms = 0;
	for (auto &nm : g.name)
		atoms += nm.size();
	return 4 + 4 * g.n() + 16 * atoms;
}

/* padding bytes of each sub_node are written as zero (the reference leaves them indeterminate) */
void encode(const graph_t &g, std::vector<uint8_t> &out) {
	size_t n = g.n();
	out.assign(encoded_size(g), 0);
	uint8_t *p = out.data();
	uint16_t n16 = (uint16_t)n;
	memcpy(p, &n16, 2);
	for (size_t i = 0; i < n; ++i) {
		p[2 + i] = g.left[i];
		p[2 + n + i] = g.right[i];
	}
	uint8_t *nb = p + 2 + 2 * n;
	uint8_t *nodes = p + 4 + 4 * n;
	uint16_t k = 0;
	for (size_t i = 0; i < n; ++i) {
		memcpy(nb + 2 * i, &k, 2);
		for (auto &a : g.name[i]) {
			memcpy(nodes + 16 * (size_t)k, &a.hmlz, 2);
			memcpy(nodes + 16 * (size_t)k + 2, &a.kind, 2);
			memcpy(nodes + 16 * (size_t)k + 8, &a.hash, 8);
			++k;
		}
	}
	memcpy(nb + 2 * n, &k, 2);
}

/* qcgd.hpp:122-146 -- reads raw bytes so that it is valid on any object, not only on ones this
 * file encoded */
uint64_t hash_graph(const uint8_t *p) {
	uint16_t n;
	memcpy(&n, p, 2);
	const uint8_t *l = p + 2, *r = p + 2 + n, *nb = p + 2 + 2 * (size_t)n, *nodes = p + 4 + 4 * (size_t)n;
	uint64_t hl = 0, hr = 0, hn = 0;
	for (uint16_t i = 0; i < n; ++i) {
		if (l[i]) hl = hash_combine(hl, i);
		if (r[i]) hr = hash_combine(hr, i);
		uint16_t b;
		memcpy(&b, nb + 2 * i, 2);
		uint64_t first_hash;
		memcpy(&first_hash, nodes + 16 * (size_t)b + 8, 8);
		hn = hash_combine(hn, first_hash);
	}
	hn = hash_combine(hn, hl);
	return hash_combine(hn, hr);
}

/* name algebra, qcgd.hpp:181-208 */
name_t name_merge(const name_t &l, const name_t &r) {
	if (l[0].kind == DOT_L && r[0].kind == DOT_R && l[1].hash == r[1].hash)
		return name_t(l.begin() + 1, l.end()); /* X.l v X.r -> X, decided on the hash only */
	name_t out;
	out.push_back(atom_pair(l[0], r[0], (int16_t)(l.size() + 1)));
	out.insert(out.end(), l.begin(), l.end());
	out.insert(out.end(), r.begin(), r.end());
	return out;
}
name_t name_left(const name_t &p) {
	if (p[0].kind >= 0)
		return name_t(p.begin() + 1, p.begin() + p[0].kind);
	name_t out;
	out.push_back(atom_wrap(p[0], DOT_L));
	out.insert(out.end(), p.begin(), p.end());
	return out;
}
name_t name_right(const name_t &p) {
	if (p[0].kind >= 0)
		return name_t(p.begin() + p[0].kind, p.end());
	name_t out;
	out.push_back(atom_wrap(p[0], DOT_R));
	out.insert(out.end(), p.begin(), p.end());
	return out;
}

/* ------------------------------------------------------------------------------------------
 * rules
 * ------------------------------------------------------------------------------------------ */
struct rule_t {
	int id = 0;
	size_t bit = 0;         /* hadamard */
	cplx go, stay;          /* do_, do_not  (qcgd.hpp:466-471) */

	rule_t(int id_, const double *params) : id(id_) {
		if (id == ORC_RULE_HADAMARD) {
			bit = (size_t)params[0];
		} else {
			go = std::polar(std::sin(params[0]), params[1]);
			stay = std::polar(std::cos(params[0]), params[2]);
		}
	}

	uint64_t hasher(const uint8_t *b, uint32_t size) const {
		return id == ORC_RULE_HADAMARD ? hash_bytes(b, size) : hash_graph(b);
	}

	/* the factor applied for one binary choice: "taken" uses go or conj(go), "not taken" uses
	 * stay or -conj(stay)  (qcgd.hpp:501-506, 575-580, 686-696) */
	cplx factor(bool taken, bool conjugated) const {
		if (taken)
			return conjugated ? std::conj(go) : go;
		return conjugated ? -std::conj(stay) : stay;
	}

	/* ---- number of children ---- */
	uint32_t num_child(const uint8_t *p, uint32_t size, uint32_t &max_child_size) const {
		max_child_size = size;
		if (id == ORC_RULE_HADAMARD)
			return 2; /* quantum_computer.hpp:36-39 */

		graph_t g = decode(p);
		size_t n = g.n();
		uint32_t count = 1;
		if (id == ORC_RULE_ERASE_CREATE || id == ORC_RULE_COIN) {
			bool want_equal = id == ORC_RULE_ERASE_CREATE; /* qcgd.hpp:481-485 vs 556-560 */
			for (size_t i = 0; i < n; ++i)
				if ((g.left[i] == g.right[i]) == want_equal)
					count *= 2;
			return count;
		}

		/* split_merge, qcgd.hpp:623-642 */
		max_child_size = 4 * size;
		bool first_split = g.left[0] && g.right[0];
		bool last_merge = !first_split && n > 1 && g.right[0] && g.left[n - 1] && !g.right[n - 1];
		if (first_split || last_merge)
			count = 2;
		for (size_t i = first_split; i < n - last_merge; ++i) {
			bool split, merge;
			site(g, i, split, merge);
			if (split || merge)
				count *= 2;
		}
		return count;
	}

	/* qcgd.hpp:174-179 */
	static void site(const graph_t &g, size_t i, bool &split, bool &merge) {
		split = g.left[i] && g.right[i];
		merge = !split && i + 1 < g.n() && g.left[i] && g.right[i + 1] && !g.left[i + 1];
	}

	/* ---- one child: bytes, and the magnitude factor accumulated in mag ---- */
	void child(const uint8_t *p, uint32_t size, uint32_t child_id, std::vector<uint8_t> &out, cplx &mag) const {
		if (id == ORC_RULE_HADAMARD) { /* quantum_computer.hpp:40-49 */
			const double s = 1 / std::sqrt(2.);
			mag *= (p[bit] && child_id) ? -s : s;
			out.assign(p, p + size);
			out[bit] ^= !child_id;
			return;
		}

		graph_t g = decode(p);
		size_t n = g.n();

		if (id == ORC_RULE_ERASE_CREATE || id == ORC_RULE_COIN) { /* qcgd.hpp:487-510, 562-584 */
			bool want_equal = id == ORC_RULE_ERASE_CREATE;
			uint32_t bits = child_id;
			for (size_t i = 0; i < n; ++i) {
				if ((g.left[i] == g.right[i]) != want_equal)
					continue;
				/* erase_create conjugates when both particles are present, coin when the left one is */
				bool conjugated = want_equal ? (g.left[i] && g.right[i]) : g.left[i];
				bool taken = bits & 1;
				mag *= factor(taken, conjugated);
				if (taken) {
					g.left[i] = !g.left[i];
					g.right[i] = !g.right[i];
				}
				bits >>= 1;
			}
			/* same size, names untouched: patch the parent bytes so that padding is carried over */
			out.assign(p, p + size);
			for (size_t i = 0; i < n; ++i) {
				out[2 + i] = g.left[i];
				out[2 + n + i] = g.right[i];
			}
			return;
		}

		/* split_merge, qcgd.hpp:643-848 */
		uint32_t bits = child_id;
		bool first_split = g.left[0] && g.right[0];
		bool last_merge = !first_split && n > 1 && g.right[0] && g.left[n - 1] && !g.right[n - 1];

		/* the wrap-around site consumes the first bit; a first split that is NOT taken leaves the
		 * bit for the general walk, which then sees node 0 as an ordinary split site (:655-659) */
		first_split = first_split && (bits & 1);
		if (first_split) {
			mag *= factor(true, false);
			bits >>= 1;
		}
		if (last_merge) {
			bool taken = bits & 1;
			mag *= factor(taken, true);
			last_merge = taken;
			bits >>= 1;
		}

		/* magnitude: every site in the walked range contributes, including the node following a
		 * taken merge -- which can never be a site itself (:673-700) */
		{
			uint32_t b = bits;
			for (size_t i = (size_t)first_split + last_merge; i < n - last_merge; ++i) {
				bool split, merge;
				site(g, i, split, merge);
				if (split || merge) {
					mag *= factor(b & 1, merge);
					b >>= 1;
				}
			}
		}

		graph_t c;
		auto emit = [&](bool l, bool r, name_t nm) {
			c.left.push_back(l);
			c.right.push_back(r);
			c.name.push_back(std::move(nm));
		};

		/* when node 0 splits and the most-left element of its name is not 0, the left half goes
		 * to the END of the child (:709-749, 838-845) */
		bool overflow = false;
		if (first_split) {
			const name_t &nm = g.name[0];
			bool most_left_zero = !(nm[0].kind >= 0 && nm[1].hmlz > 0);
			if (most_left_zero) {
				emit(true, false, name_left(nm));
				emit(false, true, name_right(nm));
			} else {
				overflow = true;
				emit(false, true, name_right(nm));
			}
		}
		if (last_merge)
			emit(true, true, name_merge(g.name[n - 1], g.name[0]));

		for (size_t i = (size_t)first_split + last_merge; i < n - last_merge; ++i) {
			bool split, merge;
			site(g, i, split, merge);
			bool taken = false;
			if (split || merge) {
				taken = bits & 1;
				bits >>= 1;
			}
			if (taken && split) {
				emit(true, false, name_left(g.name[i]));
				emit(false, true, name_right(g.name[i]));
			} else if (taken && merge) {
				emit(true, true, name_merge(g.name[i], g.name[i + 1]));
				++i; /* node i+1 was consumed (:816) */
			} else {
				emit(g.left[i], g.right[i], g.name[i]);
			}
		}
		if (overflow)
			emit(true, false, name_left(g.name[0]));

		encode(c, out);
	}
};

/* ------------------------------------------------------------------------------------------
 * modifiers
 * ------------------------------------------------------------------------------------------ */
int apply_modifier(state_t &s, int id, const double *params) {
	for (size_t i = 0; i < s.n(); ++i) {
		uint8_t *b = s.obj(i);
		cplx &mag = s.mag[i];
		switch (id) {
		case ORC_MOD_CNOT: /* quantum_computer.hpp:25-29 */
			b[(size_t)params[1]] ^= b[(size_t)params[0]];
			break;
		case ORC_MOD_XGATE: /* :52-56 */
			b[(size_t)params[0]] = !b[(size_t)params[0]];
			break;
		case ORC_MOD_YGATE: { /* :58-66 */
			size_t bit = (size_t)params[0];
			mag *= cplx(0, 1);
			if (b[bit]) mag *= -1;
			b[bit] = !b[bit];
			break;
		}
		case ORC_MOD_ZGATE: { /* :68-75 -- flips the bit as well, as the reference does */
			size_t bit = (size_t)params[0];
			if (b[bit]) mag *= -1;
			b[bit] = !b[bit];
			break;
		}
		case ORC_MOD_STEP:            /* qcgd.hpp:443-449: left particles move left, right particles move right */
		case ORC_MOD_REVERSED_STEP: { /* qcgd.hpp:451-457 */
			uint16_t n;
			memcpy(&n, b, 2);
			uint8_t *l = b + 2, *r = b + 2 + n;
			if (n == 0) break;
			if (id == ORC_MOD_REVERSED_STEP) std::swap(l, r);
			std::rotate(l, l + 1, l + n);
			std::rotate(r, r + n - 1, r + n);
			break;
		}
		case ORC_MOD_PHASE: /* bench modifier: reads the object, writes only the magnitude */
			if (b[0] & 1) mag *= std::polar(1.0, params[0]);
			break;
		default:
			return -1;
		}
	}
	return 0;
}

double g_last_seconds = 0;

/* ------------------------------------------------------------------------------------------
 * one rule iteration (SURVEY appendix A.2 / quids.hpp:448-543), simple truncation
 * ------------------------------------------------------------------------------------------ */
int simulate(const state_t &in, const rule_t &rule, state_t &out, uint64_t k, double tolerance, uint64_t *counters) {
	size_t np = in.n();

	/* 1. child counts (quids.hpp:548-569) */
	std::vector<uint32_t> num_child(np);
	for (size_t i = 0; i < np; ++i) {
		uint32_t ub;
		num_child[i] = rule.num_child(in.obj(i), in.size(i), ub);
	}

	/* 2. parent pre-truncation: the k most probable parents (quids.hpp:613-642); ties arbitrary */
	std::vector<size_t> kept(np);
	std::iota(kept.begin(), kept.end(), 0);
	if (k < np) {
		std::stable_sort(kept.begin(), kept.end(), [&](size_t a, size_t b) { return std::norm(in.mag[a]) > std::norm(in.mag[b]); });
		kept.resize(k);
		std::sort(kept.begin(), kept.end());
	}

	/* 3-4. symbolic children (quids.hpp:647-721) */
	struct sym_t {
		uint64_t hash;
		cplx mag;
		size_t parent;
		uint32_t child_id;
	};
	std::vector<sym_t> sym;
	std::vector<uint8_t> scratch;
	for (size_t p : kept)
		for (uint32_t c = 0; c < num_child[p]; ++c) {
			cplx mag = in.mag[p];
			rule.child(in.obj(p), in.size(p), c, scratch, mag);
			sym.push_back(sym_t{rule.hasher(scratch.data(), (uint32_t)scratch.size()), mag, p, c});
		}
	counters[0] = sym.size();

	/* 5. interference: equal 64-bit hash = same object, magnitudes add; the first child seen
	 * represents the group (quids.hpp:785-809) */
	std::unordered_map<uint64_t, size_t> first;
	first.reserve(sym.size());
	std::vector<size_t> reps;
	for (size_t i = 0; i < sym.size(); ++i) {
		auto ins = first.emplace(sym[i].hash, i);
		if (ins.second)
			reps.push_back(i);
		else
			sym[ins.first->second].mag += sym[i].mag;
	}

	/* 6. tolerance, strict > on re^2+im^2 (quids.hpp:819-823) */
	std::vector<size_t> alive;
	for (size_t i : reps)
		if (std::norm(sym[i].mag) > tolerance)
			alive.push_back(i);
	counters[1] = alive.size();

	/* 7. post-truncation: k most probable (quids.hpp:866-900); ties arbitrary */
	if (k < alive.size()) {
		std::stable_sort(alive.begin(), alive.end(), [&](size_t a, size_t b) { return std::norm(sym[a].mag) > std::norm(sym[b].mag); });
		alive.resize(k);
	}

	/* 8. materialise in ascending symbolic index (quids.hpp:927-967) */
	std::sort(alive.begin(), alive.end());
	out.clear();
	for (size_t i : alive) {
		cplx dummy = 1;
		rule.child(in.obj(sym[i].parent), in.size(sym[i].parent), sym[i].child_id, scratch, dummy);
		out.push(scratch.data(), (uint32_t)scratch.size(), sym[i].mag);
	}

	/* 9. normalise; total_proba keeps the pre-normalisation sum (quids.hpp:985-1017) */
	double total = 0;
	for (auto &m : out.mag)
		total += std::norm(m);
	out.total_proba = total;
	double f = std::sqrt(total);
	if (out.n() > 0 && f != 1)
		for (auto &m : out.mag)
			m /= f;
	return 0;
}

} // namespace

struct orc_state {
	state_t s;
};

extern "C" {

const char *orc_kind(void) { return "port"; }
int orc_num_threads(void) { return 1; }
int orc_set_num_threads(int) { return 1; }

orc_state *orc_state_create(void) { return new orc_state(); }
void orc_state_destroy(orc_state *s) { delete s; }

int orc_state_load(orc_state *s, uint64_t n, const uint32_t *sizes, const double *mags, const uint8_t *bytes) {
	s->s.clear();
	uint64_t off = 0;
	for (uint64_t i = 0; i < n; ++i) {
		s->s.push(bytes + off, sizes[i], cplx(mags[2 * i], mags[2 * i + 1]));
		off += sizes[i];
	}
	return 0;
}
uint64_t orc_state_num_object(const orc_state *s) { return s->s.n(); }
uint64_t orc_state_num_bytes(const orc_state *s) { return s->s.bytes.size(); }
double orc_state_total_proba(const orc_state *s) { return s->s.total_proba; }
int orc_state_store(const orc_state *s, uint32_t *sizes, double *mags, uint8_t *bytes) {
	for (size_t i = 0; i < s->s.n(); ++i) {
		sizes[i] = s->s.size(i);
		mags[2 * i] = s->s.mag[i].real();
		mags[2 * i + 1] = s->s.mag[i].imag();
	}
	if (!s->s.bytes.empty())
		memcpy(bytes, s->s.bytes.data(), s->s.bytes.size());
	return 0;
}

int orc_state_pop(orc_state *s, uint64_t n, int normalize) {
	state_t &st = s->s;
	if (n > st.n())
		return -1;
	if (n < 1) /* quids.hpp:195-196 */
		return 0;
	const size_t left = st.n() - n;
	st.bytes.resize(st.begin[left]);
	st.begin.resize(left + 1);
	st.mag.resize(left);
	if (normalize) { /* quids.hpp:985-1017 */
		st.total_proba = 0;
		for (size_t i = 0; i < left; ++i)
			st.total_proba += std::norm(st.mag[i]);
		const double f = std::sqrt(st.total_proba);
		if (left > 0 && f != 1)
			for (size_t i = 0; i < left; ++i)
				st.mag[i] /= f;
	}
	return 0;
}

int orc_qcgd_random_state(orc_state *s, uint32_t n_node, uint64_t n_graphs, uint32_t seed, double re, double im) {
	/* qcgd.hpp:1129-1136: all graphs are appended first, THEN randomised in storage order */
	graph_t g;
	for (uint32_t i = 0; i < n_node; ++i) {
		g.left.push_back(0);
		g.right.push_back(0);
		g.name.push_back(name_t{atom_element((int16_t)i)});
	}
	std::vector<uint8_t> enc;
	encode(g, enc);
	s->s.clear();
	for (uint64_t i = 0; i < n_graphs; ++i)
		s->s.push(enc.data(), (uint32_t)enc.size(), cplx(re, im));
	srand(seed);
	for (uint64_t i = 0; i < n_graphs; ++i) {
		uint8_t *b = s->s.obj(i);
		for (uint32_t j = 0; j < n_node; ++j) { /* qcgd.hpp:114-120: left then right, per node */
			b[2 + j] = rand() & 1;
			b[2 + n_node + j] = rand() & 1;
		}
	}
	return 0;
}

int orc_hash_objects(const orc_state *s, int rule_id, const double *params, uint64_t *hashes) {
	rule_t rule(rule_id, params);
	for (size_t i = 0; i < s->s.n(); ++i)
		hashes[i] = rule.hasher(s->s.obj(i), s->s.size(i));
	return 0;
}

int orc_apply_modifier(orc_state *s, int modifier_id, const double *params) { return apply_modifier(s->s, modifier_id, params); }

/* quids.hpp:208-234 with the observables of qcgd utils::serialize (qcgd.hpp:319-345): a plain sequential sum */
int orc_average_value(const orc_state *s, int observable_id, const double *params, double *value) {
	double avg = 0;
	for (size_t i = 0; i < s->s.n(); ++i) {
		const uint8_t *o = s->s.obj(i);
		const uint32_t size = s->s.size(i);
		double v = 0;
		if (observable_id >= ORC_OBS_QCGD_SIZE && observable_id <= ORC_OBS_QCGD_SQUARED_DENSITY) {
			const uint32_t n = (uint32_t)o[0] | ((uint32_t)o[1] << 8);
			double density = 0;
			for (uint32_t k = 0; k < n; ++k)
				density += o[2 + k] + o[2 + n + k];
			density /= 2 * (double)n;
			v = observable_id == ORC_OBS_QCGD_SIZE ? (double)n : observable_id == ORC_OBS_QCGD_SQUARED_SIZE ? (double)n * n : observable_id == ORC_OBS_QCGD_DENSITY ? density : density * density;
		} else if (observable_id == ORC_OBS_QUBIT) {
			const uint64_t bit = (uint64_t)params[0];
			v = bit < size && o[bit] ? 1 : 0;
		} else if (observable_id == ORC_OBS_BYTES) {
			v = size;
		} else {
			return -1;
		}
		avg += v * std::norm(s->s.mag[i]);
	}
	*value = avg;
	return 0;
}

int orc_simulate(orc_state *in, int rule_id, const double *params, orc_state *out, uint64_t max_num_object, double tolerance, uint64_t *counters) {
	if (max_num_object == 0)
		return -2;
	if (rule_id < ORC_RULE_HADAMARD || rule_id > ORC_RULE_SPLIT_MERGE)
		return -1;
	rule_t rule(rule_id, params);
	auto t0 = std::chrono::steady_clock::now();
	int rc = simulate(in->s, rule, out->s, max_num_object, tolerance, counters);
	g_last_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	return rc;
}

double orc_last_simulate_seconds(void) { return g_last_seconds; }
}
