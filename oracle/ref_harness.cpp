/*
 * ref_harness.cpp -- the UNMODIFIED QuIDS reference behind the oracle_api.h interface.
 *
 * TEST INFRASTRUCTURE ONLY.  This translation unit contains no QuIDS code: it includes the
 * reference headers where they lie (-I/root/reference/src, see oracle/Makefile) and forwards the
 * oracle_api.h calls to the reference's own public API (quids::it_t, quids::simulate, the rule
 * classes of rules/quantum_computer.hpp and rules/qcgd.hpp).  The resulting shared object is
 * written to oracle/_ref/ (git-ignored, shipped to the GPU box by gpurun) and is used
 *   - to pin oracle.cpp (tests/test_oracle.py, oracle/gen_golden.py),
 *   - as the CPU baseline of bench.py (cpu_baseline.kind = "reference", all host threads).
 */
#include "oracle_api.h"

#include <chrono>
#include <cstring>

#include "quids.hpp"
#include "rules/qcgd.hpp"
#include "rules/quantum_computer.hpp"

#include <omp.h>

/* quids::iteration::append() re-runs resize() -- which re-initialises truncated_oid with an iota over
 * the whole state (quids.hpp:279-306) -- for every object, so building an n-object state through the
 * public API costs O(n^2).  This subclass fills the reference's own (protected) arrays in one go, with
 * exactly the layout append() produces (quids.hpp:174-188); the code under test, quids::simulate, is
 * untouched and sees a plain quids::it_t&. */
struct bulk_iteration : quids::it_t {
	void load(uint64_t n, const uint32_t *sizes, const double *mags, const uint8_t *bytes) {
		uint64_t total = 0;
		for (uint64_t i = 0; i < n; ++i)
			total += sizes[i] + quids::get_alignment_offset(sizes[i]);
		num_object = n;
		resize(n);
		allocate(total);
		uint64_t src = 0, dst = 0;
		object_begin[0] = 0;
		for (uint64_t i = 0; i < n; ++i) {
			memcpy(&objects[dst], bytes + src, sizes[i]);
			magnitude[i] = quids::mag_t(mags[2 * i], mags[2 * i + 1]);
			object_size[i] = sizes[i];
			src += sizes[i];
			dst += sizes[i] + quids::get_alignment_offset(sizes[i]);
			object_begin[i + 1] = dst;
		}
	}
};

struct orc_state {
	bulk_iteration it;
};

namespace {
double g_last_seconds = 0;

/* rules are tiny and, as in the reference drivers, never deleted (quids::rule has no virtual destructor) */
quids::rule_t *make_rule(int id, const double *p) {
	switch (id) {
	case ORC_RULE_HADAMARD: return new quids::rules::quantum_computer::hadamard((size_t)p[0]);
	case ORC_RULE_ERASE_CREATE: return new quids::rules::qcgd::erase_create(p[0], p[1], p[2]);
	case ORC_RULE_COIN: return new quids::rules::qcgd::coin(p[0], p[1], p[2]);
	case ORC_RULE_SPLIT_MERGE: return new quids::rules::qcgd::split_merge(p[0], p[1], p[2]);
	}
	return nullptr;
}

void clear(quids::it_t &it) {
	if (it.num_object > 0)
		it.pop(it.num_object, false);
}
} // namespace

extern "C" {

const char *orc_kind(void) { return "reference"; }
int orc_num_threads(void) { return omp_get_max_threads(); }
int orc_set_num_threads(int n) {
	omp_set_num_threads(n > 0 ? n : omp_get_num_procs());
	return omp_get_max_threads();
}

orc_state *orc_state_create(void) { return new orc_state(); }
void orc_state_destroy(orc_state *s) { delete s; }

int orc_state_load(orc_state *s, uint64_t n, const uint32_t *sizes, const double *mags, const uint8_t *bytes) {
	if (n <= 4096) { /* small states go through the public append(), like a driver */
		clear(s->it);
		uint64_t off = 0;
		for (uint64_t i = 0; i < n; ++i) {
			const char *b = (const char *)bytes + off;
			s->it.append(b, b + sizes[i], quids::mag_t(mags[2 * i], mags[2 * i + 1]));
			off += sizes[i];
		}
	} else
		s->it.load(n, sizes, mags, bytes);
	return 0;
}

uint64_t orc_state_num_object(const orc_state *s) { return s->it.num_object; }

uint64_t orc_state_num_bytes(const orc_state *s) {
	uint64_t total = 0;
	for (size_t i = 0; i < s->it.num_object; ++i) {
		char const *b;
		uint size;
		quids::mag_t mag;
		s->it.get_object(i, b, size, mag);
		total += size;
	}
	return total;
}

double orc_state_total_proba(const orc_state *s) { return s->it.total_proba; }

int orc_state_store(const orc_state *s, uint32_t *sizes, double *mags, uint8_t *bytes) {
	uint64_t off = 0;
	for (size_t i = 0; i < s->it.num_object; ++i) {
		char const *b;
		uint size;
		quids::mag_t mag;
		s->it.get_object(i, b, size, mag);
		sizes[i] = size;
		mags[2 * i] = mag.real();
		mags[2 * i + 1] = mag.imag();
		memcpy(bytes + off, b, size);
		off += size;
	}
	return 0;
}

int orc_state_pop(orc_state *s, uint64_t n, int normalize) {
	if (n > s->it.num_object)
		return -1;
	s->it.pop(n, normalize != 0);
	return 0;
}

int orc_qcgd_random_state(orc_state *s, uint32_t n_node, uint64_t n_graphs, uint32_t seed, double re, double im) {
	clear(s->it);
	for (uint64_t i = 0; i < n_graphs; ++i) {
		char *b, *e;
		quids::rules::qcgd::utils::make_graph(b, e, n_node);
		s->it.append(b, e, quids::mag_t(re, im));
		delete[] b;
	}
	std::srand(seed);
	quids::rules::qcgd::utils::randomize(s->it);
	return 0;
}

int orc_hash_objects(const orc_state *s, int rule_id, const double *params, uint64_t *hashes) {
	quids::rule_t *rule = make_rule(rule_id, params);
	if (!rule)
		return -1;
	for (size_t i = 0; i < s->it.num_object; ++i) {
		char const *b;
		uint size;
		quids::mag_t mag;
		s->it.get_object(i, b, size, mag);
		hashes[i] = rule->hasher(b, b + size);
	}
	return 0;
}

int orc_apply_modifier(orc_state *s, int modifier_id, const double *params) {
	namespace qc = quids::rules::quantum_computer;
	switch (modifier_id) {
	case ORC_MOD_CNOT: quids::simulate(s->it, qc::cnot((uint32_t)params[0], (uint32_t)params[1])); break;
	case ORC_MOD_XGATE: quids::simulate(s->it, qc::Xgate((size_t)params[0])); break;
	case ORC_MOD_YGATE: quids::simulate(s->it, qc::Ygate((size_t)params[0])); break;
	case ORC_MOD_ZGATE: quids::simulate(s->it, qc::Zgate((size_t)params[0])); break;
	case ORC_MOD_STEP: quids::simulate(s->it, quids::rules::qcgd::step); break;
	case ORC_MOD_REVERSED_STEP: quids::simulate(s->it, quids::rules::qcgd::reversed_step); break;
	case ORC_MOD_PHASE: {
		quids::mag_t phase = std::polar(1.0, params[0]);
		quids::simulate(s->it, [phase](char *b, char *e, quids::mag_t &mag) {
			if (b[0] & 1)
				mag *= phase;
		});
		break;
	}
	default: return -1;
	}
	return 0;
}

/* the reference's own iteration::average_value with the lambdas of its utils::serialize (qcgd.hpp:319-345) */
int orc_average_value(const orc_state *s, int observable_id, const double *params, double *value) {
	namespace graphs = quids::rules::qcgd::graphs;
	auto density = [](char const *object_begin) {
		PROBA_TYPE num_nodes = graphs::num_nodes(object_begin);
		PROBA_TYPE d = 0;
		for (auto i = 0; i < num_nodes; ++i)
			d += graphs::left(object_begin, i) + graphs::right(object_begin, i);
		return d / (2 * num_nodes);
	};
	const uint64_t bit = params ? (uint64_t)params[0] : 0;
	switch (observable_id) {
	case ORC_OBS_QCGD_SIZE: *value = s->it.average_value([](char const *b, char const *) { return (PROBA_TYPE)graphs::num_nodes(b); }); break;
	case ORC_OBS_QCGD_SQUARED_SIZE: *value = s->it.average_value([](char const *b, char const *) { PROBA_TYPE n = graphs::num_nodes(b); return n * n; }); break;
	case ORC_OBS_QCGD_DENSITY: *value = s->it.average_value([&](char const *b, char const *) { return density(b); }); break;
	case ORC_OBS_QCGD_SQUARED_DENSITY: *value = s->it.average_value([&](char const *b, char const *) { PROBA_TYPE d = density(b); return d * d; }); break;
	case ORC_OBS_QUBIT: *value = s->it.average_value([bit](char const *b, char const *e) { return (PROBA_TYPE)(bit < (uint64_t)(e - b) && b[bit] ? 1 : 0); }); break;
	case ORC_OBS_BYTES: *value = s->it.average_value([](char const *b, char const *e) { return (PROBA_TYPE)(e - b); }); break;
	default: return -1;
	}
	return 0;
}

int orc_simulate(orc_state *in, int rule_id, const double *params, orc_state *out, uint64_t max_num_object, double tolerance, uint64_t *counters) {
	if (max_num_object == 0)
		return -2;
	quids::rule_t *rule = make_rule(rule_id, params);
	if (!rule)
		return -1;
	quids::tolerance = tolerance;
	quids::simple_truncation = true;
	static quids::sy_it_t sy_it; /* reused across calls like a driver would */
	auto t0 = std::chrono::steady_clock::now();
	quids::simulate(in->it, rule, out->it, sy_it, (size_t)max_num_object);
	g_last_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	counters[0] = sy_it.num_object;
	counters[1] = sy_it.num_object_after_interferences;
	return 0;
}

double orc_last_simulate_seconds(void) { return g_last_seconds; }
}
