/*
 * oracle_api.h -- C interface shared by the two CPU checkers of this repository.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * The same symbols are implemented twice, in two separate shared objects:
 *   oracle/liboracle.so        (oracle.cpp)       a from-scratch CPU restatement of the
 *                                                 rule-application step of QuIDS
 *   oracle/_ref/libquids_ref.so (ref_harness.cpp) the UNMODIFIED reference headers from
 *                                                 /root/reference/src, compiled where they lie
 * so that the restatement can be pinned against the real reference on identical inputs.
 *
 * States cross this interface "packed": object bytes back to back WITHOUT alignment padding,
 * one u32 size and one (re, im) double pair per object.
 */
#ifndef QUIDS_ORACLE_API_H
#define QUIDS_ORACLE_API_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* rule ids (params are doubles) */
enum {
	ORC_RULE_HADAMARD     = 1, /* params: [bit]                 quantum_computer.hpp:31-50  */
	ORC_RULE_ERASE_CREATE = 2, /* params: [theta, phi, xi]      qcgd.hpp:459-532            */
	ORC_RULE_COIN         = 3, /* params: [theta, phi, xi]      qcgd.hpp:534-605            */
	ORC_RULE_SPLIT_MERGE  = 4  /* params: [theta, phi, xi]      qcgd.hpp:607-1036           */
};

/* modifier ids */
enum {
	ORC_MOD_CNOT          = 1, /* params: [control, target]     quantum_computer.hpp:25-29  */
	ORC_MOD_XGATE         = 2, /* params: [bit]                 quantum_computer.hpp:52-56  */
	ORC_MOD_YGATE         = 3, /* params: [bit]                 quantum_computer.hpp:58-66  */
	ORC_MOD_ZGATE         = 4, /* params: [bit]                 quantum_computer.hpp:68-75  */
	ORC_MOD_STEP          = 5, /* params: none                  qcgd.hpp:443-449            */
	ORC_MOD_REVERSED_STEP = 6, /* params: none                  qcgd.hpp:451-457            */
	ORC_MOD_PHASE         = 7  /* params: [theta]  mag *= polar(1, theta) when (obj[0]&1); bench modifier */
};

typedef struct orc_state orc_state;

/* which implementation is behind this .so: "port" (oracle.cpp) or "reference" (ref_harness.cpp) */
const char *orc_kind(void);
/* number of host threads the implementation will use for orc_simulate */
int orc_num_threads(void);
/* ask for `n` host threads in the following orc_simulate calls (n <= 0: all the cores of the machine); returns what
 * orc_num_threads() now reports.  torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: bench.py's
 * reference arm undoes that here.  The port is single-threaded and always returns 1. */
int orc_set_num_threads(int n);

orc_state *orc_state_create(void);
void orc_state_destroy(orc_state *s);
/* replace the content of s by n packed objects */
int orc_state_load(orc_state *s, uint64_t n, const uint32_t *sizes, const double *mags, const uint8_t *bytes);
uint64_t orc_state_num_object(const orc_state *s);
uint64_t orc_state_num_bytes(const orc_state *s); /* sum of object sizes, no padding */
double orc_state_total_proba(const orc_state *s);
/* copy the state out, packed, in storage order */
int orc_state_store(const orc_state *s, uint32_t *sizes, double *mags, uint8_t *bytes);

/* iteration::pop (quids.hpp:194-203): drop the last n objects, then (normalize != 0) iteration::normalize (quids.hpp:985-1017):
 * total_proba = sum of |mag|^2 over what is left, magnitudes divided by its square root unless that is exactly 1 */
int orc_state_pop(orc_state *s, uint64_t n, int normalize);

/* n_graphs fresh n_node graphs (make_graph, qcgd.hpp:214-230) with magnitude (re, im), then
 * left/right bits drawn with glibc srand(seed); rand()&1 in the order of qcgd.hpp:114-120,232-240 */
int orc_qcgd_random_state(orc_state *s, uint32_t n_node, uint64_t n_graphs, uint32_t seed, double re, double im);

/* hash of every object as rule->hasher sees it */
int orc_hash_objects(const orc_state *s, int rule_id, const double *params, uint64_t *hashes);

/* in place, quids.hpp:973-980 */
int orc_apply_modifier(orc_state *s, int modifier_id, const double *params);

/* observables (params are doubles) */
enum {
	ORC_OBS_QCGD_SIZE            = 1, /* number of nodes                       qcgd.hpp:319-321 */
	ORC_OBS_QCGD_SQUARED_SIZE    = 2, /* its square                            qcgd.hpp:322-325 */
	ORC_OBS_QCGD_DENSITY         = 3, /* particles / (2 nodes)                 qcgd.hpp:326-335 */
	ORC_OBS_QCGD_SQUARED_DENSITY = 4, /* its square                            qcgd.hpp:336-345 */
	ORC_OBS_QUBIT                = 5, /* params: [bit]; 1 when the qubit is set (byte != 0 inside the object), else 0 */
	ORC_OBS_BYTES                = 6  /* size of the object in bytes */
};
/* iteration::average_value, quids.hpp:208-234: sum over the objects of observable(begin, end) * norm(magnitude) */
int orc_average_value(const orc_state *s, int observable_id, const double *params, double *value);

/* one rule iteration, quids.hpp:448-543, simple truncation.
 * max_num_object: UINT64_MAX = no truncation (0 = auto budget is NOT supported: returns -2).
 * counters[0] = N_c (sy_it.num_object), counters[1] = N_u (num_object_after_interferences). */
int orc_simulate(orc_state *in, int rule_id, const double *params, orc_state *out,
                 uint64_t max_num_object, double tolerance, uint64_t *counters);

/* seconds spent inside the last orc_simulate call (wall clock, steady_clock) */
double orc_last_simulate_seconds(void);

#ifdef __cplusplus
}
#endif
#endif
