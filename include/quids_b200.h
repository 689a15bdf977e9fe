/*
 * quids_b200.h -- C ABI of the B200-native QuIDS rule-application step (libquids_b200.so).
 *
 * The reference (jolatechno/QuIDS) has no FFI: its boundary is the header-only C++ API of
 * src/quids.hpp / src/quids_mpi.hpp.  The drop-in headers of this repository (the .hpp files under include/quids)
 * keep that C++ surface and call the functions below; every entry point names the reference
 * interface it stands behind (file:line into /root/reference/src).
 *
 * Conventions
 *   - every function returns 0 on success, a negative qb_status otherwise; qb_last_error() gives
 *     the message (thread local).  The C++ layer turns non-zero into std::runtime_error, the only
 *     exception the reference itself throws (utils/vector.hpp:126-127).
 *   - one qb_ctx <-> one GPU <-> one host thread.  Calls are synchronous at the API: counters are
 *     valid on return (quids.hpp:152-154,344-346 are plain public members in the reference).
 *   - plain pointers and sizes only.  Magnitudes cross as interleaved (re, im) doubles, layout
 *     identical to std::complex<double> (PROBA_TYPE = double, quids.hpp:21-23,78).
 *   - states cross in the reference's own storage layout (quids.hpp:266-276, appendix A.1 of
 *     SURVEY.md): `objects` = object bytes padded to the caller's alignment, object_begin[n+1],
 *     object_size[n], magnitude[n].
 *   - there is NO CPU fallback: every compute entry point fails with QB_ERR_CUDA when no device
 *     is usable.
 */
#ifndef QUIDS_B200_H
#define QUIDS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum qb_status {
	QB_OK = 0,
	QB_ERR_CUDA = -1,        /* a CUDA runtime call failed (includes "no device") */
	QB_ERR_ARG = -2,         /* invalid argument */
	QB_ERR_UNKNOWN_RULE = -3,
	QB_ERR_UNSUPPORTED = -4, /* the request is outside what this build implements */
	QB_ERR_CAPACITY = -5,    /* an internal limit was exceeded (child index >= 2^40, object >= 16 MiB, table full) */
	QB_ERR_COMM = -6         /* NCCL failure on the distributed path */
} qb_status;

typedef struct qb_ctx qb_ctx;   /* device context: stream, workspace              */
typedef struct qb_iter qb_iter; /* quids::iteration (quids.hpp:149-335) in HBM     */
typedef struct qb_sym qb_sym;   /* quids::symbolic_iteration (quids.hpp:338-429)   */
typedef struct qb_comm qb_comm; /* stands where MPI_Comm stands in quids_mpi.hpp   */

#define QB_NO_TRUNCATION UINT64_MAX /* max_num_object = -1 in the reference (quids.hpp:445) */
/* max_num_object = 0 is the reference's "automatic" budget (quids.hpp:459-485,510-536: truncate to what fits in free RAM).
 * Here the budget is the GPU memory left after the safety margin (cudaMemGetInfo, or qb_options.memory_budget): the most
 * probable parents whose symbolic workspace (interference table at its safe size, compacted lists, work items) fits are
 * kept, then as many of the most probable children as the next state has room for.  QB_ERR_CAPACITY only when not even
 * one parent's children fit.  On the distributed path the budget is per rank for the parents and agreed between
 * the ranks for the children (qb_simulate_dist). */

/* the mutable namespace globals of the reference that influence one call (quids.hpp:60-75) */
typedef struct qb_options {
	double tolerance;          /* quids::tolerance, strict > on re^2+im^2 (quids.hpp:62,821)        */
	uint32_t align_byte_length; /* quids::align_byte_length (quids.hpp:60,93-102)                    */
	int32_t simple_truncation; /* quids::simple_truncation (quids.hpp:69-75): 1 = keep the most probable; 0 = probabilistic: keep the
	                              smallest u / |mag|^2, u uniform in (0,1) (quids.hpp:594-608,829-845), drawn from `seed` */
	double table_load;         /* engine knob: max load factor of the interference table, 0 = default */
	int32_t profile;           /* 1: record CUDA events at the phase boundaries (qb_sym_phase_ms)     */
	float safety_margin;       /* quids::safety_margin (quids.hpp:64): fraction of GPU memory the automatic budget leaves free */
	uint32_t seed;             /* probabilistic truncation: seed of the counter-based generator (the reference seeds from rand()) */
	int32_t locality_sort;     /* engine knob: process the child groups in the order of the rule's group key so that equal
	                              objects are merged on chip before the table; 0 off, 1 when there are >= 2^16 groups (default), 2 always */
	int32_t binned_inserts;    /* engine knob: one-child-per-lane rules (split_merge, hadamard, rules written with the four reference methods)
	                              send their children to small bins that are deduplicated in shared memory and written out as a dense
	                              array of unique children (the reference's bucket partition, quids.hpp:755-809; table.cuh); 0 off,
	                              1 when there are >= 2^22 children and no history of heavy duplication (default), 2 always */
	int32_t family_routing;    /* engine knob, qb_simulate_dist: rules with families (erase_create, coin) move the PARENTS to the rank that
	                              owns their family, so that interference needs no exchange of children; 1 on (default), 0 off */
	uint64_t memory_budget;    /* engine knob: bytes the automatic budget (max_num_object = 0) may spend on the symbolic workspace and
	                              on the next state; 0 = measured (cudaMemGetInfo minus safety_margin of the GPU) */
	/* load balancing at the head of quids::mpi::simulate (quids_mpi.hpp:442-500); only qb_simulate_dist reads these */
	int32_t equalize;            /* 0 off (default of the C ABI), 1 by objects (quids::mpi::equalize_children = false), 2 by children */
	float equalize_inbalance;    /* quids::mpi::equalize_inbalance (quids_mpi.hpp:48): stop below this (max - avg) / max      */
	float min_equalize_step;     /* quids::mpi::min_equalize_step (:50): stop when a round improves the gap by less than this */
	uint64_t min_equalize_size;  /* quids::mpi::min_equalize_size (:46): do nothing when every rank holds fewer objects        */
} qb_options;

void qb_options_default(qb_options *opt);

/* phase callback = quids::debug_t mid_step_function (quids.hpp:90); labels and order of SURVEY
 * section 5 are kept.  Called on the calling thread, after the stream has been drained. */
typedef void (*qb_step_cb)(const char *label, void *user);

const char *qb_last_error(void);
int qb_version(void);
int qb_device_count(void); /* 0 when no CUDA device is usable */

int qb_ctx_create(int device, qb_ctx **out);
int qb_ctx_destroy(qb_ctx *ctx);
int qb_ctx_synchronize(qb_ctx *ctx);
/* cudaStream_t every kernel of this context is launched on (for CUDA-event timing by the caller) */
void *qb_ctx_stream(qb_ctx *ctx);
/* number of kernels this context has launched so far */
uint64_t qb_ctx_launch_count(const qb_ctx *ctx);

/* pinned host memory for fast transfers (plain cudaHostAlloc / cudaFreeHost) */
int qb_host_alloc(size_t bytes, void **out);
int qb_host_free(void *p);

/* ---- quids::iteration --------------------------------------------------------------------- */
int qb_iter_create(qb_ctx *ctx, qb_iter **out); /* iteration() quids.hpp:157-161 */
int qb_iter_destroy(qb_iter *it);
/* replaces the state (what append() builds on the host, quids.hpp:174-188) */
int qb_iter_upload(qb_iter *it, uint64_t num_object, const uint8_t *objects, uint64_t num_bytes,
                   const uint64_t *object_begin, const uint32_t *object_size, const double *magnitude,
                   double total_proba);
/* num_object, total_proba (quids.hpp:152-154) and the padded byte length object_begin[num_object] */
int qb_iter_counts(const qb_iter *it, uint64_t *num_object, uint64_t *num_bytes, double *total_proba);
/* copies the state out (what get_object() reads, quids.hpp:242-258); any pointer may be NULL */
int qb_iter_download(const qb_iter *it, uint8_t *objects, uint64_t *object_begin, uint32_t *object_size, double *magnitude);
/* device pointers of the four arrays (for zero-copy consumers; valid until the next call that writes `it`) */
/* PROBA_TYPE = float (quids.hpp:21-23): the same two transfers with complex<float> magnitudes (2 x f32 per object).
 * The state in HBM and the device arithmetic stay double, so a float build of a driver agrees with the reference's
 * float build to the reference's own rounding (1e-5 relative). */
int qb_iter_upload_f32(qb_iter *it, uint64_t num_object, const uint8_t *objects, uint64_t num_bytes,
                       const uint64_t *object_begin, const uint32_t *object_size, const float *magnitude, double total_proba);
int qb_iter_download_f32(const qb_iter *it, uint8_t *objects, uint64_t *object_begin, uint32_t *object_size, float *magnitude);
/* The same two transfers on dedicated copy streams, overlapping the rule iterations of OTHER states (double buffering:
 * upload the input of step i+1 and download the result of step i-1 while step i computes).  The host arrays must be
 * page-locked (qb_host_alloc) and stay untouched until qb_iter_wait(it) returns; every later call that uses `it` orders
 * itself after the transfers on the device, so only the HOST side ever needs qb_iter_wait.  Counters (qb_iter_counts)
 * are valid as soon as qb_iter_upload_async returns. */
int qb_iter_upload_async(qb_iter *it, uint64_t num_object, const uint8_t *objects, uint64_t num_bytes,
                         const uint64_t *object_begin, const uint32_t *object_size, const double *magnitude, double total_proba);
int qb_iter_download_async(const qb_iter *it, uint8_t *objects, uint64_t *object_begin, uint32_t *object_size, double *magnitude);
int qb_iter_wait(const qb_iter *it);
/* iteration::append (quids.hpp:174-188) for a whole state at once, HBM to HBM: the objects of `other` are appended to
 * `it` with their magnitudes (no normalisation; total_proba of `it` is kept) */
int qb_iter_append_state(qb_iter *it, const qb_iter *other);
/* the four arrays of the state in HBM (quids.hpp:266-276 layout: objects, object_begin u64[n+1], object_size u32[n], magnitude
 * 2 x f64 [n]), for device code that reads or modifies a state in place (quids/device/plugin.cuh: modifier lambdas); the
 * pointers are invalidated by the next call that resizes the state */
int qb_iter_device_ptrs(const qb_iter *it, void **objects, void **object_begin, void **object_size, void **magnitude);
/* the context a state belongs to, and the CUDA device / stream of a context (device plug-ins launch on that stream) */
qb_ctx *qb_iter_ctx(const qb_iter *it);
int qb_ctx_device(const qb_ctx *ctx);
/* pop(n, normalize) quids.hpp:194-203 */
int qb_iter_pop(qb_iter *it, uint64_t n, int normalize);
/* normalize() quids.hpp:985-1017 */
int qb_iter_normalize(qb_iter *it);

/* ---- quids::symbolic_iteration ------------------------------------------------------------- */
int qb_sym_create(qb_ctx *ctx, qb_sym **out);
int qb_sym_destroy(qb_sym *sym);
/* num_object, num_object_after_interferences (quids.hpp:344-346) */
int qb_sym_counts(const qb_sym *sym, uint64_t *num_object, uint64_t *num_object_after_interferences);

/* per-phase device times of the last qb_simulate (needs options.profile = 1) */
enum {
	QB_PHASE_NUM_CHILD = 0, /* get_num_child kernel + scan           quids.hpp:548-569        */
	QB_PHASE_PRE_TRUNCATE,  /* parent top-k (+ ordering of groups)   quids.hpp:613-642        */
	QB_PHASE_TABLE_CLEAR,   /* interference table reset                                       */
	QB_PHASE_SYMBOLIC,      /* children -> (hash, mag) -> table      quids.hpp:647-721,785-809 */
	QB_PHASE_COMPACT,       /* tolerance filter + compaction         quids.hpp:819-823        */
	QB_PHASE_TRUNCATE,      /* child top-k                           quids.hpp:866-900        */
	QB_PHASE_FINALIZE,      /* sizes, scan, populate_child_simple    quids.hpp:905-968        */
	QB_PHASE_NORMALIZE,     /* quids.hpp:985-1017                                             */
	QB_PHASE_EXCHANGE,      /* distributed: all-to-allv of records and survivors   quids_mpi.hpp:741-743,842 */
	QB_PHASE_OWNER,         /* distributed: owner-side merge, tolerance, return lists   quids_mpi.hpp:762-830 */
	QB_PHASE_INSERT,        /* binned inserts: the children's records reach the table region by region   quids.hpp:755-809 */
	QB_PHASE_COUNT
};
int qb_sym_phase_ms(const qb_sym *sym, float *ms /* [QB_PHASE_COUNT] */);
/* bytes of HBM currently held by the symbolic workspace */
uint64_t qb_sym_device_bytes(const qb_sym *sym);

/* ---- rules and modifiers -------------------------------------------------------------------- */
/* registry lookups; built in: rules "hadamard", "erase_create", "coin", "split_merge" (+ "_generic"
 * variants that run the four reference methods without the fused symbolic hook);
 * modifiers "cnot", "xgate", "ygate", "zgate", "step", "reversed_step", "phase".
 * Parameters are passed as an array of doubles, in the order of the reference constructors:
 *   hadamard(bit) quantum_computer.hpp:35 | erase_create/coin/split_merge(theta, phi, xi) qcgd.hpp:466,541,614
 *   cnot(control, target) :25 | xgate/ygate/zgate(bit) :52,58,68 | phase(theta) */
int qb_rule_id(const char *name);     /* >= 1, or QB_ERR_UNKNOWN_RULE */
int qb_modifier_id(const char *name); /* >= 1, or QB_ERR_UNKNOWN_RULE */

/* quids::simulate(it_t&, modifier_t)  quids.hpp:436-438 / apply_modifier :973-980 */
int qb_apply_modifier(qb_iter *it, int modifier_id, const double *params, uint32_t num_params);

/* iteration::average_value(observable) (quids.hpp:208-234: sum over the objects of observable(begin, end) * |mag|^2) for an
 * observable registered on the device (the reference's observable_t is a host closure; the drop-in headers keep that
 * overload, evaluated on the host mirror).  An observable produces qb_observable_values(id) <= 4 numbers per object in
 * one pass: "qcgd_stats" = {nodes, nodes^2, density, density^2} of utils::serialize (qcgd.hpp:309-372), "qcgd_size",
 * "qubit" (params: [bit]) = probability that the qubit is set, "object_bytes". */
int qb_observable_id(const char *name); /* >= 1, or QB_ERR_UNKNOWN_RULE */
int qb_observable_values(int observable_id);
int qb_iter_average_value(const qb_iter *it, int observable_id, const double *params, uint32_t num_params, double *values, uint32_t capacity);

/* quids::simulate(it_t&, rule_t const*, it_t&, sy_it_t&, size_t max_num_object, debug_t)  quids.hpp:448-543 */
int qb_simulate(qb_iter *it, int rule_id, const double *params, uint32_t num_params, qb_iter *next,
                qb_sym *sym, uint64_t max_num_object, const qb_options *opt, qb_step_cb cb, void *user);

/* rule->hasher over every object of a state (quids.hpp:143-145, qcgd.hpp:472-474) */
int qb_hash_objects(const qb_iter *it, int rule_id, const double *params, uint32_t num_params, uint64_t *hashes);

/* ---- distributed path (quids::mpi, quids_mpi.hpp) ------------------------------------------- */
/* 128-byte NCCL unique id, created on one rank and shared by the caller (torch.distributed, MPI, files...) */
int qb_comm_unique_id(uint8_t id[128]);
/* stands for the MPI_Comm argument of quids::mpi::simulate (quids_mpi.hpp:423) */
int qb_comm_create(qb_ctx *ctx, int world_size, int rank, const uint8_t id[128], qb_comm **out);
int qb_comm_destroy(qb_comm *comm);
/* quids::mpi::simulate quids_mpi.hpp:423-598: hash-ownership interference over NCCL all-to-allv,
 * global top-k, global normalisation.  next->total_proba is the global sum; node_total_proba is
 * this rank's share (quids_mpi.hpp:67,892).  max_num_object counts objects over ALL ranks; 0 = automatic
 * budget (per-rank parent budget, children: what the ranks agree their next states can hold).
 * Errors are collective: if any rank fails (a table overflow on its share, an allocation), EVERY rank
 * returns a non-zero status from the same call -- no rank is left waiting in a collective. */
int qb_simulate_dist(qb_iter *it, int rule_id, const double *params, uint32_t num_params, qb_iter *next,
                     qb_sym *sym, qb_comm *comm, uint64_t max_num_object, const qb_options *opt,
                     qb_step_cb cb, void *user, double *node_total_proba);

/* ---- object migration (quids_mpi.hpp:124-231, 903-1077): the tail of a state moves HBM -> NVLink -> HBM ------------
 * send_objects / receive_objects are called on the two ranks of a pair (not collective), like MPI_Send / MPI_Recv:
 * the sender pops the objects it sent (without normalising, quids_mpi.hpp:170).  *moved = objects really moved
 * (0 when the receiver had no room, quids_mpi.hpp:136-138,196-198). */
int qb_iter_send_objects(qb_iter *it, qb_comm *comm, uint64_t num_object_sent, int node, uint64_t *moved);
/* max_mem: most bytes (52 per object + object bytes) this rank accepts, UINT64_MAX = whatever fits the GPU */
int qb_iter_receive_objects(qb_iter *it, qb_comm *comm, int node, uint64_t max_mem, uint64_t *moved);
/* mpi_iteration::distribute_objects (quids_mpi.hpp:1031-1051) / gather_objects (:1056-1077); collective */
int qb_iter_distribute_objects(qb_iter *it, qb_comm *comm, int node_id);
int qb_iter_gather_objects(qb_iter *it, qb_comm *comm, int node_id);
/* number of children the local objects have under a rule = get_num_symbolic_object after compute_num_child
 * (quids.hpp:548-569); not collective */
/* children counted by the last rule iteration (or qb_iter_count_children) over this state: get_num_symbolic_object, quids.hpp:322-324 */
int qb_iter_num_symbolic_object(const qb_iter *it, uint64_t *num_symbolic_object);
int qb_iter_count_children(qb_iter *it, int rule_id, const double *params, uint32_t num_params, uint64_t *num_children);
/* pairing rounds of mpi_iteration::equalize (rule_id = 0, quids_mpi.hpp:903-960) or equalize_symbolic (rule_id >= 1,
 * ranks weighed by the children of that rule, :965-1026); collective.  max_rounds = 1 and inbalance = 0 is one
 * unconditional round (the public equalize()); the loop of quids::mpi::simulate (:442-500) passes
 * ceil(log2(world)) and the quids::mpi thresholds.  *rounds = pairing rounds run. */
int qb_iter_equalize(qb_iter *it, qb_comm *comm, int rule_id, const double *params, uint32_t num_params, int max_rounds,
                     uint64_t min_equalize_size, float equalize_inbalance, float min_equalize_step, int *rounds);

/* get_total_num_object / get_total_num_symbolic_object style all-reduced counters (quids_mpi.hpp:77-99,322-339) */
int qb_comm_allreduce_u64(qb_comm *comm, uint64_t *values, uint32_t n, int op_max);
int qb_comm_allreduce_f64(qb_comm *comm, double *values, uint32_t n);

#ifdef __cplusplus
}
#endif
#endif
