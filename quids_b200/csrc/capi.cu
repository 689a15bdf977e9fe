// capi.cu -- implementation of the C ABI of include/quids_b200.h: contexts, states in HBM, and the
// orchestration of one rule iteration (the GPU counterpart of quids::simulate, quids.hpp:448-543).
#include <chrono>
#include <cmath>
#include <complex>
#include <map>
#include <mutex>
#include <vector>

#include <quids/device/engine.cuh>
#include "pipeline.cuh"
#include "sort.cuh"

using namespace qb;

// ---- error reporting ----------------------------------------------------------------------------------
static thread_local std::string g_last_error;

template <class F>
static int guarded(F &&f) {
	try {
		f();
		return QB_OK;
	} catch (const qb::error &e) {
		g_last_error = e.what();
		return e.status;
	} catch (const std::exception &e) {
		g_last_error = e.what();
		return QB_ERR_ARG;
	}
}

// ---- registry -----------------------------------------------------------------------------------------
namespace qb {
static std::vector<rule_ops> &rule_registry() {
	static std::vector<rule_ops> r;
	return r;
}
static std::vector<modifier_ops> &modifier_registry() {
	static std::vector<modifier_ops> r;
	return r;
}
static std::vector<observable_ops> &observable_registry() {
	static std::vector<observable_ops> r;
	return r;
}
int register_observable(const observable_ops &ops) {
	observable_registry().push_back(ops);
	return (int)observable_registry().size();
}
const observable_ops *find_observable(int id) {
	return id >= 1 && (size_t)id <= observable_registry().size() ? &observable_registry()[id - 1] : nullptr;
}
int register_rule(const rule_ops &ops) {
	rule_registry().push_back(ops);
	return (int)rule_registry().size();
}
int register_modifier(const modifier_ops &ops) {
	modifier_registry().push_back(ops);
	return (int)modifier_registry().size();
}
const rule_ops *find_rule(int id) { return id >= 1 && id <= (int)rule_registry().size() ? &rule_registry()[id - 1] : nullptr; }
const modifier_ops *find_modifier(int id) { return id >= 1 && id <= (int)modifier_registry().size() ? &modifier_registry()[id - 1] : nullptr; }
} // namespace qb

// ---- objects behind the opaque handles ----------------------------------------------------------------
enum { // u64 words of the small device scratch
	DS_MAX_CHILD_SIZE = 0,
	DS_COUNT = 1,
	DS_OVERFLOW = 2,
	DS_TOTAL = 3,
	DS_USED = 4,
	DS_SELECT_GT = 5, // and 6
	DS_CURSOR = 7,    // region mode: slots handed out
	DS_REGIONS = 8,   // region mode: regions created
	DS_CHILD_RANGE = 9, // two u32: largest child count, ~smallest
	DS_SPILL = 10,      // binned interference: records in the spill list
	DS_KEPT = 11,       // filtered compaction: children above the tolerance (the list itself holds fewer)
	DS_FLOOR = 12,      // filtered compaction: the key below which nothing was listed
	DS_WORDS = 13
};

// Counters the host needs between two launches travel through MAPPED pinned memory, written by a one-block kernel on the
// compute stream, not through cudaMemcpyAsync: a device-to-host copy shares the copy engine of its direction with the bulk
// qb_iter_download_async transfers on the copy stream and would queue behind them (QB_PEEK_MEMCPY=1 restores the copies).
struct peek_list {
	uint64_t *dst[4];
	const uint64_t *src[4];
	uint32_t words[4];
	uint32_t n = 0;
	void add(uint64_t *host_mapped, const uint64_t *dev, uint32_t n_words = 1) {
		dst[n] = host_mapped;
		src[n] = dev;
		words[n++] = n_words;
	}
};
__global__ void peek_kernel(peek_list l) {
#pragma unroll
	for (uint32_t e = 0; e < 4; ++e)
		if (e < l.n)
			for (uint32_t i = threadIdx.x; i < l.words[e]; i += blockDim.x)
				l.dst[e][i] = l.src[e][i];
}

struct qb_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;
	int sm_count = 0;
	uint64_t launches = 0;
	uint64_t select_k = 0; // rank requested from the last select_threshold
	uint64_t *h_small = nullptr; // pinned mirror of d_small
	dev_buf d_small;
	dev_buf scan_ws;
	dev_buf partials;
	dev_buf select;
	dev_buf select_cand; // keys still in the race after two digits of a radix select (select.cuh)

	size_t l2_fetch_granularity_before = 0; // cudaLimitMaxL2FetchGranularity as found by qb_ctx_create (restored by qb_ctx_destroy)
	cudaStream_t copy_in = nullptr, copy_out = nullptr; // host <-> device transfers that overlap the compute stream (qb_iter_*_async)
	cudaEvent_t fence = nullptr;

	void use() const { QB_CUDA(cudaSetDevice(device)); }
	void sync() const { QB_CUDA(cudaStreamSynchronize(stream)); }
	void copy_streams() {
		if (copy_in) return;
		QB_CUDA(cudaStreamCreateWithFlags(&copy_in, cudaStreamNonBlocking));
		QB_CUDA(cudaStreamCreateWithFlags(&copy_out, cudaStreamNonBlocking));
		QB_CUDA(cudaEventCreateWithFlags(&fence, cudaEventDisableTiming));
	}
	// `other` waits for everything enqueued on the compute stream so far
	void order_after_compute(cudaStream_t other) {
		QB_CUDA(cudaEventRecord(fence, stream));
		QB_CUDA(cudaStreamWaitEvent(other, fence, 0));
	}

	// zeroed look-back workspace for a scan over `tiles` tiles
	scan_state scan(uint64_t tiles) {
		const size_t bytes = (tiles + 2) * sizeof(uint64_t);
		scan_ws.ensure(bytes, stream);
		QB_CUDA(cudaMemsetAsync(scan_ws.ptr, 0, bytes, stream));
		scan_state st;
		st.status = scan_ws.as<uint64_t>() + 1;
		st.ticket = scan_ws.as<unsigned int>();
		return st;
	}
	uint64_t *small(int word) { return d_small.as<uint64_t>() + word; }
	// enqueue the fetch of up to four groups of words into h_small (visible to the host after the next sync())
	void peek(const peek_list &l) {
		static const bool by_copy = getenv("QB_PEEK_MEMCPY") != nullptr;
		if (by_copy) {
			for (uint32_t e = 0; e < l.n; ++e)
				QB_CUDA(cudaMemcpyAsync(l.dst[e], l.src[e], l.words[e] * sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
			return;
		}
		peek_kernel<<<1, 32, 0, stream>>>(l);
		QB_CUDA(cudaGetLastError());
		++launches;
	}
	void fetch_small() {
		peek_list l;
		l.add(h_small, d_small.as<uint64_t>(), DS_WORDS);
		peek(l);
		sync();
	}
	int grid_cap() const { return sm_count * 8; }
};

struct qb_iter {
	qb_ctx *ctx;
	uint64_t n = 0, n_bytes = 0;
	double total_proba = 1; // quids.hpp:154
	uint64_t n_symbolic = 0; // children counted by the last compute_num_child over this state (get_num_symbolic_object, quids.hpp:322-324)
	dev_buf objects, begin, size, mag, num_childs, child_begin, num_groups, group_begin;
	// transfers in flight on the copy streams (qb_iter_upload_async / qb_iter_download_async)
	cudaEvent_t uploaded = nullptr, downloaded = nullptr;
	mutable bool upload_pending = false, download_pending = false;

	// the compute stream waits for the transfers of this state; called at the head of every entry point that touches it
	void settle() const {
		if (upload_pending) {
			QB_CUDA(cudaStreamWaitEvent(ctx->stream, uploaded, 0));
			upload_pending = false;
		}
		if (download_pending) {
			QB_CUDA(cudaStreamWaitEvent(ctx->stream, downloaded, 0));
			download_pending = false;
		}
	}

	iter_view view() const { return iter_view{objects.as<uint8_t>(), begin.as<uint64_t>(), size.as<uint32_t>(), mag.as<cplx>(), n}; }
};

struct qb_sym {
	qb_ctx *ctx;
	uint64_t n_children = 0, n_unique = 0; // quids.hpp:344-346
	dev_buf table, directory, ukey, uslot, sslot, kept, scratch, survivor_parent, survivor_child, padded, chunk_parent, sort_keys, sort_vals, sort_hist, sort_base, parent_ctx, bin_records, bin_cursor, bin_spill, bin_spill_key, sample_keys;
	cudaEvent_t ev[2 * QB_PHASE_COUNT] = {};
	bool ev_used[QB_PHASE_COUNT] = {};
	float phase_ms[QB_PHASE_COUNT] = {};
	std::map<uint64_t, double> unique_ratio; // per (rule id, parameters), see rule_history_key(): slots created / children of the last call (sizes the next table)
	std::map<uint64_t, std::pair<double, double>> region_ratio; // region mode: (slots handed out, regions created) / children of the last call
	uint64_t table_capacity = 0;        // of the last call
	int table_attempts = 0;

	uint64_t device_bytes() const {
		return table.cap + directory.cap + ukey.cap + uslot.cap + sslot.cap + kept.cap + scratch.cap + survivor_parent.cap + survivor_child.cap + padded.cap + chunk_parent.cap + sort_keys.cap + sort_vals.cap + sort_hist.cap + sort_base.cap + parent_ctx.cap + bin_records.cap + bin_cursor.cap + bin_spill.cap + bin_spill_key.cap + sample_keys.cap;
	}
};

#include "dist.inc.cuh"

// ---- helpers --------------------------------------------------------------------------------------------
namespace {

struct phase_timer {
	qb_sym *sym;
	bool on;
	phase_timer(qb_sym *s, bool on_) : sym(s), on(on_) {
		for (int p = 0; p < QB_PHASE_COUNT; ++p) {
			sym->ev_used[p] = false;
			sym->phase_ms[p] = 0;
		}
	}
	bool started[QB_PHASE_COUNT] = {};
	void begin(int phase) { // a phase may be entered several times: it spans first begin .. last end
		if (!on || started[phase]) return;
		started[phase] = true;
		if (!sym->ev[2 * phase]) {
			QB_CUDA(cudaEventCreate(&sym->ev[2 * phase]));
			QB_CUDA(cudaEventCreate(&sym->ev[2 * phase + 1]));
		}
		QB_CUDA(cudaEventRecord(sym->ev[2 * phase], sym->ctx->stream));
	}
	void restart(int phase) { started[phase] = false; }
	void end(int phase) {
		if (!on) return;
		QB_CUDA(cudaEventRecord(sym->ev[2 * phase + 1], sym->ctx->stream));
		sym->ev_used[phase] = true;
	}
	void collect() {
		if (!on) return;
		sym->ctx->sync();
		for (int p = 0; p < QB_PHASE_COUNT; ++p)
			if (sym->ev_used[p])
				QB_CUDA(cudaEventElapsedTime(&sym->phase_ms[p], sym->ev[2 * p], sym->ev[2 * p + 1]));
	}
};

struct stepper { // mid_step_function: the stream is drained before the driver's callback runs
	qb_ctx *ctx;
	qb_step_cb cb;
	void *user;
	void operator()(const char *label) const {
		if (cb) {
			ctx->sync();
			cb(label, user);
		}
	}
};

template <class F>
void exclusive_scan(qb_ctx *ctx, F f, uint64_t *out, uint64_t n) {
	if (n == 0) {
		QB_CUDA(cudaMemsetAsync(out, 0, sizeof(uint64_t), ctx->stream));
		return;
	}
	const uint64_t tiles = div_up<uint64_t>(n, SCAN_TILE);
	scan_state st = ctx->scan(tiles);
	exclusive_scan_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(f, out, n, st);
	++ctx->launches;
	QB_CUDA(cudaGetLastError());
}

// k-th largest key over this GPU's keys -- or, with a communicator, over the keys of ALL ranks: every
// digit histogram is all-reduced before the digit is picked, so all ranks walk to the same threshold.
// Leaves threshold / count_gt / need (global values) in ctx->select (device).
template <class KeyFn>
void select_threshold(qb_ctx *ctx, comm_ops *comm, KeyFn key_of, uint64_t n, uint64_t k) {
	ctx->select.ensure(sizeof(select_state), ctx->stream);
	// after two digits (24 bits) the keys still in the race are copied to a dense buffer and the four remaining digits
	// read only that: 3 passes over all the keys instead of 6 (select.cuh).  Not worth a launch for small inputs.
	const bool filter = n >= (1ull << 18);
	const uint64_t cand_capacity = filter ? n / 32 + 1024 : 0;
	if (filter)
		ctx->select_cand.ensure(sizeof(uint64_t) * cand_capacity, ctx->stream);
	select_state *st = ctx->select.as<select_state>();
	select_init_kernel<<<1, SCAN_THREADS, 0, ctx->stream>>>(st, k, cand_capacity);
	++ctx->launches;
	ctx->select_k = k;
	const int grid = grid_for(div_up<uint64_t>(n, SELECT_ITEMS), 256, ctx->grid_cap());
	const uint64_t *cand = nullptr;
	int shift = 64;
	for (int pass = 0; shift > 0; ++pass) {
		const int bits = shift >= SELECT_MAX_BITS ? SELECT_MAX_BITS : shift;
		shift -= bits;
		if (n > 0) {
			select_histogram_kernel<<<grid, 256, 0, ctx->stream>>>(key_of, n, st, shift, bits, cand);
			++ctx->launches;
		}
		if (comm)
			comm->allreduce_u64_device(st->hist, SELECT_BINS);
		select_pick_kernel<<<1, SCAN_THREADS, 0, ctx->stream>>>(st, shift, bits);
		++ctx->launches;
		if (filter && pass == 1) {
			select_filter_kernel<<<grid, 256, 0, ctx->stream>>>(key_of, n, st, ctx->select_cand.as<uint64_t>());
			++ctx->launches;
			cand = ctx->select_cand.as<uint64_t>();
		}
	}
	QB_CUDA(cudaGetLastError());
}

// keeps the selected elements of THIS GPU (out(rank, index)) and returns how many.  Ties at the
// threshold are arbitrary in the reference; here the first ones in storage order win, and across
// ranks the lower ranks are served first.
template <class KeyFn, class OutFn>
uint64_t select_keep(qb_ctx *ctx, comm_ops *comm, KeyFn key_of, uint64_t n, OutFn out) {
	uint64_t kept = ctx->select_k;
	if (comm) {
		unsigned long long *gt_eq = reinterpret_cast<unsigned long long *>(ctx->small(DS_SELECT_GT));
		QB_CUDA(cudaMemsetAsync(gt_eq, 0, 2 * sizeof(uint64_t), ctx->stream));
		if (n > 0) {
			select_count_kernel<<<grid_for(n, 256, ctx->grid_cap()), 256, 0, ctx->stream>>>(key_of, n, ctx->select.as<select_state>(), gt_eq);
			++ctx->launches;
		}
		uint64_t mine[3]; // gt, eq of this rank; the global `need`
		QB_CUDA(cudaMemcpyAsync(mine, gt_eq, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
		QB_CUDA(cudaMemcpyAsync(&mine[2], &ctx->select.as<select_state>()->k, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
		ctx->sync();
		std::vector<uint64_t> all = comm->allgather_u64(mine, 2);
		uint64_t before = 0; // ties taken by lower ranks
		for (int r = 0; r < comm->rank(); ++r)
			before += all[2 * r + 1];
		const uint64_t need = mine[2] > before ? std::min<uint64_t>(mine[2] - before, mine[1]) : 0;
		select_patch_kernel<<<1, 1, 0, ctx->stream>>>(ctx->select.as<select_state>(), need, mine[0]);
		++ctx->launches;
		kept = mine[0] + need;
	}
	if (n > 0) {
		const uint64_t tiles = div_up<uint64_t>(n, SELECT_TILE);
		const select_state *sel = ctx->select.as<select_state>();
		// QB_SELECT_TWO_PASS: developer / test knob that forces the path of inputs of 2^31 keys and more
		if (n < (1ull << 31) && !getenv("QB_SELECT_TWO_PASS")) { // both kinds in one pass (two 31-bit running counts in one look-back word)
			scan_state st = ctx->scan(tiles);
			select_compact_kernel<2><<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(key_of, n, sel, out, st);
			++ctx->launches;
		} else {
			scan_state st = ctx->scan(tiles);
			select_compact_kernel<0><<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(key_of, n, sel, out, st);
			st = ctx->scan(tiles);
			select_compact_kernel<1><<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(key_of, n, sel, out, st);
			ctx->launches += 2;
		}
	}
	QB_CUDA(cudaGetLastError());
	return kept;
}

uint64_t global_sum(comm_ops *comm, uint64_t v) { return comm ? comm->sum_u64(v) : v; }

struct key_from_mag {
	const cplx *mag;
	__device__ uint64_t operator()(uint64_t i) const { return key_of_norm(cnorm(mag[i])); }
};
// probabilistic truncation (quids.hpp:594-608, 829-845): the reference keeps the objects with the SMALLEST
// random_selector = rng() / |mag|^2, rng uniform in [0, 1).  Here u comes from a counter-based generator
// (splitmix64 of seed and element index: reproducible for a given seed), and the selector's bit pattern is
// inverted so that the same "k largest keys" select applies.
__device__ __forceinline__ uint64_t random_selector_key(double norm, uint64_t index, uint32_t seed) {
	uint64_t x = (index + 1) * 0x9e3779b97f4a7c15ull + ((uint64_t)seed << 32 | seed);
	x = mix64(x);
	const double u = ((double)(x >> 11) + 0.5) * (1.0 / 9007199254740992.0); // (0, 1)
	return ~(uint64_t)__double_as_longlong(u / norm);
}
struct key_from_mag_random {
	const cplx *mag;
	uint32_t seed;
	__device__ uint64_t operator()(uint64_t i) const { return random_selector_key(cnorm(mag[i]), i, seed); }
};
// the compacted keys of the unique children are |mag|^2 bit patterns: replace them by selector keys.  The counter of
// the generator is the object's HASH (not its table slot, which depends on insertion order and table capacity), so a
// seed picks the same objects on every run and on any number of GPUs.
__global__ void __launch_bounds__(256) randomize_keys_kernel(uint64_t *keys, table_view t, const uint32_t *slot, uint64_t n, uint32_t seed) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
		keys[i] = random_selector_key(__longlong_as_double((long long)keys[i]), t.slots[slot[i]].key, seed ^ 0x5bd1e995u);
}

struct key_from_array {
	const uint64_t *key;
	__device__ uint64_t operator()(uint64_t i) const { return key[i]; }
};
struct out_index {
	uint64_t *dst;
	__device__ void operator()(uint64_t rank, uint64_t i) const { dst[rank] = i; }
};
struct out_gather_u32 {
	uint32_t *dst;
	const uint32_t *src;
	__device__ void operator()(uint64_t rank, uint64_t i) const { dst[rank] = src[i]; }
};
struct counts_through {
	const uint32_t *num_childs;
	const uint64_t *kept;
	__device__ uint64_t operator()(uint64_t j) const { return num_childs[kept ? kept[j] : j]; }
};
struct widen_u32 {
	const uint32_t *v;
	__device__ uint64_t operator()(uint64_t j) const { return v[j]; }
};

__global__ void __launch_bounds__(256) iota_kernel(uint64_t *out, uint64_t n) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
		out[i] = i;
}

// number of positions where the sorted key changes (= runs - 1)
__global__ void __launch_bounds__(256) key_changes_kernel(const uint32_t *keys, uint64_t n, unsigned long long *count) {
	unsigned long long local = 0;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; i < n; i += stride)
		local += keys[i] != keys[i - 1];
	local = warp_sum((uint64_t)local);
	if (lane_id() == 0 && local)
		atomicAdd(count, local);
}

// stable radix sort of n (u32 key, u64 value) pairs that sit in the first halves of sym->sort_keys /
// sort_vals (each sized for 2 n); returns the sorted values (device)
// Only the top SORT_KEY_BITS of the key are sorted on (3 passes instead of 4): the order only has to bring equal keys
// together, and two different keys that agree on 24 bits merely interleave their (rare) items.
constexpr int SORT_KEY_BITS = 24;
const uint64_t *sort_items(qb_ctx *ctx, qb_sym *sym, uint64_t n, const uint32_t **sorted_keys = nullptr) {
	cudaStream_t stream = ctx->stream;
	const uint64_t tiles = div_up<uint64_t>(n, SORT_TILE);
	sym->sort_hist.ensure(sizeof(uint32_t) * SORT_BINS * tiles, stream);
	sym->sort_base.ensure(sizeof(uint64_t) * (SORT_BINS * tiles + 1), stream);
	uint32_t *keys[2] = {sym->sort_keys.as<uint32_t>(), sym->sort_keys.as<uint32_t>() + n};
	uint64_t *vals[2] = {sym->sort_vals.as<uint64_t>(), sym->sort_vals.as<uint64_t>() + n};
	static bool stage_allowed[MAX_DEVICES] = {}; // the scatter kernel sorts a tile in 58 KB of shared memory: a per-device opt-in
	if (!stage_allowed[ctx->device % MAX_DEVICES]) {
		QB_CUDA(cudaFuncSetAttribute((const void *)radix_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SORT_STAGE_BYTES));
		stage_allowed[ctx->device % MAX_DEVICES] = true;
	}
	int src = 0;
	for (int shift = 32 - SORT_KEY_BITS; shift < 32; shift += 8, src ^= 1) {
		radix_histogram_kernel<<<(unsigned)tiles, SORT_THREADS, 0, stream>>>(keys[src], n, shift, sym->sort_hist.as<uint32_t>(), tiles);
		++ctx->launches;
		exclusive_scan(ctx, widen_u32{sym->sort_hist.as<uint32_t>()}, sym->sort_base.as<uint64_t>(), SORT_BINS * tiles);
		radix_scatter_kernel<<<(unsigned)tiles, SORT_THREADS, SORT_STAGE_BYTES, stream>>>(keys[src], vals[src], n, shift, sym->sort_base.as<uint64_t>(), tiles,
		                                                                   keys[src ^ 1], vals[src ^ 1]);
		++ctx->launches;
	}
	QB_CUDA(cudaGetLastError());
	if (sorted_keys)
		*sorted_keys = keys[src];
	return vals[src];
}

// ---- binned interference (table.cuh), host side: bins for n records in the given buffers ...
struct bin_buffers {
	dev_buf *records, *cursor, *spill, *spill_key;
};
bin_view make_bins(qb_ctx *ctx, const bin_buffers &buf, uint64_t n_records, unsigned long long *spill_cursor) {
	cudaStream_t stream = ctx->stream;
	const uint64_t bins = div_up<uint64_t>(std::max<uint64_t>(n_records, 1), BIN_MEAN_RECORDS);
	QB_REQUIRE(bins < (1ull << 24), QB_ERR_CAPACITY, "binned interference: more than 2^24 bins");
	const uint64_t spill_capacity = n_records / 8 + 65536;
	buf.records->ensure(sizeof(bin_record) * bins * BIN_CAPACITY, stream);
	buf.cursor->ensure(sizeof(unsigned int) * bins, stream);
	buf.spill->ensure(sizeof(bin_record) * spill_capacity, stream);
	buf.spill_key->ensure(sizeof(unsigned int) * spill_capacity, stream);
	QB_CUDA(cudaMemsetAsync(buf.cursor->ptr, 0, sizeof(unsigned int) * bins, stream));
	QB_CUDA(cudaMemsetAsync(spill_cursor, 0, sizeof(uint64_t), stream));
	return bin_view{buf.records->as<bin_record>(), buf.cursor->as<unsigned int>(), (uint32_t)bins, buf.spill->as<bin_record>(), buf.spill_key->as<unsigned int>(),
	                spill_cursor, spill_capacity};
}

// ... and pass 2: every bin deduplicated in shared memory -> dense unique entries + the compacted (norm key, slot) list.
// Drains the stream once (the spill count sizes the sort of the spill list).  Nothing is launched if pass 1 raised the overflow flag.
void launch_bin_dedup(qb_ctx *ctx, qb_sym *sym, const bin_view &bins, const table_view &dense, double tolerance, uint64_t *ukey, uint32_t *uslot, bool max_rep) {
	cudaStream_t stream = ctx->stream;
	ctx->fetch_small();
	if (ctx->h_small[DS_OVERFLOW] != 0)
		return;
	bin_dedup_args a;
	a.bins = bins;
	a.spill_order = nullptr;
	a.spill_sorted = nullptr;
	a.n_spill = std::min<uint64_t>(ctx->h_small[DS_SPILL], bins.spill_capacity);
	if (a.n_spill > 0) { // sort the spilled records by bin: (bin << 8, index) pairs
		sym->sort_keys.ensure(2 * sizeof(uint32_t) * a.n_spill, stream);
		sym->sort_vals.ensure(2 * sizeof(uint64_t) * a.n_spill, stream);
		QB_CUDA(cudaMemcpyAsync(sym->sort_keys.ptr, bins.spill_bin, sizeof(uint32_t) * a.n_spill, cudaMemcpyDeviceToDevice, stream));
		iota_kernel<<<grid_for(a.n_spill, 256, ctx->grid_cap()), 256, 0, stream>>>(sym->sort_vals.as<uint64_t>(), a.n_spill);
		++ctx->launches;
		const uint32_t *sorted_keys = nullptr;
		a.spill_order = sort_items(ctx, sym, a.n_spill, &sorted_keys);
		a.spill_sorted = sorted_keys;
	}
	a.dense = dense;
	a.dense_cursor = dense.used;
	a.tolerance = tolerance;
	a.ukey = ukey;
	a.uslot = uslot;
	a.count = reinterpret_cast<unsigned long long *>(ctx->small(DS_COUNT));
	a.max_rep = max_rep ? 1 : 0;
	static bool smem_allowed[MAX_DEVICES] = {};
	if (!smem_allowed[ctx->device % MAX_DEVICES]) {
		QB_CUDA(cudaFuncSetAttribute((const void *)bin_dedup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BIN_DEDUP_SMEM));
		smem_allowed[ctx->device % MAX_DEVICES] = true;
	}
	bin_dedup_kernel<<<(unsigned)std::min<uint64_t>(bins.bins, (uint64_t)ctx->sm_count * BIN_DEDUP_BLOCKS_PER_SM), BIN_DEDUP_THREADS, BIN_DEDUP_SMEM, stream>>>(a);
	++ctx->launches;
	QB_CUDA(cudaGetLastError());
}

double reduce_norm_total(qb_ctx *ctx, const cplx *mag, uint64_t n) {
	const int grid = grid_for(n, SCAN_THREADS, ctx->grid_cap());
	ctx->partials.ensure(sizeof(double) * (size_t)ctx->grid_cap(), ctx->stream);
	norm_partial_kernel<<<grid, SCAN_THREADS, 0, ctx->stream>>>(mag, n, ctx->partials.as<double>());
	norm_total_kernel<<<1, SCAN_THREADS, 0, ctx->stream>>>(ctx->partials.as<double>(), grid, reinterpret_cast<double *>(ctx->small(DS_TOTAL)));
	ctx->launches += 2;
	ctx->fetch_small();
	double total;
	memcpy(&total, &ctx->h_small[DS_TOTAL], sizeof total);
	return total;
}

void scale_state(qb_ctx *ctx, cplx *mag, uint64_t n, double total) {
	const double factor = std::sqrt(total);
	if (n == 0 || factor == 1) // quids.hpp:1010
		return;
	scale_kernel<<<grid_for(n, SCAN_THREADS, ctx->grid_cap()), SCAN_THREADS, 0, ctx->stream>>>(mag, n, factor);
	++ctx->launches;
}

// what the table-size history of a symbolic iteration is filed under: the rule AND its parameters (erase_create(0.1) and
// erase_create(pi/4) on the same state create very different numbers of slots: amplitudes that vanish prune children)
uint64_t rule_history_key(int rule_id, const double *params, uint32_t num_params) {
	uint64_t h = mix64((uint64_t)rule_id + 0x9e3779b97f4a7c15ull);
	for (uint32_t i = 0; i < num_params; ++i) {
		uint64_t bits;
		memcpy(&bits, &params[i], sizeof bits);
		h = mix64(h ^ bits) + i;
	}
	return h;
}

void resolve_options(const qb_options *in, qb_options &opt) {
	qb_options_default(&opt);
	if (in)
		opt = *in;
	if (!(opt.table_load > 0 && opt.table_load <= 0.95))
		opt.table_load = 0.75;
	if (in == nullptr || opt.binned_inserts < 0 || opt.binned_inserts > 2)
		opt.binned_inserts = 1;
}

// ======================================================================================================
// one rule iteration
// ======================================================================================================
// bytes the automatic budget (max_num_object = 0) may spend: what is free now + what `sym` (and `next`) already hold
// and will reuse, minus the safety margin (quids.hpp:459-470 with cudaMemGetInfo in place of /proc/meminfo);
// qb_options.memory_budget overrides the measurement (tests, or a share of a GPU)
double automatic_budget(qb_ctx *ctx, const qb_sym *sym, const qb_iter *next, const qb_options &opt, double workspace = 0) {
	if (opt.memory_budget > 0)
		return (double)opt.memory_budget - (next ? workspace : 0.0);
	size_t free_bytes = 0, total_bytes = 0;
	QB_CUDA(cudaMemGetInfo(&free_bytes, &total_bytes));
	{ // what the stream-ordered pool holds but nobody uses counts as free: the next allocation takes it first
		cudaMemPool_t pool = nullptr;
		uint64_t reserved = 0, used = 0;
		if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess && cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
		    cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
			free_bytes += reserved - used;
		cudaGetLastError();
	}
	double reusable = next ? (double)(next->objects.cap + next->begin.cap + next->size.cap + next->mag.cap) : (double)sym->device_bytes();
	return (double)free_bytes + reusable - (double)opt.safety_margin * (double)total_bytes;
}

// stages 1-6 of an iteration on THIS GPU: child counts, parent pre-truncation, index ranges, children -> interference
// table (regions / bins / hashed), compaction of the table into (norm key, slot) lists
struct local_table {
	uint64_t n_parents = 0;
	const uint64_t *kept = nullptr;
	uint64_t n_children = 0;
	uint32_t max_child_size = 0;
	uint32_t uniform_fanout = 0; // != 0: every parent of the state has this many children
	double workspace = 0; // automatic budget: bytes the symbolic workspace of the kept parents was counted for
	uint64_t n_unique = 0; // unique children above the tolerance (N_u)
	uint64_t n_listed = 0; // entries of the (norm key, slot) list: n_unique, or fewer when the compaction was filtered (pipeline.cuh)
	bool filtered = false; // the list only holds the entries whose key is at least `floor_key`
	uint64_t floor_key = 0;
	uint64_t scan_n = 0;   // slots the compaction scans (0: the list was not made by table_compact_kernel)
	double compaction_tolerance = 0;
	table_view table{};
	int empty_from = -1; // >= 0: nothing to do from label `empty_from` on (no parents / no children)
};

local_table build_local_table(qb_iter *it, uint64_t rule_id, const rule_ops *ops, const void *rule, qb_sym *sym, uint64_t max_num_object, bool automatic,
                              const qb_options &opt, double compaction_tolerance, phase_timer &timer, const stepper &step, engine_launch &L, comm_ops *comm,
                              bool *past_collectives = nullptr) {
	qb_ctx *ctx = it->ctx;
	cudaStream_t stream = ctx->stream;
	local_table R;

	// ---- 1. number of children per parent (quids.hpp:548-569) --------------------------------------
	step("num_child");
	sym->n_children = sym->n_unique = 0;
	const uint64_t n_global = global_sum(comm, it->n);
	if (n_global == 0) {
		step("truncate_symbolic - prepare");
		step("truncate_symbolic");
		R.empty_from = 0;
		return R;
	}
	QB_CUDA(cudaMemsetAsync(ctx->d_small.ptr, 0, DS_WORDS * sizeof(uint64_t), stream));
	if (it->n > 0) {
		timer.begin(QB_PHASE_NUM_CHILD);
		it->num_childs.ensure(sizeof(uint32_t) * it->n, stream);
		L.num_childs = it->num_childs.as<uint32_t>();
		if (ops->has_groups) {
			it->num_groups.ensure(sizeof(uint32_t) * it->n, stream);
			L.num_groups = it->num_groups.as<uint32_t>();
		}
		L.max_child_size = reinterpret_cast<unsigned int *>(ctx->small(DS_MAX_CHILD_SIZE));
		L.child_count_range = reinterpret_cast<unsigned int *>(ctx->small(DS_CHILD_RANGE));
		ops->launch_num_child(rule, L);
		timer.end(QB_PHASE_NUM_CHILD);
	}

	// child index ranges of a set of kept parents (quids.hpp:666-671: a serial loop in the reference)
	uint64_t n_groups = 0;
	uint32_t max_child_size = 0, uniform_fanout = 0;
	const uint64_t *group_begin = nullptr;
	auto index_children = [&](const uint64_t *kept, uint64_t n_parents) -> uint64_t {
		it->child_begin.ensure(sizeof(uint64_t) * (n_parents + 1), stream);
		exclusive_scan(ctx, counts_through{it->num_childs.as<uint32_t>(), kept}, it->child_begin.as<uint64_t>(), n_parents);
		group_begin = it->child_begin.as<uint64_t>();
		if (ops->has_groups) { // children are produced in groups that share work: a second index space
			it->group_begin.ensure(sizeof(uint64_t) * (n_parents + 1), stream);
			exclusive_scan(ctx, counts_through{it->num_groups.as<uint32_t>(), kept}, it->group_begin.as<uint64_t>(), n_parents);
			group_begin = it->group_begin.as<uint64_t>();
		}
		peek_list l;
		l.add(&ctx->h_small[DS_COUNT], it->child_begin.as<uint64_t>() + n_parents);
		l.add(&ctx->h_small[DS_USED], group_begin + n_parents);
		l.add(&ctx->h_small[DS_MAX_CHILD_SIZE], ctx->small(DS_MAX_CHILD_SIZE));
		l.add(&ctx->h_small[DS_CHILD_RANGE], ctx->small(DS_CHILD_RANGE));
		ctx->peek(l);
		ctx->sync();
		{
			const uint32_t largest = (uint32_t)ctx->h_small[DS_CHILD_RANGE], smallest = ~(uint32_t)(ctx->h_small[DS_CHILD_RANGE] >> 32);
			uniform_fanout = largest == smallest ? largest : 0;
		}
		n_groups = ctx->h_small[DS_USED];
		max_child_size = (uint32_t)ctx->h_small[DS_MAX_CHILD_SIZE];
		return ctx->h_small[DS_COUNT];
	};

	// ---- 2a. automatic budget (max_num_object = 0, quids.hpp:459-485): keep the most probable parents whose
	//          symbolic workspace fits in the GPU memory left after the safety margin.  The reference bisects
	//          over get_truncated_mem_size (:574-589); here the count shrinks by the ratio budget / need until it fits.
	bool parents_chosen = false;
	R.n_parents = it->n;
	if (automatic && it->n > 0) {
		const double budget = automatic_budget(ctx, sym, nullptr, opt);
		uint64_t k = it->n;
		const uint64_t *kept = nullptr;
		double need = 0;
		for (int round = 0; round < 64; ++round) {
			const uint64_t children = index_children(kept, k);
			// interference table at its safe size + the compacted (key, slot) lists + sorted work items and parent contexts
			need = (std::ceil((double)children / opt.table_load) + 2) * sizeof(table_slot) + 12.0 * (double)children +
			       (ops->has_group_key ? 24.0 * (double)n_groups + (double)ops->ctx_bytes * (double)k : 0.0) + 16.0 * (double)k +
			       (comm ? 140.0 * (double)children : 0.0); // distributed path: send + receive buffers, owner table and its compacted lists
			if (need <= budget || k <= 1)
				break;
			k = std::max<uint64_t>(1, std::min<uint64_t>(k - 1, (uint64_t)((double)k * std::min(0.9, budget / need))));
			sym->kept.ensure(sizeof(uint64_t) * k, stream);
			key_from_mag keys{it->mag.as<cplx>()};
			select_threshold(ctx, nullptr, keys, it->n, k);
			k = select_keep(ctx, nullptr, keys, it->n, out_index{sym->kept.as<uint64_t>()});
			kept = sym->kept.as<uint64_t>();
		}
		QB_REQUIRE(need <= budget, QB_ERR_CAPACITY,
		           "max_num_object = 0 (automatic budget): the children of a single parent need about " + std::to_string((uint64_t)(need / 1e6)) +
		               " MB of workspace, " + std::to_string((uint64_t)(std::max(0.0, budget) / 1e6)) + " MB are available after the safety margin");
		if (k < it->n) {
			if (comm) { // distributed path: every rank keeps what ITS memory holds (as the reference does, quids_mpi.hpp:505-537)
				parents_chosen = true;
				R.n_parents = k;
				R.kept = kept;
			} else
				max_num_object = k;
		}
		R.workspace = need;
	}

	// ---- 2. parent pre-truncation: the max_num_object most probable parents (quids.hpp:613-642);
	//         over ALL ranks on the distributed path, so that the result equals the single-GPU one ---------
	step("truncate_symbolic - prepare");
	step("truncate_symbolic");
	if (!parents_chosen && max_num_object < n_global) {
		timer.begin(QB_PHASE_PRE_TRUNCATE);
		sym->kept.ensure(sizeof(uint64_t) * std::max<uint64_t>(1, std::min<uint64_t>(it->n, max_num_object)), stream);
		if (opt.simple_truncation) {
			key_from_mag keys{it->mag.as<cplx>()};
			select_threshold(ctx, comm, keys, it->n, max_num_object);
			R.n_parents = select_keep(ctx, comm, keys, it->n, out_index{sym->kept.as<uint64_t>()});
		} else {
			key_from_mag_random keys{it->mag.as<cplx>(), opt.seed + (comm ? 0x9e3779b9u * (uint32_t)comm->rank() : 0u)};
			select_threshold(ctx, comm, keys, it->n, max_num_object);
			R.n_parents = select_keep(ctx, comm, keys, it->n, out_index{sym->kept.as<uint64_t>()});
		}
		R.kept = sym->kept.as<uint64_t>();
		timer.end(QB_PHASE_PRE_TRUNCATE);
	}
	// ---- 3. child index ranges ---------------------------------------------------------------------------
	step("prepare_index");
	if (R.n_parents > 0) {
		timer.begin(QB_PHASE_NUM_CHILD);
		R.n_children = index_children(R.kept, R.n_parents);
		timer.end(QB_PHASE_NUM_CHILD);
	}
	R.max_child_size = max_child_size;
	R.uniform_fanout = uniform_fanout;
	sym->n_children = R.n_children;
	it->n_symbolic = R.n_children;
	if (global_sum(comm, R.n_children) == 0) {
		R.empty_from = 1;
		return R;
	}
	QB_REQUIRE(R.n_children <= REP_MAX_INDEX, QB_ERR_CAPACITY, "more than 2^40 children in one iteration");
	QB_REQUIRE(max_child_size <= REP_MAX_SIZE, QB_ERR_CAPACITY, "child objects of 16 MiB or more are not supported");
	step("symbolic_iteration");
	if (past_collectives) // nothing below talks to the other ranks: what fails from here on is agreed on by the caller
		*past_collectives = true;
	if (R.n_children == 0) // this rank has nothing to contribute (distributed path only)
		return R;

	// ---- 4-6. interference: table sized from the last call of this rule, full size as fallback -------------
	// Capacity: enough for every child to be unique at the configured load (safe), unless the previous
	// call of the same rule showed how many slots are really created: then 2.6x that prediction.  An
	// insert that probes too long raises `overflow` and the whole step is redone at full size.
	const uint64_t n_children = R.n_children;
	uint64_t full_capacity = std::max<uint64_t>(1024, (uint64_t)std::ceil((double)n_children / opt.table_load));
	uint64_t capacity = full_capacity, directory = 0, full_directory = 0;
	bool have_history = false, region_mode = false;
	double expected_runs = 0; // region mode: runs of equal work items (= regions) the sorted order is expected to hold
	{
		auto hint = sym->unique_ratio.find(rule_id);
		have_history = hint != sym->unique_ratio.end();
		if (have_history)
			capacity = std::min<uint64_t>(full_capacity, std::max<uint64_t>(1024, (uint64_t)((hint->second * (double)n_children * 1.3 + 1024) / 0.5)));
	}
	L.child_begin = it->child_begin.as<uint64_t>();
	L.group_begin = group_begin;
	L.kept = R.kept;
	L.n_parents = R.n_parents;
	L.n_children = n_children;
	L.n_groups = n_groups;
	sym->chunk_parent.ensure(sizeof(uint64_t) * (ops->symbolic_chunks(n_groups) + 2), stream);
	L.chunk_parent = sym->chunk_parent.as<uint64_t>();
	if (ops->needs_scratch) {
		L.scratch_stride = (max_child_size + 15u) & ~15u;
		if (L.scratch_stride == 0)
			L.scratch_stride = 16;
		sym->scratch.ensure((size_t)ops->symbolic_grid(ctx->sm_count) * SYMBOLIC_THREADS * L.scratch_stride, stream);
		L.scratch = sym->scratch.as<uint8_t>();
	}
	// sorted order: all groups that produce the same objects become consecutive work items, so that the
	// rule can merge them in shared memory before the table sees them (rule_api.cuh, has_group_key)
	const bool sorted_order = ops->has_group_key && opt.locality_sort != 0 && (opt.locality_sort > 1 || n_groups >= (1u << 16)) &&
	                          R.n_parents < (1ull << (64 - ITEM_GROUP_BITS));
	if (sorted_order) {
		timer.begin(QB_PHASE_PRE_TRUNCATE);
		sym->sort_keys.ensure(2 * sizeof(uint32_t) * n_groups, stream);
		sym->sort_vals.ensure(2 * sizeof(uint64_t) * n_groups, stream);
		sym->parent_ctx.ensure(ops->ctx_bytes * R.n_parents, stream);
		L.parent_ctx = sym->parent_ctx.ptr;
		L.item_keys = sym->sort_keys.as<uint32_t>();
		L.item_vals = sym->sort_vals.as<uint64_t>();
		ops->launch_group_items(rule, L);
		const uint32_t *sorted_keys = nullptr;
		L.items = sort_items(ctx, sym, n_groups, &sorted_keys);
		// region mode (table.cuh): the rule sends every run to a region of consecutive slots found through a directory
		// hashed by the run's identity.  Slots: one per child is always enough and needs no load factor.
		region_mode = ops->region_size_limit > 0 && max_child_size < ops->region_size_limit;
		if (region_mode) {
			// slots are handed to the warps in chunks of REGION_CHUNK: the tail of a chunk that cannot hold the next region
			// (< group_capacity slots) and the last chunk of every warp stay empty
			const uint64_t chunk_slack = std::min<uint64_t>(div_up<uint64_t>(n_groups, ITEM_CHUNK), (uint64_t)ctx->sm_count * 32) * REGION_CHUNK;
			full_capacity = std::max<uint64_t>(1024, n_children + n_children / (REGION_CHUNK / ops->group_capacity - 1) + chunk_slack + REGION_CHUNK);
			full_directory = 2 * n_groups + 1024;
			auto hint = sym->region_ratio.find(rule_id);
			have_history = hint != sym->region_ratio.end();
			if (have_history) {
				capacity = std::min<uint64_t>(full_capacity, (uint64_t)(hint->second.first * (double)n_children * 1.3) + 4096);
				directory = std::min<uint64_t>(full_directory, (uint64_t)(hint->second.second * (double)n_children * 2.6) + 1024);
				expected_runs = hint->second.second * (double)n_children;
			} else {
				capacity = full_capacity;
				directory = full_directory;
			}
		}
		if (!have_history) {
			// in sorted order the table typically receives one group's worth of objects per run of equal keys (plus one
			// split per chunk of work items), far fewer than one per child: a prediction that needs no history.  It is
			// not a bound (groups with equal keys may still hold different objects): an overflow falls back to the full size.
			QB_CUDA(cudaMemsetAsync(ctx->small(DS_COUNT), 0, sizeof(uint64_t), stream));
			key_changes_kernel<<<grid_for(n_groups, 256, ctx->grid_cap()), 256, 0, stream>>>(sorted_keys, n_groups,
			                                                                              reinterpret_cast<unsigned long long *>(ctx->small(DS_COUNT)));
			++ctx->launches;
			ctx->fetch_small();
			const double flushes = 1.25 * (double)(ctx->h_small[DS_COUNT] + 1 + div_up<uint64_t>(n_groups, ITEM_CHUNK)) + 4096;
			expected_runs = (double)(ctx->h_small[DS_COUNT] + 1);
			if (region_mode) {
				const uint64_t chunk_slack = std::min<uint64_t>(div_up<uint64_t>(n_groups, ITEM_CHUNK), (uint64_t)ctx->sm_count * 32) * REGION_CHUNK;
				capacity = std::min<uint64_t>(full_capacity, std::max<uint64_t>(1024, (uint64_t)(flushes * ops->group_capacity * 1.15) + chunk_slack));
				directory = std::min<uint64_t>(full_directory, (uint64_t)(2 * flushes));
			} else {
				capacity = std::min<uint64_t>(full_capacity, std::max<uint64_t>(1024, (uint64_t)(flushes * ops->group_capacity / opt.table_load)));
			}
		}
		timer.end(QB_PHASE_PRE_TRUNCATE);
	}
	// mid_step_function labels of compute_collisions (quids.hpp:754,784,810), single-GPU path: the reference's phases map to
	// prepare = table (and directory) clear, insert = the fused child-generation + table-insert kernel, finalize = compaction by
	// tolerance; "symbolic_iteration" before them covers the ordering of the work items.  A redo after a table overflow does not
	// repeat the labels (their order is part of the API): its time is attributed to "finalize".
	const bool collision_labels = comm == nullptr || comm->local_interference;
	// binned interference (table.cuh) pays when most children are unique and the table would be far larger than L2; a state
	// whose children mostly coincide (history of the rule: < 0.2 slots per child) keeps the hashed table, which then stays hot
	const bool duplicate_heavy = !sorted_order && have_history && sym->unique_ratio[rule_id] < 0.2;
	const bool binned_allowed = !sorted_order && !ops->warp_groups && opt.binned_inserts != 0 && n_children <= (1ull << 29) &&
	                            (opt.binned_inserts > 1 || (n_children >= (1ull << 22) && !duplicate_heavy));
	bool compaction_filtered = false;
	for (sym->table_attempts = 1;; ++sym->table_attempts) {
		QB_REQUIRE(capacity + 1 <= 0xffffffffull, QB_ERR_CAPACITY, "interference table would need more than 2^32 slots");
		if (collision_labels && sym->table_attempts == 1)
			step("compute_collisions - prepare");
		timer.begin(QB_PHASE_TABLE_CLEAR);
		// binned interference (table.cuh): one-child-per-lane rules with many children and no history of heavy duplication send
		// their children to small bins first; each bin is then deduplicated in shared memory and written out as a DENSE array
		// of unique children -- the table below is that array (no clear, no compaction pass).  First attempt only: a bin that
		// overflows falls back to the hashed global table.
		L.bins = bin_view{nullptr, nullptr, 0, nullptr, nullptr, nullptr, 0};
		const bool binned = binned_allowed && sym->table_attempts == 1;
		if (binned) {
			capacity = n_children; // room for "every child is unique": the dense array is neither cleared nor scanned, room costs nothing
			L.bins = make_bins(ctx, bin_buffers{&sym->bin_records, &sym->bin_cursor, &sym->bin_spill, &sym->bin_spill_key}, n_children,
			                   reinterpret_cast<unsigned long long *>(ctx->small(DS_SPILL)));
		}
		const size_t table_bytes = (capacity + 1) * sizeof(table_slot);
		if (!binned && table_bytes > sym->table.cap && table_bytes > ((size_t)4 << 30)) {
			// a table of many GB has to grow: what the binned path of another rule left behind is dead weight until that rule runs again
			sym->bin_records.free_async(stream);
			sym->bin_spill.free_async(stream);
			sym->bin_spill_key.free_async(stream);
		}
		sym->table.ensure(table_bytes, stream);
		if (binned) // only the dedicated slot of the hash 0 is accumulated into
			QB_CUDA(cudaMemsetAsync(sym->table.as<table_slot>() + capacity, 0, sizeof(table_slot), stream));
		else if (!region_mode) // regions are written whole by the runs that create them, what stays unused is zeroed by its warp (table.cuh)
			QB_CUDA(cudaMemsetAsync(sym->table.ptr, 0, table_bytes, stream));
		QB_CUDA(cudaMemsetAsync(ctx->small(DS_COUNT), 0, 4 * sizeof(uint64_t), stream)); // count, overflow, total, used
		R.table = table_view{sym->table.as<table_slot>(), capacity, reinterpret_cast<unsigned int *>(ctx->small(DS_OVERFLOW)),
		                     reinterpret_cast<unsigned long long *>(ctx->small(DS_USED))};
		if (region_mode) {
			sym->directory.ensure(directory * sizeof(region_entry), stream);
			QB_CUDA(cudaMemsetAsync(sym->directory.ptr, 0, directory * sizeof(region_entry), stream));
			QB_CUDA(cudaMemsetAsync(ctx->small(DS_CURSOR), 0, 2 * sizeof(uint64_t), stream)); // cursor, regions
			R.table.dir = sym->directory.as<region_entry>();
			R.table.dir_capacity = directory;
			R.table.cursor = reinterpret_cast<unsigned long long *>(ctx->small(DS_CURSOR));
			R.table.regions = reinterpret_cast<unsigned long long *>(ctx->small(DS_REGIONS));
		}
		timer.end(QB_PHASE_TABLE_CLEAR);

		// children -> (hash, magnitude) -> table (quids.hpp:705-719 fused with :785-809)
		if (collision_labels && sym->table_attempts == 1)
			step("compute_collisions - insert");
		timer.begin(QB_PHASE_SYMBOLIC);
		L.table = R.table;
		// short runs (a grown state: about one group per region) go through the batch kernel, long runs (many parents of one
		// family) through the accumulating one; QB_ITEMS_BATCH = 0 / 1 forces the choice (developer knob for A/B runs)
		bool batch_mode = region_mode && ops->has_region_batch && (double)n_groups <= 3.0 * expected_runs;
		if (const char *force = getenv("QB_ITEMS_BATCH"))
			batch_mode = region_mode && ops->has_region_batch && atoi(force) != 0;
		if (sorted_order && batch_mode)
			ops->launch_symbolic_items_batch(rule, L);
		else if (sorted_order)
			ops->launch_symbolic_items(rule, L);
		else
			ops->launch_symbolic(rule, L);
		QB_CUDA(cudaGetLastError());
		timer.end(QB_PHASE_SYMBOLIC);
		if (binned) {
			// pass 2: every bin deduplicated in shared memory -> dense unique children + the compacted (norm key, slot) list
			if (collision_labels && sym->table_attempts == 1)
				step("compute_collisions - finalize");
			timer.begin(QB_PHASE_INSERT);
			const uint64_t bound = std::min<uint64_t>(n_children, capacity + 1);
			sym->ukey.ensure(sizeof(uint64_t) * bound, stream);
			sym->uslot.ensure(sizeof(uint32_t) * bound, stream);
			launch_bin_dedup(ctx, sym, L.bins, R.table, compaction_tolerance, sym->ukey.as<uint64_t>(), sym->uslot.as<uint32_t>(), false);
			timer.end(QB_PHASE_INSERT);
		} else {
		// unique children above the tolerance (quids.hpp:819-823)
		if (collision_labels && sym->table_attempts == 1)
			step("compute_collisions - finalize");
		timer.begin(QB_PHASE_COMPACT);
		uint64_t scan_slots = capacity; // hashed table: every slot may be occupied; regions: only the slots handed out
		if (region_mode) {
			ctx->fetch_small();
			scan_slots = std::min<uint64_t>(capacity, ctx->h_small[DS_CURSOR]);
		}
		const uint64_t bound = std::min<uint64_t>(n_children, scan_slots + 1);
		sym->ukey.ensure(sizeof(uint64_t) * bound, stream);
		sym->uslot.ensure(sizeof(uint32_t) * bound, stream);
		// hashed table: every slot + the dedicated slot of the hash 0; regions: exactly the slots handed out (nothing beyond was written or zeroed)
		const uint64_t scan_n = region_mode ? scan_slots : scan_slots + 1;
		const uint64_t tiles = div_up<uint64_t>(std::max<uint64_t>(scan_n, 1), COMPACT_TILE);
		const unsigned compact_grid = (unsigned)std::min<uint64_t>(tiles, (uint64_t)ctx->sm_count * 16);
		unsigned long long *listed = reinterpret_cast<unsigned long long *>(ctx->small(DS_COUNT)), *kept = reinterpret_cast<unsigned long long *>(ctx->small(DS_KEPT));
		// filtered compaction (pipeline.cuh): when the step keeps k survivors out of far more slots, a lower bound of the k-th
		// largest key is taken from a random sample of the slots first, and only the entries above it are listed
		// Distributed path with the parents routed by family: the top-k is global, but a rank only has to list what can be among
		// the global survivors.  Families are spread by hash, so a rank holds about k / world of them: its floor is taken at
		// 1.5 k / world of ITS keys, and simulate() checks it against the global threshold once that is known (and redoes the
		// list of a rank whose floor turned out too high).
		constexpr uint32_t SAMPLES = 1u << 20;
		const bool routed = comm && comm->local_interference;
		const bool simple_topk = (!comm || routed) && !automatic && opt.simple_truncation && max_num_object != QB_NO_TRUNCATION;
		const uint64_t filter_min = getenv("QB_COMPACT_FILTER_MIN") ? strtoull(getenv("QB_COMPACT_FILTER_MIN"), nullptr, 10) : (1ull << 24); // (test knob)
		uint64_t local_k = max_num_object;
		if (routed && simple_topk) {
			const double factor = getenv("QB_DIST_FLOOR_FACTOR") ? atof(getenv("QB_DIST_FLOOR_FACTOR")) : 1.5; // (test knob: a small factor forces the redo)
			local_k = (uint64_t)(factor * (double)max_num_object / comm->world()) + 1024;
		}
		bool filtered = false;
		if (simple_topk && scan_n >= filter_min && local_k < scan_n / 8 && !getenv("QB_NO_COMPACT_FILTER")) {
			const double expected = (double)local_k * SAMPLES / (double)scan_n; // sample keys at or above the k-th largest key
			const uint64_t rank = (uint64_t)(expected + 6.0 * std::sqrt(expected) + 16.0);
			if (rank < SAMPLES / 2) {
				sym->sample_keys.ensure(sizeof(uint64_t) * SAMPLES, stream);
				sample_norm_keys_kernel<<<SAMPLES / 256, 256, 0, stream>>>(R.table, scan_n, compaction_tolerance, sym->sample_keys.as<uint64_t>(), SAMPLES,
				                                                           0x51ed270b0a3f2c1dull * (sym->table_attempts + 1));
				++ctx->launches;
				select_threshold(ctx, nullptr, key_from_array{sym->sample_keys.as<uint64_t>()}, SAMPLES, rank);
				QB_CUDA(cudaMemcpyAsync(ctx->small(DS_FLOOR), &ctx->select.as<select_state>()->prefix, sizeof(uint64_t), cudaMemcpyDeviceToDevice, stream));
				filtered = true;
			}
		}
		if (filtered) {
			QB_CUDA(cudaMemsetAsync(kept, 0, sizeof(uint64_t), stream));
			table_compact_kernel<true><<<compact_grid, SCAN_THREADS, 0, stream>>>(R.table, scan_n, compaction_tolerance, sym->ukey.as<uint64_t>(),
			                                                                     sym->uslot.as<uint32_t>(), listed, ctx->small(DS_FLOOR), kept);
			++ctx->launches;
			ctx->fetch_small();
			if (ctx->h_small[DS_COUNT] < std::min<uint64_t>(local_k, ctx->h_small[DS_KEPT])) { // the sample misled (6 sigma): the full list after all
				filtered = false;
				QB_CUDA(cudaMemsetAsync(listed, 0, sizeof(uint64_t), stream));
			}
		}
		if (!filtered) {
			table_compact_kernel<false><<<compact_grid, SCAN_THREADS, 0, stream>>>(R.table, scan_n, compaction_tolerance, sym->ukey.as<uint64_t>(),
			                                                                      sym->uslot.as<uint32_t>(), listed, nullptr, nullptr);
			++ctx->launches;
		}
		R.scan_n = scan_n;
		compaction_filtered = filtered;
		QB_CUDA(cudaGetLastError());
		timer.end(QB_PHASE_COMPACT);
		}
		ctx->fetch_small();
		if (getenv("QB_TABLE_TRACE"))
			fprintf(stderr, "[qb table] rule %s attempt %d: children %llu groups %llu capacity %llu (full %llu) sorted %d regions %d (directory %llu, created %llu, slots %llu) -> used %llu kept %llu overflow %llu\n",
			        ops->name, sym->table_attempts, (unsigned long long)n_children, (unsigned long long)n_groups, (unsigned long long)capacity, (unsigned long long)full_capacity,
			        (int)sorted_order, (int)region_mode, (unsigned long long)directory, (unsigned long long)ctx->h_small[DS_REGIONS], (unsigned long long)ctx->h_small[DS_CURSOR],
			        (unsigned long long)ctx->h_small[DS_USED], (unsigned long long)ctx->h_small[DS_COUNT], (unsigned long long)ctx->h_small[DS_OVERFLOW]);
		if (ctx->h_small[DS_OVERFLOW] == 0)
			break;
		QB_REQUIRE(capacity < full_capacity || directory < full_directory, QB_ERR_CAPACITY, "interference table overflow at full size");
		// the prediction was too small (the kernels stop early once an insert gives up): redo at the always-safe full size
		capacity = full_capacity;
		directory = full_directory;
		for (int p : {QB_PHASE_TABLE_CLEAR, QB_PHASE_SYMBOLIC, QB_PHASE_INSERT, QB_PHASE_COMPACT})
			timer.restart(p); // report the attempt that counted
	}
	R.n_listed = ctx->h_small[DS_COUNT];
	R.n_unique = compaction_filtered ? ctx->h_small[DS_KEPT] : R.n_listed;
	R.filtered = compaction_filtered;
	R.floor_key = compaction_filtered ? ctx->h_small[DS_FLOOR] : 0;
	R.compaction_tolerance = compaction_tolerance;
	sym->table_capacity = capacity;
	if (region_mode)
		sym->region_ratio[rule_id] = std::make_pair((double)ctx->h_small[DS_CURSOR] / (double)n_children, (double)ctx->h_small[DS_REGIONS] / (double)n_children);
	else
		sym->unique_ratio[rule_id] = (double)ctx->h_small[DS_USED] / (double)n_children;
	return R;
}

// stages 8-9 on THIS GPU: sizes, offsets, populate_child_simple, normalisation.  The survivors are
// given either as slots of the local table, or (distributed path) as (representative, magnitude) records.
struct survivor_source {
	table_view table{};
	const uint32_t *slot = nullptr;
	const survivor_record *records = nullptr;
};

void finalize_and_normalize(qb_iter *it, const rule_ops *ops, const void *rule, qb_iter *next, qb_sym *sym, const qb_options &opt, const local_table &R,
                            const survivor_source &src, uint64_t n_survivors, phase_timer &timer, const stepper &step, engine_launch &L, comm_ops *comm,
                            double *node_total_proba, bool fail_injected = false);

void finish_empty(qb_ctx *ctx, qb_iter *next, const stepper &step, phase_timer &timer, int from) {
	// the label sequences of the reference's early outs (quids.hpp:650-651,729-731,908-909,989-990)
	static const char *labels[] = {"prepare_index", "symbolic_iteration", "compute_collisions - prepare", "compute_collisions - insert",
	                               "compute_collisions - finalize", "truncate - prepare", "truncate", "prepare_final", "final", "normalize", "end"};
	for (int i = from; i < 11; ++i)
		step(labels[i]);
	next->n = 0;
	next->n_bytes = 0;
	next->total_proba = 0;
	next->begin.ensure(sizeof(uint64_t), ctx->stream);
	QB_CUDA(cudaMemsetAsync(next->begin.ptr, 0, sizeof(uint64_t), ctx->stream));
	ctx->sync();
	timer.collect();
}

// test knob QB_DIST_INJECT_FAILURE="<rank>:<phase>": that rank fails in that phase of the distributed iteration (phases of the
// record-exchange path: local, partition, owner, return, finalize; of the family-routed path: route, local, finalize)
void inject_failure(int rank, const char *phase) {
	const char *inject = getenv("QB_DIST_INJECT_FAILURE");
	if (inject && atoi(inject) == rank && strchr(inject, ':') && !strcmp(strchr(inject, ':') + 1, phase))
		throw qb::error(QB_ERR_CAPACITY, std::string("injected failure in phase ") + phase);
}

// the complete (norm key, slot) list of a table whose compaction was filtered (the floor turned out to be too high, or nothing
// is truncated after all)
void compact_unfiltered(qb_ctx *ctx, qb_sym *sym, local_table &R) {
	cudaStream_t stream = ctx->stream;
	unsigned long long *listed = reinterpret_cast<unsigned long long *>(ctx->small(DS_COUNT));
	QB_CUDA(cudaMemsetAsync(listed, 0, sizeof(uint64_t), stream));
	const uint64_t tiles = div_up<uint64_t>(std::max<uint64_t>(R.scan_n, 1), COMPACT_TILE);
	table_compact_kernel<false><<<(unsigned)std::min<uint64_t>(tiles, (uint64_t)ctx->sm_count * 16), SCAN_THREADS, 0, stream>>>(
	    R.table, R.scan_n, R.compaction_tolerance, sym->ukey.as<uint64_t>(), sym->uslot.as<uint32_t>(), listed, nullptr, nullptr);
	++ctx->launches;
	QB_CUDA(cudaGetLastError());
	ctx->fetch_small();
	R.n_listed = ctx->h_small[DS_COUNT];
	R.filtered = false;
	R.floor_key = 0;
}

// One rule iteration with the interference complete on this GPU: the single-GPU path (comm = nullptr), and the distributed path
// after the parents were routed by family (route.inc.cuh) -- there only the scalars are collective: counts, the digit histograms
// of the global top-k, the norm.
void simulate(qb_iter *it, uint64_t rule_id, const rule_ops *ops, const void *rule, qb_iter *next, qb_sym *sym, uint64_t max_num_object, const qb_options &opt,
              qb_step_cb cb, void *user, comm_ops *comm = nullptr, double *node_total_proba = nullptr) {
	qb_ctx *ctx = it->ctx;
	QB_REQUIRE(next->ctx == ctx && sym->ctx == ctx, QB_ERR_ARG, "iteration, next iteration and symbolic iteration belong to different contexts");
	QB_REQUIRE(next != it, QB_ERR_ARG, "next_iteration must be a different object from iteration");
	const bool automatic = max_num_object == 0; // quids.hpp:459-485, 510-536: keep what fits in the memory left
	if (automatic)
		max_num_object = QB_NO_TRUNCATION;
	ctx->use();
	it->settle();
	next->settle();
	cudaStream_t stream = ctx->stream;
	stepper step{ctx, cb, user};
	phase_timer timer(sym, opt.profile != 0);
	comm_ops::pending_error err; // distributed path: a failure of this rank is agreed on at the next collective (dist.inc.cuh)

	engine_launch L;
	memset(&L, 0, sizeof L);
	L.stream = stream;
	L.sm_count = ctx->sm_count;
	L.launch_counter = &ctx->launches;
	L.it = it->view();

	local_table R;
	if (comm) {
		bool past_collectives = false;
		try {
			R = build_local_table(it, rule_id, ops, rule, sym, max_num_object, automatic, opt, opt.tolerance, timer, step, L, comm, &past_collectives);
			inject_failure(comm->rank(), "local");
		} catch (const qb::error &e) {
			if (!past_collectives)
				throw;
			err.status = e.status;
			err.what = e.what();
		}
	} else {
		R = build_local_table(it, rule_id, ops, rule, sym, max_num_object, automatic, opt, opt.tolerance, timer, step, L, nullptr);
	}
	if (err.ok() && R.empty_from >= 0) {
		finish_empty(ctx, next, step, timer, R.empty_from);
		if (node_total_proba) *node_total_proba = 0;
		return;
	}
	sym->n_unique = R.n_unique; // (the compute_collisions labels were emitted inside build_local_table, around the phases they name)

	// ---- 7. child truncation: the max_num_object most probable (quids.hpp:866-900), over all ranks on the distributed path
	uint64_t room = 0;
	if (automatic) {
		// quids.hpp:510-536: as many children as the next state can hold.  Per survivor: its bytes (bounded by the largest
		// child, padded) + object_begin 8 + size 4 + magnitude 16 + the finalisation's parent 8, child id 4, padded size 4, slot 4
		auto fit = [&] {
			const double budget = automatic_budget(ctx, sym, next, opt, R.workspace);
			const double per_object = (double)((R.max_child_size + 7u) & ~7u) + 48.0;
			room = budget > 0 ? (uint64_t)(budget / per_object) : 0;
			QB_REQUIRE(room >= 1, QB_ERR_CAPACITY, "max_num_object = 0 (automatic budget): no room left for a next state after the interference table");
		};
		if (comm)
			err.run(fit);
		else
			fit();
		max_num_object = room;
	}
	uint64_t n_unique_global = R.n_unique;
	if (comm) { // agreement point of the local interference step; the automatic budget becomes what the tightest rank allows
		const uint64_t mine[2] = {err.ok() ? R.n_unique : 0, automatic ? room : ~0ull};
		std::vector<uint64_t> all = comm->allgather_agreed(mine, 2, err, "the interference step");
		n_unique_global = 0;
		uint64_t min_room = ~0ull;
		for (int r = 0; r < comm->world(); ++r) {
			n_unique_global += all[2 * r];
			min_room = std::min(min_room, all[2 * r + 1]);
		}
		if (automatic) // families are spread by hash, so are the survivors: a factor 2 of head room covers the spread
			max_num_object = std::max<uint64_t>(1, (uint64_t)((double)min_room * comm->world() / 2));
	}
	step("truncate - prepare");
	step("truncate");
	if (R.filtered && !(max_num_object < n_unique_global)) // nothing is truncated after all (the other ranks hold fewer children than expected): the whole list
		compact_unfiltered(ctx, sym, R);
	uint64_t n_survivors = R.n_unique;
	survivor_source src;
	src.table = R.table;
	src.slot = sym->uslot.as<uint32_t>();
	if (max_num_object < n_unique_global) {
		timer.begin(QB_PHASE_TRUNCATE);
		if (!opt.simple_truncation && R.n_unique > 0) {
			randomize_keys_kernel<<<grid_for(R.n_unique, 256, ctx->grid_cap()), 256, 0, stream>>>(sym->ukey.as<uint64_t>(), R.table, sym->uslot.as<uint32_t>(),
			                                                                                    R.n_unique, opt.seed);
			++ctx->launches;
		}
		key_from_array keys{sym->ukey.as<uint64_t>()};
		for (;;) {
			select_threshold(ctx, comm, keys, R.n_listed, max_num_object);
			if (!comm)
				break; // one GPU: the floor was checked against the counts when the list was made
			// routed by family: every rank listed its keys above ITS floor; the global threshold must not lie below any of them
			uint64_t threshold = 0;
			QB_CUDA(cudaMemcpyAsync(&threshold, &ctx->select.as<select_state>()->prefix, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
			ctx->sync();
			const uint64_t too_high = R.filtered && R.floor_key > threshold ? 1 : 0;
			if (comm->sum_u64(too_high) == 0)
				break;
			if (too_high)
				compact_unfiltered(ctx, sym, R);
		}
		sym->sslot.ensure(sizeof(uint32_t) * std::max<uint64_t>(1, std::min<uint64_t>(R.n_listed, max_num_object)), stream);
		n_survivors = select_keep(ctx, comm, keys, R.n_listed, out_gather_u32{sym->sslot.as<uint32_t>(), sym->uslot.as<uint32_t>()});
		src.slot = sym->sslot.as<uint32_t>();
		timer.end(QB_PHASE_TRUNCATE);
	}
	if (n_survivors == 0 && !comm) {
		finish_empty(ctx, next, step, timer, 7);
		return;
	}
	bool fail_finalize = false;
	if (comm) {
		try {
			inject_failure(comm->rank(), "finalize");
		} catch (const qb::error &) {
			fail_finalize = true;
		}
	}
	finalize_and_normalize(it, ops, rule, next, sym, opt, R, src, n_survivors, timer, step, L, comm, node_total_proba, fail_finalize);
}

void finalize_and_normalize(qb_iter *it, const rule_ops *ops, const void *rule, qb_iter *next, qb_sym *sym, const qb_options &opt, const local_table &R,
                            const survivor_source &src, uint64_t n_survivors, phase_timer &timer, const stepper &step, engine_launch &L, comm_ops *comm,
                            double *node_total_proba, bool fail_injected) {
	qb_ctx *ctx = it->ctx;
	cudaStream_t stream = ctx->stream;
	comm_ops::pending_error err; // distributed path: a failure of this rank's finalisation is agreed on at the normalisation sum
	auto guarded_phase = [&](auto &&f) {
		if (comm)
			err.run(f);
		else
			f();
	};

	// ---- 8. finalisation (quids.hpp:905-968) ------------------------------------------------------------
	step("prepare_final");
	timer.begin(QB_PHASE_FINALIZE);
	next->n = n_survivors;
	next->n_bytes = 0;
	double local_total = 0;
	guarded_phase([&] {
	if (fail_injected)
		throw qb::error(QB_ERR_CAPACITY, "injected failure in phase finalize");
	next->begin.ensure(sizeof(uint64_t) * (n_survivors + 1), stream);
	if (n_survivors > 0) {
		next->size.ensure(sizeof(uint32_t) * n_survivors, stream);
		next->mag.ensure(sizeof(cplx) * n_survivors, stream);
		sym->padded.ensure(sizeof(uint32_t) * n_survivors, stream);
		sym->survivor_parent.ensure(sizeof(uint64_t) * n_survivors, stream);
		sym->survivor_child.ensure(sizeof(uint32_t) * n_survivors, stream);
		ctx->partials.ensure(sizeof(double) * (size_t)ctx->grid_cap(), stream);
		const int meta_grid = grid_for(n_survivors, SCAN_THREADS, ctx->grid_cap());
		if (src.records) {
			finalize_record_args a;
			a.records = src.records;
			a.n_survivors = n_survivors;
			a.child_begin = it->child_begin.as<uint64_t>();
			a.kept = R.kept;
			a.n_parents = R.n_parents;
			a.uniform_fanout = R.uniform_fanout;
			a.align = opt.align_byte_length;
			a.next_size = next->size.as<uint32_t>();
			a.next_padded = sym->padded.as<uint32_t>();
			a.next_mag = next->mag.as<cplx>();
			a.survivor_parent = sym->survivor_parent.as<uint64_t>();
			a.survivor_child = sym->survivor_child.as<uint32_t>();
			a.partial_norm = ctx->partials.as<double>();
			finalize_meta_records_kernel<<<meta_grid, SCAN_THREADS, 0, stream>>>(a);
		} else {
			finalize_args a;
			a.table = src.table;
			a.survivor_slot = src.slot;
			a.n_survivors = n_survivors;
			a.child_begin = it->child_begin.as<uint64_t>();
			a.kept = R.kept;
			a.n_parents = R.n_parents;
			a.uniform_fanout = R.uniform_fanout;
			a.align = opt.align_byte_length;
			a.next_size = next->size.as<uint32_t>();
			a.next_padded = sym->padded.as<uint32_t>();
			a.next_mag = next->mag.as<cplx>();
			a.survivor_parent = sym->survivor_parent.as<uint64_t>();
			a.survivor_child = sym->survivor_child.as<uint32_t>();
			a.partial_norm = ctx->partials.as<double>();
			finalize_meta_kernel<<<meta_grid, SCAN_THREADS, 0, stream>>>(a);
		}
		norm_total_kernel<<<1, SCAN_THREADS, 0, stream>>>(ctx->partials.as<double>(), meta_grid, reinterpret_cast<double *>(ctx->small(DS_TOTAL)));
		ctx->launches += 2;
		exclusive_scan(ctx, widen_u32{sym->padded.as<uint32_t>()}, next->begin.as<uint64_t>(), n_survivors);
		peek_list l;
		l.add(&ctx->h_small[DS_COUNT], next->begin.as<uint64_t>() + n_survivors);
		l.add(&ctx->h_small[DS_TOTAL], ctx->small(DS_TOTAL));
		ctx->peek(l);
		ctx->sync();
		next->n_bytes = ctx->h_small[DS_COUNT];
		memcpy(&local_total, &ctx->h_small[DS_TOTAL], sizeof local_total);
		next->objects.ensure(next->n_bytes + 16, stream);
	} else {
		QB_CUDA(cudaMemsetAsync(next->begin.ptr, 0, sizeof(uint64_t), stream));
	}
	});

	step("final");
	guarded_phase([&] {
	if (n_survivors > 0) {
		L.n_survivors = n_survivors;
		L.survivor_parent = sym->survivor_parent.as<uint64_t>();
		L.survivor_child = sym->survivor_child.as<uint32_t>();
		L.next_objects = next->objects.as<uint8_t>();
		L.next_begin = next->begin.as<uint64_t>();
		L.next_size = next->size.as<uint32_t>();
		ops->launch_populate(rule, L);
		QB_CUDA(cudaGetLastError());
	}
	});
	timer.end(QB_PHASE_FINALIZE);

	// ---- 9. normalisation (quids.hpp:985-1017, quids_mpi.hpp:870-895); total_proba keeps the pre-normalisation sum
	step("normalize");
	timer.begin(QB_PHASE_NORMALIZE);
	const double total = comm ? comm->sum_f64_agreed(local_total, err, "the finalisation") : local_total;
	next->total_proba = total;
	if (node_total_proba)
		*node_total_proba = total > 0 ? local_total / total : 0; // quids_mpi.hpp:892
	scale_state(ctx, next->mag.as<cplx>(), n_survivors, total);
	timer.end(QB_PHASE_NORMALIZE);
	ctx->sync();
	QB_CUDA(cudaGetLastError());
	timer.collect();
	step("end");
}

#include "migrate.inc.cuh"
#include "route.inc.cuh"

// ======================================================================================================
// one rule iteration over the GPUs of a communicator (see dist.inc.cuh for the protocol)
// ======================================================================================================
void simulate_dist(qb_iter *it, uint64_t rule_id, const rule_ops *ops, const void *rule, qb_iter *next, qb_sym *sym, qb_comm *cm, uint64_t max_num_object,
                   const qb_options &opt, qb_step_cb cb, void *user, double *node_total_proba) {
	qb_ctx *ctx = it->ctx;
	QB_REQUIRE(next->ctx == ctx && sym->ctx == ctx && cm->ctx == ctx, QB_ERR_ARG, "handles belong to different contexts");
	QB_REQUIRE(next != it, QB_ERR_ARG, "next_iteration must be a different object from iteration");
	// max_num_object = 0 (the reference's default, quids_mpi.hpp:423): automatic budget.  Parents: every rank keeps the most
	// probable of ITS parents whose workspace fits its GPU (per rank, as in the reference, :505-537).  Children: the global
	// top-k with k = what the ranks can hold, agreed by all ranks (below, after the owner merge).
	const bool automatic = max_num_object == 0;
	if (automatic)
		max_num_object = QB_NO_TRUNCATION;
	ctx->use();
	it->settle();
	next->settle();
	cudaStream_t stream = ctx->stream;
	stepper step{ctx, cb, user};
	phase_timer timer(sym, opt.profile != 0);
	comm_ops comm{cm};
	comm_ops::pending_error err; // see dist.inc.cuh: a failure on one rank is agreed on at the next count exchange
	const uint32_t world = (uint32_t)cm->world;
	const bool trace = getenv("QB_DIST_TRACE") != nullptr; // developer aid: wall-clock of every distributed sub-step (drains the stream)
	auto injected = [&](const char *phase) { inject_failure(cm->rank, phase); };
	auto t_last = std::chrono::steady_clock::now();
	auto mark = [&](const char *what) {
		if (!trace) return;
		ctx->sync();
		auto now = std::chrono::steady_clock::now();
		fprintf(stderr, "[qb dist rank %d] %-28s %8.3f ms\n", cm->rank, what, std::chrono::duration<double, std::milli>(now - t_last).count());
		t_last = now;
	};

	// 0. load balancing (quids_mpi.hpp:442-500): pairs of ranks level their objects or their children
	if (opt.equalize) {
		step(opt.equalize == 2 ? "equalize_child" : "equalize_object");
		int max_rounds = 0;
		while ((1u << max_rounds) < world) ++max_rounds; // utils::log_2_upper_bound(size), quids_mpi.hpp:432
		equalize_loop(it, cm, ops, rule, opt.equalize == 2, opt.min_equalize_size, opt.equalize_inbalance, opt.min_equalize_step, max_rounds);
		mark("equalize");
	}

	// 0b. rules with families (rule_api.cuh): the PARENTS go to the rank that owns their family, and the iteration needs no
	//     exchange of children at all (route.inc.cuh); otherwise the locally unique children are exchanged (below)
	if (ops->has_family && ops->region_size_limit > 0 && opt.family_routing != 0) {
		if (!cm->route)
			cm->route = new route_buffers();
		step("compute_collisions - com");
		timer.begin(QB_PHASE_EXCHANGE);
		const bool routed = route_by_family(it, cm, ops, rule, comm, *cm->route, trace);
		timer.end(QB_PHASE_EXCHANGE);
		mark("route parents by family");
		if (routed) {
			comm.local_interference = true;
			float exchange_ms = 0;
			if (opt.profile) { // simulate() resets the phase table: carry the routing time over
				ctx->sync();
				QB_CUDA(cudaEventElapsedTime(&exchange_ms, sym->ev[2 * QB_PHASE_EXCHANGE], sym->ev[2 * QB_PHASE_EXCHANGE + 1]));
			}
			simulate(it, rule_id, ops, rule, next, sym, automatic ? 0 : max_num_object, opt, cb, user, &comm, node_total_proba);
			sym->phase_ms[QB_PHASE_EXCHANGE] = exchange_ms;
			mark("local iteration on the owned families");
			return;
		}
	}

	engine_launch L;
	memset(&L, 0, sizeof L);
	L.stream = stream;
	L.sm_count = ctx->sm_count;
	L.launch_counter = &ctx->launches;
	L.it = it->view();

	// 1. local children, merged locally; the tolerance applies to GLOBAL sums only: keep every occupied slot.
	//    (The collectives inside -- global counts, the all-reduced select of the parents -- come BEFORE the data-dependent part,
	//    the interference table; what fails there is caught and agreed on at the count exchange of step 3.)
	local_table R;
	bool past_collectives = false;
	try {
		R = build_local_table(it, rule_id, ops, rule, sym, max_num_object, automatic, opt, -1.0, timer, step, L, &comm, &past_collectives);
		injected("local");
	} catch (const qb::error &e) {
		if (!past_collectives)
			throw;
		err.status = e.status;
		err.what = e.what();
	}
	if (err.ok() && R.empty_from >= 0) {
		finish_empty(ctx, next, step, timer, R.empty_from);
		if (node_total_proba) *node_total_proba = 0;
		return;
	}
	mark("local table");
	step("compute_collisions - prepare");

	// 2. partition the locally unique children by owner
	timer.begin(QB_PHASE_OWNER);
	const uint64_t n_local = err.ok() ? R.n_unique : 0;
	// records grouped by (owner, region of the owner's hashed table); a large exchange is merged through bins on the owner's
	// side (below), where the order of arrival does not matter: plain grouping by owner then
	const uint32_t sub = (opt.binned_inserts != 0 && n_local >= (1ull << 21)) ? 1u : owner_sub_buckets(world), bins = world * sub;
	unsigned long long *counts = nullptr, *cursor = nullptr;
	std::vector<uint64_t> send_counts(world, 0);
	err.run([&] {
		cm->cursors.ensure(sizeof(uint64_t) * 2 * (bins + 1), stream);
		counts = cm->cursors.as<unsigned long long>();
		cursor = counts + bins + 1;
		QB_CUDA(cudaMemsetAsync(counts, 0, sizeof(uint64_t) * 2 * (bins + 1), stream));
		injected("partition");
		if (n_local > 0) {
			const int grid_local = grid_for(n_local, 256, ctx->grid_cap());
			owner_count_kernel<<<grid_local, 256, sizeof(unsigned int) * bins, stream>>>(R.table, sym->uslot.as<uint32_t>(), n_local, world, sub, counts);
			++ctx->launches;
			std::vector<uint64_t> bin_counts(bins, 0), offsets(bins, 0);
			QB_CUDA(cudaMemcpyAsync(bin_counts.data(), counts, sizeof(uint64_t) * bins, cudaMemcpyDeviceToHost, stream));
			ctx->sync();
			for (uint32_t b = 0; b < bins; ++b) {
				send_counts[b / sub] += bin_counts[b];
				if (b > 0)
					offsets[b] = offsets[b - 1] + bin_counts[b - 1];
			}
			QB_CUDA(cudaMemcpyAsync(cursor, offsets.data(), sizeof(uint64_t) * bins, cudaMemcpyHostToDevice, stream));
			cm->send.ensure(sizeof(exchange_record) * n_local, stream);
			owner_scatter_kernel<<<grid_for(div_up<uint64_t>(n_local, SCATTER_TILE) * 256, 256, ctx->grid_cap()), 256, 2 * sizeof(unsigned long long) * bins, stream>>>(
			    R.table, sym->uslot.as<uint32_t>(), n_local, world, sub, cursor, cm->send.as<exchange_record>());
			++ctx->launches;
			ctx->sync(); // `offsets` lives on this stack frame
		}
	});
	if (!err.ok())
		std::fill(send_counts.begin(), send_counts.end(), 0);

	mark("partition by owner");
	// 3. all-to-allv of the records (its count exchange is the agreement point of steps 1 and 2)
	timer.end(QB_PHASE_OWNER);
	step("compute_collisions - com");
	timer.begin(QB_PHASE_EXCHANGE);
	uint64_t n_recv = 0;
	std::vector<uint64_t> recv_counts = comm.alltoallv(cm->send.ptr, send_counts, cm->recv, sizeof(exchange_record), n_recv, &err, "the local interference step");
	timer.end(QB_PHASE_EXCHANGE);
	std::vector<uint64_t> recv_begin(world + 1, 0);
	for (uint32_t r = 0; r < world; ++r)
		recv_begin[r + 1] = recv_begin[r] + recv_counts[r];

	mark("exchange records");
	// 4. owner: merge what arrived, apply the tolerance
	step("compute_collisions - insert");
	timer.begin(QB_PHASE_OWNER);
	table_view owner{};
	uint64_t n_owner_unique = 0;
	uint64_t *owner_keys = nullptr;  // compacted (norm key, slot) lists of the owner's unique objects
	uint32_t *owner_slots = nullptr;
	err.run([&] {
		injected("owner");
		if (n_recv == 0)
			return;
		// every rank sends each of its objects once, so an object arrives up to `world` times: the table is sized from the share
		// of the records that created a slot in the previous call (x 1.25), at most for "every record is a new object"; a
		// table that turns out too small is redone at that size
		const uint64_t full_capacity = std::max<uint64_t>(1024, (uint64_t)std::ceil((double)n_recv / 0.5));
		uint64_t capacity = full_capacity;
		if (cm->owner_unique_ratio > 0)
			capacity = std::min<uint64_t>(full_capacity, std::max<uint64_t>(1024, (uint64_t)((cm->owner_unique_ratio * 1.25 * (double)n_recv + 1024) / 0.5)));
		QB_REQUIRE(full_capacity + 1 <= 0xffffffffull, QB_ERR_CAPACITY, "owner table would need more than 2^32 slots");
		// binned owner merge (table.cuh): the records are streamed into bins and every bin is deduplicated in shared memory --
		// the owner's "table" is then the dense array of the unique objects; a bin that overflows falls back to the hashed table
		if (opt.binned_inserts != 0 && (opt.binned_inserts > 1 || n_recv >= (1ull << 22)) && n_recv <= (1ull << 29)) {
			// The local table, its (key, slot) lists and the local bins are dead once the records are in the send buffer: the
			// owner's dense array, lists and bins take their memory (at 1e7 parents per GPU that is 19 GB not allocated twice)
			capacity = n_recv;
			sym->table.ensure((capacity + 1) * sizeof(table_slot), stream);
			sym->ukey.ensure(sizeof(uint64_t) * (n_recv + 1), stream);
			sym->uslot.ensure(sizeof(uint32_t) * (n_recv + 1), stream);
			QB_CUDA(cudaMemsetAsync(sym->table.as<table_slot>() + capacity, 0, sizeof(table_slot), stream));
			QB_CUDA(cudaMemsetAsync(ctx->small(DS_COUNT), 0, 4 * sizeof(uint64_t), stream));
			owner = table_view{sym->table.as<table_slot>(), capacity, reinterpret_cast<unsigned int *>(ctx->small(DS_OVERFLOW)),
			                   reinterpret_cast<unsigned long long *>(ctx->small(DS_USED))};
			const bin_view ob = make_bins(ctx, bin_buffers{&sym->bin_records, &sym->bin_cursor, &sym->bin_spill, &sym->bin_spill_key}, n_recv,
			                              reinterpret_cast<unsigned long long *>(ctx->small(DS_SPILL)));
			record_bin_kernel<<<grid_for(n_recv, 256, ctx->grid_cap()), 256, 0, stream>>>(ob, owner, cm->recv.as<exchange_record>(), n_recv);
			++ctx->launches;
			launch_bin_dedup(ctx, sym, ob, owner, opt.tolerance, sym->ukey.as<uint64_t>(), sym->uslot.as<uint32_t>(), true);
			ctx->fetch_small();
			if (ctx->h_small[DS_OVERFLOW] == 0) {
				n_owner_unique = ctx->h_small[DS_COUNT];
				owner_keys = sym->ukey.as<uint64_t>();
				owner_slots = sym->uslot.as<uint32_t>();
				return;
			}
			capacity = full_capacity; // fall back to the hashed table below
		}
		cm->okey.ensure(sizeof(uint64_t) * (n_recv + 1), stream);
		cm->oslot.ensure(sizeof(uint32_t) * (n_recv + 1), stream);
		owner_keys = cm->okey.as<uint64_t>();
		owner_slots = cm->oslot.as<uint32_t>();
		for (int attempt = 0;; ++attempt) {
			cm->owner_table.ensure((capacity + 1) * sizeof(table_slot), stream);
			QB_CUDA(cudaMemsetAsync(cm->owner_table.ptr, 0, (capacity + 1) * sizeof(table_slot), stream));
			QB_CUDA(cudaMemsetAsync(ctx->small(DS_COUNT), 0, 4 * sizeof(uint64_t), stream));
			owner = table_view{cm->owner_table.as<table_slot>(), capacity, reinterpret_cast<unsigned int *>(ctx->small(DS_OVERFLOW)),
			                   reinterpret_cast<unsigned long long *>(ctx->small(DS_USED))};
			record_insert_kernel<<<grid_for(n_recv, 256, ctx->grid_cap()), 256, 0, stream>>>(owner, cm->recv.as<exchange_record>(), n_recv);
			++ctx->launches;
			const uint64_t tiles = div_up<uint64_t>(capacity + 1, COMPACT_TILE);
			table_compact_kernel<false><<<(unsigned)std::min<uint64_t>(tiles, (uint64_t)ctx->sm_count * 16), SCAN_THREADS, 0, stream>>>(
			    owner, capacity + 1, opt.tolerance, cm->okey.as<uint64_t>(), cm->oslot.as<uint32_t>(), reinterpret_cast<unsigned long long *>(ctx->small(DS_COUNT)),
			    nullptr, nullptr);
			++ctx->launches;
			QB_CUDA(cudaGetLastError());
			ctx->fetch_small();
			if (ctx->h_small[DS_OVERFLOW] == 0)
				break;
			QB_REQUIRE(capacity < full_capacity && attempt == 0, QB_ERR_CAPACITY, "owner table overflow");
			capacity = full_capacity;
		}
		cm->owner_unique_ratio = (double)ctx->h_small[DS_USED] / (double)n_recv;
		n_owner_unique = ctx->h_small[DS_COUNT];
	});
	timer.end(QB_PHASE_OWNER);
	step("compute_collisions - finalize");
	// automatic budget, children (quids.hpp:510-536): what this rank's next state has room for; all ranks then keep the global
	// top-k with k = half of (ranks x the smallest room): the survivors spread evenly over the ranks of their representatives
	// (pseudo-random choice, dist.inc.cuh), a factor 2 of head room covers the spread
	uint64_t room = 0;
	if (automatic)
		err.run([&] {
			const double budget = automatic_budget(ctx, sym, next, opt, R.workspace);
			const double per_object = (double)((R.max_child_size + 7u) & ~7u) + 48.0 + sizeof(survivor_record) * 2;
			room = budget > 0 ? (uint64_t)(budget / per_object) : 0;
			QB_REQUIRE(room >= 1, QB_ERR_CAPACITY, "max_num_object = 0 (automatic budget): no room left for a next state on this rank");
		});
	uint64_t n_unique_global = 0;
	{
		const uint64_t mine[2] = {n_owner_unique, automatic ? room : ~0ull};
		std::vector<uint64_t> all = comm.allgather_agreed(mine, 2, err, "the owner-side interference step");
		uint64_t min_room = ~0ull;
		for (uint32_t r = 0; r < world; ++r) {
			n_unique_global += all[2 * r];
			min_room = std::min(min_room, all[2 * r + 1]);
		}
		if (automatic)
			max_num_object = std::max<uint64_t>(1, (uint64_t)((double)min_room * world / 2));
	}
	sym->n_unique = n_owner_unique; // this rank's share; qb_comm_allreduce_u64 gives the total (get_total_num_object_after_interferences)

	mark("owner merge + compact");
	// 5. truncation: the max_num_object most probable over ALL ranks
	step("truncate - prepare");
	step("truncate");
	uint64_t n_owner_survivors = n_owner_unique;
	const uint32_t *owner_survivor_slot = owner_slots;
	if (max_num_object < n_unique_global) {
		timer.begin(QB_PHASE_TRUNCATE);
		sym->sslot.ensure(sizeof(uint32_t) * std::max<uint64_t>(1, std::min<uint64_t>(n_owner_unique, max_num_object)), stream);
		if (!opt.simple_truncation && n_owner_unique > 0) {
			randomize_keys_kernel<<<grid_for(n_owner_unique, 256, ctx->grid_cap()), 256, 0, stream>>>(owner_keys, owner, owner_slots,
			                                                                                         n_owner_unique, opt.seed);
			++ctx->launches;
		}
		key_from_array keys{owner_keys};
		select_threshold(ctx, &comm, keys, n_owner_unique, max_num_object);
		n_owner_survivors = select_keep(ctx, &comm, keys, n_owner_unique, out_gather_u32{sym->sslot.as<uint32_t>(), owner_slots});
		owner_survivor_slot = sym->sslot.as<uint32_t>();
		timer.end(QB_PHASE_TRUNCATE);
	}

	mark("global select");
	// 6. survivors go back to the rank of their representative
	timer.begin(QB_PHASE_OWNER);
	std::vector<uint64_t> back_counts(world, 0);
	err.run([&] {
		injected("return");
		QB_CUDA(cudaMemsetAsync(counts, 0, sizeof(uint64_t) * 2 * (world + 1), stream));
		if (n_owner_survivors > 0) {
			// recv_begin on the device
			dev_buf &begin_dev = cm->recv_begin; // kept across calls: a local buffer would cost a cudaMalloc and a synchronising cudaFree per step
			begin_dev.ensure(sizeof(uint64_t) * (world + 1), stream);
			QB_CUDA(cudaMemcpyAsync(begin_dev.ptr, recv_begin.data(), sizeof(uint64_t) * (world + 1), cudaMemcpyHostToDevice, stream));
			const int grid_back = grid_for(n_owner_survivors, 256, ctx->grid_cap());
			return_count_kernel<<<grid_back, 256, sizeof(unsigned int) * world, stream>>>(owner, owner_survivor_slot, n_owner_survivors, begin_dev.as<uint64_t>(), world, counts);
			++ctx->launches;
			QB_CUDA(cudaMemcpyAsync(back_counts.data(), counts, sizeof(uint64_t) * world, cudaMemcpyDeviceToHost, stream));
			ctx->sync();
			std::vector<uint64_t> offsets(world, 0);
			for (uint32_t r = 1; r < world; ++r)
				offsets[r] = offsets[r - 1] + back_counts[r - 1];
			QB_CUDA(cudaMemcpyAsync(cursor, offsets.data(), sizeof(uint64_t) * world, cudaMemcpyHostToDevice, stream));
			cm->ret_send.ensure(sizeof(survivor_record) * n_owner_survivors, stream);
			return_scatter_kernel<<<grid_for(div_up<uint64_t>(n_owner_survivors, SCATTER_TILE) * 256, 256, ctx->grid_cap()), 256, 2 * sizeof(unsigned long long) * world, stream>>>(owner, owner_survivor_slot, n_owner_survivors, begin_dev.as<uint64_t>(), world,
			                                                     cm->recv.as<exchange_record>(), cursor, cm->ret_send.as<survivor_record>());
			++ctx->launches;
			ctx->sync();
		}
	});
	if (!err.ok())
		std::fill(back_counts.begin(), back_counts.end(), 0);
	timer.end(QB_PHASE_OWNER);
	step("compute_collisions - com");
	timer.begin(QB_PHASE_EXCHANGE);
	uint64_t n_survivors = 0;
	comm.alltoallv(cm->ret_send.ptr, back_counts, cm->ret_recv, sizeof(survivor_record), n_survivors, &err, "the return of the survivors");
	timer.end(QB_PHASE_EXCHANGE);

	mark("return survivors");
	// 7. every rank rebuilds the survivors whose representative it generated, then global normalisation
	survivor_source src;
	src.records = cm->ret_recv.as<survivor_record>();
	bool fail_finalize = false;
	try {
		injected("finalize");
	} catch (const qb::error &) {
		fail_finalize = true;
	}
	finalize_and_normalize(it, ops, rule, next, sym, opt, R, src, n_survivors, timer, step, L, &comm, node_total_proba, fail_finalize);
	mark("finalize + normalize");
}

} // namespace

// ======================================================================================================
// C ABI
// ======================================================================================================
extern "C" {

void qb_options_default(qb_options *opt) {
	memset(opt, 0, sizeof *opt);
	opt->tolerance = 1e-30;     // TOLERANCE, quids.hpp:30-32
	opt->align_byte_length = 8; // ALIGNMENT_BYTE_LENGTH, quids.hpp:27-29
	opt->simple_truncation = 1;
	opt->table_load = 0;
	opt->profile = 0;
	opt->locality_sort = 1;
	opt->binned_inserts = 1;
	opt->family_routing = 1;
	opt->safety_margin = 0.2f; // SAFETY_MARGIN, quids.hpp:33-35
	opt->memory_budget = 0;
	opt->equalize = 0;
	opt->equalize_inbalance = 0.1f; // EQUALIZE_INBALANCE, quids_mpi.hpp:28-30
	opt->min_equalize_step = 0.2f;  // MIN_INBALANCE_STEP, quids_mpi.hpp:31-33
	opt->min_equalize_size = 100;   // MIN_EQUALIZE_SIZE, quids_mpi.hpp:13-15
}

const char *qb_last_error(void) { return g_last_error.c_str(); }
int qb_version(void) { return 100; }

int qb_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

int qb_ctx_create(int device, qb_ctx **out) {
	return guarded([&] {
		QB_REQUIRE(out, QB_ERR_ARG, "qb_ctx_create: null output");
		int count = 0;
		QB_CUDA(cudaGetDeviceCount(&count));
		QB_REQUIRE(device >= 0 && device < count, QB_ERR_CUDA, "qb_ctx_create: no such CUDA device (there is no CPU fallback)");
		QB_CUDA(cudaSetDevice(device));
		qb_ctx *ctx = new qb_ctx();
		ctx->device = device;
		// the interference table is hit at random, one 32-byte slot at a time: do not let L2 widen the fills.  The limit is
		// device-wide (it also applies to other libraries in this process): the value found here is put back by qb_ctx_destroy,
		// and QB_KEEP_L2_FETCH_GRANULARITY=1 leaves it alone altogether
		if (!getenv("QB_KEEP_L2_FETCH_GRANULARITY")) {
			if (cudaDeviceGetLimit(&ctx->l2_fetch_granularity_before, cudaLimitMaxL2FetchGranularity) != cudaSuccess)
				ctx->l2_fetch_granularity_before = 0;
			cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
			cudaGetLastError();
		}
		QB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
		QB_CUDA(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
		{ // dev_buf allocates stream-ordered from the device's default pool: keep what is freed cached instead of returning it to
		  // the driver at every synchronisation (the tables of consecutive rule iterations are GB-sized and change size)
			cudaMemPool_t pool = nullptr;
			QB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
			uint64_t keep = ~0ull;
			QB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
		}
		QB_CUDA(cudaHostAlloc((void **)&ctx->h_small, DS_WORDS * sizeof(uint64_t), cudaHostAllocMapped));
		ctx->d_small.ensure(DS_WORDS * sizeof(uint64_t), ctx->stream);
		QB_CUDA(cudaMemsetAsync(ctx->d_small.ptr, 0, DS_WORDS * sizeof(uint64_t), ctx->stream));
		ctx->sync();
		*out = ctx;
	});
}

int qb_ctx_destroy(qb_ctx *ctx) {
	return guarded([&] {
		if (!ctx) return;
		cudaSetDevice(ctx->device);
		cudaStreamSynchronize(ctx->stream);
		if (ctx->h_small) cudaFreeHost(ctx->h_small);
		ctx->d_small.release();
		ctx->scan_ws.release();
		ctx->partials.release();
		ctx->select.release();
		ctx->select_cand.release();
		cudaStreamDestroy(ctx->stream);
		if (ctx->l2_fetch_granularity_before) {
			cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, ctx->l2_fetch_granularity_before);
			cudaGetLastError();
		}
		delete ctx;
	});
}

int qb_ctx_synchronize(qb_ctx *ctx) {
	return guarded([&] {
		QB_REQUIRE(ctx, QB_ERR_ARG, "null context");
		ctx->sync();
	});
}
void *qb_ctx_stream(qb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
uint64_t qb_ctx_launch_count(const qb_ctx *ctx) { return ctx ? ctx->launches : 0; }

int qb_host_alloc(size_t bytes, void **out) {
	return guarded([&] { QB_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault)); });
}
int qb_host_free(void *p) {
	return guarded([&] {
		if (p) QB_CUDA(cudaFreeHost(p));
	});
}

// ---- iteration ---------------------------------------------------------------------------------------
int qb_iter_create(qb_ctx *ctx, qb_iter **out) {
	return guarded([&] {
		QB_REQUIRE(ctx && out, QB_ERR_ARG, "qb_iter_create: null argument");
		ctx->use();
		qb_iter *it = new qb_iter();
		it->ctx = ctx;
		it->begin.ensure(sizeof(uint64_t), ctx->stream);
		QB_CUDA(cudaMemsetAsync(it->begin.ptr, 0, sizeof(uint64_t), ctx->stream)); // object_begin[0] = 0, quids.hpp:160
		ctx->sync();
		*out = it;
	});
}

int qb_iter_destroy(qb_iter *it) {
	return guarded([&] {
		if (!it) return;
		it->ctx->use();
		if (it->uploaded) {
			cudaEventSynchronize(it->uploaded);
			cudaEventSynchronize(it->downloaded);
			cudaEventDestroy(it->uploaded);
			cudaEventDestroy(it->downloaded);
		}
		it->ctx->sync();
		delete it;
	});
}

int qb_iter_upload(qb_iter *it, uint64_t n, const uint8_t *objects, uint64_t num_bytes, const uint64_t *object_begin, const uint32_t *object_size,
                   const double *magnitude, double total_proba) {
	return guarded([&] {
		QB_REQUIRE(it, QB_ERR_ARG, "null iteration");
		QB_REQUIRE(n == 0 || (object_begin && object_size && magnitude), QB_ERR_ARG, "qb_iter_upload: null array");
		QB_REQUIRE(num_bytes == 0 || objects, QB_ERR_ARG, "qb_iter_upload: null objects");
		qb_ctx *ctx = it->ctx;
		ctx->use();
		it->settle();
		cudaStream_t s = ctx->stream;
		it->objects.ensure(num_bytes + 16, s);
		it->begin.ensure(sizeof(uint64_t) * (n + 1), s);
		it->size.ensure(sizeof(uint32_t) * (n ? n : 1), s);
		it->mag.ensure(sizeof(cplx) * (n ? n : 1), s);
		if (num_bytes) QB_CUDA(cudaMemcpyAsync(it->objects.ptr, objects, num_bytes, cudaMemcpyHostToDevice, s));
		if (n) {
			QB_CUDA(cudaMemcpyAsync(it->begin.ptr, object_begin, sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, s));
			QB_CUDA(cudaMemcpyAsync(it->size.ptr, object_size, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, s));
			QB_CUDA(cudaMemcpyAsync(it->mag.ptr, magnitude, sizeof(cplx) * n, cudaMemcpyHostToDevice, s));
		} else {
			QB_CUDA(cudaMemsetAsync(it->begin.ptr, 0, sizeof(uint64_t), s));
		}
		ctx->sync();
		it->n = n;
		it->n_bytes = num_bytes;
		it->total_proba = total_proba;
	});
}

// ---- PROBA_TYPE = float at the boundary (quids.hpp:21-23): magnitudes cross the C ABI as complex<float>; the state in HBM
// and all device arithmetic stay double (at least the reference's float accuracy: parity 1e-5) ---------------
__global__ void __launch_bounds__(256) widen_mag_kernel(const float2 *in, cplx *out, uint64_t n) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const float2 v = in[i];
		out[i] = cplx{(double)v.x, (double)v.y};
	}
}
__global__ void __launch_bounds__(256) narrow_mag_kernel(const cplx *in, float2 *out, uint64_t n) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const cplx v = in[i];
		out[i] = make_float2((float)v.re, (float)v.im);
	}
}

int qb_iter_upload_f32(qb_iter *it, uint64_t n, const uint8_t *objects, uint64_t num_bytes, const uint64_t *object_begin, const uint32_t *object_size,
                       const float *magnitude, double total_proba) {
	return guarded([&] {
		QB_REQUIRE(it, QB_ERR_ARG, "null iteration");
		QB_REQUIRE(n == 0 || (object_begin && object_size && magnitude), QB_ERR_ARG, "qb_iter_upload_f32: null array");
		QB_REQUIRE(num_bytes == 0 || objects, QB_ERR_ARG, "qb_iter_upload_f32: null objects");
		qb_ctx *ctx = it->ctx;
		ctx->use();
		it->settle();
		cudaStream_t s = ctx->stream;
		it->objects.ensure(num_bytes + 16, s);
		it->begin.ensure(sizeof(uint64_t) * (n + 1), s);
		it->size.ensure(sizeof(uint32_t) * (n ? n : 1), s);
		it->mag.ensure(sizeof(cplx) * (n ? n : 1), s);
		if (num_bytes) QB_CUDA(cudaMemcpyAsync(it->objects.ptr, objects, num_bytes, cudaMemcpyHostToDevice, s));
		if (n) {
			dev_buf staging;
			staging.ensure(sizeof(float2) * n, s);
			QB_CUDA(cudaMemcpyAsync(it->begin.ptr, object_begin, sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, s));
			QB_CUDA(cudaMemcpyAsync(it->size.ptr, object_size, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, s));
			QB_CUDA(cudaMemcpyAsync(staging.ptr, magnitude, sizeof(float2) * n, cudaMemcpyHostToDevice, s));
			widen_mag_kernel<<<grid_for(n, 256, ctx->grid_cap()), 256, 0, s>>>(staging.as<float2>(), it->mag.as<cplx>(), n);
			++ctx->launches;
			ctx->sync();
		} else {
			QB_CUDA(cudaMemsetAsync(it->begin.ptr, 0, sizeof(uint64_t), s));
			ctx->sync();
		}
		it->n = n;
		it->n_bytes = num_bytes;
		it->total_proba = total_proba;
	});
}

int qb_iter_download_f32(const qb_iter *it, uint8_t *objects, uint64_t *object_begin, uint32_t *object_size, float *magnitude) {
	return guarded([&] {
		QB_REQUIRE(it, QB_ERR_ARG, "null iteration");
		qb_ctx *ctx = it->ctx;
		ctx->use();
		it->settle();
		cudaStream_t s = ctx->stream;
		if (objects && it->n_bytes) QB_CUDA(cudaMemcpyAsync(objects, it->objects.ptr, it->n_bytes, cudaMemcpyDeviceToHost, s));
		if (object_begin) QB_CUDA(cudaMemcpyAsync(object_begin, it->begin.ptr, sizeof(uint64_t) * (it->n + 1), cudaMemcpyDeviceToHost, s));
		if (object_size && it->n) QB_CUDA(cudaMemcpyAsync(object_size, it->size.ptr, sizeof(uint32_t) * it->n, cudaMemcpyDeviceToHost, s));
		dev_buf staging;
		if (magnitude && it->n) {
			staging.ensure(sizeof(float2) * it->n, s);
			narrow_mag_kernel<<<grid_for(it->n, 256, ctx->grid_cap()), 256, 0, s>>>(it->mag.as<cplx>(), staging.as<float2>(), it->n);
			++ctx->launches;
			QB_CUDA(cudaMemcpyAsync(magnitude, staging.ptr, sizeof(float2) * it->n, cudaMemcpyDeviceToHost, s));
		}
		ctx->sync();
	});
}

// ---- transfers that overlap the rule iterations: pinned host memory <-> HBM on dedicated copy streams --------------
// A copy engine runs one copy to completion before it looks at another stream: behind a 2.5 GB object array the 8-byte
// counter fetches of the rule iteration running meanwhile (fetch_small and friends, on the compute stream, same engine
// per direction) would wait for the whole array, and the iteration with them. Bulk transfers are therefore cut into
// pieces between which the engine can serve the compute stream (QB_COPY_CHUNK_MB, default 8; 0 = one copy per array).
static size_t copy_chunk_bytes() {
	static const size_t bytes = [] {
		const char *e = getenv("QB_COPY_CHUNK_MB");
		const long mb = e ? atol(e) : 8;
		return mb > 0 ? (size_t)mb << 20 : (size_t)0;
	}();
	return bytes;
}
static void copy_in_pieces(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s) {
	const size_t piece = copy_chunk_bytes();
	if (piece == 0 || bytes <= piece) {
		QB_CUDA(cudaMemcpyAsync(dst, src, bytes, kind, s));
		return;
	}
	for (size_t done = 0; done < bytes; done += piece)
		QB_CUDA(cudaMemcpyAsync((char *)dst + done, (const char *)src + done, std::min(piece, bytes - done), kind, s));
}

static void transfer_events(qb_iter *it) {
	it->ctx->copy_streams();
	if (!it->uploaded) {
		QB_CUDA(cudaEventCreateWithFlags(&it->uploaded, cudaEventDisableTiming));
		QB_CUDA(cudaEventCreateWithFlags(&it->downloaded, cudaEventDisableTiming));
	}
}

int qb_iter_upload_async(qb_iter *it, uint64_t n, const uint8_t *objects, uint64_t num_bytes, const uint64_t *object_begin, const uint32_t *object_size,
                         const double *magnitude, double total_proba) {
	return guarded([&] {
		QB_REQUIRE(it, QB_ERR_ARG, "null iteration");
		QB_REQUIRE(n == 0 || (object_begin && object_size && magnitude), QB_ERR_ARG, "qb_iter_upload_async: null array");
		QB_REQUIRE(num_bytes == 0 || objects, QB_ERR_ARG, "qb_iter_upload_async: null objects");
		qb_ctx *ctx = it->ctx;
		ctx->use();
		transfer_events(it);
		cudaStream_t s = ctx->copy_in;
		// buffers grow on the compute stream (stream-ordered allocation), BEFORE the copy stream is ordered after it
		it->objects.ensure(num_bytes + 16, ctx->stream);
		it->begin.ensure(sizeof(uint64_t) * (n + 1), ctx->stream);
		it->size.ensure(sizeof(uint32_t) * (n ? n : 1), ctx->stream);
		it->mag.ensure(sizeof(cplx) * (n ? n : 1), ctx->stream);
		// the old contents may still be in use: by kernels already enqueued, or by a download in flight
		ctx->order_after_compute(s);
		if (it->download_pending)
			QB_CUDA(cudaStreamWaitEvent(s, it->downloaded, 0));
		if (num_bytes) copy_in_pieces(it->objects.ptr, objects, num_bytes, cudaMemcpyHostToDevice, s);
		if (n) {
			copy_in_pieces(it->begin.ptr, object_begin, sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, s);
			copy_in_pieces(it->size.ptr, object_size, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, s);
			copy_in_pieces(it->mag.ptr, magnitude, sizeof(cplx) * n, cudaMemcpyHostToDevice, s);
		} else {
			QB_CUDA(cudaMemsetAsync(it->begin.ptr, 0, sizeof(uint64_t), s));
		}
		QB_CUDA(cudaEventRecord(it->uploaded, s));
		it->upload_pending = true;
		it->n = n;
		it->n_bytes = num_bytes;
		it->total_proba = total_proba;
	});
}

int qb_iter_download_async(const qb_iter *cit, uint8_t *objects, uint64_t *object_begin, uint32_t *object_size, double *magnitude) {
	return guarded([&] {
		QB_REQUIRE(cit, QB_ERR_ARG, "null iteration");
		qb_iter *it = const_cast<qb_iter *>(cit);
		qb_ctx *ctx = it->ctx;
		ctx->use();
		transfer_events(it);
		cudaStream_t s = ctx->copy_out;
		ctx->order_after_compute(s); // the state as the calls made so far leave it
		if (it->upload_pending)
			QB_CUDA(cudaStreamWaitEvent(s, it->uploaded, 0));
		if (objects && it->n_bytes) copy_in_pieces(objects, it->objects.ptr, it->n_bytes, cudaMemcpyDeviceToHost, s);
		if (object_begin) copy_in_pieces(object_begin, it->begin.ptr, sizeof(uint64_t) * (it->n + 1), cudaMemcpyDeviceToHost, s);
		if (object_size && it->n) copy_in_pieces(object_size, it->size.ptr, sizeof(uint32_t) * it->n, cudaMemcpyDeviceToHost, s);
		if (magnitude && it->n) copy_in_pieces(magnitude, it->mag.ptr, sizeof(cplx) * it->n, cudaMemcpyDeviceToHost, s);
		QB_CUDA(cudaEventRecord(it->downloaded, s));
		it->download_pending = true;
	});
}

int qb_iter_wait(const qb_iter *it) {
	return guarded([&] {
		QB_REQUIRE(it, QB_ERR_ARG, "null iteration");
		it->ctx->use();
		if (it->uploaded) {
			QB_CUDA(cudaEventSynchronize(it->uploaded));
			QB_CUDA(cudaEventSynchronize(it->downloaded));
		}
		it->upload_pending = it->download_pending = false;
	});
}

int qb_iter_counts(const qb_iter *it, uint64_t *num_object, uint64_t *num_bytes, double *total_proba) {
	return guarded([&] {
		QB_REQUIRE(it, QB_ERR_ARG, "null iteration");
		if (num_object) *num_object = it->n;
		if (num_bytes) *num_bytes = it->n_bytes;
		if (total_proba) *total_proba = it->total_proba;
	});
}

int qb_iter_download(const qb_iter *it, uint8_t *objects, uint64_t *object_begin, uint32_t *object_size, double *magnitude) {
	return guarded([&] {
		QB_REQUIRE(it, QB_ERR_ARG, "null iteration");
		qb_ctx *ctx = it->ctx;
		ctx->use();
		it->settle();
		cudaStream_t s = ctx->stream;
		if (objects && it->n_bytes) QB_CUDA(cudaMemcpyAsync(objects, it->objects.ptr, it->n_bytes, cudaMemcpyDeviceToHost, s));
		if (object_begin) QB_CUDA(cudaMemcpyAsync(object_begin, it->begin.ptr, sizeof(uint64_t) * (it->n + 1), cudaMemcpyDeviceToHost, s));
		if (object_size && it->n) QB_CUDA(cudaMemcpyAsync(object_size, it->size.ptr, sizeof(uint32_t) * it->n, cudaMemcpyDeviceToHost, s));
		if (magnitude && it->n) QB_CUDA(cudaMemcpyAsync(magnitude, it->mag.ptr, sizeof(cplx) * it->n, cudaMemcpyDeviceToHost, s));
		ctx->sync();
	});
}

int qb_iter_device_ptrs(const qb_iter *it, void **objects, void **object_begin, void **object_size, void **magnitude) {
	return guarded([&] {
		QB_REQUIRE(it, QB_ERR_ARG, "null iteration");
		it->ctx->use();
		it->settle(); // what the caller enqueues on the context's stream comes after the transfers in flight
		if (objects) *objects = it->objects.ptr;
		if (object_begin) *object_begin = it->begin.ptr;
		if (object_size) *object_size = it->size.ptr;
		if (magnitude) *magnitude = it->mag.ptr;
	});
}

qb_ctx *qb_iter_ctx(const qb_iter *it) { return it ? it->ctx : nullptr; }
int qb_ctx_device(const qb_ctx *ctx) { return ctx ? ctx->device : -1; }

int qb_iter_append_state(qb_iter *it, const qb_iter *other) {
	return guarded([&] {
		QB_REQUIRE(it && other && it != other && it->ctx == other->ctx, QB_ERR_ARG, "qb_iter_append_state: bad handle");
		qb_ctx *ctx = it->ctx;
		ctx->use();
		it->settle();
		other->settle();
		const uint64_t n = other->n, bytes = other->n_bytes;
		if (n == 0) return;
		cudaStream_t s = ctx->stream;
		it->objects.ensure(it->n_bytes + bytes + 16, s, true, it->n_bytes);
		it->begin.ensure(sizeof(uint64_t) * (it->n + n + 1), s, true, sizeof(uint64_t) * (it->n + 1));
		it->size.ensure(sizeof(uint32_t) * (it->n + n), s, true, sizeof(uint32_t) * it->n);
		it->mag.ensure(sizeof(cplx) * (it->n + n), s, true, sizeof(cplx) * it->n);
		QB_CUDA(cudaMemcpyAsync(it->objects.as<uint8_t>() + it->n_bytes, other->objects.ptr, bytes, cudaMemcpyDeviceToDevice, s));
		QB_CUDA(cudaMemcpyAsync(it->begin.as<uint64_t>() + it->n + 1, other->begin.as<uint64_t>() + 1, sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, s));
		QB_CUDA(cudaMemcpyAsync(it->size.as<uint32_t>() + it->n, other->size.ptr, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, s));
		QB_CUDA(cudaMemcpyAsync(it->mag.as<cplx>() + it->n, other->mag.ptr, sizeof(cplx) * n, cudaMemcpyDeviceToDevice, s));
		if (it->n_bytes) {
			rebase_begin_kernel<<<grid_for(n, 256, ctx->grid_cap()), 256, 0, s>>>(it->begin.as<uint64_t>() + it->n + 1, n, 0, it->n_bytes);
			++ctx->launches;
		}
		ctx->sync();
		it->n += n;
		it->n_bytes += bytes;
	});
}

int qb_iter_normalize(qb_iter *it) {
	return guarded([&] {
		QB_REQUIRE(it, QB_ERR_ARG, "null iteration");
		qb_ctx *ctx = it->ctx;
		ctx->use();
		it->settle();
		it->total_proba = 0; // quids.hpp:986
		if (it->n == 0) return;
		it->total_proba = reduce_norm_total(ctx, it->mag.as<cplx>(), it->n);
		scale_state(ctx, it->mag.as<cplx>(), it->n, it->total_proba);
		ctx->sync();
	});
}

int qb_iter_pop(qb_iter *it, uint64_t n, int normalize) {
	int rc = guarded([&] {
		QB_REQUIRE(it, QB_ERR_ARG, "null iteration");
		QB_REQUIRE(n <= it->n, QB_ERR_ARG, "qb_iter_pop: more objects than the state holds");
		if (n < 1) return; // quids.hpp:195-196
		qb_ctx *ctx = it->ctx;
		ctx->use();
		it->settle();
		it->n -= n;
		peek_list l;
		l.add(&ctx->h_small[DS_COUNT], it->begin.as<uint64_t>() + it->n);
		ctx->peek(l);
		ctx->sync();
		it->n_bytes = ctx->h_small[DS_COUNT];
	});
	if (rc == QB_OK && n >= 1 && normalize)
		rc = qb_iter_normalize(it);
	return rc;
}

// ---- symbolic iteration --------------------------------------------------------------------------------
int qb_sym_create(qb_ctx *ctx, qb_sym **out) {
	return guarded([&] {
		QB_REQUIRE(ctx && out, QB_ERR_ARG, "qb_sym_create: null argument");
		qb_sym *sym = new qb_sym();
		sym->ctx = ctx;
		*out = sym;
	});
}

int qb_sym_destroy(qb_sym *sym) {
	return guarded([&] {
		if (!sym) return;
		sym->ctx->use();
		sym->ctx->sync();
		for (cudaEvent_t e : sym->ev)
			if (e) cudaEventDestroy(e);
		delete sym;
	});
}

int qb_sym_counts(const qb_sym *sym, uint64_t *num_object, uint64_t *num_object_after_interferences) {
	return guarded([&] {
		QB_REQUIRE(sym, QB_ERR_ARG, "null symbolic iteration");
		if (num_object) *num_object = sym->n_children;
		if (num_object_after_interferences) *num_object_after_interferences = sym->n_unique;
	});
}

int qb_sym_phase_ms(const qb_sym *sym, float *ms) {
	return guarded([&] {
		QB_REQUIRE(sym && ms, QB_ERR_ARG, "null argument");
		for (int p = 0; p < QB_PHASE_COUNT; ++p)
			ms[p] = sym->phase_ms[p];
	});
}

uint64_t qb_sym_device_bytes(const qb_sym *sym) { return sym ? sym->device_bytes() : 0; }

// ---- rules, modifiers ------------------------------------------------------------------------------------
int qb_rule_id(const char *name) {
	if (name)
		for (size_t i = 0; i < rule_registry().size(); ++i)
			if (!strcmp(rule_registry()[i].name, name))
				return (int)i + 1;
	g_last_error = std::string("unknown rule: ") + (name ? name : "(null)");
	return QB_ERR_UNKNOWN_RULE;
}

int qb_modifier_id(const char *name) {
	if (name)
		for (size_t i = 0; i < modifier_registry().size(); ++i)
			if (!strcmp(modifier_registry()[i].name, name))
				return (int)i + 1;
	g_last_error = std::string("unknown modifier: ") + (name ? name : "(null)");
	return QB_ERR_UNKNOWN_RULE;
}

int qb_apply_modifier(qb_iter *it, int modifier_id, const double *params, uint32_t num_params) {
	return guarded([&] {
		QB_REQUIRE(it, QB_ERR_ARG, "null iteration");
		const modifier_ops *ops = find_modifier(modifier_id);
		QB_REQUIRE(ops, QB_ERR_UNKNOWN_RULE, "unknown modifier id");
		alignas(16) unsigned char storage[RULE_STORAGE_BYTES];
		int rc = ops->make(params, num_params, storage);
		QB_REQUIRE(rc == QB_OK, rc, std::string("bad parameters for modifier ") + ops->name);
		qb_ctx *ctx = it->ctx;
		ctx->use();
		it->settle();
		if (it->n == 0) return;
		ops->launch(storage, it->view(), ctx->stream, ctx->sm_count);
		++ctx->launches;
		QB_CUDA(cudaGetLastError());
		ctx->sync();
	});
}

int qb_observable_id(const char *name) {
	if (name)
		for (size_t i = 0; i < observable_registry().size(); ++i)
			if (!strcmp(observable_registry()[i].name, name))
				return (int)i + 1;
	g_last_error = std::string("unknown observable: ") + (name ? name : "(null)");
	return QB_ERR_UNKNOWN_RULE;
}

int qb_observable_values(int observable_id) {
	const observable_ops *ops = find_observable(observable_id);
	return ops ? ops->values : QB_ERR_UNKNOWN_RULE;
}

// iteration::average_value (quids.hpp:208-234) for a registered device observable: sum of observable(object) * |mag|^2
int qb_iter_average_value(const qb_iter *cit, int observable_id, const double *params, uint32_t num_params, double *values, uint32_t capacity) {
	return guarded([&] {
		qb_iter *it = const_cast<qb_iter *>(cit);
		QB_REQUIRE(it && values, QB_ERR_ARG, "qb_iter_average_value: null argument");
		const observable_ops *ops = find_observable(observable_id);
		QB_REQUIRE(ops, QB_ERR_UNKNOWN_RULE, "unknown observable id");
		QB_REQUIRE(capacity >= (uint32_t)ops->values, QB_ERR_ARG, std::string("observable ") + ops->name + " produces " + std::to_string(ops->values) + " values");
		alignas(16) unsigned char storage[RULE_STORAGE_BYTES];
		int rc = ops->make(params, num_params, storage);
		QB_REQUIRE(rc == QB_OK, rc, std::string("bad parameters for observable ") + ops->name);
		for (int k = 0; k < ops->values; ++k)
			values[k] = 0;
		if (it->n == 0)
			return;
		qb_ctx *ctx = it->ctx;
		ctx->use();
		it->settle();
		ctx->partials.ensure(sizeof(double) * (size_t)ctx->grid_cap() * OBSERVABLE_MAX_VALUES, ctx->stream);
		const int grid = ops->launch(storage, it->view(), ctx->partials.as<double>(), ctx->stream, ctx->sm_count);
		++ctx->launches;
		QB_REQUIRE(grid <= ctx->grid_cap(), QB_ERR_CUDA, "observable grid larger than the partial buffer");
		for (int k = 0; k < ops->values; ++k) { // fixed-order second stage, one value after the other
			norm_total_kernel<<<1, SCAN_THREADS, 0, ctx->stream>>>(ctx->partials.as<double>() + (size_t)k * grid, grid, reinterpret_cast<double *>(ctx->small(DS_TOTAL)));
			++ctx->launches;
			ctx->fetch_small();
			memcpy(&values[k], &ctx->h_small[DS_TOTAL], sizeof(double));
		}
		QB_CUDA(cudaGetLastError());
	});
}

int qb_simulate(qb_iter *it, int rule_id, const double *params, uint32_t num_params, qb_iter *next, qb_sym *sym, uint64_t max_num_object,
                const qb_options *opt_in, qb_step_cb cb, void *user) {
	return guarded([&] {
		QB_REQUIRE(it && next && sym, QB_ERR_ARG, "qb_simulate: null handle");
		const rule_ops *ops = find_rule(rule_id);
		QB_REQUIRE(ops, QB_ERR_UNKNOWN_RULE, "unknown rule id");
		alignas(16) unsigned char storage[RULE_STORAGE_BYTES];
		int rc = ops->make(params, num_params, storage);
		QB_REQUIRE(rc == QB_OK, rc, std::string("bad parameters for rule ") + ops->name);
		qb_options opt;
		resolve_options(opt_in, opt);
		simulate(it, rule_history_key(rule_id, params, num_params), ops, storage, next, sym, max_num_object, opt, cb, user);
	});
}

int qb_hash_objects(const qb_iter *it, int rule_id, const double *params, uint32_t num_params, uint64_t *hashes) {
	return guarded([&] {
		QB_REQUIRE(it && (hashes || it->n == 0), QB_ERR_ARG, "qb_hash_objects: null argument");
		const rule_ops *ops = find_rule(rule_id);
		QB_REQUIRE(ops, QB_ERR_UNKNOWN_RULE, "unknown rule id");
		alignas(16) unsigned char storage[RULE_STORAGE_BYTES];
		int rc = ops->make(params, num_params, storage);
		QB_REQUIRE(rc == QB_OK, rc, std::string("bad parameters for rule ") + ops->name);
		if (it->n == 0) return;
		qb_ctx *ctx = it->ctx;
		ctx->use();
		it->settle();
		dev_buf out;
		out.ensure(sizeof(uint64_t) * it->n, ctx->stream);
		engine_launch L;
		memset(&L, 0, sizeof L);
		L.stream = ctx->stream;
		L.sm_count = ctx->sm_count;
		L.launch_counter = &ctx->launches;
		L.it = it->view();
		L.hashes = out.as<uint64_t>();
		ops->launch_hash(storage, L);
		QB_CUDA(cudaGetLastError());
		QB_CUDA(cudaMemcpyAsync(hashes, out.ptr, sizeof(uint64_t) * it->n, cudaMemcpyDeviceToHost, ctx->stream));
		ctx->sync();
	});
}

// ---- distributed path (dist.inc.cuh) -------------------------------------------------------------------
int qb_comm_unique_id(uint8_t id[128]) {
	return guarded([&] {
		QB_REQUIRE(id, QB_ERR_ARG, "null id");
		QB_REQUIRE(nccl().ok, QB_ERR_COMM, "NCCL is not usable: " + nccl().why);
		static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
		ncclUniqueId uid;
		QB_NCCL(nccl().GetUniqueId(&uid));
		memcpy(id, &uid, 128);
	});
}

int qb_comm_create(qb_ctx *ctx, int world_size, int rank, const uint8_t id[128], qb_comm **out) {
	return guarded([&] {
		QB_REQUIRE(ctx && id && out && world_size >= 1 && rank >= 0 && rank < world_size, QB_ERR_ARG, "qb_comm_create: bad argument");
		QB_REQUIRE(nccl().ok, QB_ERR_COMM, "NCCL is not usable: " + nccl().why);
		ctx->use();
		ncclUniqueId uid;
		memcpy(&uid, id, 128);
		qb_comm *c = new qb_comm();
		c->ctx = ctx;
		c->world = world_size;
		c->rank = rank;
		QB_NCCL(nccl().CommInitRank(&c->nccl, world_size, uid, rank));
		*out = c;
	});
}

int qb_comm_destroy(qb_comm *comm) {
	return guarded([&] {
		if (!comm) return;
		comm->ctx->use();
		comm->ctx->sync();
		if (comm->nccl) nccl().CommDestroy(comm->nccl);
		delete comm->route;
		delete comm;
	});
}

int qb_simulate_dist(qb_iter *it, int rule_id, const double *params, uint32_t num_params, qb_iter *next, qb_sym *sym, qb_comm *comm,
                     uint64_t max_num_object, const qb_options *opt_in, qb_step_cb cb, void *user, double *node_total_proba) {
	return guarded([&] {
		QB_REQUIRE(it && next && sym && comm, QB_ERR_ARG, "qb_simulate_dist: null handle");
		const rule_ops *ops = find_rule(rule_id);
		QB_REQUIRE(ops, QB_ERR_UNKNOWN_RULE, "unknown rule id");
		alignas(16) unsigned char storage[RULE_STORAGE_BYTES];
		int rc = ops->make(params, num_params, storage);
		QB_REQUIRE(rc == QB_OK, rc, std::string("bad parameters for rule ") + ops->name);
		qb_options opt;
		resolve_options(opt_in, opt);
		if (comm->world == 1) { // quids_mpi.hpp:439-440
			simulate(it, rule_history_key(rule_id, params, num_params), ops, storage, next, sym, max_num_object, opt, cb, user);
			if (node_total_proba) *node_total_proba = 1;
			return;
		}
		simulate_dist(it, rule_history_key(rule_id, params, num_params), ops, storage, next, sym, comm, max_num_object, opt, cb, user, node_total_proba);
	});
}

int qb_iter_send_objects(qb_iter *it, qb_comm *comm, uint64_t num_object_sent, int node, uint64_t *moved) {
	return guarded([&] {
		QB_REQUIRE(it && comm && it->ctx == comm->ctx, QB_ERR_ARG, "qb_iter_send_objects: bad handle");
		it->ctx->use();
		it->settle();
		const uint64_t n = send_objects(it, comm, num_object_sent, node);
		if (moved) *moved = n;
	});
}

int qb_iter_receive_objects(qb_iter *it, qb_comm *comm, int node, uint64_t max_mem, uint64_t *moved) {
	return guarded([&] {
		QB_REQUIRE(it && comm && it->ctx == comm->ctx, QB_ERR_ARG, "qb_iter_receive_objects: bad handle");
		it->ctx->use();
		it->settle();
		const uint64_t n = receive_objects(it, comm, node, max_mem);
		if (moved) *moved = n;
	});
}

int qb_iter_distribute_objects(qb_iter *it, qb_comm *comm, int node_id) {
	return guarded([&] {
		QB_REQUIRE(it && comm && it->ctx == comm->ctx, QB_ERR_ARG, "qb_iter_distribute_objects: bad handle");
		QB_REQUIRE(node_id >= 0 && node_id < comm->world, QB_ERR_ARG, "qb_iter_distribute_objects: bad node id");
		it->ctx->use();
		it->settle();
		if (comm->rank == node_id) {
			const uint64_t initial = it->n;
			for (int node = 1; node < comm->world; ++node) {
				const int node_to_send = node <= node_id ? node - 1 : node; // skip this node, quids_mpi.hpp:1040
				const uint64_t share = (uint64_t)(((unsigned __int128)initial * (node + 1)) / comm->world - ((unsigned __int128)initial * node) / comm->world);
				send_objects(it, comm, share, node_to_send);
			}
		} else {
			receive_objects(it, comm, node_id, ~0ull);
		}
	});
}

int qb_iter_gather_objects(qb_iter *it, qb_comm *comm, int node_id) {
	return guarded([&] {
		QB_REQUIRE(it && comm && it->ctx == comm->ctx, QB_ERR_ARG, "qb_iter_gather_objects: bad handle");
		QB_REQUIRE(node_id >= 0 && node_id < comm->world, QB_ERR_ARG, "qb_iter_gather_objects: bad node id");
		it->ctx->use();
		it->settle();
		if (comm->rank == node_id) {
			for (int node = 1; node < comm->world; ++node)
				receive_objects(it, comm, node <= node_id ? node - 1 : node, ~0ull);
		} else {
			send_objects(it, comm, it->n, node_id);
		}
	});
}

int qb_iter_num_symbolic_object(const qb_iter *it, uint64_t *num_symbolic_object) {
	return guarded([&] {
		QB_REQUIRE(it && num_symbolic_object, QB_ERR_ARG, "qb_iter_num_symbolic_object: null argument");
		*num_symbolic_object = it->n_symbolic;
	});
}

int qb_iter_count_children(qb_iter *it, int rule_id, const double *params, uint32_t num_params, uint64_t *num_children) {
	return guarded([&] {
		QB_REQUIRE(it && num_children, QB_ERR_ARG, "qb_iter_count_children: null argument");
		const rule_ops *ops = find_rule(rule_id);
		QB_REQUIRE(ops, QB_ERR_UNKNOWN_RULE, "unknown rule id");
		alignas(16) unsigned char storage[RULE_STORAGE_BYTES];
		int rc = ops->make(params, num_params, storage);
		QB_REQUIRE(rc == QB_OK, rc, std::string("bad parameters for rule ") + ops->name);
		it->ctx->use();
		it->settle();
		*num_children = count_children(it, ops, storage);
	});
}

int qb_iter_equalize(qb_iter *it, qb_comm *comm, int rule_id, const double *params, uint32_t num_params, int max_rounds, uint64_t min_equalize_size,
                     float equalize_inbalance, float min_equalize_step, int *rounds) {
	return guarded([&] {
		QB_REQUIRE(it && comm && it->ctx == comm->ctx, QB_ERR_ARG, "qb_iter_equalize: bad handle");
		const rule_ops *ops = nullptr;
		alignas(16) unsigned char storage[RULE_STORAGE_BYTES];
		if (rule_id != 0) {
			ops = find_rule(rule_id);
			QB_REQUIRE(ops, QB_ERR_UNKNOWN_RULE, "unknown rule id");
			int rc = ops->make(params, num_params, storage);
			QB_REQUIRE(rc == QB_OK, rc, std::string("bad parameters for rule ") + ops->name);
		}
		it->ctx->use();
		it->settle();
		const int r = comm->world > 1 ? equalize_loop(it, comm, ops, storage, ops != nullptr, min_equalize_size, equalize_inbalance, min_equalize_step, max_rounds) : 0;
		if (rounds) *rounds = r;
	});
}

int qb_comm_allreduce_u64(qb_comm *comm, uint64_t *values, uint32_t n, int op_max) {
	return guarded([&] {
		QB_REQUIRE(comm && values, QB_ERR_ARG, "null argument");
		comm->ctx->use();
		comm_ops ops{comm};
		std::vector<uint64_t> all = ops.allgather_u64(values, n);
		for (uint32_t i = 0; i < n; ++i) {
			uint64_t acc = 0;
			for (int r = 0; r < comm->world; ++r)
				acc = op_max ? std::max(acc, all[(size_t)r * n + i]) : acc + all[(size_t)r * n + i];
			values[i] = acc;
		}
	});
}

int qb_comm_allreduce_f64(qb_comm *comm, double *values, uint32_t n) {
	return guarded([&] {
		QB_REQUIRE(comm && values, QB_ERR_ARG, "null argument");
		comm->ctx->use();
		comm_ops ops{comm};
		for (uint32_t i = 0; i < n; ++i)
			values[i] = ops.sum_f64(values[i]);
	});
}
}
