// dist.inc.cuh -- the distributed path: quids::mpi::simulate (quids_mpi.hpp:423-598) over NCCL.
// Textually included by capi.cu (it needs the handle structs and the pipeline stages defined there).
//
// The reference ships EVERY child's (hash, magnitude) to the rank that owns the hash
// (MPI_Alltoallv x2, quids_mpi.hpp:741-743), merges there, and ships a magnitude back per child
// (:842).  Here every GPU first merges its own children in its local interference table (the very
// kernel of the single-GPU path), so only the locally unique (hash, magnitude, representative)
// records cross NVLink -- for the workloads of BASELINE.json that is 1-10 % of the children -- and
// only the SURVIVORS come back.  Steps (one process per GPU):
//   1. local table                       build_local_table(), tolerance not applied yet
//   2. owner(hash) = mulhi(mix64(hash), world); records partitioned by owner on the device
//   3. counts: ncclAllGather;  records: grouped ncclSend/ncclRecv  (all-to-allv)
//   4. owner: second interference table over the received records, tolerance, global N_u
//   5. truncation: radix select with ncclAllReduce of every digit histogram = the GLOBAL top-k
//      (the reference keeps max_num_object / local_size per rank, quids_mpi.hpp:537,590, which is
//      not the single-node result; north_star asks for the latter)
//   6. survivors go back to the rank of their representative (grouped send/recv), which rebuilds
//      them next to their parents (finalisation of the single-GPU path)
//   7. normalisation with the all-reduced total (quids_mpi.hpp:870-895)
// NCCL is loaded with dlopen the first time a communicator is created: libquids_b200.so itself has
// no NCCL dependency, and inside a PyTorch process the already-loaded libnccl.so.2 is reused.
#pragma once

#include <dlfcn.h>
#include <nccl.h>

struct nccl_api {
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
	bool ok = false;
	std::string why;
};

static nccl_api &nccl() {
	static nccl_api api = [] {
		nccl_api a;
		void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
		if (!h) {
			a.why = std::string("cannot load libnccl.so.2: ") + dlerror();
			return a;
		}
		auto sym = [&](const char *name) {
			void *p = dlsym(h, name);
			if (!p) a.why += std::string(" missing ") + name;
			return p;
		};
		a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
		a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
		a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
		a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
		a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
		a.Send = (decltype(a.Send))sym("ncclSend");
		a.Recv = (decltype(a.Recv))sym("ncclRecv");
		a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
		a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
		a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
		a.ok = a.why.empty();
		return a;
	}();
	return api;
}

#define QB_NCCL(call)                                                                                                              \
	do {                                                                                                                           \
		ncclResult_t qb_nccl_ = (call);                                                                                            \
		if (qb_nccl_ != ncclSuccess)                                                                                               \
			throw ::qb::error(QB_ERR_COMM, std::string(#call) + ": " + nccl().GetErrorString(qb_nccl_) + " (" + __FILE__ + ":" + \
			                                   std::to_string(__LINE__) + ")");                                                   \
	} while (0)

struct route_buffers { // route.inc.cuh: kept in the communicator across calls
	dev_buf family, owner, counts, cursor;
	dev_buf s_size, s_padded, s_mag, s_src, s_begin, s_bytes; // this rank's objects grouped by owner
	dev_buf r_size, r_padded, r_mag, r_begin, r_bytes;        // what arrived: the state this rank owns
};

struct qb_comm {
	qb_ctx *ctx = nullptr;
	int world = 1, rank = 0;
	ncclComm_t nccl = nullptr;
	dev_buf scratch; // small device staging for host-value collectives
	dev_buf send, recv, owner_table, okey, oslot, ret_send, ret_recv, cursors, recv_begin;
	double owner_unique_ratio = 0; // slots created / records received by the owner table of the last call (0 = no call yet)
	route_buffers *route = nullptr; // route.inc.cuh: parents grouped by family owner, and what arrived (created on first use)
};

namespace {

struct comm_ops {
	qb_comm *c;
	bool local_interference = false; // the parents were routed by family (route.inc.cuh): interference is complete on every rank
	qb_ctx *ctx() const { return c->ctx; }
	int world() const { return c->world; }
	int rank() const { return c->rank; }

	void allreduce_u64_device(void *ptr, size_t n, bool max = false) {
		QB_NCCL(nccl().AllReduce(ptr, ptr, n, ncclUint64, max ? ncclMax : ncclSum, c->nccl, c->ctx->stream));
	}
	std::vector<uint64_t> allgather_u64(const uint64_t *values, size_t n) { // n values per rank -> world * n
		c->scratch.ensure(sizeof(uint64_t) * n * (c->world + 1), c->ctx->stream);
		uint64_t *mine = c->scratch.as<uint64_t>(), *all = mine + n;
		QB_CUDA(cudaMemcpyAsync(mine, values, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, c->ctx->stream));
		QB_NCCL(nccl().AllGather(mine, all, n, ncclUint64, c->nccl, c->ctx->stream));
		std::vector<uint64_t> out(n * c->world);
		QB_CUDA(cudaMemcpyAsync(out.data(), all, sizeof(uint64_t) * out.size(), cudaMemcpyDeviceToHost, c->ctx->stream));
		c->ctx->sync();
		return out;
	}
	// ---- agreement on errors.  A failure that only ONE rank sees (a table that overflows on its share of the data, a failed
	// allocation) must not leave the other ranks waiting in the next collective for ever: every phase of the distributed
	// iteration runs under a pending_error, and the status word travels with the next count exchange the protocol does
	// anyway (no extra synchronisation); if any rank reports a failure, ALL ranks throw, with the failing rank's status.
	struct pending_error {
		int status = QB_OK;
		std::string what;
		bool ok() const { return status == QB_OK; }
		template <class F>
		void run(F &&f) { // runs f unless an earlier phase already failed; remembers the first failure
			if (!ok())
				return;
			try {
				f();
			} catch (const qb::error &e) {
				status = e.status;
				what = e.what();
			} catch (const std::exception &e) {
				status = QB_ERR_ARG;
				what = e.what();
			}
		}
	};
	// all-gather of n values per rank with the status word appended; throws on every rank if any rank failed
	std::vector<uint64_t> allgather_agreed(const uint64_t *values, size_t n, pending_error &err, const char *phase) {
		std::vector<uint64_t> mine(values, values + n);
		mine.push_back((uint64_t)(uint32_t)err.status);
		std::vector<uint64_t> all = allgather_u64(mine.data(), n + 1);
		std::vector<uint64_t> out;
		out.reserve(n * c->world);
		int failed_rank = -1, failed_status = QB_OK;
		for (int r = 0; r < c->world; ++r) {
			const int st = (int)(uint32_t)all[(size_t)r * (n + 1) + n];
			if (st != QB_OK && failed_rank < 0) {
				failed_rank = r;
				failed_status = st;
			}
			out.insert(out.end(), all.begin() + (size_t)r * (n + 1), all.begin() + (size_t)r * (n + 1) + n);
		}
		if (failed_rank >= 0) {
			if (!err.ok())
				throw qb::error(err.status, err.what + " [rank " + std::to_string(c->rank) + ", " + phase + "; every rank of the communicator stops]");
			throw qb::error(failed_status, std::string("quids::mpi::simulate: rank ") + std::to_string(failed_rank) + " failed during " + phase +
			                                   " (status " + std::to_string(failed_status) + "); this rank stops too");
		}
		return out;
	}
	uint64_t sum_u64_agreed(uint64_t v, pending_error &err, const char *phase) {
		uint64_t total = 0;
		for (uint64_t x : allgather_agreed(&v, 1, err, phase))
			total += x;
		return total;
	}
	double sum_f64_agreed(double v, pending_error &err, const char *phase) {
		uint64_t bits;
		memcpy(&bits, &v, 8);
		double total = 0;
		for (uint64_t x : allgather_agreed(&bits, 1, err, phase)) {
			double d;
			memcpy(&d, &x, 8);
			total += d;
		}
		return total;
	}
	uint64_t sum_u64(uint64_t v) {
		uint64_t total = 0;
		for (uint64_t x : allgather_u64(&v, 1))
			total += x;
		return total;
	}
	double sum_f64(double v) { // summed in rank order on every rank: the same bits everywhere
		uint64_t bits;
		memcpy(&bits, &v, 8);
		double total = 0;
		for (uint64_t x : allgather_u64(&bits, 1)) {
			double d;
			memcpy(&d, &x, 8);
			total += d;
		}
		return total;
	}
	// all-to-allv of fixed-size records: send_counts[r] records go to rank r (contiguous, rank order)
	std::vector<uint64_t> alltoallv(const void *send, const std::vector<uint64_t> &send_counts, dev_buf &recv, size_t record_bytes, uint64_t &n_recv,
	                                pending_error *err = nullptr, const char *phase = "") {
		pending_error none;
		std::vector<uint64_t> matrix = allgather_agreed(send_counts.data(), c->world, err ? *err : none, phase); // matrix[src * world + dst]
		std::vector<uint64_t> recv_counts(c->world);
		n_recv = 0;
		for (int src = 0; src < c->world; ++src) {
			recv_counts[src] = matrix[(size_t)src * c->world + c->rank];
			n_recv += recv_counts[src];
		}
		// the receive buffer may have to grow, and that can fail on one rank only: agree once more before anything is posted
		pending_error grow;
		grow.run([&] { recv.ensure(record_bytes * std::max<uint64_t>(1, n_recv), c->ctx->stream); });
		sum_u64_agreed(0, grow, "allocation of a receive buffer");
		QB_NCCL(nccl().GroupStart());
		uint64_t send_off = 0, recv_off = 0;
		const void *self_from = nullptr;
		void *self_to = nullptr;
		for (int r = 0; r < c->world; ++r) {
			if (r == c->rank) { // what stays on this GPU is a plain device copy, not a send to oneself
				self_from = (const char *)send + send_off * record_bytes;
				self_to = recv.as<char>() + recv_off * record_bytes;
			} else {
				if (send_counts[r])
					QB_NCCL(nccl().Send((const char *)send + send_off * record_bytes, send_counts[r] * record_bytes, ncclUint8, r, c->nccl, c->ctx->stream));
				if (recv_counts[r])
					QB_NCCL(nccl().Recv(recv.as<char>() + recv_off * record_bytes, recv_counts[r] * record_bytes, ncclUint8, r, c->nccl, c->ctx->stream));
			}
			send_off += send_counts[r];
			recv_off += recv_counts[r];
		}
		QB_NCCL(nccl().GroupEnd());
		if (send_counts[c->rank])
			QB_CUDA(cudaMemcpyAsync(self_to, self_from, send_counts[c->rank] * record_bytes, cudaMemcpyDeviceToDevice, c->ctx->stream));
		return recv_counts;
	}
};

// ---- records that cross NVLink ---------------------------------------------------------------------------
struct __align__(32) exchange_record { // a locally unique child, on its way to the owner of its hash
	unsigned long long hash;
	double re, im;
	unsigned long long rep; // representative in the symbolic order of the SENDING rank
};

__device__ __forceinline__ uint32_t owner_of(uint64_t hash, uint32_t world) { return (uint32_t)__umul64hi(mix64(hash ^ 0x9e3779b97f4a7c15ull), (uint64_t)world); }

// Inside the segment a rank sends to one owner, the records are grouped by the TOP BITS of the mix that also places them in
// the owner's table (table_home = mulhi(mix64(hash), capacity) grows with mix64(hash)): the owner then walks every
// source's segment as one sweep over its table, and the slots it touches at any moment are a few tens of MB -- L2
// resident (21 ps per insert) instead of anywhere in a GB of DRAM (100 ps, scripts/table_bench.cu).
// number of regions (a power of two, at most 256) such that world * regions <= 2048 bins: 32 KB of shared memory in the
// scatter kernel
inline uint32_t owner_sub_buckets(uint32_t world) {
	uint32_t sub = 256;
	while (sub > 1 && world * sub > 2048)
		sub >>= 1;
	return sub;
}
__device__ __forceinline__ uint32_t owner_bin(uint64_t hash, uint32_t world, uint32_t sub) {
	const uint32_t region = sub > 1 ? (uint32_t)__umul64hi(mix64(hash), (uint64_t)sub) : 0u; // the top log2(sub) bits
	return owner_of(hash, world) * sub + region;
}

// The partition kernels below count and rank elements by a key with only `world` (2..8) values: one shared-memory atomic
// per ELEMENT would have 256 threads hammering 2 addresses (1.5 ms for 1.5e7 records, QB_DIST_TRACE).  The lanes of a
// warp with the same key are merged first: one atomic per (warp, key).
struct warp_group {
	unsigned peers;  // lanes of this warp with the same key (0 for a lane without element)
	uint32_t rank;   // this lane's rank among them
	uint32_t size;
	bool leader;
};
__device__ __forceinline__ warp_group warp_group_by(bool valid, uint32_t key) {
	warp_group g{0, 0, 0, false};
	const unsigned active = __ballot_sync(0xffffffffu, valid);
	if (valid) {
		g.peers = __match_any_sync(active, key);
		g.rank = __popc(g.peers & ((1u << lane_id()) - 1));
		g.size = __popc(g.peers);
		g.leader = g.rank == 0;
	}
	return g;
}

// counts[bin] over the locally unique children, bin = owner * sub + sub-bucket (shared-memory histogram per CTA: a global
// atomic per element would serialise in L2)
__global__ void __launch_bounds__(256) owner_count_kernel(table_view t, const uint32_t *uslot, uint64_t n, uint32_t world, uint32_t sub, unsigned long long *counts) {
	extern __shared__ unsigned int s_counts[];
	const uint32_t bins = world * sub;
	for (uint32_t i = threadIdx.x; i < bins; i += blockDim.x)
		s_counts[i] = 0;
	__syncthreads();
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < n; base += stride) {
		const uint64_t i = base + threadIdx.x;
		const bool valid = i < n;
		const uint32_t b = valid ? owner_bin(t.slots[uslot[i]].key, world, sub) : 0;
		const warp_group g = warp_group_by(valid, b);
		if (g.leader)
			atomicAdd(&s_counts[b], g.size);
	}
	__syncthreads();
	for (uint32_t i = threadIdx.x; i < bins; i += blockDim.x)
		if (s_counts[i])
			atomicAdd(&counts[i], (unsigned long long)s_counts[i]);
}

constexpr int SCATTER_TILE = 2048; // elements one CTA places per round: at most one global atomic per bin and tile

// records grouped by bin: cursor[bin] starts at the bin's offset (owner-major, so every owner's records are contiguous).
// Per tile: histogram in shared memory, one global atomicAdd per bin reserves the tile's range, ranks inside the tile
// come from shared-memory atomics (one per warp and bin).
__global__ void __launch_bounds__(256) owner_scatter_kernel(table_view t, const uint32_t *uslot, uint64_t n, uint32_t world, uint32_t sub, unsigned long long *cursor,
                                                            exchange_record *out) {
	extern __shared__ unsigned long long s_base[]; // [bins] tile base (u64), then [bins] counts and [bins] running ranks (u32)
	const uint32_t bins = world * sub;
	unsigned int *s_count = reinterpret_cast<unsigned int *>(s_base + bins), *s_rank = s_count + bins;
	for (uint64_t tile = (uint64_t)blockIdx.x * SCATTER_TILE; tile < n; tile += (uint64_t)gridDim.x * SCATTER_TILE) {
		for (uint32_t i = threadIdx.x; i < bins; i += blockDim.x)
			s_count[i] = s_rank[i] = 0;
		__syncthreads();
		const uint64_t end = min(tile + (uint64_t)SCATTER_TILE, n);
		for (uint64_t base = tile; base < end; base += blockDim.x) {
			const uint64_t i = base + threadIdx.x;
			const bool valid = i < end;
			const uint32_t b = valid ? owner_bin(t.slots[uslot[i]].key, world, sub) : 0;
			const warp_group g = warp_group_by(valid, b);
			if (g.leader)
				atomicAdd(&s_count[b], g.size);
		}
		__syncthreads();
		for (uint32_t b = threadIdx.x; b < bins; b += blockDim.x)
			s_base[b] = s_count[b] ? atomicAdd(&cursor[b], (unsigned long long)s_count[b]) : 0;
		__syncthreads();
		for (uint64_t base = tile; base < end; base += blockDim.x) {
			const uint64_t i = base + threadIdx.x;
			const bool valid = i < end;
			table_slot s{};
			if (valid)
				s = t.slots[uslot[i]];
			const uint32_t b = valid ? owner_bin(s.key, world, sub) : 0;
			const warp_group g = warp_group_by(valid, b);
			unsigned int first = 0;
			if (g.leader)
				first = atomicAdd(&s_rank[b], g.size);
			if (valid) {
				first = __shfl_sync(g.peers, first, __ffs(g.peers) - 1);
				out[s_base[b] + first + g.rank] = exchange_record{s.key, s.re, s.im, s.rep};
			}
		}
		__syncthreads();
	}
}

// owner side: the representative of a slot is the POSITION of one of the records merged into it (its
// sender and its original representative are looked up there later).  Which one: the record with the
// largest pseudo-random byte (atomicMax), NOT the one that created the slot -- the receive buffer is in
// rank order, so "first come" would hand most survivors, hence most of the finalisation, to rank 0.
__device__ __forceinline__ uint64_t owner_rep_pack(uint64_t position, uint64_t hash) { return ((mix64(position ^ hash) >> 56) << 56) | (position + 1); }
__device__ __forceinline__ uint64_t owner_rep_position(uint64_t rep) { return (rep & ((1ull << 56) - 1)) - 1; }

// Two records per thread and round: their key loads go out before either is resolved.  Not more: the records of a source
// arrive grouped by table region (owner_bin), and the narrower the window of records in flight, the smaller the part of
// the table it touches -- it has to stay L2 resident.
__global__ void __launch_bounds__(256) record_insert_kernel(table_view t, const exchange_record *records, uint64_t n) {
	constexpr int N = 2;
	const uint64_t threads = (uint64_t)gridDim.x * blockDim.x;
	unsigned int created = 0;
	for (uint64_t first = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; first < n; first += threads * N) {
		exchange_record r[N];
		uint64_t at[N];
		unsigned pending = 0;
#pragma unroll
		for (int q = 0; q < N; ++q) {
			const uint64_t i = first + (uint64_t)q * threads;
			if (i < n) {
				r[q] = records[i];
				if (r[q].hash == 0) { // the dedicated slot
					table_slot *s = t.slots + t.capacity;
					atomicAdd(&s->re, r[q].re);
					atomicAdd(&s->im, r[q].im);
					atomicMax(&s->rep, (unsigned long long)owner_rep_pack(i, 0));
				} else {
					at[q] = table_home(r[q].hash, t.capacity);
					pending |= 1u << q;
				}
			}
		}
		for (uint32_t round = 0; pending; ++round) {
			unsigned long long seen[N];
#pragma unroll
			for (int q = 0; q < N; ++q)
				if (pending & (1u << q))
					seen[q] = __ldcg(&t.slots[at[q]].key);
#pragma unroll
			for (int q = 0; q < N; ++q)
				if (pending & (1u << q)) {
					table_slot *s = t.slots + at[q];
					if (seen[q] == 0) {
						seen[q] = atomicCAS(&s->key, 0ull, (unsigned long long)r[q].hash);
						if (seen[q] == 0) {
							++created;
							seen[q] = r[q].hash;
						}
					}
					if (seen[q] == r[q].hash) {
						atomicAdd(&s->re, r[q].re);
						atomicAdd(&s->im, r[q].im);
						atomicMax(&s->rep, (unsigned long long)owner_rep_pack(first + (uint64_t)q * threads, r[q].hash));
						pending &= ~(1u << q);
					} else if (++at[q] == t.capacity) {
						at[q] = 0;
					}
				}
			if (round > TABLE_MAX_PROBES) {
				*t.overflow = 1;
				break;
			}
		}
	}
	created = (unsigned int)warp_sum((uint64_t)created);
	if (lane_id() == 0 && created)
		atomicAdd(t.used, (unsigned long long)created);
}

// owner side, binned (table.cuh): every received record goes to the bin of its hash with the POSITION it arrived at as its
// representative (plus the pseudo-random byte of owner_rep_pack: bin_dedup_kernel keeps the largest, a fair choice between
// the ranks); one streaming pass, no table
__global__ void __launch_bounds__(256) record_bin_kernel(bin_view b, table_view t, const exchange_record *records, uint64_t n) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const ulonglong2 lo = __ldcs(reinterpret_cast<const ulonglong2 *>(records + i));
		const ulonglong2 hi = __ldcs(reinterpret_cast<const ulonglong2 *>(records + i) + 1);
		const uint64_t hash = lo.x;
		const cplx mag{__longlong_as_double((long long)lo.y), __longlong_as_double((long long)hi.x)};
		if (hash == 0) { // the dedicated slot
			table_slot *s = t.slots + t.capacity;
			atomicAdd(&s->re, mag.re);
			atomicAdd(&s->im, mag.im);
			atomicMax(&s->rep, (unsigned long long)owner_rep_pack(i, 0));
		} else {
			bin_put(b, t.overflow, hash, mag, owner_rep_pack(i, hash));
		}
	}
}

// survivors (owner slots) -> which rank their representative came from; counts per rank
__global__ void __launch_bounds__(256) return_count_kernel(table_view t, const uint32_t *slot, uint64_t n, const uint64_t *recv_begin, uint32_t world,
                                                           unsigned long long *counts) {
	extern __shared__ unsigned int s_counts[];
	for (uint32_t i = threadIdx.x; i < world; i += blockDim.x)
		s_counts[i] = 0;
	__syncthreads();
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < n; base += stride) {
		const uint64_t i = base + threadIdx.x;
		const bool valid = i < n;
		const uint32_t src = valid ? (uint32_t)upper_bound_u64(recv_begin, world + 1, owner_rep_position(t.slots[slot[i]].rep)) - 1 : 0;
		const warp_group g = warp_group_by(valid, src);
		if (g.leader)
			atomicAdd(&s_counts[src], g.size);
	}
	__syncthreads();
	for (uint32_t i = threadIdx.x; i < world; i += blockDim.x)
		if (s_counts[i])
			atomicAdd(&counts[i], (unsigned long long)s_counts[i]);
}

__global__ void __launch_bounds__(256) return_scatter_kernel(table_view t, const uint32_t *slot, uint64_t n, const uint64_t *recv_begin, uint32_t world,
                                                             const exchange_record *received, unsigned long long *cursor, survivor_record *out) {
	extern __shared__ unsigned long long s_base[];
	unsigned int *s_count = reinterpret_cast<unsigned int *>(s_base + world), *s_rank = s_count + world;
	for (uint64_t tile = (uint64_t)blockIdx.x * SCATTER_TILE; tile < n; tile += (uint64_t)gridDim.x * SCATTER_TILE) {
		for (uint32_t i = threadIdx.x; i < world; i += blockDim.x)
			s_count[i] = s_rank[i] = 0;
		__syncthreads();
		const uint64_t end = min(tile + (uint64_t)SCATTER_TILE, n);
		for (uint64_t base = tile; base < end; base += blockDim.x) {
			const uint64_t i = base + threadIdx.x;
			const bool valid = i < end;
			const uint32_t src = valid ? (uint32_t)upper_bound_u64(recv_begin, world + 1, owner_rep_position(t.slots[slot[i]].rep)) - 1 : 0;
			const warp_group g = warp_group_by(valid, src);
			if (g.leader)
				atomicAdd(&s_count[src], g.size);
		}
		__syncthreads();
		for (uint32_t o = threadIdx.x; o < world; o += blockDim.x)
			s_base[o] = s_count[o] ? atomicAdd(&cursor[o], (unsigned long long)s_count[o]) : 0;
		__syncthreads();
		for (uint64_t base = tile; base < end; base += blockDim.x) {
			const uint64_t i = base + threadIdx.x;
			const bool valid = i < end;
			table_slot s{};
			if (valid)
				s = t.slots[slot[i]];
			const uint64_t position = valid ? owner_rep_position(s.rep) : 0;
			const uint32_t src = valid ? (uint32_t)upper_bound_u64(recv_begin, world + 1, position) - 1 : 0;
			const warp_group g = warp_group_by(valid, src);
			unsigned int first = 0;
			if (g.leader)
				first = atomicAdd(&s_rank[src], g.size);
			if (valid) {
				first = __shfl_sync(g.peers, first, __ffs(g.peers) - 1);
				out[s_base[src] + first + g.rank] = survivor_record{received[position].rep, s.re, s.im};
			}
		}
		__syncthreads();
	}
}

} // namespace
