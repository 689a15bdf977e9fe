// migrate.inc.cuh -- object migration between the GPUs of a communicator: the NCCL counterpart of
// mpi_iteration::send_objects / receive_objects (quids_mpi.hpp:124-231), equalize / equalize_symbolic
// (:903-1026), distribute_objects (:1031-1051) and gather_objects (:1056-1077).
// Textually included by capi.cu.
//
// A migration moves the TAIL of a state (the reference pops from the tail too): the four arrays of the
// last n objects travel HBM -> NVLink -> HBM in one NCCL group, no host staging.  object_begin travels
// as stored and is rebased on arrival, so the layout (alignment padding included) is preserved whatever
// align_byte_length the two sides use.  The handshake of the reference is kept: the sender announces
// (count, bytes), the receiver answers whether it has room (quids_mpi.hpp:136-138, 196-198).
#pragma once
// (included inside the anonymous namespace of capi.cu)

__global__ void __launch_bounds__(256) rebase_begin_kernel(uint64_t *begin, uint64_t n, uint64_t from, uint64_t to) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
		begin[i] = begin[i] - from + to;
}

// first index i in [0, n] with child_begin[i] >= target (std::lower_bound of quids_mpi.hpp:1017), one thread
__global__ void lower_bound_kernel(const uint64_t *child_begin, uint64_t n, uint64_t target, uint64_t *out) {
	uint64_t lo = 0, hi = n;
	while (lo < hi) {
		const uint64_t mid = (lo + hi) / 2;
		if (child_begin[mid] < target)
			lo = mid + 1;
		else
			hi = mid;
	}
	*out = lo;
}

// a few u64 words between the host and a peer, through the communicator's device scratch
void send_words(qb_comm *c, const uint64_t *words, size_t n, int peer) {
	cudaStream_t s = c->ctx->stream;
	c->scratch.ensure(sizeof(uint64_t) * 16, s);
	QB_CUDA(cudaMemcpyAsync(c->scratch.ptr, words, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, s));
	QB_NCCL(nccl().Send(c->scratch.ptr, n, ncclUint64, peer, c->nccl, s));
	c->ctx->sync();
}
void recv_words(qb_comm *c, uint64_t *words, size_t n, int peer) {
	cudaStream_t s = c->ctx->stream;
	c->scratch.ensure(sizeof(uint64_t) * 16, s);
	QB_NCCL(nccl().Recv(c->scratch.as<uint64_t>() + 8, n, ncclUint64, peer, c->nccl, s));
	QB_CUDA(cudaMemcpyAsync(words, c->scratch.as<uint64_t>() + 8, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, s));
	c->ctx->sync();
}

// send the last n objects of `it` to `peer` and pop them (no normalisation, quids_mpi.hpp:170); returns the number sent
uint64_t send_objects(qb_iter *it, qb_comm *c, uint64_t n, int peer) {
	qb_ctx *ctx = it->ctx;
	cudaStream_t s = ctx->stream;
	QB_REQUIRE(n <= it->n, QB_ERR_ARG, "send_objects: more objects than the state holds");
	QB_REQUIRE(peer >= 0 && peer < c->world && peer != c->rank, QB_ERR_ARG, "send_objects: bad peer rank");
	const uint64_t first = it->n - n;
	uint64_t base = it->n_bytes;
	if (n) {
		QB_CUDA(cudaMemcpyAsync(&ctx->h_small[DS_COUNT], it->begin.as<uint64_t>() + first, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
		ctx->sync();
		base = ctx->h_small[DS_COUNT];
	}
	const uint64_t header[4] = {n, it->n_bytes - base, base, 0};
	send_words(c, header, 4, peer);
	if (n == 0) // quids_mpi.hpp:129-130
		return 0;
	uint64_t accept = 0;
	recv_words(c, &accept, 1, peer);
	if (!accept)
		return 0;
	QB_NCCL(nccl().GroupStart());
	QB_NCCL(nccl().Send(it->mag.as<cplx>() + first, n * sizeof(cplx), ncclUint8, peer, c->nccl, s));
	QB_NCCL(nccl().Send(it->begin.as<uint64_t>() + first + 1, n, ncclUint64, peer, c->nccl, s));
	QB_NCCL(nccl().Send(it->size.as<uint32_t>() + first, n, ncclUint32, peer, c->nccl, s));
	if (header[1])
		QB_NCCL(nccl().Send(it->objects.as<uint8_t>() + base, header[1], ncclUint8, peer, c->nccl, s));
	QB_NCCL(nccl().GroupEnd());
	ctx->sync();
	it->n = first;
	it->n_bytes = base;
	return n;
}

// receive objects from `peer` at the tail of `it`; max_bytes bounds what may arrive (52 B of metadata per object +
// the object bytes, the reference's ITERATION_MEMORY_SIZE accounting, quids_mpi.hpp:196), ~0 = whatever fits the GPU
uint64_t receive_objects(qb_iter *it, qb_comm *c, int peer, uint64_t max_bytes) {
	qb_ctx *ctx = it->ctx;
	cudaStream_t s = ctx->stream;
	QB_REQUIRE(peer >= 0 && peer < c->world && peer != c->rank, QB_ERR_ARG, "receive_objects: bad peer rank");
	uint64_t header[4] = {0, 0, 0, 0};
	recv_words(c, header, 4, peer);
	const uint64_t n = header[0], bytes = header[1], base = header[2];
	if (n == 0)
		return 0;
	const uint64_t need = n * 52 + bytes;
	size_t free_bytes = 0, total_bytes = 0;
	QB_CUDA(cudaMemGetInfo(&free_bytes, &total_bytes));
	// growing a buffer keeps the old copy alive until the new one is filled: count the state twice
	const uint64_t grow = need + it->n * 52 + it->n_bytes + (1u << 20);
	uint64_t accept = (need < max_bytes && grow < free_bytes) ? 1 : 0;
	send_words(c, &accept, 1, peer);
	if (!accept)
		return 0;
	it->objects.ensure(it->n_bytes + bytes + 16, s, true, it->n_bytes);
	it->begin.ensure(sizeof(uint64_t) * (it->n + n + 1), s, true, sizeof(uint64_t) * (it->n + 1));
	it->size.ensure(sizeof(uint32_t) * (it->n + n), s, true, sizeof(uint32_t) * it->n);
	it->mag.ensure(sizeof(cplx) * (it->n + n), s, true, sizeof(cplx) * it->n);
	QB_NCCL(nccl().GroupStart());
	QB_NCCL(nccl().Recv(it->mag.as<cplx>() + it->n, n * sizeof(cplx), ncclUint8, peer, c->nccl, s));
	QB_NCCL(nccl().Recv(it->begin.as<uint64_t>() + it->n + 1, n, ncclUint64, peer, c->nccl, s));
	QB_NCCL(nccl().Recv(it->size.as<uint32_t>() + it->n, n, ncclUint32, peer, c->nccl, s));
	if (bytes)
		QB_NCCL(nccl().Recv(it->objects.as<uint8_t>() + it->n_bytes, bytes, ncclUint8, peer, c->nccl, s));
	QB_NCCL(nccl().GroupEnd());
	if (base != it->n_bytes) {
		rebase_begin_kernel<<<grid_for(n, 256, ctx->grid_cap()), 256, 0, s>>>(it->begin.as<uint64_t>() + it->n + 1, n, base, it->n_bytes);
		++ctx->launches;
	}
	ctx->sync();
	QB_CUDA(cudaGetLastError());
	it->n += n;
	it->n_bytes += bytes;
	return n;
}

// utils::make_equal_pairs (utils/mpi_utils.hpp:9-24): the i-th heaviest rank is paired with the i-th lightest.
// Ties are broken by rank so that every rank computes the same pairing without a scatter.
std::vector<int> make_equal_pairs(const std::vector<uint64_t> &weight) {
	const int size = (int)weight.size();
	std::vector<int> ids(size), pair(size);
	for (int i = 0; i < size; ++i)
		ids[i] = i;
	std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) { return weight[a] > weight[b]; });
	for (int i = 0; i < size; ++i)
		pair[ids[i]] = ids[size - 1 - i];
	return pair;
}

// number of children of every local object under `rule` + their inclusive scan (compute_num_child, quids.hpp:548-569)
uint64_t count_children(qb_iter *it, const rule_ops *ops, const void *rule) {
	qb_ctx *ctx = it->ctx;
	cudaStream_t stream = ctx->stream;
	it->n_symbolic = 0;
	if (it->n == 0)
		return 0;
	engine_launch L;
	memset(&L, 0, sizeof L);
	L.stream = stream;
	L.sm_count = ctx->sm_count;
	L.launch_counter = &ctx->launches;
	L.it = it->view();
	QB_CUDA(cudaMemsetAsync(ctx->d_small.ptr, 0, DS_WORDS * sizeof(uint64_t), stream));
	it->num_childs.ensure(sizeof(uint32_t) * it->n, stream);
	L.num_childs = it->num_childs.as<uint32_t>();
	if (ops->has_groups) {
		it->num_groups.ensure(sizeof(uint32_t) * it->n, stream);
		L.num_groups = it->num_groups.as<uint32_t>();
	}
	L.max_child_size = reinterpret_cast<unsigned int *>(ctx->small(DS_MAX_CHILD_SIZE));
	L.child_count_range = reinterpret_cast<unsigned int *>(ctx->small(DS_CHILD_RANGE));
	ops->launch_num_child(rule, L);
	it->child_begin.ensure(sizeof(uint64_t) * (it->n + 1), stream);
	exclusive_scan(ctx, counts_through{it->num_childs.as<uint32_t>(), nullptr}, it->child_begin.as<uint64_t>(), it->n);
	QB_CUDA(cudaMemcpyAsync(&ctx->h_small[DS_COUNT], it->child_begin.as<uint64_t>() + it->n, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
	ctx->sync();
	it->n_symbolic = ctx->h_small[DS_COUNT];
	return it->n_symbolic;
}

struct balance_stats {
	uint64_t max_objects = 0, max_weight = 0;
	double avg_weight = 0;
	std::vector<uint64_t> weight;
};

// ONE pairing round (mpi_iteration::equalize, quids_mpi.hpp:903-960, when ops == nullptr; equalize_symbolic,
// :965-1026, otherwise): the heavier rank of every pair sends half of the difference from its tail.
// Returns the number of objects this rank sent (> 0) or received (< 0).
int64_t equalize_round(qb_iter *it, qb_comm *c, const rule_ops *ops, const void *rule, const balance_stats &st) {
	const std::vector<int> pair = make_equal_pairs(st.weight);
	const int other = pair[c->rank];
	if (other == c->rank) // quids_mpi.hpp:941-942
		return 0;
	const uint64_t mine = st.weight[c->rank], theirs = st.weight[other];
	if (mine > theirs) {
		uint64_t n_send;
		if (!ops) {
			n_send = (mine - theirs) / 2; // quids_mpi.hpp:954-955
		} else { // quids_mpi.hpp:1013-1019: the objects holding the last (mine - theirs) / 2 children
			qb_ctx *ctx = it->ctx;
			const uint64_t target = mine - (mine - theirs) / 2;
			lower_bound_kernel<<<1, 1, 0, ctx->stream>>>(it->child_begin.as<uint64_t>(), it->n, target, ctx->small(DS_USED));
			++ctx->launches;
			ctx->fetch_small();
			const uint64_t limit = ctx->h_small[DS_USED] > 0 ? ctx->h_small[DS_USED] - 1 : 0;
			n_send = it->n - limit;
		}
		return (int64_t)send_objects(it, c, n_send, other);
	}
	if (mine < theirs)
		return -(int64_t)receive_objects(it, c, other, ~0ull);
	return 0;
}

balance_stats gather_balance(qb_iter *it, qb_comm *c, uint64_t weight) {
	comm_ops comm{c};
	const uint64_t mine[2] = {weight, it->n};
	const std::vector<uint64_t> all = comm.allgather_u64(mine, 2);
	balance_stats st;
	st.weight.resize(c->world);
	double sum = 0;
	for (int r = 0; r < c->world; ++r) {
		st.weight[r] = all[2 * r];
		st.max_weight = std::max(st.max_weight, all[2 * r]);
		st.max_objects = std::max(st.max_objects, all[2 * r + 1]);
		sum += (double)all[2 * r];
	}
	st.avg_weight = sum / c->world;
	return st;
}

// the load-balancing loop at the head of quids::mpi::simulate (quids_mpi.hpp:442-500): up to ceil(log2(world))
// pairing rounds, stopped when the state is small, balanced within `inbalance`, or no longer improving.
// by_children: weigh the ranks by the children `rule` will produce (equalize_children, the reference's default).
int equalize_loop(qb_iter *it, qb_comm *c, const rule_ops *ops, const void *rule, bool by_children, uint64_t min_size, float inbalance_limit, float min_step,
                  int max_rounds) {
	int rounds = 0;
	float previous_diff = 0;
	double avg0 = -1;
	for (int i = 0; i < max_rounds; ++i) {
		const uint64_t weight = by_children ? count_children(it, ops, rule) : it->n;
		const balance_stats st = gather_balance(it, c, weight);
		if (avg0 < 0) avg0 = st.avg_weight; // the reference computes the average once, before the loop (:446, :475)
		const float diff = (float)st.max_weight - (float)avg0;
		const float inbalance = st.max_weight ? diff / (float)st.max_weight : 0.f;
		if (st.max_objects < min_size || inbalance < inbalance_limit || (i > 0 && diff > previous_diff * (1 - min_step)))
			break;
		equalize_round(it, c, by_children ? ops : nullptr, rule, st);
		previous_diff = diff;
		++rounds;
	}
	return rounds;
}

