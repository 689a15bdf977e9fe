// route.inc.cuh -- distributed interference WITHOUT an exchange of children: parents move to the rank that owns their FAMILY.
// Textually included by capi.cu (inside its anonymous namespace, after dist.inc.cuh and migrate.inc.cuh).
//
// The reference ships every child's (hash, magnitude) to the rank that owns the hash and a magnitude back per child
// (quids_mpi.hpp:741-743, 842); dist.inc.cuh ships the locally unique children instead.  On the state the QCGD loop settles
// in, erase_create still leaves 0.5 unique children per child: 1.2e9 records of 32 bytes per GPU, 40 GB over NVLink and a
// second interference table on the owner's side per rule iteration -- communication bound, and out of memory at 1e7 parents
// per GPU.  For a rule with FAMILIES (rule_api.cuh: sets of objects closed under the rule; erase_create and coin keep a
// graph's node count, eligible nodes, other particles and names) interference never crosses a family, so it is enough that
// all members of a family sit on one GPU:
//     owner(parent) = mulhi(mix64(family_key(parent)), world)
// and what crosses NVLink is the PARENTS -- a few hundred bytes each, once -- as an all-to-allv of the four arrays of the state
// (sizes, magnitudes, padded sizes, object bytes: "hash-ownership redistribution ... of object bytes over NVLink", north_star).
// Afterwards every rank runs the single-GPU iteration on what it owns; only the scalars stay collective: child counts, the
// digit histograms of the global top-k (select_threshold), the norm.  The state a rank holds after the call is the set of
// objects it owns, not the one it passed in -- like the reference's own load balancer (quids_mpi.hpp:442-500), which also
// moves parents between ranks inside mpi::simulate.
#pragma once

// owner of every object (from its family key), and per owner: number of objects, padded bytes; counts[2 * world] = largest object
__global__ void __launch_bounds__(256) route_count_kernel(const uint64_t *family, const uint64_t *begin, const uint32_t *size, uint64_t n, uint32_t world, uint32_t *owner,
                                                          unsigned long long *counts) {
	extern __shared__ unsigned long long s_route[]; // [world] objects, [world] bytes, [1] largest size
	for (uint32_t i = threadIdx.x; i < 2 * world + 1; i += blockDim.x)
		s_route[i] = 0;
	__syncthreads();
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	unsigned long long largest = 0;
	for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < n; base += stride) {
		const uint64_t i = base + threadIdx.x;
		const bool valid = i < n;
		uint32_t o = 0;
		unsigned long long bytes = 0;
		if (valid) {
			o = (uint32_t)__umul64hi(mix64(family[i]), (uint64_t)world);
			owner[i] = o;
			bytes = begin[i + 1] - begin[i];
			largest = max(largest, (unsigned long long)size[i]);
		}
		const warp_group g = warp_group_by(valid, o);
		if (valid) { // one shared-memory atomic per (warp, owner) for the objects; the bytes differ per object
			if (g.leader)
				atomicAdd(&s_route[o], (unsigned long long)g.size);
			atomicAdd(&s_route[world + o], bytes);
		}
	}
	atomicMax(&s_route[2 * world], largest);
	__syncthreads();
	for (uint32_t i = threadIdx.x; i < 2 * world; i += blockDim.x)
		if (s_route[i])
			atomicAdd(&counts[i], s_route[i]);
	if (threadIdx.x == 0)
		atomicMax(&counts[2 * world], s_route[2 * world]);
}

// every object gets a position in its owner's segment of the send arrays (cursor[owner] starts at the segment's first
// position); per tile: counts in shared memory, one global atomic per (tile, owner), ranks from shared-memory atomics
__global__ void __launch_bounds__(256) route_slot_kernel(const uint32_t *owner, const uint64_t *begin, const uint32_t *size, const cplx *mag, uint64_t n, uint32_t world,
                                                         unsigned long long *cursor, uint32_t *s_size, uint32_t *s_padded, cplx *s_mag, uint64_t *s_src) {
	extern __shared__ unsigned long long s_base[]; // [world] tile base, then [world] counts and [world] running ranks (u32)
	unsigned int *s_count = reinterpret_cast<unsigned int *>(s_base + world), *s_rank = s_count + world;
	for (uint64_t tile = (uint64_t)blockIdx.x * SCATTER_TILE; tile < n; tile += (uint64_t)gridDim.x * SCATTER_TILE) {
		for (uint32_t i = threadIdx.x; i < world; i += blockDim.x)
			s_count[i] = s_rank[i] = 0;
		__syncthreads();
		const uint64_t end = min(tile + (uint64_t)SCATTER_TILE, n);
		for (uint64_t base = tile; base < end; base += blockDim.x) {
			const uint64_t i = base + threadIdx.x;
			const bool valid = i < end;
			const uint32_t o = valid ? owner[i] : 0;
			const warp_group g = warp_group_by(valid, o);
			if (g.leader)
				atomicAdd(&s_count[o], g.size);
		}
		__syncthreads();
		for (uint32_t o = threadIdx.x; o < world; o += blockDim.x)
			s_base[o] = s_count[o] ? atomicAdd(&cursor[o], (unsigned long long)s_count[o]) : 0;
		__syncthreads();
		for (uint64_t base = tile; base < end; base += blockDim.x) {
			const uint64_t i = base + threadIdx.x;
			const bool valid = i < end;
			const uint32_t o = valid ? owner[i] : 0;
			const warp_group g = warp_group_by(valid, o);
			unsigned int first = 0;
			if (g.leader)
				first = atomicAdd(&s_rank[o], g.size);
			if (valid) {
				first = __shfl_sync(g.peers, first, __ffs(g.peers) - 1);
				const uint64_t at = s_base[o] + first + g.rank;
				s_size[at] = size[i];
				s_padded[at] = (uint32_t)(begin[i + 1] - begin[i]);
				s_mag[at] = mag[i];
				s_src[at] = i;
			}
		}
		__syncthreads();
	}
}

// object bytes (padding included) into the send buffer: one warp per object, 8-byte words when both sides allow
__global__ void __launch_bounds__(256) route_copy_kernel(const uint8_t *objects, const uint64_t *begin, const uint64_t *s_src, const uint64_t *s_begin, uint64_t n,
                                                         uint8_t *out) {
	const unsigned lane = lane_id();
	const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	for (uint64_t at = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; at < n; at += warps) {
		const uint64_t src = s_src[at];
		const uint8_t *from = objects + begin[src];
		uint8_t *to = out + s_begin[at];
		const uint32_t bytes = (uint32_t)(s_begin[at + 1] - s_begin[at]);
		const uint32_t have = min(bytes, (uint32_t)(begin[src + 1] - begin[src])); // the last object of a segment may carry 8 more bytes of padding (below)
		if (((reinterpret_cast<uintptr_t>(from) | reinterpret_cast<uintptr_t>(to) | bytes | have) & 7) == 0) {
			for (uint32_t w = lane; w < bytes / 8; w += 32)
				reinterpret_cast<uint2 *>(to)[w] = w < have / 8 ? reinterpret_cast<const uint2 *>(from)[w] : make_uint2(0, 0);
		} else {
			for (uint32_t b = lane; b < bytes; b += 32)
				to[b] = b < have ? from[b] : (uint8_t)0;
		}
	}
}

// NCCL moves a buffer 16 bytes at a time only when it starts 16-byte aligned (otherwise 8 or 4: the routing ran at 240 GB/s per
// GPU instead of 560).  The segment a rank sends to one owner is a sum of padded object sizes, multiples of 8 with the usual
// align_byte_length = 8: a segment of 16 k + 8 bytes gets 8 more bytes of (zero) padding after its LAST object, so that every
// segment starts 16-byte aligned on both sides.  The object keeps them in the state that arrives (object_begin counts them).
struct route_bumps {
	uint64_t last[MAX_DEVICES]; // position of the segment's last object in the send arrays, ~0 = nothing to add
};
__global__ void route_bump_kernel(uint32_t *s_padded, route_bumps bumps, uint32_t world) {
	if (threadIdx.x < world && bumps.last[threadIdx.x] != ~0ull)
		s_padded[bumps.last[threadIdx.x]] += 8;
}
inline uint64_t routed_segment_bytes(uint64_t bytes, uint64_t objects) { return objects > 0 && bytes % 16 == 8 ? bytes + 8 : bytes; }

// Moves every object of `it` to the rank that owns its family.  Returns false (nothing moved, on EVERY rank) when some rank
// holds an object the rule has no family for (region_size_limit).  Collective; failures are agreed on (pending_error).
bool route_by_family(qb_iter *it, qb_comm *cm, const rule_ops *ops, const void *rule, comm_ops &comm, route_buffers &rb, bool trace) {
	qb_ctx *ctx = it->ctx;
	cudaStream_t stream = ctx->stream;
	const uint32_t world = (uint32_t)cm->world;
	const uint64_t n = it->n;
	comm_ops::pending_error err;
	std::vector<uint64_t> mine(2 * world + 1, 0);
	auto t_last = std::chrono::steady_clock::now();
	auto mark = [&](const char *what) { // QB_DIST_TRACE: wall time of the routing's own phases
		if (!trace)
			return;
		ctx->sync();
		auto now = std::chrono::steady_clock::now();
		fprintf(stderr, "[qb route rank %d]   %-34s %8.3f ms\n", cm->rank, what, std::chrono::duration<double, std::milli>(now - t_last).count());
		t_last = now;
	};
	err.run([&] {
		inject_failure(cm->rank, "route");
		rb.counts.ensure(sizeof(uint64_t) * (2 * world + 1), stream);
		QB_CUDA(cudaMemsetAsync(rb.counts.ptr, 0, sizeof(uint64_t) * (2 * world + 1), stream));
		if (n > 0) {
			rb.family.ensure(sizeof(uint64_t) * n, stream);
			rb.owner.ensure(sizeof(uint32_t) * n, stream);
			engine_launch L;
			memset(&L, 0, sizeof L);
			L.stream = stream;
			L.sm_count = ctx->sm_count;
			L.launch_counter = &ctx->launches;
			L.it = it->view();
			L.hashes = rb.family.as<uint64_t>();
			ops->launch_family(rule, L);
			route_count_kernel<<<grid_for(n, 256, ctx->grid_cap()), 256, sizeof(uint64_t) * (2 * world + 1), stream>>>(
			    rb.family.as<uint64_t>(), it->begin.as<uint64_t>(), it->size.as<uint32_t>(), n, world, rb.owner.as<uint32_t>(), rb.counts.as<unsigned long long>());
			++ctx->launches;
			QB_CUDA(cudaGetLastError());
		}
		QB_CUDA(cudaMemcpyAsync(mine.data(), rb.counts.ptr, sizeof(uint64_t) * (2 * world + 1), cudaMemcpyDeviceToHost, stream));
		ctx->sync();
	});
	mark("family keys + counts");
	// matrix[src][0..world) objects for each owner, [world..2 world) bytes, [2 world] largest object of src
	const uint32_t row = 2 * world + 1;
	std::vector<uint64_t> matrix = comm.allgather_agreed(mine.data(), row, err, "the family count of the parents");
	mark("count exchange");
	uint64_t largest = 0;
	for (uint32_t r = 0; r < world; ++r)
		largest = std::max(largest, matrix[(size_t)r * row + 2 * world]);
	if (ops->region_size_limit == 0 || largest >= ops->region_size_limit)
		return false; // some object has no family: every rank takes the record-exchange path instead

	std::vector<uint64_t> send_obj(world), send_bytes(world), recv_obj(world), recv_bytes(world), seg(world + 1, 0);
	uint64_t n_recv = 0, bytes_recv = 0, bytes_send = 0;
	route_bumps bumps;
	QB_REQUIRE(world <= (uint32_t)MAX_DEVICES, QB_ERR_ARG, "family routing: more ranks than MAX_DEVICES");
	for (uint32_t r = 0; r < world; ++r) {
		send_obj[r] = mine[r];
		send_bytes[r] = routed_segment_bytes(mine[world + r], mine[r]);
		bumps.last[r] = send_bytes[r] != mine[world + r] ? seg[r] + send_obj[r] - 1 : ~0ull;
		recv_obj[r] = matrix[(size_t)r * row + cm->rank];
		recv_bytes[r] = routed_segment_bytes(matrix[(size_t)r * row + world + cm->rank], recv_obj[r]);
		seg[r + 1] = seg[r] + send_obj[r];
		n_recv += recv_obj[r];
		bytes_recv += recv_bytes[r];
		bytes_send += send_bytes[r];
	}
	if (trace)
		fprintf(stderr, "[qb route rank %d] %llu objects (%.1f MB) -> owns %llu objects (%.1f MB); stays here: %llu\n", cm->rank, (unsigned long long)n, bytes_send / 1e6,
		        (unsigned long long)n_recv, bytes_recv / 1e6, (unsigned long long)send_obj[cm->rank]);
	err.run([&] {
		// this rank's objects grouped by owner
		rb.s_size.ensure(sizeof(uint32_t) * std::max<uint64_t>(1, n), stream);
		rb.s_padded.ensure(sizeof(uint32_t) * std::max<uint64_t>(1, n), stream);
		rb.s_mag.ensure(sizeof(cplx) * std::max<uint64_t>(1, n), stream);
		rb.s_src.ensure(sizeof(uint64_t) * std::max<uint64_t>(1, n), stream);
		rb.s_begin.ensure(sizeof(uint64_t) * (n + 1), stream);
		rb.s_bytes.ensure(bytes_send + 16, stream);
		rb.cursor.ensure(sizeof(uint64_t) * world, stream);
		if (n > 0) {
			QB_CUDA(cudaMemcpyAsync(rb.cursor.ptr, seg.data(), sizeof(uint64_t) * world, cudaMemcpyHostToDevice, stream));
			route_slot_kernel<<<grid_for(div_up<uint64_t>(n, SCATTER_TILE) * 256, 256, ctx->grid_cap()), 256, 2 * sizeof(unsigned long long) * world, stream>>>(
			    rb.owner.as<uint32_t>(), it->begin.as<uint64_t>(), it->size.as<uint32_t>(), it->mag.as<cplx>(), n, world, rb.cursor.as<unsigned long long>(),
			    rb.s_size.as<uint32_t>(), rb.s_padded.as<uint32_t>(), rb.s_mag.as<cplx>(), rb.s_src.as<uint64_t>());
			++ctx->launches;
			route_bump_kernel<<<1, MAX_DEVICES, 0, stream>>>(rb.s_padded.as<uint32_t>(), bumps, world);
			++ctx->launches;
			exclusive_scan(ctx, widen_u32{rb.s_padded.as<uint32_t>()}, rb.s_begin.as<uint64_t>(), n);
			route_copy_kernel<<<grid_for(n * 32, 256, ctx->grid_cap()), 256, 0, stream>>>(it->objects.as<uint8_t>(), it->begin.as<uint64_t>(), rb.s_src.as<uint64_t>(),
			                                                                                rb.s_begin.as<uint64_t>(), n, rb.s_bytes.as<uint8_t>());
			++ctx->launches;
			QB_CUDA(cudaGetLastError());
			ctx->sync(); // `seg` lives on this stack frame
		}
		// the state this rank owns
		rb.r_size.ensure(sizeof(uint32_t) * std::max<uint64_t>(1, n_recv), stream);
		rb.r_padded.ensure(sizeof(uint32_t) * std::max<uint64_t>(1, n_recv), stream);
		rb.r_mag.ensure(sizeof(cplx) * std::max<uint64_t>(1, n_recv), stream);
		rb.r_begin.ensure(sizeof(uint64_t) * (n_recv + 1), stream);
		rb.r_bytes.ensure(bytes_recv + 16, stream);
	});
	mark("grouping by owner (slots, scan, copy)");
	comm.sum_u64_agreed(0, err, "the grouping of the parents by family owner"); // nothing is posted unless every rank is ready
	mark("agreement");

	// all-to-allv of the four arrays; what stays here is a device copy.  Two NCCL groups: the object bytes alone first (one
	// send and one receive per peer, 99 % of the volume), then the three small arrays -- with all four in one group the twelve
	// operations per peer pair shared the channels and the exchange ran at 240 GB/s per GPU (4 GPUs, QB_DIST_TRACE), against
	// 560 GB/s for the one-array exchange of the record path
	for (int pass = 0; pass < 2; ++pass) {
		uint64_t so = 0, sb = 0, ro = 0, rbytes = 0;
		QB_NCCL(nccl().GroupStart());
		for (uint32_t r = 0; r < world; ++r) {
			if ((int)r != cm->rank) {
				if (send_obj[r]) {
					if (pass == 0) {
						if (send_bytes[r])
							QB_NCCL(nccl().Send(rb.s_bytes.as<uint8_t>() + sb, send_bytes[r], ncclUint8, r, cm->nccl, stream));
					} else {
						QB_NCCL(nccl().Send(rb.s_size.as<uint32_t>() + so, send_obj[r], ncclUint32, r, cm->nccl, stream));
						QB_NCCL(nccl().Send(rb.s_padded.as<uint32_t>() + so, send_obj[r], ncclUint32, r, cm->nccl, stream));
						QB_NCCL(nccl().Send(rb.s_mag.as<cplx>() + so, send_obj[r] * sizeof(cplx), ncclUint8, r, cm->nccl, stream));
					}
				}
				if (recv_obj[r]) {
					if (pass == 0) {
						if (recv_bytes[r])
							QB_NCCL(nccl().Recv(rb.r_bytes.as<uint8_t>() + rbytes, recv_bytes[r], ncclUint8, r, cm->nccl, stream));
					} else {
						QB_NCCL(nccl().Recv(rb.r_size.as<uint32_t>() + ro, recv_obj[r], ncclUint32, r, cm->nccl, stream));
						QB_NCCL(nccl().Recv(rb.r_padded.as<uint32_t>() + ro, recv_obj[r], ncclUint32, r, cm->nccl, stream));
						QB_NCCL(nccl().Recv(rb.r_mag.as<cplx>() + ro, recv_obj[r] * sizeof(cplx), ncclUint8, r, cm->nccl, stream));
					}
				}
			}
			so += send_obj[r];
			sb += send_bytes[r];
			ro += recv_obj[r];
			rbytes += recv_bytes[r];
		}
		QB_NCCL(nccl().GroupEnd());
	}
	{
		uint64_t self_so = 0, self_sb = 0, self_ro = 0, self_rb = 0;
		for (int r = 0; r < cm->rank; ++r) {
			self_so += send_obj[r];
			self_sb += send_bytes[r];
			self_ro += recv_obj[r];
			self_rb += recv_bytes[r];
		}
		const uint64_t k = send_obj[cm->rank];
		if (k) {
			QB_CUDA(cudaMemcpyAsync(rb.r_size.as<uint32_t>() + self_ro, rb.s_size.as<uint32_t>() + self_so, sizeof(uint32_t) * k, cudaMemcpyDeviceToDevice, stream));
			QB_CUDA(cudaMemcpyAsync(rb.r_padded.as<uint32_t>() + self_ro, rb.s_padded.as<uint32_t>() + self_so, sizeof(uint32_t) * k, cudaMemcpyDeviceToDevice, stream));
			QB_CUDA(cudaMemcpyAsync(rb.r_mag.as<cplx>() + self_ro, rb.s_mag.as<cplx>() + self_so, sizeof(cplx) * k, cudaMemcpyDeviceToDevice, stream));
			if (send_bytes[cm->rank])
				QB_CUDA(cudaMemcpyAsync(rb.r_bytes.as<uint8_t>() + self_rb, rb.s_bytes.as<uint8_t>() + self_sb, send_bytes[cm->rank], cudaMemcpyDeviceToDevice, stream));
		}
	}
	mark("all-to-allv of the four arrays");
	// object_begin of what arrived: every segment carries its objects back to back with their padding
	exclusive_scan(ctx, widen_u32{rb.r_padded.as<uint32_t>()}, rb.r_begin.as<uint64_t>(), n_recv);
	// the arrays that arrived become the state (the old ones become next call's receive buffers)
	it->objects.swap(rb.r_bytes);
	it->begin.swap(rb.r_begin);
	it->size.swap(rb.r_size);
	it->mag.swap(rb.r_mag);
	it->n = n_recv;
	it->n_bytes = bytes_recv;
	QB_CUDA(cudaGetLastError());
	return true;
}
