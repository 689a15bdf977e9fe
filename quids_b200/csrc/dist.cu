// dist.cu -- distributed path (quids::mpi::simulate, quids_mpi.hpp:423-598) over NCCL.
#include "common.cuh"

static thread_local std::string g_dist_error;

extern "C" {

int qb_comm_unique_id(uint8_t id[128]) { return QB_ERR_UNSUPPORTED; }
int qb_comm_create(qb_ctx *, int, int, const uint8_t[128], qb_comm **) { return QB_ERR_UNSUPPORTED; }
int qb_comm_destroy(qb_comm *) { return QB_ERR_UNSUPPORTED; }
int qb_simulate_dist(qb_iter *, int, const double *, uint32_t, qb_iter *, qb_sym *, qb_comm *, uint64_t, const qb_options *, qb_step_cb, void *, double *) {
	return QB_ERR_UNSUPPORTED;
}
int qb_comm_allreduce_u64(qb_comm *, uint64_t *, uint32_t, int) { return QB_ERR_UNSUPPORTED; }
int qb_comm_allreduce_f64(qb_comm *, double *, uint32_t) { return QB_ERR_UNSUPPORTED; }
}
