// quids_mpi.hpp -- drop-in for the distributed API of the reference (src/quids_mpi.hpp), with NCCL over
// NVLink in place of MPI: one process per GPU, hash-ownership interference inside libquids_b200.so
// (qb_simulate_dist, see quids_b200/csrc/dist.inc.cuh for the protocol).
//
//   quids::mpi::mpi_it_t, mpi_sy_it_t                                         quids_mpi.hpp:59-61
//   quids::mpi::simulate(mpi_it_t&, rule_t const*, mpi_it_t&, mpi_sy_it_t&, communicator, size_t = 0, debug_t = {})   :423
// The communicator argument keeps its position; its type is quids::mpi::communicator (an NCCL
// communicator over the GPUs of the job) instead of MPI_Comm.
//
// Differences from the reference, on purpose:
//   * truncation keeps the max_num_object most probable objects over ALL ranks (the single-node result);
//     the reference keeps max_num_object / local_size per rank (quids_mpi.hpp:537,590).
//   * object migration (send_objects / receive_objects / equalize / distribute_objects / gather_objects,
//     quids_mpi.hpp:124-231,903-1077) moves the tail of a state HBM -> NVLink -> HBM in one NCCL group; the
//     node arguments are ranks of the communicator, which comes LAST-but-one as in the reference.
#pragma once

#include <chrono>
#include <cstdio>
#include <fstream>
#include <thread>

#include "quids.hpp"

namespace quids::mpi {
	// knobs of the reference's load balancer (quids_mpi.hpp:46-56), kept so that drivers assigning them compile
	inline size_t min_equalize_size = 100;
	inline float equalize_inbalance = 0.1f;
	inline float min_equalize_step = 0.2f;
	inline bool equalize_children = true;

	/// the job's communicator: stands where MPI_Comm stands in the reference
	class communicator {
	public:
		int rank = 0, size = 1;

		/// from an NCCL unique id that the caller has shared between the ranks by any means
		communicator(int world_size, int rank_, const uint8_t id[128]) : rank(rank_), size(world_size) {
			quids::detail::check(qb_comm_create(quids::detail::context(), world_size, rank_, id, &handle_));
		}
		communicator(const communicator &) = delete;
		communicator &operator=(const communicator &) = delete;
		~communicator() { qb_comm_destroy(handle_); }

		/// RANK / WORLD_SIZE from the environment (torchrun, mpirun wrappers...); the NCCL id travels through the
		/// file QUIDS_COMM_FILE (default /tmp/quids_nccl_id.<MASTER_PORT>), written by rank 0
		static communicator *from_env() {
			const char *r = std::getenv("RANK"), *w = std::getenv("WORLD_SIZE"), *port = std::getenv("MASTER_PORT"), *f = std::getenv("QUIDS_COMM_FILE");
			const int rank = r ? std::atoi(r) : 0, world = w ? std::atoi(w) : 1;
			const std::string path = f ? f : std::string("/tmp/quids_nccl_id.") + (port ? port : "0");
			uint8_t id[128];
			if (rank == 0) {
				quids::detail::check(qb_comm_unique_id(id));
				std::ofstream(path + ".tmp", std::ios::binary).write(reinterpret_cast<const char *>(id), 128);
				std::rename((path + ".tmp").c_str(), path.c_str());
			} else {
				for (int tries = 0;; ++tries) {
					std::ifstream in(path, std::ios::binary);
					if (in.read(reinterpret_cast<char *>(id), 128) && in.gcount() == 128)
						break;
					if (tries > 6000)
						throw std::runtime_error("quids::mpi: no NCCL id at " + path);
					std::this_thread::sleep_for(std::chrono::milliseconds(10));
				}
			}
			communicator *c = new communicator(world, rank, id);
			if (rank == 0)
				std::remove(path.c_str());
			return c;
		}

		size_t sum(size_t v) const {
			uint64_t x = v;
			quids::detail::check(qb_comm_allreduce_u64(handle_, &x, 1, 0));
			return x;
		}
		PROBA_TYPE sum(PROBA_TYPE v) const {
			double x = v;
			quids::detail::check(qb_comm_allreduce_f64(handle_, &x, 1));
			return x;
		}
		qb_comm *handle() const { return handle_; }

	private:
		qb_comm *handle_ = nullptr;
	};

	typedef class mpi_iteration mpi_it_t;
	typedef class mpi_symbolic_iteration mpi_sy_it_t;

	/// this rank's share of the wave function (quids_mpi.hpp:64-310)
	class mpi_iteration : public quids::iteration {
	public:
		/// share of the total probability held by this rank (quids_mpi.hpp:67)
		PROBA_TYPE node_total_proba = 1;

		mpi_iteration() {}
		mpi_iteration(char *object_begin_, char *object_end_) : quids::iteration(object_begin_, object_end_) {}

		size_t get_total_num_object(communicator const &comm) const { return comm.sum(num_object); }          // quids_mpi.hpp:77-87
		/// children counted by the last rule iteration over this state, summed over the ranks (quids_mpi.hpp:92-96)
		size_t get_total_num_symbolic_object(communicator const &comm) const {
			uint64_t n = 0;
			quids::detail::check(qb_iter_num_symbolic_object(handle_, &n));
			return comm.sum((size_t)n);
		}
		PROBA_TYPE average_value(const quids::observable_t observable) const { return quids::iteration::average_value(observable); }
		/// global average of an observable (quids_mpi.hpp:101-116): local averages weighted by the local share, summed
		PROBA_TYPE average_value(const quids::observable_t observable, communicator const &comm) const {
			return comm.sum(quids::iteration::average_value(observable));
		}

		/// send the last num_object_sent objects to rank `node`, which must call receive_objects (quids_mpi.hpp:124-172)
		void send_objects(size_t num_object_sent, int node, communicator const &comm, bool /*send_num_child*/ = false) {
			to_device();
			quids::detail::check(qb_iter_send_objects(handle_, comm.handle(), num_object_sent, node, nullptr));
			after_migration();
		}
		/// receive at the tail what rank `node` sends (quids_mpi.hpp:180-231); max_mem = -1: whatever fits the GPU
		void receive_objects(int node, communicator const &comm, bool /*receive_num_child*/ = false, size_t max_mem = -1) {
			to_device();
			quids::detail::check(qb_iter_receive_objects(handle_, comm.handle(), node, (uint64_t)max_mem, nullptr));
			after_migration();
		}
		/// one pairing round: the i-th fullest rank gives half of the difference to the i-th emptiest (quids_mpi.hpp:903-960)
		void equalize(communicator const &comm) {
			to_device();
			quids::detail::check(qb_iter_equalize(handle_, comm.handle(), 0, nullptr, 0, 1, 0, -1.f, 0.f, nullptr));
			after_migration();
		}
		/// spread the objects of rank node_id evenly over all ranks (quids_mpi.hpp:1031-1051)
		void distribute_objects(communicator const &comm, int node_id = 0) {
			to_device();
			quids::detail::check(qb_iter_distribute_objects(handle_, comm.handle(), node_id));
			after_migration();
		}
		/// bring every object to rank node_id (quids_mpi.hpp:1056-1077)
		void gather_objects(communicator const &comm, int node_id = 0) {
			to_device();
			quids::detail::check(qb_iter_gather_objects(handle_, comm.handle(), node_id));
			after_migration();
			node_total_proba = comm.rank == node_id; // quids_mpi.hpp:1076
		}

	private:
		friend void simulate(mpi_it_t &, quids::rule_t const *, mpi_it_t &, mpi_sy_it_t &, communicator &, size_t, quids::debug_t);
		void after_migration() { // counts changed on the device; total_proba is a property of the whole wave function and stays
			const PROBA_TYPE proba = total_proba;
			after_device_write();
			total_proba = proba;
		}
	};

	class mpi_symbolic_iteration : public quids::symbolic_iteration {
	public:
		size_t get_total_num_object(communicator const &comm) const { return comm.sum(num_object); }                                                   // quids_mpi.hpp:322-330
		size_t get_total_num_object_after_interferences(communicator const &comm) const { return comm.sum(num_object_after_interferences); } // :331-339

	private:
		friend void simulate(mpi_it_t &, quids::rule_t const *, mpi_it_t &, mpi_sy_it_t &, communicator &, size_t, quids::debug_t);
	};

	/// quids::mpi::simulate (quids_mpi.hpp:423-598).  max_num_object counts objects over all ranks; -1 = no truncation;
	/// 0 (the default, as in the reference) = automatic budget: every rank keeps the most probable of its parents whose
	/// workspace fits its GPU, and the ranks agree on how many children the next states can hold (capi.cu simulate_dist).
	void inline simulate(mpi_it_t &iteration, quids::rule_t const *rule, mpi_it_t &next_iteration, mpi_sy_it_t &symbolic_iteration, communicator &comm,
	                     size_t max_num_object = 0, quids::debug_t mid_step_function = [](const char *) {}) {
		iteration.to_device();
		qb_options opt = quids::detail::options();
		opt.equalize = equalize_children ? 2 : 1; // quids_mpi.hpp:442-500
		opt.equalize_inbalance = equalize_inbalance;
		opt.min_equalize_step = min_equalize_step;
		opt.min_equalize_size = min_equalize_size;
		const uint64_t k = max_num_object == std::numeric_limits<size_t>::max() ? QB_NO_TRUNCATION : (uint64_t)max_num_object;
		double node = 1;
		quids::detail::check(qb_simulate_dist(iteration.handle_, rule->id(), rule->params().data(), (uint32_t)rule->params().size(), next_iteration.handle_,
		                                      symbolic_iteration.handle_, comm.handle(), k, &opt, mid_step_function ? quids::detail::forward_step : nullptr,
		                                      &mid_step_function, &node));
		symbolic_iteration.refresh();
		iteration.after_migration(); // the load balancer may have moved parents
		next_iteration.after_device_write();
		next_iteration.node_total_proba = node;
	}
}
