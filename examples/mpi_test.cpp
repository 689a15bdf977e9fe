// Distributed driver (one process per GPU, NCCL): the scenario of the reference's MPI example -- a register
// built on ONE rank, gates applied there, objects spread over the ranks, global statistics, the gates undone
// with interference ACROSS ranks, everything gathered again -- written against the drop-in API.
// Expected: the gathered final state is the initial one (SURVEY 8c fixture 3).
//
//   python -m torch.distributed.run --nproc-per-node 2 --no-python ./mpi_test.out
//   (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_PORT come from the launcher)
#include "quids_mpi.hpp"
#include "rules/quantum_computer.hpp"

#include <iostream>
#include <memory>
#include <sstream>
#include <vector>

namespace qc = quids::rules::quantum_computer;
using quids::mpi::communicator;
using quids::mpi::mpi_it_t;

// one block of text per rank, in rank order (a collective stands for the barrier)
static void show(const char *title, mpi_it_t const &part, communicator const &comm) {
	std::ostringstream block;
	std::streambuf *saved = std::cout.rdbuf(block.rdbuf());
	qc::utils::print(part);
	std::cout.rdbuf(saved);
	if (comm.rank == 0)
		std::cout << "\n" << title << ":\n";
	for (int turn = 0; turn < comm.size; ++turn) {
		comm.sum((size_t)0);
		if (turn == comm.rank)
			std::cout << "    node " << comm.rank << "/" << comm.size << ":\n" << block.str() << std::flush;
	}
	comm.sum((size_t)0);
}

int main() {
	quids::tolerance = 1e-8;
	quids::align_byte_length = 0;
	std::unique_ptr<communicator> comm(communicator::from_env());
	const int home = comm->size > 1 ? 1 : 0; // the rank that owns the register at the start and at the end

	std::vector<std::unique_ptr<quids::rule_t>> H;
	for (size_t bit = 0; bit < 3; ++bit)
		H.emplace_back(new qc::hadamard(bit));
	const quids::modifier_t flip2 = qc::Xgate(2);

	mpi_it_t a, b;
	quids::mpi::mpi_sy_it_t scratch;
	if (comm->rank == home) {
		const std::vector<std::pair<std::vector<char>, quids::mag_t>> kets = {
		    {{1, 1, 0, 0}, {0.5, 0}}, {{0, 1, 1, 0, 1}, {0, 0.5}}, {{0, 1, 1, 0, 1, 0}, {0.5, -0.5}}};
		for (auto const &[bits, amplitude] : kets)
			a.append(bits.data(), bits.data() + bits.size(), amplitude);
	}
	show("initial state", a, *comm);

	// local gates on the home rank: H1 H2 H0 X2 (the other ranks hold empty states)
	mpi_it_t *cur = &a, *nxt = &b;
	for (int bit : {1, 2, 0}) {
		quids::simulate(*cur, H[bit].get(), *nxt, scratch);
		std::swap(cur, nxt);
	}
	quids::simulate(*cur, flip2);
	show("applied some gates", *cur, *comm);

	cur->distribute_objects(*comm, home);
	show("distributed all objects", *cur, *comm);

	const double mean_size = cur->average_value([](const char *first, const char *last) { return (PROBA_TYPE)(last - first); }, *comm);
	const size_t everywhere = cur->get_total_num_object(*comm);
	if (comm->rank == 0)
		std::cout << "\nthe average size is " << mean_size << "\nthe total number of objects is " << everywhere << "\n";

	// undo: X2, then H0 H2 H1 with the interference resolved over all ranks
	quids::simulate(*cur, flip2);
	for (int bit : {0, 2, 1}) {
		quids::mpi::simulate(*cur, H[bit].get(), *nxt, scratch, *comm);
		std::swap(cur, nxt);
	}
	if (comm->rank == 0)
		std::cout << "\nP=" << cur->total_proba;
	show("applied all gates in reverse order", *cur, *comm);

	cur->gather_objects(*comm, home);
	show("gathered all objects", *cur, *comm);
	return 0;
}
