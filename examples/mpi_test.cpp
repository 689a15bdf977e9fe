// Driver for the distributed path: every rank builds its share of a QCGD state, three distributed rule
// iterations, global statistics (the role of the reference's examples/mpi_test.cpp, without the object
// migration utilities that SURVEY 8(f) lists as next).
//
//   torchrun-style launch: RANK / WORLD_SIZE / LOCAL_RANK / MASTER_PORT in the environment, one process per GPU
//   python -m torch.distributed.run --nproc-per-node 2 --no-python ./mpi_test.out
#include "quids_mpi.hpp"
#include "rules/qcgd.hpp"

#include <iostream>

namespace qcgd = quids::rules::qcgd;

int main() {
	quids::tolerance = 1e-18;
	quids::mpi::communicator *comm = quids::mpi::communicator::from_env();

	quids::mpi::mpi_it_t state, buffer;
	quids::mpi::mpi_sy_it_t symbolic;
	// the same seed everywhere, then rank r keeps every size-th graph: a partition of one global state
	std::srand(3);
	const int n_graphs = 64;
	for (int i = 0; i < n_graphs; ++i) {
		char *begin, *end;
		qcgd::utils::make_graph(begin, end, 8);
		qcgd::graphs::randomize(begin);
		if (i % comm->size == comm->rank)
			state.append(begin, end, 1 / std::sqrt((double)n_graphs));
		delete[] begin;
	}

	quids::rule_t *erase_create = new qcgd::erase_create(0.3333), *split_merge = new qcgd::split_merge(0.25, 0.25, 0.25);
	quids::mpi::simulate(state, erase_create, buffer, symbolic, *comm, 5000);
	quids::simulate(buffer, qcgd::step);
	quids::mpi::simulate(buffer, split_merge, state, symbolic, *comm, 5000);
	quids::mpi::simulate(state, erase_create, buffer, symbolic, *comm, 5000);

	const size_t total = buffer.get_total_num_object(*comm), children = symbolic.get_total_num_object(*comm);
	const double nodes = buffer.average_value([](char const *b, char const *) { return (double)qcgd::graphs::num_nodes(b); }, *comm);
	if (comm->rank == 0)
		std::cout << "objects: " << total << ", children of the last step: " << children << ", P=" << buffer.total_proba << ", <nodes>=" << nodes << "\n";
	std::cout << "rank " << comm->rank << ": " << buffer.num_object << " objects, share of the probability " << buffer.node_total_proba << "\n";
	delete comm;
	return 0;
}
