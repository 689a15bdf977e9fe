// Driver: quantum causal graph dynamics on one random graph, forward then reversed
// (the scenario of the reference's examples/qcgd_test.cpp; the final state must be the initial graph, P = 1).
//
//   usage: qcgd_test [n_node = 6] [seed = 1]
//
// Builds against either header set (see quantum_computer_test.cpp).
#include "quids.hpp"
#include "rules/qcgd.hpp"

#include <iomanip>
#include <iostream>
#include <string>

namespace qcgd = quids::rules::qcgd;

// small states are printed graph by graph; large ones through order-independent observables
// (which of several equally probable graphs come first in a print-out is unspecified)
static void show(const char *title, quids::it_t const &state) {
	std::cout << title << " (" << state.num_object << " graphs, P=" << std::fixed << std::setprecision(5) << state.total_proba << "):\n";
	if (state.num_object <= 64) {
		qcgd::utils::print(state);
	} else {
		const double nodes = state.average_value([](char const *b, char const *) { return (double)qcgd::graphs::num_nodes(b); });
		const double particles = state.average_value([](char const *b, char const *) {
			double count = 0;
			for (int i = 0; i < qcgd::graphs::num_nodes(b); ++i)
				count += qcgd::graphs::left(b, i) + qcgd::graphs::right(b, i);
			return count;
		});
		std::cout << std::setprecision(9) << "\t<nodes> = " << nodes << ", <particles> = " << particles << "\n";
	}
	std::cout << "\n";
}

int main(int argc, char *argv[]) {
	const std::string n_node = argc > 1 ? argv[1] : "6", seed = argc > 2 ? argv[2] : "1";
	quids::tolerance = 1e-15;
	quids::simple_truncation = true;

	quids::sy_it_t symbolic;
	quids::it_t state, buffer;
	qcgd::flags::read_n_iter(("1,seed=" + seed).c_str());
	qcgd::flags::read_state(n_node.c_str(), state);

	quids::rule_t *erase_create = new qcgd::erase_create(0.3333);
	quids::rule_t *erase_create_phase = new qcgd::erase_create(0.25, 0.25);
	quids::rule_t *split_merge = new qcgd::split_merge(0.25, 0.25, 0.25);
	quids::rule_t *reversed_split_merge = new qcgd::split_merge(0.25, 0.25, -0.25);
	const size_t no_truncation = -1;

	show("initial state", state);
	quids::simulate(state, qcgd::step);
	show("after step", state);
	quids::simulate(state, erase_create_phase, buffer, symbolic, no_truncation);
	show("after erase_create(0.25, 0.25)", buffer);
	quids::simulate(buffer, erase_create_phase, state, symbolic, no_truncation);
	quids::simulate(state, erase_create, buffer, symbolic, no_truncation);
	show("after a second one and erase_create(0.3333)", buffer);
	quids::simulate(buffer, split_merge, state, symbolic, no_truncation);
	show("after split_merge", state);
	quids::simulate(state, qcgd::step);
	quids::simulate(state, split_merge, buffer, symbolic, no_truncation);
	show("after step and split_merge", buffer);
	std::cout << "symbolic objects: " << symbolic.num_object << ", after interferences: " << symbolic.num_object_after_interferences << "\n\n";

	quids::simulate(buffer, reversed_split_merge, state, symbolic, no_truncation);
	quids::simulate(state, qcgd::reversed_step);
	quids::simulate(state, reversed_split_merge, buffer, symbolic, no_truncation);
	quids::simulate(buffer, erase_create, state, symbolic, no_truncation);
	quids::simulate(state, qcgd::reversed_step);
	show("after the reversed sequence", state);
	return 0;
}
