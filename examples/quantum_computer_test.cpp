// Driver: a small circuit on two bit strings of different length, then the inverse circuit
// (the scenario of the reference's examples/quantum_computer_test.cpp).
//
// This file only uses the public QuIDS API, so it builds against EITHER header set:
//   this repository:  g++ -std=c++17 -I../include -I../include/quids quantum_computer_test.cpp -L../quids_b200 -lquids_b200
//   the reference:    g++ --std=c++2a -O3 -fopenmp -I/root/reference/src quantum_computer_test.cpp
// examples/Makefile builds both; tests/test_examples.py compares their transcripts.
#include "quids.hpp"
#include "rules/quantum_computer.hpp"

#include <cmath>
#include <iostream>

namespace qc = quids::rules::quantum_computer;

static void show(const char *title, quids::it_t const &state) {
	std::cout << title << ":\n";
	qc::utils::print(state);
	std::cout << "\n";
}

int main() {
	quids::align_byte_length = 0;

	quids::rule_t *h1 = new qc::hadamard(1), *h2 = new qc::hadamard(2);
	quids::modifier_t cnot13 = qc::cnot(1, 3), x2 = qc::Xgate(2), y0 = qc::Ygate(0), z3 = qc::Zgate(3);

	quids::sy_it_t symbolic;
	quids::it_t state, buffer;
	char first[4] = {1, 1, 0, 0}, second[5] = {0, 1, 1, 0, 1};
	const double amplitude = 1 / std::sqrt(2);
	state.append(first, first + 4, amplitude);
	state.append(second, second + 5, {0, amplitude});
	show("initial state", state);

	quids::simulate(state, h1, buffer, symbolic);
	show("hadamard on qubit 1", buffer);
	quids::simulate(buffer, h2, state, symbolic);
	show("hadamard on qubit 2", state);
	quids::simulate(state, cnot13);
	show("cnot on qubit 3 controlled by qubit 1", state);
	quids::simulate(state, x2);
	show("X on qubit 2", state);
	quids::simulate(state, y0);
	show("Y on qubit 0", state);
	quids::simulate(state, z3);
	show("Z on qubit 3", state);

	// the same gates backwards: Z and Y and X and CNOT undo themselves (up to a sign), then H2, H1
	quids::simulate(state, z3);
	quids::simulate(state, y0);
	quids::simulate(state, x2);
	quids::simulate(state, cnot13);
	quids::simulate(state, h2, buffer, symbolic);
	quids::simulate(buffer, h1, state, symbolic);
	show("after the inverse circuit", state);
	std::cout << "objects: " << state.num_object << ", symbolic objects of the last step: " << symbolic.num_object
	          << ", after interferences: " << symbolic.num_object_after_interferences << "\n";
	return 0;
}
