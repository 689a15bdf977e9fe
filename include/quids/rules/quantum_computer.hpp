// rules/quantum_computer.hpp -- drop-in for the reference's src/rules/quantum_computer.hpp: the same
// class / factory names and constructor arguments; the bodies run on the GPU (device code:
// quids_b200/csrc/rules_qc.cuh).  Objects are bit strings, one byte (0/1) per qubit.
#pragma once

#include <cmath>
#include <iostream>

#include "../quids.hpp"

namespace quids::rules::quantum_computer {
	namespace utils {
		/// prints "re +/- im i  bits" for every object (quantum_computer.hpp:10-22)
		inline void print(quids::it_t const &iter) {
			for (size_t oid = 0; oid < iter.num_object; ++oid) {
				uint size;
				mag_t mag;
				char const *begin;
				iter.get_object(oid, begin, size, mag);
				std::cout << "\t" << mag.real() << (mag.imag() < 0 ? " - " : " + ") << std::abs(mag.imag()) << "i  ";
				for (uint i = 0; i < size; ++i)
					std::cout << (begin[i] ? '1' : '0');
				std::cout << "\n";
			}
		}
	}

	/// controlled not: bit ^= control_bit (quantum_computer.hpp:25-29)
	modifier_t inline cnot(uint32_t control_bit, uint32_t bit) { return modifier_t("cnot", {(double)control_bit, (double)bit}); }

	/// Hadamard gate on one qubit: two children per object (quantum_computer.hpp:31-50)
	class hadamard : public quids::rule {
	public:
		hadamard(size_t bit_) : quids::rule("hadamard", {(double)bit_}) {}
	};

	modifier_t inline Xgate(size_t bit) { return modifier_t("xgate", {(double)bit}); } // quantum_computer.hpp:52-56
	modifier_t inline Ygate(size_t bit) { return modifier_t("ygate", {(double)bit}); } // quantum_computer.hpp:58-66
	/// as in the reference, Z also flips the bit (quantum_computer.hpp:68-75)
	modifier_t inline Zgate(size_t bit) { return modifier_t("zgate", {(double)bit}); }
}
