// rules_qc.cuh -- device implementations of the quantum_computer rules and modifiers
// (reference: src/rules/quantum_computer.hpp).  Objects are one byte per qubit (0/1).
#pragma once

#include <quids/device/rule_api.cuh>

namespace qb {
namespace qc {

// hadamard(bit), quantum_computer.hpp:31-50.  Two children of the parent's size: child 0 flips the
// bit, child 1 keeps it; the magnitude is scaled by +-1/sqrt(2), minus only for child 1 of a set bit.
// Default hasher (libstdc++ murmur over the raw bytes).
struct hadamard : rule_base<hadamard> {
	uint64_t bit;
	double inv_sqrt2; // 1/std::sqrt(2.) computed on the host like the reference does (:41)

	__device__ void get_num_child(const uint8_t *, uint32_t parent_size, uint32_t &num_child, uint32_t &max_child_size) const {
		num_child = 2;
		max_child_size = parent_size;
	}
	__device__ double factor(const uint8_t *parent, uint32_t child_id) const { return (parent[bit] && child_id) ? -inv_sqrt2 : inv_sqrt2; }

	__device__ void populate_child(const uint8_t *parent, uint32_t parent_size, uint8_t *child, uint32_t child_id, uint32_t &size, cplx &mag) const {
		mag = cscale(mag, factor(parent, child_id));
		size = parent_size;
		for (uint32_t i = 0; i < parent_size; ++i)
			child[i] = parent[i];
		child[bit] ^= (uint8_t)!child_id;
	}
};

// same rule with the fused symbolic hook: the child is hashed as "parent with one byte changed",
// nothing is written in the symbolic phase
struct hadamard_fused : hadamard {
	static constexpr bool needs_scratch = false;
	static constexpr bool has_edit_child = true;

	__device__ void edit_child(const uint8_t *, uint32_t, uint8_t *child, uint32_t child_id) const { child[bit] ^= (uint8_t)!child_id; }

	__device__ uint64_t symbolic(const uint8_t *parent, uint32_t parent_size, const no_ctx &, uint32_t child_id, uint8_t *, uint32_t &size,
	                             cplx &mag) const {
		mag = cscale(mag, factor(parent, child_id));
		size = parent_size;
		const uint8_t flip = (uint8_t)!child_id;
		const uint32_t b = (uint32_t)bit;
		if ((reinterpret_cast<uintptr_t>(parent) & 7) == 0 && (parent_size & 7) == 0) {
			// a register of 8 k qubits on an 8-byte boundary: the murmur words are the object's own 64-bit words (little
			// endian), and the flipped qubit is one XOR into the word that holds it -- 3 loads for 24 qubits instead of 24
			const uint64_t *words = reinterpret_cast<const uint64_t *>(parent);
			const uint32_t flip_word = b >> 3;
			const uint64_t flip_bits = (uint64_t)flip << (8 * (b & 7));
			uint64_t h = 0xc70f6907ull ^ ((uint64_t)parent_size * MURMUR_MUL);
			for (uint32_t k = 0; k < parent_size / 8; ++k) {
				const uint64_t w = words[k] ^ (k == flip_word ? flip_bits : 0);
				h ^= shift_mix(w * MURMUR_MUL) * MURMUR_MUL;
				h *= MURMUR_MUL;
			}
			h = shift_mix(h) * MURMUR_MUL;
			return shift_mix(h);
		}
		return murmur_bytes([=](uint32_t i) { return (uint8_t)(parent[i] ^ (i == b ? flip : 0)); }, parent_size);
	}
};

inline int make_hadamard(const double *params, uint32_t num_params, void *storage) {
	if (num_params < 1)
		return QB_ERR_ARG;
	hadamard_fused r;
	r.bit = (uint64_t)params[0];
	r.inv_sqrt2 = 1 / std::sqrt(2.);
	memcpy(storage, &r, sizeof r);
	return QB_OK;
}

// ---- modifiers (quantum_computer.hpp:25-29, 52-75) ------------------------------------------------
struct cnot {
	uint32_t control, target;
	__device__ void operator()(uint8_t *b, uint32_t, cplx &) const { b[target] ^= b[control]; }
};
struct xgate {
	uint64_t bit;
	__device__ void operator()(uint8_t *b, uint32_t, cplx &) const { b[bit] = !b[bit]; }
};
struct ygate { // mag *= i; if set: mag *= -1; flip
	uint64_t bit;
	__device__ void operator()(uint8_t *b, uint32_t, cplx &mag) const {
		cplx m{-mag.im, mag.re};
		if (b[bit])
			m = cplx{-m.re, -m.im};
		mag = m;
		b[bit] = !b[bit];
	}
};
struct zgate { // flips the bit as well, like the reference (:68-75)
	uint64_t bit;
	__device__ void operator()(uint8_t *b, uint32_t, cplx &mag) const {
		if (b[bit])
			mag = cplx{-mag.re, -mag.im};
		b[bit] = !b[bit];
	}
};
// bench modifier of SURVEY 8(d) C2: reads the object, writes only the magnitude
struct phase {
	cplx rot;
	__device__ void operator()(uint8_t *b, uint32_t, cplx &mag) const {
		if (b[0] & 1)
			mag = cmul(mag, rot);
	}
};

// ---- observables ------------------------------------------------------------------------------------------------------
// probability that a qubit is set: observable = object[bit] (a driver would write it as a lambda for average_value)
struct bit_observable {
	static constexpr int values = 1;
	uint64_t bit;
	__device__ void operator()(const uint8_t *object, uint32_t size, double *out) const { out[0] = bit < size && object[bit] ? 1.0 : 0.0; }
};
// size of the object in bytes
struct bytes_observable {
	static constexpr int values = 1;
	__device__ void operator()(const uint8_t *, uint32_t size, double *out) const { out[0] = (double)size; }
};

} // namespace qc
} // namespace qb
