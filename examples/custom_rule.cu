// custom_rule.cu -- USER-WRITTEN device rules and modifier lambdas, built OUT OF TREE against the installed headers
// (include/quids/device/) and libquids_b200.so: nothing here is compiled into the library.
//
//   reference                                                   here
//   class my_rule : public quids::rule { four virtual methods } struct my_rule : qb::rule_base<my_rule> { the same four methods, __device__ }
//   (quids.hpp:105-146)                                         + QB_REGISTER_RULE(name, my_rule, make)            (plugin.cuh)
//   quids::simulate(state, [](char *b, char *e, mag_t &m){..})  quids::simulate(state, [=] __device__ (char *b, char *e, quids::device_mag_t &m){..})
//   (quids.hpp:86,436-438)                                                                                         (lambda.cuh)
//
// Build (examples/Makefile does both):
//   executable   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 --extended-lambda -I../include custom_rule.cu \
//                     -L../quids_b200 -lquids_b200 -Xlinker -rpath=$PWD/../quids_b200 -o custom_rule.out
//   plug-in .so  ... -DCUSTOM_RULE_NO_MAIN -shared -Xcompiler -fPIC -o libcustom_rule.so      (loaded by tests/test_plugin.py:
//                the rules register themselves when the module is loaded, and are then driven BY NAME through the C ABI)
#include <cmath>
#include <cstdio>
#include <vector>

#include <quids/device/lambda.cuh>
#include <quids/rules/quantum_computer.hpp>

namespace user {

// ---- (1) the Hadamard gate as a user would port it from quantum_computer.hpp:31-50: the four reference methods only
//          (get_num_child, populate_child; populate_child_simple and hasher keep their defaults, as in the reference) ----
struct my_hadamard : qb::rule_base<my_hadamard> {
	uint32_t bit;
	double s; // 1 / sqrt(2.), computed on the host

	__device__ void get_num_child(const uint8_t *, uint32_t parent_size, uint32_t &num_child, uint32_t &max_child_size) const {
		num_child = 2;
		max_child_size = parent_size;
	}
	__device__ void populate_child(const uint8_t *parent, uint32_t parent_size, uint8_t *child, uint32_t child_id, uint32_t &size, qb::cplx &mag) const {
		for (uint32_t i = 0; i < parent_size; ++i)
			child[i] = parent[i];
		size = parent_size;
		mag = qb::cscale(mag, (parent[bit] && child_id) ? -s : s);
		child[bit] ^= (uint8_t)!child_id;
	}
};
int make_my_hadamard(const double *params, uint32_t num_params, void *storage) {
	if (num_params < 1)
		return QB_ERR_ARG;
	my_hadamard r;
	r.bit = (uint32_t)params[0];
	r.s = 1 / std::sqrt(2.);
	memcpy(storage, &r, sizeof r);
	return QB_OK;
}

// ---- (2) a rule the library does not ship: the rotation Ry(theta) of one qubit,
//          |0> -> cos(theta/2) |0> + sin(theta/2) |1>,   |1> -> -sin(theta/2) |0> + cos(theta/2) |1>
//          Two children per object; with theta = 0 or pi one of them has magnitude 0 and is dropped by the tolerance. ----
struct ry_gate : qb::rule_base<ry_gate> {
	uint32_t bit;
	double c, s;

	__device__ void get_num_child(const uint8_t *, uint32_t parent_size, uint32_t &num_child, uint32_t &max_child_size) const {
		num_child = 2;
		max_child_size = parent_size;
	}
	// child_id = value of the qubit in the child
	__device__ void populate_child(const uint8_t *parent, uint32_t parent_size, uint8_t *child, uint32_t child_id, uint32_t &size, qb::cplx &mag) const {
		for (uint32_t i = 0; i < parent_size; ++i)
			child[i] = parent[i];
		size = parent_size;
		const bool from = parent[bit] != 0, to = child_id != 0;
		mag = qb::cscale(mag, from == to ? c : (to ? s : -s));
		child[bit] = (uint8_t)to;
	}
};
int make_ry(const double *params, uint32_t num_params, void *storage) {
	if (num_params < 2)
		return QB_ERR_ARG;
	ry_gate r;
	r.bit = (uint32_t)params[0];
	r.c = std::cos(params[1] / 2);
	r.s = std::sin(params[1] / 2);
	memcpy(storage, &r, sizeof r);
	return QB_OK;
}

// ---- (3) a modifier registered by name: swap two qubits ----
struct swap_bits {
	uint32_t a, b;
	__device__ void operator()(uint8_t *object, uint32_t, qb::cplx &) const {
		const uint8_t t = object[a];
		object[a] = object[b];
		object[b] = t;
	}
};
int make_swap(const double *params, uint32_t num_params, void *storage) {
	if (num_params < 2)
		return QB_ERR_ARG;
	swap_bits m{(uint32_t)params[0], (uint32_t)params[1]};
	memcpy(storage, &m, sizeof m);
	return QB_OK;
}

} // namespace user

QB_REGISTER_RULE(user_hadamard, user::my_hadamard, user::make_my_hadamard);
QB_REGISTER_RULE(user_ry, user::ry_gate, user::make_ry);
QB_REGISTER_MODIFIER(user_swap, user::swap_bits, user::make_swap);

// ---- (4) modifier LAMBDAS on a C-ABI state handle (what a binding in another language would call) ----
extern "C" int user_phase_lambda(qb_iter *state, double theta) { // the bench's phase modifier: mag *= e^{i theta} when obj[0] & 1
	const qb::device_mag_t phase(std::cos(theta), std::sin(theta));
	try {
		qb::apply_device_modifier(state, [=] __device__(char *begin, char *, qb::device_mag_t &mag) {
			if (begin[0] & 1)
				mag *= phase;
		});
	} catch (const qb::error &e) {
		fprintf(stderr, "user_phase_lambda: %s\n", e.what());
		return e.status;
	}
	return QB_OK;
}
extern "C" int user_xgate_lambda(qb_iter *state, unsigned bit) { // quantum_computer.hpp:52-56 as a lambda
	try {
		qb::apply_device_modifier(state, [=] __device__(char *begin, char *, qb::device_mag_t &) { begin[bit] = !begin[bit]; });
	} catch (const qb::error &e) {
		fprintf(stderr, "user_xgate_lambda: %s\n", e.what());
		return e.status;
	}
	return QB_OK;
}

#ifndef CUSTOM_RULE_NO_MAIN
// ---- a driver written against the drop-in header API, with the user's rules and a lambda ----
class Ry : public quids::rule { // host handle, like the rule classes of rules/quantum_computer.hpp
public:
	Ry(size_t bit, double theta) : quids::rule("user_ry", {(double)bit, theta}) {}
};

static int check(bool ok, const char *what) {
	printf("%s %s\n", ok ? "ok" : "FAILED", what);
	return ok ? 0 : 1;
}

int main() {
	namespace qc = quids::rules::quantum_computer;
	quids::align_byte_length = 0;
	quids::tolerance = 1e-20;
	quids::it_t state, buffer;
	quids::sy_it_t sy_it;
	char zero[6] = {0, 0, 0, 0, 0, 0};
	state.append(zero, zero + 6);
	int failed = 0;

	// Ry(theta) on every qubit: the product state (cos|0> + sin|1>)^6, 64 objects
	const double theta = 0.8;
	for (size_t bit = 0; bit < 6; ++bit) {
		Ry rule(bit, theta);
		quids::simulate(bit % 2 ? buffer : state, &rule, bit % 2 ? state : buffer, sy_it, (size_t)-1);
	}
	failed += check(state.num_object == 64, "user_ry: 6 rotations -> 64 objects");
	double worst = 0;
	for (size_t oid = 0; oid < state.num_object; ++oid) {
		char const *b;
		uint size;
		quids::mag_t mag;
		state.get_object(oid, b, size, mag);
		int ones = 0;
		for (uint i = 0; i < size; ++i)
			ones += b[i];
		const double expect = std::pow(std::cos(theta / 2), 6 - ones) * std::pow(std::sin(theta / 2), ones);
		worst = std::max(worst, std::abs(mag - quids::mag_t(expect, 0)));
	}
	failed += check(worst < 1e-14, "user_ry: magnitudes = cos^(6-k) sin^k");

	// a modifier lambda through quids::simulate: a phase on the objects whose qubit 2 is set, then its inverse
	const quids::device_mag_t phase(std::cos(0.3), std::sin(0.3)), back(std::cos(0.3), -std::sin(0.3));
	quids::simulate(state, [=] __device__(char *begin, char *, quids::device_mag_t &mag) { if (begin[2]) mag *= phase; });
	double moved = 0;
	for (size_t oid = 0; oid < state.num_object; ++oid) {
		char const *b;
		uint size;
		quids::mag_t mag;
		state.get_object(oid, b, size, mag);
		moved = std::max(moved, std::abs(mag.imag()));
	}
	failed += check(moved > 1e-3, "lambda modifier: the phase reached the state");
	quids::simulate(state, [=] __device__(char *begin, char *, quids::device_mag_t &mag) { if (begin[2]) mag *= back; });

	// the registered modifier, twice (a swap is its own inverse), then the rotations undone: back to |000000>
	quids::simulate(state, quids::modifier_t("user_swap", {1, 4}));
	quids::simulate(state, quids::modifier_t("user_swap", {1, 4}));
	for (size_t bit = 0; bit < 6; ++bit) {
		Ry rule(bit, -theta);
		quids::simulate(bit % 2 ? buffer : state, &rule, bit % 2 ? state : buffer, sy_it, (size_t)-1);
	}
	failed += check(state.num_object == 1, "inverse rotations interfere back to one object");
	char const *b;
	uint size;
	quids::mag_t mag;
	state.get_object(0, b, size, mag);
	bool zeros = size == 6;
	for (uint i = 0; i < size; ++i)
		zeros = zeros && b[i] == 0;
	failed += check(zeros && std::abs(mag - quids::mag_t(1, 0)) < 1e-13, "... which is |000000> with magnitude 1");
	qc::utils::print(state);
	return failed;
}
#endif
