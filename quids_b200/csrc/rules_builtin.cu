// rules_builtin.cu -- instantiates the engine kernels for the rules and modifiers shipped with the
// reference (src/rules/quantum_computer.hpp, src/rules/qcgd.hpp) and registers them.
//
// Every rule is registered twice: under its name with the fused symbolic hook (no child is written
// in the symbolic phase), and as "<name>_generic", which runs the four reference methods only
// (populate_child into scratch + hasher) -- the shape any user-written rule starts from.
#include <cmath>
#include <complex>

#include <quids/device/engine.cuh>
#include "rules_qc.cuh"
#include "rules_qcgd.cuh"

namespace qb {

static int make_hadamard_generic(const double *params, uint32_t num_params, void *storage) {
	if (num_params < 1)
		return QB_ERR_ARG;
	qc::hadamard r;
	r.bit = (uint64_t)params[0];
	r.inv_sqrt2 = 1 / std::sqrt(2.);
	memcpy(storage, &r, sizeof r);
	return QB_OK;
}

QB_REGISTER_RULE(hadamard, qc::hadamard_fused, qc::make_hadamard);
QB_REGISTER_RULE(erase_create, qcgd::flip_rule_fused<true>, qcgd::make_qcgd_rule<qcgd::flip_rule_fused<true>>);
QB_REGISTER_RULE(coin, qcgd::flip_rule_fused<false>, qcgd::make_qcgd_rule<qcgd::flip_rule_fused<false>>);
QB_REGISTER_RULE(split_merge, qcgd::split_merge_fused, qcgd::make_qcgd_rule<qcgd::split_merge_fused>);
QB_REGISTER_RULE(hadamard_generic, qc::hadamard, make_hadamard_generic);
QB_REGISTER_RULE(erase_create_generic, qcgd::flip_rule<true>, qcgd::make_qcgd_rule<qcgd::flip_rule<true>>);
QB_REGISTER_RULE(coin_generic, qcgd::flip_rule<false>, qcgd::make_qcgd_rule<qcgd::flip_rule<false>>);
QB_REGISTER_RULE(split_merge_generic, qcgd::split_merge, qcgd::make_qcgd_rule<qcgd::split_merge>);

// ---- modifiers ------------------------------------------------------------------------------------------
static int make_cnot(const double *p, uint32_t n, void *storage) {
	if (n < 2) return QB_ERR_ARG;
	qc::cnot m{(uint32_t)p[0], (uint32_t)p[1]};
	memcpy(storage, &m, sizeof m);
	return QB_OK;
}
template <class M>
static int make_bit(const double *p, uint32_t n, void *storage) {
	if (n < 1) return QB_ERR_ARG;
	M m{(uint64_t)p[0]};
	memcpy(storage, &m, sizeof m);
	return QB_OK;
}
static int make_phase(const double *p, uint32_t n, void *storage) {
	if (n < 1) return QB_ERR_ARG;
	std::complex<double> rot = std::polar(1.0, p[0]);
	qc::phase m{cplx{rot.real(), rot.imag()}};
	memcpy(storage, &m, sizeof m);
	return QB_OK;
}
template <class M>
static int make_empty(const double *, uint32_t, void *storage) {
	M m;
	memcpy(storage, &m, sizeof m);
	return QB_OK;
}

QB_REGISTER_MODIFIER(cnot, qc::cnot, make_cnot);
QB_REGISTER_MODIFIER(xgate, qc::xgate, make_bit<qc::xgate>);
QB_REGISTER_MODIFIER(ygate, qc::ygate, make_bit<qc::ygate>);
QB_REGISTER_MODIFIER(zgate, qc::zgate, make_bit<qc::zgate>);
QB_REGISTER_MODIFIER(step, qcgd::step_modifier<false>, make_empty<qcgd::step_modifier<false>>);
QB_REGISTER_MODIFIER(reversed_step, qcgd::step_modifier<true>, make_empty<qcgd::step_modifier<true>>);
QB_REGISTER_MODIFIER(phase, qc::phase, make_phase);

// ---- observables ----------------------------------------------------------------------------------------------------------
QB_REGISTER_OBSERVABLE(qcgd_stats, qcgd::stats_observable, make_empty<qcgd::stats_observable>);
QB_REGISTER_OBSERVABLE(qcgd_size, qcgd::size_observable, make_empty<qcgd::size_observable>);
QB_REGISTER_OBSERVABLE(qubit, qc::bit_observable, make_bit<qc::bit_observable>);
QB_REGISTER_OBSERVABLE(object_bytes, qc::bytes_observable, make_empty<qc::bytes_observable>);

} // namespace qb
