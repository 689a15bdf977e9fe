// plugin.cuh -- the DEVICE side of the plug-in boundary: what a user of the reference writes as a subclass of quids::rule
// (four virtual methods, quids.hpp:105-146) or as a modifier lambda (quids.hpp:86, 436-438, 973-980) is written here as
// __device__ code and compiled by the USER, out of tree, against these headers and libquids_b200.so:
//
//     nvcc -gencode arch=compute_100a,code=sm_100a -std=c++20 --extended-lambda --expt-relaxed-constexpr \
//          -I<repo>/include my_rules.cu -L<repo>/quids_b200 -lquids_b200 -Xlinker -rpath=<repo>/quids_b200 [-shared -Xcompiler -fPIC]
//
//   * RULES: a trivially copyable struct deriving from qb::rule_base<T> with the reference's four methods as __device__
//     members (rule_api.cuh; only get_num_child and populate_child are mandatory, as in the reference), plus a `make`
//     function that builds it from the constructor arguments, registered with
//         QB_REGISTER_RULE(name, type, make);
//     in any translation unit of the user's program or shared library.  The macro instantiates the engine's kernels for
//     the type in THAT translation unit and hands their launchers to the library's registry (qb::register_rule, exported by
//     libquids_b200.so) when the module is loaded; from then on quids::rule("name", {args...}) / qb_rule_id("name") /
//     qb.Rule("name", ...) drive it like a built-in.  QB_REGISTER_MODIFIER and QB_REGISTER_OBSERVABLE work the same way.
//   * MODIFIER LAMBDAS: qb::apply_device_modifier(state, f) runs any __device__ callable
//         f(char *object_begin, char *object_end, device_mag_t &mag)     (the reference's modifier_t signature)
//     over a state in HBM -- e.g. an extended lambda `[=] __device__ (char *b, char *e, qb::device_mag_t &mag) {...}` -- and
//     lambda.cuh adds the matching overload quids::simulate(it_t &, F) for drivers written against quids.hpp.
//     device_mag_t = cuda::std::complex<double>: the device-usable twin of std::complex<double> (same layout; std::complex
//     itself is built on _Complex, which nvcc does not accept in device code).
// See examples/custom_rule.cu.
#pragma once

#include <cuda/std/complex>

#include "engine.cuh"

namespace qb {

typedef cuda::std::complex<double> device_mag_t;

// the reference's modifier signature on top of the engine's functor shape.  Mag = cuda::std::complex<double> sees the magnitude
// in place; any other complex type (cuda::std::complex<float> for PROBA_TYPE = float) works on a converted copy.
template <class F, class Mag>
struct modifier_from_callable {
	F f;
	__device__ void operator()(uint8_t *object, uint32_t size, cplx &mag) const {
		char *begin = reinterpret_cast<char *>(object);
		if constexpr (sizeof(Mag) == sizeof(cplx)) {
			f(begin, begin + size, *reinterpret_cast<Mag *>(&mag));
		} else {
			Mag m(static_cast<typename Mag::value_type>(mag.re), static_cast<typename Mag::value_type>(mag.im));
			f(begin, begin + size, m);
			mag = cplx{(double)m.real(), (double)m.imag()};
		}
	}
};

// quids::simulate(it_t &, modifier_t) for a __device__ callable (quids.hpp:436-438, 973-980): one pass over the state in HBM,
// in place, object sizes unchanged, no normalisation.  Synchronous like every entry point of the C ABI.
template <class Mag = device_mag_t, class F>
inline void apply_device_modifier(qb_iter *state, F f) {
	QB_REQUIRE(state, QB_ERR_ARG, "apply_device_modifier: null state");
	uint64_t n = 0;
	int rc = qb_iter_counts(state, &n, nullptr, nullptr);
	QB_REQUIRE(rc == QB_OK, rc, qb_last_error());
	if (n == 0)
		return;
	void *objects = nullptr, *begin = nullptr, *size = nullptr, *mag = nullptr;
	rc = qb_iter_device_ptrs(state, &objects, &begin, &size, &mag);
	QB_REQUIRE(rc == QB_OK, rc, qb_last_error());
	qb_ctx *ctx = qb_iter_ctx(state);
	int device = qb_ctx_device(ctx), sm_count = 0;
	QB_CUDA(cudaSetDevice(device));
	QB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
	cudaStream_t stream = static_cast<cudaStream_t>(qb_ctx_stream(ctx));
	const iter_view view{static_cast<uint8_t *>(objects), static_cast<const uint64_t *>(begin), static_cast<const uint32_t *>(size), static_cast<cplx *>(mag), n};
	typedef modifier_from_callable<F, Mag> wrapped;
	modifier_glue<wrapped>::launch(&static_cast<const wrapped &>(wrapped{f}), view, stream, sm_count);
	QB_CUDA(cudaGetLastError());
	QB_CUDA(cudaStreamSynchronize(stream));
}

} // namespace qb
