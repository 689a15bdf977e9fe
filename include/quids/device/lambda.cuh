// lambda.cuh -- the in-place modifier LAMBDA path of the reference (quids.hpp:86: modifier_t is any callable
// void(char *begin, char *end, mag_t &mag); :436-438 quids::simulate(it_t &, modifier_t); :973-980 apply_modifier) for
// drivers written against include/quids/quids.hpp and compiled with nvcc:
//
//     quids::simulate(state, [=] __device__ (char *begin, char *end, quids::device_mag_t &mag) { if (begin[bit]) mag *= phase; });
//
// quids::device_mag_t = cuda::std::complex<PROBA_TYPE> is the device-usable twin of quids::mag_t (identical layout; std::complex
// is built on _Complex, which device code cannot use).  Build with nvcc ... -std=c++17 --extended-lambda.
// A HOST closure cannot run in a kernel and there is no CPU fallback: passing one does not compile.
#pragma once

#include "../quids.hpp"
#include "plugin.cuh"

namespace quids {
	typedef cuda::std::complex<PROBA_TYPE> device_mag_t;

	template <class F, class = std::enable_if_t<!std::is_convertible<F, modifier_t const &>::value && !std::is_convertible<F, rule_t const *>::value>>
	void simulate(it_t &iteration, F modifier) {
#if defined(__CUDACC_EXTENDED_LAMBDA__)
		static_assert(__nv_is_extended_device_lambda_closure_type(F) || __nv_is_extended_host_device_lambda_closure_type(F) || std::is_trivially_copyable<F>::value,
		              "quids::simulate(it_t &, F): F must be a __device__ callable (an extended lambda [=] __device__ (...), or a trivially copyable functor with a __device__ operator())");
#endif
		try {
			qb::apply_device_modifier<device_mag_t>(iteration.device_handle(), modifier);
		} catch (const qb::error &e) {
			throw std::runtime_error(std::string("quids: ") + e.what());
		}
		iteration.device_modified();
	}
}
