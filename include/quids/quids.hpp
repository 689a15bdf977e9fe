// quids.hpp -- drop-in replacement for the header-only API of jolatechno/QuIDS (reference: src/quids.hpp),
// running the rule-application step on a B200 through the C ABI of quids_b200.h (libquids_b200.so).
//
// Source compatible with the reference's drivers: the same names, argument meaning and defaults
//     quids::it_t, sy_it_t, rule_t, modifier_t, observable_t, debug_t, mag_t          quids.hpp:78-90
//     quids::simulate(it_t&, modifier_t)                                               quids.hpp:436
//     quids::simulate(it_t&, rule_t const*, it_t&, sy_it_t&, size_t = 0, debug_t = {})  quids.hpp:448
//     globals tolerance, align_byte_length, safety_margin, simple_truncation, ...      quids.hpp:60-75
// What differs, because host closures and vtables cannot run on the device:
//   * a rule_t is a host HANDLE: the name of a device rule registered in the library (rule_api.cuh,
//     QB_REGISTER_RULE) plus the constructor arguments of the reference class.  The rule classes of
//     rules/quantum_computer.hpp and rules/qcgd.hpp keep their names and constructors.
//   * a modifier_t is likewise a handle (the reference's is a std::function host closure); the
//     factories cnot/Xgate/Ygate/Zgate and qcgd::step / reversed_step keep their names.  The LAMBDA path
//     (simulate(it, callable)) takes __device__ callables: include quids/device/lambda.cuh and compile the
//     driver with nvcc (extended lambdas).  A HOST lambda is not accepted: there is no CPU fallback.
//   * user-written rules and modifiers are device code too: quids/device/plugin.cuh, examples/custom_rule.cu.
//   * simple_truncation defaults to true (the probabilistic mode is not reproducible even in the
//     reference).  false selects the probabilistic truncation of the reference
//     (keep the smallest u / |mag|^2) with a counter-based generator seeded by quids::truncation_seed.
//   * states live in HBM; append/get_object/average_value work on a host mirror that is synchronised
//     lazily (uploaded before a simulate, downloaded on the first read after one).
//
// Build a driver with:  g++ -std=c++17 -I<repo>/include driver.cpp -L<repo>/quids_b200 -lquids_b200
#pragma once

typedef unsigned uint;

#include <complex>
#include <cstddef>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <limits>
#include <stdexcept>
#include <type_traits>
#include <string>
#include <vector>

#include "quids_b200.h"

#ifndef PROBA_TYPE
	#define PROBA_TYPE double
#endif
#ifndef ALIGNMENT_BYTE_LENGTH
	#define ALIGNMENT_BYTE_LENGTH 8
#endif
#ifndef TOLERANCE
	#define TOLERANCE 1e-30
#endif
#ifndef SAFETY_MARGIN
	#define SAFETY_MARGIN 0.2
#endif
#ifndef EQUALIZE_FACTOR
	#define EQUALIZE_FACTOR 0.25
#endif
#ifndef LOAD_BALANCING_BUCKET_PER_THREAD
	#define LOAD_BALANCING_BUCKET_PER_THREAD 32
#endif

namespace quids {
	// PROBA_TYPE = double (default) or float (quids.hpp:21-23).  With float the host mirror, the magnitudes crossing the C ABI
	// and every value a driver sees are complex<float>; the state in HBM and the device arithmetic stay double.
	static_assert(std::is_same<PROBA_TYPE, double>::value || std::is_same<PROBA_TYPE, float>::value, "PROBA_TYPE must be double or float");

	// ---- the mutable namespace globals drivers assign (quids.hpp:60-75) -----------------------------
	inline uint align_byte_length = ALIGNMENT_BYTE_LENGTH;
	inline PROBA_TYPE tolerance = TOLERANCE;
	inline float safety_margin = SAFETY_MARGIN;               // used by the automatic budget (max_num_object = 0)
	inline float equalize_factor = EQUALIZE_FACTOR;           // kept for source compatibility (host-memory heuristic of the reference)
	inline int load_balancing_bucket_per_thread = LOAD_BALANCING_BUCKET_PER_THREAD; // no effect: no CPU bucket partition here
	inline bool simple_truncation = true;
	inline unsigned truncation_seed = 0; // probabilistic truncation only (no counterpart in the reference, which seeds from rand())

	namespace utils { // utils/vector.hpp:29-33, kept so that drivers assigning them still compile
		inline float upsize_policy = 1.1f;
		inline float downsize_policy = 0.85f;
		inline size_t min_vector_size = 1000;
	}

	typedef std::complex<PROBA_TYPE> mag_t;
	typedef class iteration it_t;
	typedef class symbolic_iteration sy_it_t;
	typedef class rule rule_t;
	typedef class modifier modifier_t;
	typedef std::function<PROBA_TYPE(char const *object_begin, char const *object_end)> observable_t;
	typedef std::function<void(const char *step)> debug_t;

	uint inline get_alignment_offset(const uint size) { // quids.hpp:93-102
		if (align_byte_length <= 1)
			return 0;
		uint alignment_offset = align_byte_length - size % align_byte_length;
		return alignment_offset == align_byte_length ? 0 : alignment_offset;
	}

	namespace detail {
		inline void check(int status) {
			if (status != QB_OK)
				throw std::runtime_error(std::string("quids: ") + qb_last_error());
		}
		// one context (GPU, stream) per process; QUIDS_DEVICE or LOCAL_RANK selects the device
		inline qb_ctx *context() {
			static qb_ctx *ctx = [] {
				int device = 0;
				if (const char *e = std::getenv("QUIDS_DEVICE")) device = std::atoi(e);
				else if (const char *l = std::getenv("LOCAL_RANK")) device = std::atoi(l);
				qb_ctx *c = nullptr;
				check(qb_ctx_create(device, &c));
				return c;
			}();
			return ctx;
		}
		inline qb_options options() {
			qb_options o;
			qb_options_default(&o);
			o.tolerance = tolerance;
			o.align_byte_length = align_byte_length;
			o.simple_truncation = simple_truncation ? 1 : 0;
			o.safety_margin = safety_margin;
			o.seed = truncation_seed;
			return o;
		}
		inline void forward_step(const char *label, void *user) { (*static_cast<debug_t *>(user))(label); }
	}

	/// a rule: handle on a device rule registered under `name`, with the reference constructor's arguments
	class rule {
	public:
		rule(const char *name, std::initializer_list<double> params) : params_(params) {
			id_ = qb_rule_id(name);
			if (id_ < 1)
				throw std::runtime_error(std::string("quids: no device rule registered under the name ") + name);
		}
		virtual ~rule() {}
		int id() const { return id_; }
		const std::vector<double> &params() const { return params_; }

	private:
		int id_;
		std::vector<double> params_;
	};

	/// a modifier: handle on a registered device modifier (in-place, same size, quids.hpp:86,973-980)
	class modifier {
	public:
		modifier() : id_(0) {} // "no modifier" (placeholder in flags::simulator_t); applying it throws
		modifier(const char *name, std::initializer_list<double> params = {}) : params_(params) {
			id_ = qb_modifier_id(name);
			if (id_ < 1)
				throw std::runtime_error(std::string("quids: no device modifier registered under the name ") + name);
		}
		int id() const { return id_; }
		const std::vector<double> &params() const { return params_; }

	private:
		int id_;
		std::vector<double> params_;
	};

	/// an observable evaluated ON THE DEVICE: handle on a registered device observable (quids_b200.h: "qcgd_stats",
	/// "qcgd_size", "qubit", "object_bytes").  it_t::average_value accepts it next to the reference's host closure
	/// (observable_t); the state is then reduced in HBM instead of being downloaded.
	class device_observable {
	public:
		device_observable(const char *name, std::initializer_list<double> params = {}) : params_(params) {
			id_ = qb_observable_id(name);
			if (id_ < 1)
				throw std::runtime_error(std::string("quids: no device observable registered under the name ") + name);
		}
		int id() const { return id_; }
		int values() const { return qb_observable_values(id_); }
		const std::vector<double> &params() const { return params_; }

	private:
		int id_;
		std::vector<double> params_;
	};

	/// iteration (wave function), quids.hpp:149-335: state in HBM + lazily synchronised host mirror
	class iteration {
	public:
		size_t num_object = 0;
		PROBA_TYPE total_proba = 1;

		iteration() { detail::check(qb_iter_create(detail::context(), &handle_)); }
		iteration(char *object_begin_, char *object_end_) : iteration() { append(object_begin_, object_end_); }
		iteration(const iteration &) = delete; // the reference's cannot be copied either
		iteration &operator=(const iteration &) = delete;
		~iteration() { qb_iter_destroy(handle_); }

		/// quids.hpp:174-188
		void append(char const *object_begin_, char const *object_end_, mag_t const mag = 1) {
			to_host();
			const size_t size = object_end_ - object_begin_;
			const size_t offset = object_begin.back();
			objects.insert(objects.end(), object_begin_, object_end_);
			objects.resize(offset + size + get_alignment_offset(size), 0);
			magnitude.push_back(mag);
			object_size.push_back(size);
			object_begin.push_back(objects.size());
			++num_object;
			device_valid_ = false;
		}
		/// quids.hpp:194-203
		void pop(size_t n = 1, bool normalize_ = true) {
			if (n < 1)
				return;
			to_device();
			detail::check(qb_iter_pop(handle_, n, normalize_ ? 1 : 0));
			after_device_write();
		}
		/// quids.hpp:208-234 (host-side read-out: the observable is a host closure)
		PROBA_TYPE average_value(const observable_t observable) const {
			to_host();
			PROBA_TYPE avg = 0;
			for (size_t oid = 0; oid < num_object; ++oid)
				avg += observable(&objects[object_begin[oid]], &objects[object_begin[oid]] + object_size[oid]) * std::norm(magnitude[oid]);
			return avg;
		}
		/// the same sum for a device observable: one reduction kernel over the state in HBM, nothing is downloaded.
		/// Observables with several values per object fill `values` (up to 4) and return the first.
		PROBA_TYPE average_value(const device_observable &observable, double *values = nullptr) const {
			to_device();
			double out[4] = {0, 0, 0, 0};
			detail::check(qb_iter_average_value(handle_, observable.id(), observable.params().data(), (uint32_t)observable.params().size(), out, 4));
			if (values)
				for (int k = 0; k < observable.values(); ++k)
					values[k] = out[k];
			return (PROBA_TYPE)out[0];
		}
		/// read-write access, quids.hpp:242-246: the pointers alias the host mirror; the state is uploaded again before the next simulate
		void get_object(size_t const object_id, char *&object_begin_, uint &object_size_, mag_t *&mag) {
			to_host();
			device_valid_ = false;
			object_size_ = object_size[object_id];
			mag = &magnitude[object_id];
			object_begin_ = &objects[object_begin[object_id]];
		}
		/// read-only access, quids.hpp:254-258
		void get_object(size_t const object_id, char const *&object_begin_, uint &object_size_, mag_t &mag) const {
			to_host();
			object_size_ = object_size[object_id];
			mag = magnitude[object_id];
			object_begin_ = &objects[object_begin[object_id]];
		}

		/// the C-ABI handle (state uploaded first), for code that wants to call quids_b200.h directly
		qb_iter *device_handle() const {
			to_device();
			return handle_;
		}
		/// tell the mirror that device code changed the state behind device_handle() (objects / magnitudes in place, or the
		/// whole state): the public counters are refreshed and the host copy is fetched again on the next read
		void device_modified() { after_device_write(); }

	protected:
		friend void simulate(it_t &iteration, modifier_t const &rule);
		friend void simulate(it_t &iteration, rule_t const *rule, it_t &next_iteration, sy_it_t &symbolic_iteration, size_t max_num_object, debug_t mid_step_function);

		qb_iter *handle_ = nullptr;
		// host mirror, in the reference's storage layout (quids.hpp:266-276)
		mutable std::vector<char> objects;
		mutable std::vector<size_t> object_begin{0};
		mutable std::vector<uint> object_size;
		mutable std::vector<mag_t> magnitude;
		mutable bool host_valid_ = true, device_valid_ = true;

		void to_device() const {
			if (device_valid_)
				return;
			static_assert(sizeof(size_t) == sizeof(uint64_t) && sizeof(uint) == sizeof(uint32_t), "LP64 expected");
			if constexpr (std::is_same<PROBA_TYPE, float>::value)
				detail::check(qb_iter_upload_f32(handle_, num_object, reinterpret_cast<const uint8_t *>(objects.data()), object_begin[num_object],
				                                 reinterpret_cast<const uint64_t *>(object_begin.data()), object_size.data(),
				                                 reinterpret_cast<const float *>(magnitude.data()), total_proba));
			else
				detail::check(qb_iter_upload(handle_, num_object, reinterpret_cast<const uint8_t *>(objects.data()), object_begin[num_object],
				                             reinterpret_cast<const uint64_t *>(object_begin.data()), object_size.data(),
				                             reinterpret_cast<const double *>(magnitude.data()), total_proba));
			device_valid_ = true;
		}
		void to_host() const {
			if (host_valid_)
				return;
			uint64_t n = 0, bytes = 0;
			double proba = 0;
			detail::check(qb_iter_counts(handle_, &n, &bytes, &proba));
			objects.resize(bytes);
			object_begin.resize(n + 1);
			object_size.resize(n);
			magnitude.resize(n);
			if constexpr (std::is_same<PROBA_TYPE, float>::value)
				detail::check(qb_iter_download_f32(handle_, reinterpret_cast<uint8_t *>(objects.data()), reinterpret_cast<uint64_t *>(object_begin.data()),
				                                   object_size.data(), reinterpret_cast<float *>(magnitude.data())));
			else
				detail::check(qb_iter_download(handle_, reinterpret_cast<uint8_t *>(objects.data()), reinterpret_cast<uint64_t *>(object_begin.data()),
				                               object_size.data(), reinterpret_cast<double *>(magnitude.data())));
			if (n == 0)
				object_begin[0] = 0;
			host_valid_ = true;
		}
		void after_device_write() { // the device holds the truth: refresh the public counters, drop the mirror
			uint64_t n = 0, bytes = 0;
			double proba = 0;
			detail::check(qb_iter_counts(handle_, &n, &bytes, &proba));
			num_object = n;
			total_proba = proba;
			host_valid_ = false;
			device_valid_ = true;
		}
	};

	/// symbolic iteration (computation intermediary), quids.hpp:338-429: interference table and scratch, reused across calls
	class symbolic_iteration {
	public:
		symbolic_iteration() { detail::check(qb_sym_create(detail::context(), &handle_)); }
		symbolic_iteration(const symbolic_iteration &) = delete;
		symbolic_iteration &operator=(const symbolic_iteration &) = delete;
		~symbolic_iteration() { qb_sym_destroy(handle_); }

		size_t num_object = 0;
		size_t num_object_after_interferences = 0;

	protected:
		friend void simulate(it_t &iteration, rule_t const *rule, it_t &next_iteration, sy_it_t &symbolic_iteration, size_t max_num_object, debug_t mid_step_function);
		qb_sym *handle_ = nullptr;
		void refresh() {
			uint64_t a = 0, b = 0;
			detail::check(qb_sym_counts(handle_, &a, &b));
			num_object = a;
			num_object_after_interferences = b;
		}
	};

	/// apply a modifier to a wave function, quids.hpp:436-438
	void inline simulate(it_t &iteration, modifier_t const &rule) {
		if (rule.id() < 1)
			throw std::runtime_error("quids: simulate called with an empty modifier");
		iteration.to_device();
		detail::check(qb_apply_modifier(iteration.handle_, rule.id(), rule.params().data(), (uint32_t)rule.params().size()));
		iteration.host_valid_ = false;
	}

	/// apply a dynamic to a wave function, quids.hpp:448-543.
	/// max_num_object: maximum number of objects kept; -1 (SIZE_MAX) = no maximum; 0 = as many as fit in the
	/// GPU memory left after quids::safety_margin (parents first, then children, like quids.hpp:459-485,510-536).
	void inline simulate(it_t &iteration, rule_t const *rule, it_t &next_iteration, sy_it_t &symbolic_iteration, size_t max_num_object = 0,
	                     debug_t mid_step_function = [](const char *) {}) {
		iteration.to_device();
		qb_options opt = detail::options();
		detail::check(qb_simulate(iteration.handle_, rule->id(), rule->params().data(), (uint32_t)rule->params().size(), next_iteration.handle_,
		                          symbolic_iteration.handle_, max_num_object == std::numeric_limits<size_t>::max() ? QB_NO_TRUNCATION : (uint64_t)max_num_object,
		                          &opt, mid_step_function ? detail::forward_step : nullptr, &mid_step_function));
		symbolic_iteration.refresh();
		next_iteration.after_device_write();
	}
}
