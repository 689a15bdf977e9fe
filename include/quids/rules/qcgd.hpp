// rules/qcgd.hpp -- drop-in for the reference's src/rules/qcgd.hpp (Quantum Causal Graph Dynamics).
//
// Same names: graphs::{num_nodes,left,right,node_name_begin,node_name,hash_graph,randomize}, the rules
// erase_create / coin / split_merge(theta, phi, xi), the modifiers step / reversed_step, the host
// utilities utils::{make_graph,randomize,print,serialize} and the flag parser flags::*.
// The rules and modifiers are handles on device code (quids_b200/csrc/rules_qcgd.cuh); everything
// else here is host-side convenience working on the host mirror of a state, written for this
// repository (the reference's versions live at the line numbers cited below).
//
// Object layout (qcgd.hpp:63-112), n nodes:
//     u16 n | bool left[n] | bool right[n] | u16 name_begin[n+1] | sub_node names[name_begin[n]]
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <ctime>
#include <iomanip>
#include <iostream>
#include <numeric>
#include <string>
#include <tuple>
#include <vector>

#include "../quids.hpp"

namespace quids::rules::qcgd {
	namespace utils {
		/// qcgd.hpp:11-25
		inline void hash_combine(std::size_t &seed, size_t const value_64) {
			const size_t mul = 0xc6a4a7935bd1e995;
			seed = (seed * mul ^ (value_64 >> 47)) * mul;
			seed = (seed ^ value_64) * mul + 0xe6546b64;
		}
	}

	namespace graphs {
		enum { dot_l_t = -3, dot_r_t, element_t, pair_t }; // qcgd.hpp:29-34

		/// one node of a name tree in prefix order (qcgd.hpp:35-61); 16 bytes, 4 of them padding
		struct sub_node {
			int16_t hmlz_and_element; // < 0: the most-left element of the subtree is 0
			int16_t right_or_type;    // dot_l_t / dot_r_t / element_t, or the offset to the right subtree of a pair
			size_t hash;
		};
		static_assert(sizeof(sub_node) == 16, "sub_node layout");

		inline uint16_t num_nodes(char const *object) {
			uint16_t n;
			std::memcpy(&n, object, sizeof n);
			return n;
		}
		inline bool *left(char *object) { return reinterpret_cast<bool *>(object + 2); }
		inline bool const *left(char const *object) { return reinterpret_cast<bool const *>(object + 2); }
		inline bool *right(char *object) { return reinterpret_cast<bool *>(object + 2 + num_nodes(object)); }
		inline bool const *right(char const *object) { return reinterpret_cast<bool const *>(object + 2 + num_nodes(object)); }
		inline uint16_t const *node_name_begin(char const *object) { return reinterpret_cast<uint16_t const *>(object + 2 + 2 * num_nodes(object)); }
		inline sub_node const *node_name(char const *object) { return reinterpret_cast<sub_node const *>(object + 4 + 4 * num_nodes(object)); }
		inline bool left(char const *object, int node) { return left(object)[node]; }
		inline bool right(char const *object, int node) { return right(object)[node]; }

		/// left/right particles drawn with rand() & 1, node by node (qcgd.hpp:114-120)
		inline void randomize(char *object) {
			const uint16_t n = num_nodes(object);
			for (int i = 0; i < n; ++i) {
				left(object)[i] = rand() & 1;
				right(object)[i] = rand() & 1;
			}
		}

		/// the hash the QCGD rules deduplicate on (qcgd.hpp:122-146); the device computes the same value
		inline size_t hash_graph(char const *object) {
			size_t hl = 0, hr = 0, hn = 0;
			const uint16_t n = num_nodes(object);
			for (int i = 0; i < n; ++i) {
				if (left(object, i)) utils::hash_combine(hl, i);
				if (right(object, i)) utils::hash_combine(hr, i);
				sub_node first;
				std::memcpy(&first, node_name(object) + node_name_begin(object)[i], sizeof first);
				utils::hash_combine(hn, first.hash);
			}
			utils::hash_combine(hn, hl);
			utils::hash_combine(hn, hr);
			return hn;
		}
	}

	namespace utils {
		inline size_t max_print_num_graphs = -1;

		/// a fresh graph: `size` nodes, no particles, node i named by the element i (qcgd.hpp:214-230)
		inline void make_graph(char *&object_begin, char *&object_end, uint16_t size) {
			const size_t bytes = 4 + 20 * (size_t)size;
			object_begin = new char[bytes]();
			object_end = object_begin + bytes;
			std::memcpy(object_begin, &size, 2);
			uint16_t *name_begin = reinterpret_cast<uint16_t *>(object_begin + 2 + 2 * size);
			graphs::sub_node *names = reinterpret_cast<graphs::sub_node *>(object_begin + 4 + 4 * size);
			for (uint16_t i = 0; i <= size; ++i)
				name_begin[i] = i;
			for (uint16_t i = 0; i < size; ++i) {
				names[i].hmlz_and_element = i == 0 ? -1 : i + 1;
				names[i].right_or_type = graphs::element_t;
				names[i].hash = i;
			}
		}

		/// qcgd.hpp:232-240
		inline void randomize(quids::it_t &iter) {
			for (size_t gid = 0; gid < iter.num_object; ++gid) {
				uint size;
				mag_t *mag;
				char *begin;
				iter.get_object(gid, begin, size, mag);
				graphs::randomize(begin);
			}
		}

		namespace detail {
			inline void print_name(std::ostream &out, graphs::sub_node const *node) {
				graphs::sub_node n;
				std::memcpy(&n, node, sizeof n);
				if (n.right_or_type == graphs::element_t) {
					out << std::abs(n.hmlz_and_element) - 1;
				} else if (n.right_or_type == graphs::dot_l_t || n.right_or_type == graphs::dot_r_t) {
					out << "(";
					print_name(out, node + 1);
					out << (n.right_or_type == graphs::dot_l_t ? ").l" : ").r");
				} else {
					out << "(";
					print_name(out, node + 1);
					out << ")∨(";
					print_name(out, node + n.right_or_type);
					out << ")";
				}
			}
		}

		/// one line per graph, most probable first, in the reference's format (qcgd.hpp:242-307)
		inline void print(quids::it_t const &iter, std::ostream &out = std::cout) {
			std::vector<size_t> order(iter.num_object);
			std::iota(order.begin(), order.end(), 0);
			std::vector<PROBA_TYPE> proba(iter.num_object);
			for (size_t gid = 0; gid < iter.num_object; ++gid) {
				uint size;
				char const *begin;
				mag_t mag;
				iter.get_object(gid, begin, size, mag);
				proba[gid] = std::norm(mag);
			}
			std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return proba[a] > proba[b]; });

			const size_t shown = std::min(iter.num_object, max_print_num_graphs);
			for (size_t k = 0; k < shown; ++k) {
				uint size;
				char const *begin;
				mag_t mag;
				iter.get_object(order[k], begin, size, mag);
				const PROBA_TYPE re = std::abs(mag.real()) < quids::tolerance ? 0 : mag.real();
				const PROBA_TYPE im = std::abs(mag.imag()) < quids::tolerance ? 0 : mag.imag();
				out << std::fixed << std::setprecision(5) << "\t" << re << (im <= -1e-5 ? " - " : " + ") << std::abs(im) << "i  ";
				const uint16_t n = graphs::num_nodes(begin);
				for (int i = 0; i < n; ++i) {
					out << "-|" << (graphs::left(begin, i) ? "<" : " ") << "|";
					detail::print_name(out, graphs::node_name(begin) + graphs::node_name_begin(begin)[i]);
					out << "|" << (graphs::right(begin, i) ? ">" : " ") << "|-";
				}
				out << "\n";
			}
			if (shown < iter.num_object)
				out << "\t...and " << iter.num_object - shown << " other graphs\n";
		}

		/// summary statistics as JSON (qcgd.hpp:309-366), including the reference's std_dev_size quirk (:346)
		inline void serialize(quids::it_t const &iter, quids::sy_it_t const &sy_it, uint indentation = 0, std::ostream &out = std::cout) {
			PROBA_TYPE interference_ratio = 1, deletion_ratio = 1;
			if (sy_it.num_object > 0) {
				interference_ratio = (PROBA_TYPE)sy_it.num_object_after_interferences / (PROBA_TYPE)sy_it.num_object;
				deletion_ratio = (PROBA_TYPE)iter.num_object / (PROBA_TYPE)sy_it.num_object_after_interferences;
			}
			auto density = [](char const *b) {
				const PROBA_TYPE n = graphs::num_nodes(b);
				PROBA_TYPE d = 0;
				for (int i = 0; i < n; ++i)
					d += graphs::left(b, i) + graphs::right(b, i);
				return d / (2 * n);
			};
			PROBA_TYPE avg_size, avg_size2, avg_density, avg_density2;
			if constexpr (std::is_same<PROBA_TYPE, double>::value) {
				// the four averages in ONE reduction over the state in HBM (device observable "qcgd_stats"); nothing is downloaded
				static const quids::device_observable stats("qcgd_stats");
				double v[4];
				iter.average_value(stats, v);
				avg_size = v[0], avg_size2 = v[1], avg_density = v[2], avg_density2 = v[3];
			} else { // float build: accumulate in PROBA_TYPE on the host mirror, like the reference
				avg_size = iter.average_value([](char const *b, char const *) { return (PROBA_TYPE)graphs::num_nodes(b); });
				avg_size2 = iter.average_value([](char const *b, char const *) { return (PROBA_TYPE)graphs::num_nodes(b) * graphs::num_nodes(b); });
				avg_density = iter.average_value([&](char const *b, char const *) { return density(b); });
				avg_density2 = iter.average_value([&](char const *b, char const *) { return density(b) * density(b); });
			}
			PROBA_TYPE std_dev_size = avg_size2 - avg_size * avg_size;
			std_dev_size = std_dev_size < quids::tolerance ? 0 : std::sqrt(avg_size2);
			PROBA_TYPE std_dev_density = avg_density2 - avg_density * avg_density;
			std_dev_density = std_dev_density < quids::tolerance ? 0 : std::sqrt(std_dev_density);

			const std::string tab(indentation, '\t');
			out << "{\n"
			    << tab << "\t\"total_proba\" : " << iter.total_proba << ",\n"
			    << tab << "\t\"num_graphs\" : " << iter.num_object << ",\n"
			    << tab << "\t\"avg_size\" : " << avg_size << ",\n"
			    << tab << "\t\"std_dev_size\" : " << std_dev_size << ",\n"
			    << tab << "\t\"avg_density\" : " << avg_density << ",\n"
			    << tab << "\t\"std_dev_density\" : " << std_dev_density << ",\n"
			    << tab << "\t\"interference_ratio\" : " << interference_ratio << ",\n"
			    << tab << "\t\"deletion_ratio\" : " << deletion_ratio << "\n"
			    << tab << "}";
		}
	}

	/// modifiers (qcgd.hpp:443-457): particles move one node along their direction / back
	inline const modifier_t step("step");
	inline const modifier_t reversed_step("reversed_step");

	/// rules (qcgd.hpp:459-1036): theta mixes "do" and "do not", phi and xi are their phases
	class erase_create : public quids::rule {
	public:
		erase_create(PROBA_TYPE theta, PROBA_TYPE phi = 0, PROBA_TYPE xi = 0) : quids::rule("erase_create", {theta, phi, xi}) {}
	};
	class coin : public quids::rule {
	public:
		coin(PROBA_TYPE theta, PROBA_TYPE phi = 0, PROBA_TYPE xi = 0) : quids::rule("coin", {theta, phi, xi}) {}
	};
	class split_merge : public quids::rule {
	public:
		split_merge(PROBA_TYPE theta, PROBA_TYPE phi = 0, PROBA_TYPE xi = 0) : quids::rule("split_merge", {theta, phi, xi}) {}
	};

	/// string flags of the production driver (qcgd.hpp:1038-1177):
	///   "<n_iter>[,key=value...] | <n_node>[,n_graphs=][,real=][,imag=][;...] | <rule>[,theta=][,phi=][,xi=][,n_iter=][;...]"
	namespace flags {
		/// (iterations, is a rule, modifier, rule, reversed modifier, reversed rule)
		typedef std::vector<std::tuple<int, bool, quids::modifier_t, quids::rule_t *, quids::modifier_t, quids::rule_t *>> simulator_t;

		namespace detail {
			/// removes and returns the text before the first `separator` (everything if there is none)
			inline std::string take(std::string &input, const std::string &separator) {
				const size_t end = input.find(separator);
				std::string head = input.substr(0, end);
				input = end == std::string::npos ? "" : input.substr(end + separator.size());
				return head;
			}
			/// value of "key" in a comma separated list, "" if absent
			inline std::string value_of(const std::string &input, const std::string &key) {
				size_t at = input.find(key);
				if (at == std::string::npos)
					return "";
				at += key.size();
				return input.substr(at, input.find(',', at) - at);
			}
			inline int int_or(const std::string &input, const std::string &key, int fallback) {
				const std::string v = value_of(input, key);
				return v.empty() ? fallback : std::atoi(v.c_str());
			}
			inline float float_or(const std::string &input, const std::string &key, float fallback) {
				const std::string v = value_of(input, key);
				return v.empty() ? fallback : (float)std::atof(v.c_str());
			}
		}

		/// (n_iter, reversed_n_iter, max_num_object); also sets the globals the reference's parser sets (qcgd.hpp:1086-1117)
		inline std::tuple<uint, uint, size_t> read_n_iter(const char *argv) {
			std::string args = argv;
			const int n_iters = std::atoi(detail::take(args, ",").c_str());
			const std::string seed = detail::value_of(args, "seed=");
			std::srand(seed.empty() ? (unsigned)std::time(0) : (unsigned)std::atoi(seed.c_str()));
			const int reversed_n_iters = detail::int_or(args, "reversed_n_iter=", 0);
			utils::max_print_num_graphs = detail::int_or(args, "max_print_num_graphs=", (int)utils::max_print_num_graphs);
			quids::tolerance = detail::float_or(args, "tolerance=", quids::tolerance);
			quids::safety_margin = detail::float_or(args, "safety_margin=", quids::safety_margin);
			quids::align_byte_length = detail::int_or(args, "align=", quids::align_byte_length);
			quids::simple_truncation = detail::int_or(args, "simple_truncate=", quids::simple_truncation);
			quids::load_balancing_bucket_per_thread = detail::int_or(args, "load_balancing_bucket_per_thread=", quids::load_balancing_bucket_per_thread);
			const size_t max_num_object = detail::int_or(args, "max_num_object=", 0); // -1 -> SIZE_MAX = no truncation, 0 = automatic
			return {n_iters, reversed_n_iters, max_num_object};
		}

		/// appends n_graphs fresh n_node graphs per ';' entry, magnitude (real, imag)/sqrt(n_graphs) in float arithmetic,
		/// then randomises every graph of the state (qcgd.hpp:1119-1137)
		inline void read_state(const char *argv, quids::it_t &state) {
			std::string args = argv;
			for (std::string entry; !(entry = detail::take(args, ";")).empty();) {
				const int n_node = std::atoi(detail::take(entry, ",").c_str());
				const int n_graphs = detail::int_or(entry, "n_graphs=", 1);
				const float real = detail::float_or(entry, "real=", 1) / std::sqrt((float)n_graphs);
				const float imag = detail::float_or(entry, "imag=", 0) / std::sqrt((float)n_graphs);
				for (int i = 0; i < n_graphs; ++i) {
					char *begin, *end;
					utils::make_graph(begin, end, n_node);
					state.append(begin, end, {real, imag});
					delete[] begin;
				}
			}
			utils::randomize(state);
		}

		/// theta, phi, xi are given in units of pi (qcgd.hpp:1139-1166)
		inline simulator_t read_rule(const char *argv, debug_t = [](const char *) {}) {
			std::string args = argv;
			simulator_t simulator;
			const modifier_t none;
			for (std::string entry; !(entry = detail::take(args, ";")).empty();) {
				const std::string name = detail::take(entry, ",");
				const float theta = M_PI * detail::float_or(entry, "theta=", 0.25);
				const float phi = M_PI * detail::float_or(entry, "phi=", 0);
				const float xi = M_PI * detail::float_or(entry, "xi=", 0);
				const int n_iter = detail::int_or(entry, "n_iter=", 1);
				if (name == "split_merge")
					simulator.push_back({n_iter, true, none, new split_merge(theta, phi, xi), none, new split_merge(theta, phi, -xi)});
				else if (name == "erase_create")
					simulator.push_back({n_iter, true, none, new erase_create(theta, phi, xi), none, new erase_create(theta, phi, -xi)});
				else if (name == "coin")
					simulator.push_back({n_iter, true, none, new coin(theta, phi, xi), none, new coin(theta, phi, -xi)});
				else if (name == "step")
					simulator.push_back({n_iter, false, step, nullptr, reversed_step, nullptr});
				else if (name == "reversed_step")
					simulator.push_back({n_iter, false, reversed_step, nullptr, step, nullptr});
			}
			return simulator;
		}

		/// (n_iter, reversed_n_iter, simulator, max_num_object) from "iterations | state | rules" (qcgd.hpp:1168-1176)
		inline std::tuple<uint, uint, simulator_t, size_t> parse_simulation(const char *argv, it_t &state, debug_t mid_step_function = [](const char *) {}) {
			std::string args = argv;
			auto [n_iter, reversed_n_iters, max_num_object] = read_n_iter(detail::take(args, "|").c_str());
			read_state(detail::take(args, "|").c_str(), state);
			return {n_iter, reversed_n_iters, read_rule(args.c_str(), mid_step_function), max_num_object};
		}
	}
}
