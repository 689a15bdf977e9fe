// table.cuh -- the interference table: hash-keyed deduplication that merges magnitudes.
//
// Replaces the bucket partition + per-bucket robin_hood::unordered_map<size_t,size_t> of
// symbolic_iteration::compute_collisions (quids.hpp:726-824): children with equal 64-bit hash are
// ONE object whose magnitude is the complex sum of the group; one representative (here: the child
// whose insert created the slot) says how to rebuild the object.  As in the reference, object
// bytes are never compared (quids.hpp:799).
//
// Layout: open addressing, linear probing, one 32-byte slot = one DRAM sector:
//     { u64 key (the hash; 0 = empty) | f64 re | f64 im | u64 rep }
// rep = ((child index + 1) << 24) | child size, so that a set rep is never 0.  A child whose hash
// is 0 goes to the dedicated slot `capacity` (its occupancy is rep != 0).
#pragma once

#include "common.cuh"

namespace qb {

struct __align__(32) table_slot {
	unsigned long long key;
	double re, im;
	unsigned long long rep;
};

constexpr int REP_SIZE_BITS = 24;
constexpr uint64_t REP_MAX_INDEX = (1ull << (64 - REP_SIZE_BITS)) - 2;
constexpr uint32_t REP_MAX_SIZE = (1u << REP_SIZE_BITS) - 1;

__host__ __device__ __forceinline__ uint64_t rep_pack(uint64_t child_index, uint32_t size) { return ((child_index + 1) << REP_SIZE_BITS) | size; }
__host__ __device__ __forceinline__ uint64_t rep_index(uint64_t rep) { return (rep >> REP_SIZE_BITS) - 1; }
__host__ __device__ __forceinline__ uint32_t rep_size(uint64_t rep) { return (uint32_t)(rep & REP_MAX_SIZE); }

struct table_view {
	table_slot *slots;     // capacity + 1 slots
	uint64_t capacity;     // regular slots
	unsigned int *overflow; // set to 1 if an insert found no free slot
};

__device__ __forceinline__ uint64_t table_home(uint64_t hash, uint64_t capacity) { return __umul64hi(mix64(hash), capacity); }

__device__ __forceinline__ void table_insert(const table_view &t, uint64_t hash, cplx mag, uint64_t rep) {
	table_slot *s;
	if (hash == 0) {
		s = t.slots + t.capacity;
		atomicCAS(&s->rep, 0ull, (unsigned long long)rep);
	} else {
		uint64_t i = table_home(hash, t.capacity);
		uint64_t probes = 0;
		while (true) {
			s = t.slots + i;
			unsigned long long seen = atomicCAS(&s->key, 0ull, (unsigned long long)hash);
			if (seen == 0) { // this child created the slot: it is the representative
				s->rep = rep;
				break;
			}
			if (seen == hash)
				break;
			if (++i == t.capacity)
				i = 0;
			if (++probes > t.capacity) {
				*t.overflow = 1;
				return;
			}
		}
	}
	// results unused -> RED.ADD.F64, fire and forget
	atomicAdd(&s->re, mag.re);
	atomicAdd(&s->im, mag.im);
}

__device__ __forceinline__ bool slot_occupied(const table_slot &s, bool is_zero_slot) { return is_zero_slot ? s.rep != 0 : s.key != 0; }

} // namespace qb
