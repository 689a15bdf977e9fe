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

// REGIONS (sorted order of rules whose groups are dense blocks of the object space, e.g. erase_create / coin: a child
// is its parent with any subset of the eligible nodes toggled, so the objects of one group are ALL 2^levels settings of
// the group's tree nodes).  Instead of one hashed slot per object, a DIRECTORY hashed by the group's identity gives the
// group a region of 2^levels CONSECUTIVE slots (allocated from a cursor), object s of the group at slot base + s:
// one probe per run of groups instead of one per object, and the magnitudes of a run reach the table as consecutive
// 32-byte sectors (one 4 KB burst) instead of 128 random ones.  The slots keep the same layout (key = the object's
// hash, written by the run that created the region), so compaction, selection, finalisation and the distributed
// exchange read them as before.
struct __align__(16) region_entry {
	unsigned long long key;  // identity of the group's objects (0 = free)
	unsigned long long base; // first slot + 1 (0 = the creator has not published it yet)
};

struct table_view {
	table_slot *slots;     // capacity + 1 slots
	uint64_t capacity;     // regular slots
	unsigned int *overflow; // set to 1 if an insert gave up probing
	unsigned long long *used; // number of slots created (counted per CTA)
	// region mode (dir != nullptr): `capacity` slots are handed out in regions through `cursor`
	region_entry *dir;
	uint64_t dir_capacity;
	unsigned long long *cursor; // slots handed out so far
	unsigned long long *regions; // regions created so far
};

__device__ __forceinline__ uint64_t table_home(uint64_t hash, uint64_t capacity) { return __umul64hi(mix64(hash), capacity); }

constexpr uint32_t TABLE_MAX_PROBES = 4096; // longer probe sequences mean the table was sized too small: the host retries larger

// warp-uniform: has any insert of this launch given up?  Checked at chunk boundaries and every 64 probes, so that a
// table that turns out too small costs milliseconds, not one full-length probe sequence per child.
__device__ __forceinline__ bool table_overflowed(const table_view &t) { return __shfl_sync(0xffffffffu, *(volatile unsigned int *)t.overflow, 0) != 0; }
__device__ __forceinline__ bool table_overflowed_lane(const table_view &t) { return *(volatile unsigned int *)t.overflow != 0; }

// the probe loop, entered with the key already observed in slot `i` (0 = the slot looked empty).
// Returns true when this call created the slot (first child with that hash).
__device__ __forceinline__ bool table_insert_from(const table_view &t, uint64_t hash, cplx mag, uint64_t rep, uint64_t i, unsigned long long seen) {
	table_slot *s;
	bool created = false;
	uint32_t probes = 0;
	while (true) {
		s = t.slots + i;
		// keys never change once set, so a plain L2 load that sees a key is final; only an empty slot
		// needs the compare-and-swap.  Most children of a grown state find their key present.
		if (seen == 0) {
			seen = atomicCAS(&s->key, 0ull, (unsigned long long)hash);
			if (seen == 0) { // this child created the slot: it is the representative
				s->rep = rep;
				created = true;
				break;
			}
		}
		if (seen == hash)
			break;
		if (++i == t.capacity)
			i = 0;
		if (++probes > TABLE_MAX_PROBES) {
			*t.overflow = 1;
			return false;
		}
		if ((probes & 63) == 0 && table_overflowed_lane(t))
			return false;
		seen = __ldcg(&t.slots[i].key);
	}
	// results unused -> RED.ADD.F64, fire and forget
	atomicAdd(&s->re, mag.re);
	atomicAdd(&s->im, mag.im);
	return created;
}

__device__ __forceinline__ bool table_insert_zero_hash(const table_view &t, cplx mag, uint64_t rep) {
	table_slot *s = t.slots + t.capacity;
	const bool created = atomicCAS(&s->rep, 0ull, (unsigned long long)rep) == 0;
	atomicAdd(&s->re, mag.re);
	atomicAdd(&s->im, mag.im);
	return created;
}

__device__ __forceinline__ bool table_insert(const table_view &t, uint64_t hash, cplx mag, uint64_t rep) {
	if (hash == 0)
		return table_insert_zero_hash(t, mag, rep);
	const uint64_t home = table_home(hash, t.capacity);
	// (claiming the home slot with the compare-and-swap straight away, without looking first, was measured: no gain on
	// split_merge, 15 % slower on the hadamard doubling step -- a failed or redundant CAS costs more than the load it saves)
	return table_insert_from(t, hash, mag, rep, home, __ldcg(&t.slots[home].key));
}

// Up to N inserts by one thread (entries [0, count) are valid).  The table is far larger than any
// cache, so an insert is one or more DRAM round trips: every ROUND issues the key loads of all
// pending entries first (N independent requests in flight), then resolves them -- match: add the
// magnitude; empty: try to claim the slot; other key: move to the next slot and stay pending.
// The number of round trips is the LONGEST probe sequence of the batch, not the sum.
// mag_of(i) is only called when entry i is resolved (keeps the magnitudes out of registers).
template <int N, class MagOf, class RepOf>
__device__ __forceinline__ uint32_t table_insert_batch(const table_view &t, int count, const uint64_t (&hash)[N], MagOf mag_of, RepOf rep_of) {
	uint64_t slot[N];
	uint32_t pending = 0, created = 0;
#pragma unroll
	for (int i = 0; i < N; ++i)
		if (i < count) {
			if (hash[i] == 0) {
				created += table_insert_zero_hash(t, mag_of(i), rep_of(i));
			} else {
				slot[i] = table_home(hash[i], t.capacity);
				pending |= 1u << i;
			}
		}
	for (uint32_t round = 0; pending; ++round) {
		unsigned long long seen[N];
#pragma unroll
		for (int i = 0; i < N; ++i)
			if (pending & (1u << i))
				seen[i] = __ldcg(&t.slots[slot[i]].key);
#pragma unroll
		for (int i = 0; i < N; ++i)
			if (pending & (1u << i)) {
				table_slot *s = t.slots + slot[i];
				if (seen[i] == 0) {
					seen[i] = atomicCAS(&s->key, 0ull, (unsigned long long)hash[i]);
					if (seen[i] == 0) { // this child created the slot: it is the representative
						s->rep = rep_of(i);
						++created;
						seen[i] = hash[i];
					}
				}
				if (seen[i] == hash[i]) {
					const cplx mag = mag_of(i);
					atomicAdd(&s->re, mag.re); // results unused -> RED.ADD.F64, fire and forget
					atomicAdd(&s->im, mag.im);
					pending &= ~(1u << i);
				} else if (++slot[i] == t.capacity) {
					slot[i] = 0;
				}
			}
		if (round > TABLE_MAX_PROBES) {
			*t.overflow = 1;
			break;
		}
		if ((round & 63) == 63 && table_overflowed_lane(t))
			break;
	}
	return created;
}

// Slots are handed out to the warps in chunks (one atomic on the shared cursor per REGION_CHUNK slots instead of one per
// region: every creator hitting the same address was the top stall of the kernel); a warp fills its chunk front to back.
//
// The slots of region mode are NOT cleared before the kernel (round 1 did: a memset of the whole table, then read-modify-write
// RED.F64 into the zeroed sectors = 32 B written + 32 B fetched + 32 B written back per slot).  The run that CREATES a region
// writes its slots whole -- hash, summed magnitude, representative: full 32-byte sectors, no fetch -- and only then publishes
// the region's base in the directory (region_publish); a later run of the same objects waits for the base and adds with
// RED.F64.  What a warp leaves unused of its chunk (the tail that cannot hold the next region, the rest of its last chunk)
// is zeroed by that warp (region_retire), so that every slot below the cursor is either a written object or empty.
constexpr uint32_t REGION_CHUNK = 1024;
struct region_chunk {
	unsigned long long next, end; // this warp's private range of slots
};
struct region_grant {
	unsigned long long base;       // first slot of the region, ~0 = the table is too small (overflow raised)
	region_entry *entry;           // created: the directory entry to publish once the slots are written
	unsigned long long retire_from, retire_count; // slots of the warp's previous chunk to zero (region_retire)
	bool created;
};

__device__ __forceinline__ unsigned long long load_acquire(const unsigned long long *p) {
	unsigned long long v;
	asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}

// The region of the objects identified by `key` (`leaves` consecutive slots).  ONE lane calls this per run.
__device__ __forceinline__ region_grant region_acquire(const table_view &t, region_chunk &mine, uint64_t key, uint32_t leaves) {
	region_grant g{~0ull, nullptr, 0, 0, false};
	if (key == 0)
		key = 1;
	uint64_t i = __umul64hi(mix64(key), t.dir_capacity);
	for (uint32_t probes = 0; probes <= TABLE_MAX_PROBES; ++probes) {
		region_entry *e = t.dir + i;
		unsigned long long seen = __ldcg(&e->key);
		if (seen == 0) {
			seen = atomicCAS(&e->key, 0ull, (unsigned long long)key);
			if (seen == 0) { // this run creates the region
				if (mine.next + leaves > mine.end) { // what is left of the old chunk stays empty
					g.retire_from = mine.next;
					g.retire_count = mine.end - mine.next;
					const unsigned long long want = leaves > REGION_CHUNK ? leaves : REGION_CHUNK;
					mine.next = atomicAdd(t.cursor, want);
					mine.end = mine.next + want;
					if (mine.end > t.capacity) { // the new chunk does not fit: nothing of it may be touched
						mine.end = mine.next;
						*t.overflow = 1;
						atomicExch(&e->base, ~0ull); // whoever waits for this region gives up too
						return g;
					}
				}
				g.base = mine.next;
				mine.next += leaves;
				g.entry = e;
				g.created = true;
				return g;
			}
		}
		if (seen == key) {
			unsigned long long base;
			// published (release) once the creator has written the slots; the acquire load orders this run's additions after them
			while ((base = load_acquire(&e->base)) == 0)
				if (table_overflowed_lane(t))
					return g;
			g.base = base == ~0ull ? ~0ull : base - 1;
			return g;
		}
		if (++i == t.dir_capacity)
			i = 0;
		if ((probes & 63) == 63 && table_overflowed_lane(t))
			return g;
	}
	*t.overflow = 1;
	return g;
}

// the creator's slots are written: let the other runs of the same objects in (one lane, after a __syncwarp of the writers)
__device__ __forceinline__ void region_publish(const region_grant &g) {
	// st.release.gpu: the warp's slot writes (ordered before this lane by __syncwarp) become visible before the base does
	asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(&g.entry->base), "l"(g.base + 1) : "memory");
}

// the directory entry a run will probe when it is flushed, pulled into L2 when the run OPENS (the probe is a dependent DRAM
// round trip of one lane while 31 wait: 5.5 % of the kernel's stall samples, ncu profiles/r2c)
__device__ __forceinline__ void region_prefetch(const table_view &t, uint64_t key) {
	if (key == 0)
		key = 1;
	asm volatile("prefetch.global.L2 [%0];" ::"l"(t.dir + __umul64hi(mix64(key), t.dir_capacity)));
}

// zero slots [from, from + count) (whole warp): the unused part of a chunk
__device__ __forceinline__ void region_retire(const table_view &t, unsigned long long from, unsigned long long count) {
	const unsigned long long end = from + count < t.capacity ? from + count : t.capacity;
	for (unsigned long long i = from + lane_id(); i < end; i += 32) {
		ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(t.slots + i);
		dst[0] = make_ulonglong2(0, 0);
		dst[1] = make_ulonglong2(0, 0);
	}
}

__device__ __forceinline__ bool slot_occupied(const table_slot &s, bool is_zero_slot) { return is_zero_slot ? s.rep != 0 : s.key != 0; }

// ---- BINNED inserts (rules whose children come in no useful order: split_merge, hadamard, any rule written with the four
// reference methods only).  A table larger than L2 hit at random costs one DRAM round trip per probe and one read-modify-write
// of a 32-byte sector per child: 60-70 ps per insert, latency bound (profiles/table_bench_r1.txt, the kernel sat at 16 % of
// DRAM throughput).  The reference's answer is to partition the children by hash prefix so that every bucket's map is
// cache resident (utils/algorithm.hpp:170-227, quids.hpp:755-809); this is its counterpart:
//   pass 1  the child-generation kernel does not touch the table: a child's (hash, magnitude, representative) record goes to
//           the bin of its table REGION -- bin = mulhi(mix64(hash), bins), the top bits of the very mix that places it in the
//           table (table_home grows with mix64(hash)), so bin b holds exactly the children of slots [b, b + 1) * capacity / bins.
//           One L2 atomic on the bin's cursor (cursors are 32 bytes apart) and one 32-byte store per child, fire and forget:
//           the kernel is bound by the rule's own arithmetic, not by table round trips;
//   pass 2  bin_insert_kernel streams the bins IN ORDER with all CTAs: at any moment the slots being touched are a few
//           consecutive regions, tens of MB, L2 resident -- each slot's sector comes from DRAM once and goes back once.
// A bin that is full (equal hashes concentrate: every record of a hash lands in the same bin) sends its record straight to
// the table, as before: the bins are an ordering device, the table's semantics are untouched.
struct __align__(32) bin_record {
	unsigned long long hash;
	double re, im;
	unsigned long long rep;
};
constexpr uint32_t BIN_CURSOR_STRIDE = 4; // u64 words between two cursors: one 32-byte sector each

struct bin_view {
	bin_record *records;        // bins * bin_capacity records; nullptr = no binning
	unsigned long long *cursor; // cursor[bin * BIN_CURSOR_STRIDE] = records sent to the bin so far (may exceed bin_capacity)
	uint32_t bins;
	uint32_t bin_capacity;
};

__device__ __forceinline__ bool bin_emit(const bin_view &b, const table_view &t, uint64_t hash, cplx mag, uint64_t rep) {
	if (hash == 0)
		return table_insert_zero_hash(t, mag, rep);
	const uint32_t bin = (uint32_t)__umul64hi(mix64(hash), (uint64_t)b.bins);
	const unsigned long long at = atomicAdd(&b.cursor[(size_t)bin * BIN_CURSOR_STRIDE], 1ull);
	if (at < b.bin_capacity) {
		ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(b.records + (size_t)bin * b.bin_capacity + at);
		__stcs(dst, make_ulonglong2(hash, (unsigned long long)__double_as_longlong(mag.re)));
		__stcs(dst + 1, make_ulonglong2((unsigned long long)__double_as_longlong(mag.im), rep));
		return false;
	}
	return table_insert(t, hash, mag, rep); // the bin is full
}

constexpr int BIN_INSERT_THREADS = 256;
constexpr int BIN_INSERT_BATCH = 2; // records per thread and round: their key loads go out together

// all CTAs walk the bins in order, BIN_INSERT_THREADS * BIN_INSERT_BATCH consecutive records per CTA and round
static __global__ void __launch_bounds__(BIN_INSERT_THREADS) bin_insert_kernel(bin_view b, table_view t) {
	constexpr int N = BIN_INSERT_BATCH;
	constexpr uint32_t TILE = BIN_INSERT_THREADS * N;
	const uint32_t tiles_per_bin = (b.bin_capacity + TILE - 1) / TILE;
	const uint64_t tiles = (uint64_t)b.bins * tiles_per_bin;
	uint32_t created = 0;
	for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
		const uint32_t bin = (uint32_t)(tile / tiles_per_bin);
		const uint32_t first = (uint32_t)(tile % tiles_per_bin) * TILE;
		const unsigned long long sent = __ldcg(&b.cursor[(size_t)bin * BIN_CURSOR_STRIDE]);
		const uint32_t filled = sent < b.bin_capacity ? (uint32_t)sent : b.bin_capacity;
		if (first >= filled)
			continue;
		if (table_overflowed_lane(t))
			break;
		const bin_record *records = b.records + (size_t)bin * b.bin_capacity;
		uint64_t hash[N];
		cplx mag[N];
		uint64_t rep[N];
		int count = 0;
#pragma unroll
		for (int q = 0; q < N; ++q) {
			const uint32_t i = first + q * BIN_INSERT_THREADS + threadIdx.x;
			hash[q] = 0;
			if (i < filled) {
				const ulonglong2 lo = __ldcs(reinterpret_cast<const ulonglong2 *>(records + i));
				const ulonglong2 hi = __ldcs(reinterpret_cast<const ulonglong2 *>(records + i) + 1);
				hash[q] = lo.x;
				mag[q] = cplx{__longlong_as_double((long long)lo.y), __longlong_as_double((long long)hi.x)};
				rep[q] = hi.y;
				count = q + 1;
			}
		}
		// entries past `filled` keep hash 0 with a zero magnitude: they must not reach the dedicated slot of the hash 0
		// (records with hash 0 never enter a bin, bin_emit), so the batch skips them
		uint64_t slot[N];
		uint32_t pending = 0;
#pragma unroll
		for (int q = 0; q < N; ++q)
			if (q < count && hash[q] != 0) {
				slot[q] = table_home(hash[q], t.capacity);
				pending |= 1u << q;
			}
		for (uint32_t round = 0; pending; ++round) {
			unsigned long long seen[N];
#pragma unroll
			for (int q = 0; q < N; ++q)
				if (pending & (1u << q))
					seen[q] = __ldcg(&t.slots[slot[q]].key);
#pragma unroll
			for (int q = 0; q < N; ++q)
				if (pending & (1u << q)) {
					table_slot *s = t.slots + slot[q];
					if (seen[q] == 0) {
						seen[q] = atomicCAS(&s->key, 0ull, (unsigned long long)hash[q]);
						if (seen[q] == 0) {
							s->rep = rep[q];
							++created;
							seen[q] = hash[q];
						}
					}
					if (seen[q] == hash[q]) {
						atomicAdd(&s->re, mag[q].re);
						atomicAdd(&s->im, mag[q].im);
						pending &= ~(1u << q);
					} else if (++slot[q] == t.capacity) {
						slot[q] = 0;
					}
				}
			if (round > TABLE_MAX_PROBES) {
				*t.overflow = 1;
				break;
			}
			if ((round & 63) == 63 && table_overflowed_lane(t))
				break;
		}
	}
	created = (uint32_t)warp_sum((uint64_t)created);
	if (lane_id() == 0 && created)
		atomicAdd(t.used, (unsigned long long)created);
}

} // namespace qb
