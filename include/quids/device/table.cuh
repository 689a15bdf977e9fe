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

// The region of the objects identified by `key` (`leaves` consecutive slots), accumulating kernel: ONE lane calls
// region_acquire per run.  (Two halves -- region_probe_begin claims the home entry of the directory blindly, without looking at
// the result of the compare-and-swap, region_acquire_finish reads it -- so that a caller may put work between them; the batch
// kernel has its own probe, one lane per run of a 32-item batch, rules_qcgd.cuh region_batch.)
struct region_probe {
	uint64_t key, index;
	unsigned long long seen; // what the home entry held before the compare-and-swap (0 = this run claimed it)
};
__device__ __forceinline__ region_probe region_probe_begin(const table_view &t, uint64_t key) {
	region_probe p;
	p.key = key ? key : 1;
	p.index = __umul64hi(mix64(p.key), t.dir_capacity);
	p.seen = atomicCAS(&t.dir[p.index].key, 0ull, (unsigned long long)p.key);
	return p;
}

__device__ __forceinline__ region_grant region_acquire_finish(const table_view &t, region_chunk &mine, const region_probe &p, uint32_t leaves) {
	region_grant g{~0ull, nullptr, 0, 0, false};
	const uint64_t key = p.key;
	uint64_t i = p.index;
	unsigned long long seen = p.seen;
	for (uint32_t probes = 0; probes <= TABLE_MAX_PROBES; ++probes) {
		region_entry *e = t.dir + i;
		if (probes) {
			seen = __ldcg(&e->key);
			if (seen == 0)
				seen = atomicCAS(&e->key, 0ull, (unsigned long long)key);
		}
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

__device__ __forceinline__ region_grant region_acquire(const table_view &t, region_chunk &mine, uint64_t key, uint32_t leaves) {
	return region_acquire_finish(t, mine, region_probe_begin(t, key), leaves);
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

// ---- BINNED interference (rules whose children come in no useful order: split_merge, hadamard, any rule written with the
// four reference methods only).  A table larger than L2 hit at random costs one DRAM round trip per probe and a read-modify-write
// of a 32-byte sector per child: 60-70 ps per insert, latency bound.  The reference's answer is to partition the children by
// hash prefix so that every bucket's map is cache resident (utils/algorithm.hpp:170-227, quids.hpp:755-809).  Two GPU
// counterparts were built and measured (profiles/r2_*):
//   A  bins = regions of the global table, inserted region by region so that the slots in flight stay in L2: no gain when
//      most children are unique -- a FIRST touch of a slot costs the same DRAM round trip in any order (dropped);
//   B  (this one) bins small enough that a bin's interference table fits in SHARED memory:
//   pass 1  the child-generation kernel does not touch any table: a child's (hash, magnitude, representative) record goes to
//           bin = mulhi(mix64(hash), bins): one L2 atomic on the bin's cursor and one 32-byte store, fire and forget -- the
//           kernel is bound by the rule's own arithmetic.  ~770 records per bin; the bins' open cache lines (one per bin) stay
//           in L2, which merges the 32-byte stores into full lines before they reach DRAM;
//   pass 2  bin_dedup_kernel: one CTA per bin builds the bin's table in 64 KB of shared memory (shared-memory atomics, no
//           DRAM traffic at all), then writes the UNIQUE children as a dense array of table slots together with the compacted
//           (norm key, slot) list of those above the tolerance: no table clear, no compaction pass, every byte streamed once.
// Equal hashes always land in the same bin, so a bin can receive any number of records; what does not fit its fixed space goes
// to a spill list that is sorted by bin (sort.cuh) and read back by the bin's CTA.  A bin with more UNIQUE hashes than its
// table holds raises the overflow flag and the step is redone through the global table.
struct __align__(32) bin_record {
	unsigned long long hash;
	double re, im;
	unsigned long long rep;
};
constexpr uint32_t BIN_TABLE_SLOTS = 2048;                                  // shared-memory table of one bin (4 arrays of 8 bytes: 64 KB, three CTAs per SM)
constexpr uint32_t BIN_MEAN_RECORDS = 768;                                  // records per bin the bin count aims at (load 0.375 if all unique)
constexpr uint32_t BIN_CAPACITY = BIN_MEAN_RECORDS + 8 * 28 + 64;           // mean + 8 sigma of a Poisson(768) + slack
constexpr uint32_t BIN_MAX_UNIQUE = BIN_TABLE_SLOTS - BIN_TABLE_SLOTS / 8;  // beyond this load the bin gives up

struct bin_view {
	bin_record *records;         // bins * BIN_CAPACITY records; nullptr = no binning
	unsigned int *cursor;        // records sent to each bin so far (may exceed BIN_CAPACITY: the excess is in the spill list)
	uint32_t bins;
	bin_record *spill;           // records that found their bin full
	unsigned int *spill_bin;     // ... and the bin of each, shifted left by 8 (sort key, sort.cuh sorts on the top 24 bits)
	unsigned long long *spill_cursor;
	uint64_t spill_capacity;
};

__device__ __forceinline__ void bin_store(bin_record *dst, uint64_t hash, cplx mag, uint64_t rep) {
	ulonglong2 *d = reinterpret_cast<ulonglong2 *>(dst);
	__stcg(d, make_ulonglong2(hash, (unsigned long long)__double_as_longlong(mag.re)));
	__stcg(d + 1, make_ulonglong2((unsigned long long)__double_as_longlong(mag.im), rep));
}

// a record with a non-zero hash goes to its bin, or to the spill list when the bin's fixed space is full
__device__ __forceinline__ void bin_put(const bin_view &b, unsigned int *overflow, uint64_t hash, cplx mag, uint64_t rep) {
	const uint32_t bin = (uint32_t)__umul64hi(mix64(hash), (uint64_t)b.bins);
	const unsigned int at = atomicAdd(&b.cursor[bin], 1u);
	if (at < BIN_CAPACITY) {
		bin_store(b.records + (size_t)bin * BIN_CAPACITY + at, hash, mag, rep);
	} else {
		const unsigned long long s = atomicAdd(b.spill_cursor, 1ull);
		if (s < b.spill_capacity) {
			bin_store(b.spill + s, hash, mag, rep);
			b.spill_bin[s] = bin << 8;
		} else {
			*overflow = 1;
		}
	}
}

// returns true when the record created a slot right away (only the dedicated slot of the hash 0 does)
__device__ __forceinline__ bool bin_emit(const bin_view &b, const table_view &t, uint64_t hash, cplx mag, uint64_t rep) {
	if (hash == 0) // 0 marks an empty slot of the shared-memory tables: the hash 0 keeps its dedicated global slot
		return table_insert_zero_hash(t, mag, rep);
	bin_put(b, t.overflow, hash, mag, rep);
	return false;
}

constexpr int BIN_DEDUP_THREADS = 256;
constexpr int BIN_DEDUP_BLOCKS_PER_SM = 3; // phases of different bins overlap (clear / load + insert / write out)
constexpr size_t BIN_DEDUP_SMEM = (size_t)BIN_TABLE_SLOTS * 32;

struct bin_dedup_args {
	bin_view bins;
	const uint64_t *spill_order;    // sorted spill list: indices into bins.spill ...
	const uint32_t *spill_sorted;   // ... and their keys (bin << 8), ascending
	uint64_t n_spill;
	table_view dense;               // slots: the unique children, capacity = room; slot `capacity` is the dedicated slot of the hash 0
	unsigned long long *dense_cursor; // unique children written so far (= table.used)
	double tolerance;
	uint64_t *ukey;
	uint32_t *uslot;
	unsigned long long *count;      // entries of (ukey, uslot) so far
	int max_rep;                    // 0: the record that creates a slot is its representative; 1: the record with the largest `rep`
	                                // (the owner side of the distributed exchange packs a pseudo-random byte on top: a fair choice)
};

// first index in [0, n) with a[i] >= v (a ascending)
__device__ __forceinline__ uint64_t lower_bound_u32(const uint32_t *a, uint64_t n, uint32_t v) {
	uint64_t lo = 0, hi = n;
	while (lo < hi) {
		const uint64_t mid = (lo + hi) >> 1;
		if (a[mid] < v)
			lo = mid + 1;
		else
			hi = mid;
	}
	return lo;
}

static __global__ void __launch_bounds__(BIN_DEDUP_THREADS, BIN_DEDUP_BLOCKS_PER_SM) bin_dedup_kernel(bin_dedup_args a) {
	extern __shared__ __align__(16) unsigned long long s_bin[];
	unsigned long long *s_key = s_bin, *s_rep = s_bin + 3 * BIN_TABLE_SLOTS;
	double *s_re = reinterpret_cast<double *>(s_bin + BIN_TABLE_SLOTS), *s_im = reinterpret_cast<double *>(s_bin + 2 * BIN_TABLE_SLOTS);
	__shared__ unsigned int s_warp_occ[BIN_DEDUP_THREADS / 32], s_warp_keep[BIN_DEDUP_THREADS / 32], s_unique, s_failed;
	__shared__ unsigned long long s_base_dense, s_base_keep;
	const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
	constexpr uint32_t ROWS = BIN_TABLE_SLOTS / BIN_DEDUP_THREADS; // slots per thread in the compaction

	auto insert = [&](uint64_t hash, double re, double im, uint64_t rep) {
		uint32_t i = (uint32_t)(mix64(hash) >> 13) & (BIN_TABLE_SLOTS - 1); // other bits than the ones that chose the bin
		for (uint32_t probes = 0; probes < BIN_TABLE_SLOTS; ++probes) {
			unsigned long long k = s_key[i];
			if (k == 0) {
				if (s_unique >= BIN_MAX_UNIQUE) { // (a racy read is fine: the limit leaves an eighth of the table free)
					s_failed = 1;
					return;
				}
				k = atomicCAS(&s_key[i], 0ull, (unsigned long long)hash);
				if (k == 0) {
					if (!a.max_rep)
						s_rep[i] = rep; // this child created the slot: it is the representative
					atomicAdd(&s_unique, 1u);
					k = hash;
				}
			}
			if (k == hash) {
				atomicAdd(&s_re[i], re);
				atomicAdd(&s_im[i], im);
				if (a.max_rep)
					atomicMax(&s_rep[i], (unsigned long long)rep);
				return;
			}
			i = (i + 1) & (BIN_TABLE_SLOTS - 1);
		}
		s_failed = 1;
	};

	for (uint32_t bin = blockIdx.x; bin < a.bins.bins; bin += gridDim.x) {
		// empty table: keys, re, im (the representative of a slot is written by whoever creates it)
		for (uint32_t i = threadIdx.x; i < (a.max_rep ? 4 : 3) * BIN_TABLE_SLOTS / 2; i += BIN_DEDUP_THREADS)
			reinterpret_cast<ulonglong2 *>(s_bin)[i] = make_ulonglong2(0, 0);
		if (threadIdx.x == 0) {
			s_unique = 0;
			s_failed = table_overflowed_lane(a.dense) ? 1 : 0; // another bin gave up: the host redoes the step anyway
		}
		__syncthreads();
		if (s_failed)
			break;
		const unsigned int sent = __ldcg(&a.bins.cursor[bin]);
		const uint32_t filled = sent < BIN_CAPACITY ? sent : BIN_CAPACITY;
		const bin_record *records = a.bins.records + (size_t)bin * BIN_CAPACITY;
		// four records per thread in flight: all their loads are issued before the first insert (the shared-memory atomics of an
		// insert would otherwise fence every load behind them: one DRAM latency per record instead of one per batch)
		constexpr int BATCH = 4;
		for (uint32_t i0 = threadIdx.x; i0 < filled; i0 += BATCH * BIN_DEDUP_THREADS) {
			ulonglong2 lo[BATCH], hi[BATCH];
#pragma unroll
			for (int q = 0; q < BATCH; ++q) {
				const uint32_t i = i0 + q * BIN_DEDUP_THREADS;
				if (i < filled) {
					lo[q] = __ldcs(reinterpret_cast<const ulonglong2 *>(records + i));
					hi[q] = __ldcs(reinterpret_cast<const ulonglong2 *>(records + i) + 1);
				}
			}
#pragma unroll
			for (int q = 0; q < BATCH; ++q)
				if (i0 + q * BIN_DEDUP_THREADS < filled)
					insert(lo[q].x, __longlong_as_double((long long)lo[q].y), __longlong_as_double((long long)hi[q].x), hi[q].y);
		}
		if (sent > BIN_CAPACITY && a.n_spill) { // the part of the bin that went to the spill list
			const uint64_t first = lower_bound_u32(a.spill_sorted, a.n_spill, bin << 8), last = lower_bound_u32(a.spill_sorted, a.n_spill, (bin + 1) << 8);
			for (uint64_t j = first + threadIdx.x; j < last; j += BIN_DEDUP_THREADS) {
				const bin_record *r = a.bins.spill + a.spill_order[j];
				const ulonglong2 lo = __ldcs(reinterpret_cast<const ulonglong2 *>(r));
				const ulonglong2 hi = __ldcs(reinterpret_cast<const ulonglong2 *>(r) + 1);
				insert(lo.x, __longlong_as_double((long long)lo.y), __longlong_as_double((long long)hi.x), hi.y);
			}
		}
		__syncthreads();
		if (s_failed) { // more unique hashes than the table holds: the host redoes the step through the global table
			if (threadIdx.x == 0)
				*a.dense.overflow = 1;
			break;
		}
		// occupied slots -> dense array; those above the tolerance -> (norm key, slot) list.  Warp w owns slots
		// [w * 32 * ROWS, (w + 1) * 32 * ROWS), row r of lane l = that + r * 32 + l.
		uint32_t occ_bits = 0, keep_bits = 0;
		const uint32_t first_slot = warp * 32 * ROWS + lane;
#pragma unroll
		for (uint32_t r = 0; r < ROWS; ++r) {
			const uint32_t i = first_slot + r * 32;
			const bool occupied = s_key[i] != 0;
			const bool keep = occupied && cnorm(cplx{s_re[i], s_im[i]}) > a.tolerance;
			occ_bits |= (uint32_t)occupied << r;
			keep_bits |= (uint32_t)keep << r;
		}
		uint32_t warp_occ = 0, warp_keep = 0, before_occ[ROWS], before_keep[ROWS];
#pragma unroll
		for (uint32_t r = 0; r < ROWS; ++r) {
			const unsigned vo = __ballot_sync(0xffffffffu, (occ_bits >> r) & 1), vk = __ballot_sync(0xffffffffu, (keep_bits >> r) & 1);
			const unsigned lt = (1u << lane) - 1;
			before_occ[r] = warp_occ + __popc(vo & lt);
			before_keep[r] = warp_keep + __popc(vk & lt);
			warp_occ += __popc(vo);
			warp_keep += __popc(vk);
		}
		if (lane == 0) {
			s_warp_occ[warp] = warp_occ;
			s_warp_keep[warp] = warp_keep;
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			unsigned int occ = 0, keep = 0;
			for (int w = 0; w < BIN_DEDUP_THREADS / 32; ++w) {
				const unsigned int o = s_warp_occ[w], k = s_warp_keep[w];
				s_warp_occ[w] = occ;
				s_warp_keep[w] = keep;
				occ += o;
				keep += k;
			}
			s_base_dense = occ ? atomicAdd(a.dense_cursor, (unsigned long long)occ) : 0;
			s_base_keep = keep ? atomicAdd(a.count, (unsigned long long)keep) : 0;
			if (s_base_dense + occ > a.dense.capacity) {
				*a.dense.overflow = 1;
				s_failed = 1;
			}
		}
		__syncthreads();
		if (s_failed)
			break;
		const uint64_t base_dense = s_base_dense + s_warp_occ[warp], base_keep = s_base_keep + s_warp_keep[warp];
#pragma unroll
		for (uint32_t r = 0; r < ROWS; ++r) {
			if ((occ_bits >> r) & 1) {
				const uint32_t i = first_slot + r * 32;
				const uint64_t slot = base_dense + before_occ[r];
				ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(a.dense.slots + slot);
				const double re = s_re[i], im = s_im[i];
				dst[0] = make_ulonglong2(s_key[i], (unsigned long long)__double_as_longlong(re));
				dst[1] = make_ulonglong2((unsigned long long)__double_as_longlong(im), s_rep[i]);
				if ((keep_bits >> r) & 1) {
					const uint64_t at = base_keep + before_keep[r];
					a.ukey[at] = (uint64_t)__double_as_longlong(cnorm(cplx{re, im}));
					a.uslot[at] = (uint32_t)slot;
				}
			}
		}
		__syncthreads(); // the table is cleared again at the top of the loop
	}
	// the dedicated slot of the hash 0 (written by bin_emit through table_insert_zero_hash) joins the list
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		const table_slot *z = a.dense.slots + a.dense.capacity;
		if (z->rep != 0) {
			const double norm = cnorm(cplx{z->re, z->im});
			if (norm > a.tolerance) {
				const unsigned long long at = atomicAdd(a.count, 1ull);
				a.ukey[at] = (uint64_t)__double_as_longlong(norm);
				a.uslot[at] = (uint32_t)a.dense.capacity;
			}
		}
	}
}

} // namespace qb
