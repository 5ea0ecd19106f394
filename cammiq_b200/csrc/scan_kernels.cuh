// sm_100a kernels of the read-matching path.
//
//   pack_tiles_kernel      reads as the caller holds them (ASCII, or 2-bit bytes packed by the host)
//                          -> the scan's tile layout: 16 bases per 32-bit word, first base in the
//                          top bits, a fixed (odd) number of words per read.  ASCII is decoded and
//                          validated here (A1).  Streaming and coalesced; bound by instruction
//                          issue, not by HBM.
//   scan_reads_kernel      A2-A8 in one launch, persistent grid.  Every WARP owns tiles of 32
//                          reads and runs four steps per tile, with no block-wide barrier:
//                            stage    one TMA bulk copy (cp.async.bulk + mbarrier) brings the
//                                     tile's words into the warp's shared-memory buffer (one per
//                                     warp; a second one that prefetches the next tile is a build
//                                     option that measured slower).  The kernel keeps its shared
//                                     memory small on purpose: the in-flight probes live in the
//                                     SM's L1, which is what is left of the 256 KB after the
//                                     carve-out the host sets explicitly;
//                            phase 1  the tile's read positions are cut into strips of 8; a lane
//                                     takes a strip, extracts its first h-mer from three packed
//                                     words, rolls the hashes of BOTH strands through the strip
//                                     in registers and issues the strip's probes (one membership-
//                                     filter word in L2 per position; the 16 key bytes of the
//                                     table bucket when the index has no filter; filter word AND
//                                     bucket keys when the filter is only a sieve); the words of a
//                                     half strip are tested while the next half strip's loads are
//                                     in flight; positives are compacted into the warp's queue by
//                                     ballot, no atomics;
//                            phase 2  one queued candidate per lane: probe the prefix table in
//                                     HBM (one 32-byte sector = one bucket; the next bucket only
//                                     if the home bucket is flagged overflowed), descend the
//                                     path-compressed trie (a unary chain of up to 32 bases is one
//                                     node), append leaves to the owning read's hit list;
//                            phase 3  lane r: leaf set of read r -> decision -> counters.  Reads
//                                     with long hit lists are deduplicated by the whole warp
//                                     through a hash set; genome counters are combined across
//                                     the warp with match_any before they reach shared memory.
//   reduce_partials_kernel per-block genome-count partials -> 64-bit totals (no atomics)
//   init_pairs_kernel, aggregate_pairs_kernel, compact_pairs_kernel
//                          query64_sc's pair map: per-read pair records -> (pair, count) entries
//   ilp_inputs_kernel      ILP set-up coefficients over the leaf arrays (query.cpp:1154-1181)
//   random_*_kernel        measured lookup roofline (random gathers from L2 / HBM)
//
// Reference behaviour reproduced: query.cpp:480-527 (scan of every position, both strands),
// hashtrie.cpp:350-369 (find64_p), query.cpp:529-636 / 964-1067 (decision), query.cpp:447-450
// (reverse complement), query.cpp:1860-1883 (base codes).
#ifndef CAMMIQ_SCAN_KERNELS_CUH
#define CAMMIQ_SCAN_KERNELS_CUH

#include <cstdint>
#include <cuda_runtime.h>

#include "flat_index.hpp"

#ifndef CAMMIQ_MIN_BLOCKS
#define CAMMIQ_MIN_BLOCKS 3
#endif

namespace cammiq {

#ifndef CAMMIQ_SCAN_THREADS
#define CAMMIQ_SCAN_THREADS 256
#endif
static const int kScanThreads = CAMMIQ_SCAN_THREADS; // 8 warps, each streaming its own 32-read tiles
static const int kWarpsPerBlock = kScanThreads / 32;
#ifndef CAMMIQ_STRIP
#define CAMMIQ_STRIP 8
#endif
static const int kStrip = CAMMIQ_STRIP;       // read positions per lane and phase-1 round (that many probes in flight per lane)
static const int kQueueCap = 512;             // per-warp candidate queue; a round adds at most 32 * kStrip
static const int kHitSeg = 4;                 // per-read hit slots in shared memory (the rest spills to global)
#ifndef CAMMIQ_TILE_BUFS
#define CAMMIQ_TILE_BUFS 1
#endif
static const int kTileBufs = CAMMIQ_TILE_BUFS; // 2: the next tile's copy overlaps the scan of this one
static const int kLightHits = 16;             // longer hit lists are deduplicated by the whole warp
static const int kProbeUnroll = 4;            // micro-benchmark unroll
static const int kMaxBlocksPerSM = CAMMIQ_MIN_BLOCKS > 4 ? CAMMIQ_MIN_BLOCKS : 4;
static const uint32_t kMaxSmemGenomes = 1023; // 2*(G+1) u32 block counters stay below 8 KB of shared memory
static const uint32_t kSetEmpty = 0xFFFFFFFFu;

struct ScanParams {
	// index
	const TableBucket *table;
	uint64_t table_mask;
	uint32_t table_shift;     // home bucket of a key = tableBucket(filter hash B of its canonical h-mer, table_shift)
	const uint2 *filter;      // NULL: no filter, phase 1 probes the table
	uint32_t filter_words;    // number of 64-bit filter words
	uint32_t filter_sel;      // selector pairs of the filter test in use (flat_index.hpp, filterTest)
	const uint32_t *nodes_u, *nodes_d;
	const uint32_t *leaf_u_ref;
	const uint2 *leaf_d_ref;
	uint32_t h;
	uint32_t n_genomes;
	// reads in the tile layout pack_tiles_kernel writes: read r = words[r * words_per_read ..),
	// 16 bases per word, first base in the top bits, zero past the read's end; whole tiles of 32
	// reads are present (the padding reads are zero).  lengths[r] = 0 marks an invalid read.
	const uint32_t *words;
	const uint8_t *lengths;
	uint64_t n_reads;
	uint32_t words_per_read;  // ceil(longest/16) + 2, odd (the lanes' reads start in different banks)
	// outputs
	int smem_counters;        // 1: block-private genome counters + partials, 0: global atomics
	uint32_t *partials;       // [gridDim.x][2*(G+1)]
	unsigned long long *counts; // [2*(G+1)+4]: cnt_u | cnt_d | nundet nconf n_invalid (reserved)
	uint32_t *rcount_u, *rcount_d;
	unsigned long long *pair_records; // SC: (a<<32|b) per D_PAIR read
	unsigned long long *pair_count;   // SC: records held (NOT in the counter block: callers sum that block across devices)
	uint32_t *hit_spill;      // [total warps][32][spill_stride]
	uint32_t spill_stride;    // per-read overflow capacity: 4*(longest - h + 1) - kHitSeg hits at most
	uint32_t *dedup_sets;     // [total warps][dedup_slots] hash-set scratch of the cooperative dedup
	uint32_t dedup_slots;     // power of two >= 2 * (kHitSeg + spill_stride)
	uint32_t light_hits;      // hit lists up to this length are deduplicated by their lane (kLightHits)
	uint32_t *tile_counter;   // dynamic tile hand-out: 0 at launch, reset by the warp that takes the last tile
	unsigned long long *probe_count; // [5]: probes, candidates, leaf hits, chained bucket loads, bucket-key loads behind the sieve
	// optional per-read outputs
	uint8_t *read_class;
	uint32_t *read_rid_a, *read_rid_b;
	uint32_t leaf_cap;
	uint32_t *read_nleaf_u, *read_nleaf_d, *read_leaf_u, *read_leaf_d;
};

// ------------------------------------------------------------------------------ helpers

// L2 residency is the whole point of the filter: its words are loaded with an evict_last
// policy while every streaming access of the kernel (table sectors, leaf ids, rcount
// updates, the read tiles) is issued evict_first, so the 64 MB filter is what the L2 keeps.
__device__ __forceinline__ unsigned long long policyEvictLast() {
	unsigned long long p;
	asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
	return p;
}
__device__ __forceinline__ unsigned long long policyEvictFirst() {
	unsigned long long p;
	asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
	return p;
}
// 32 bytes = one sector = one prefix-table bucket, fetched with a single 256-bit load
// (LDG.E.256 on sm_100a), read-only path, no L1 allocation (every probe is a fresh sector).
__device__ __forceinline__ void loadBucket(const TableBucket *b, unsigned long long &k0,
		unsigned long long &k1, unsigned long long &r0, unsigned long long &r1) {
#ifdef CAMMIQ_STREAM_HINTS
	asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;"
		: "=l"(k0), "=l"(k1), "=l"(r0), "=l"(r1) : "l"(b), "l"(policyEvictFirst()));
#else
	asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
		: "=l"(k0), "=l"(k1), "=l"(r0), "=l"(r1) : "l"(b));
#endif
}
// the two keys of a bucket alone (first half of the sector): phase 1 without a filter
__device__ __forceinline__ void loadBucketKeys(const TableBucket *b, unsigned long long &k0, unsigned long long &k1) {
	asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(k0), "=l"(k1) : "l"(b));
}

__device__ __forceinline__ uint2 loadFilterWord(const uint2 *f, unsigned long long policy) {
	uint2 v;
#ifdef CAMMIQ_FILTER_L1_ALLOC
	asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(f), "l"(policy));
#else
	asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(f), "l"(policy));
#endif
	return v;
}
__device__ __forceinline__ uint32_t loadStreamU32(const uint32_t *a) {
	uint32_t v;
#ifdef CAMMIQ_STREAM_HINTS
	asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(policyEvictFirst()));
#else
	asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(a));
#endif
	return v;
}
__device__ __forceinline__ uint2 loadStreamU32x2(const uint2 *a) {
	uint2 v;
#ifdef CAMMIQ_STREAM_HINTS
	asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(a), "l"(policyEvictFirst()));
#else
	asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(a));
#endif
	return v;
}
__device__ __forceinline__ void redAddStream(uint32_t *a, unsigned long long policy) {
	asm volatile("red.global.add.L2::cache_hint.u32 [%0], 1, %1;" ::"l"(a), "l"(policy) : "memory");
}

__device__ __forceinline__ void prefetchL2(const void *a) {
	asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
}

__device__ __forceinline__ uint32_t smemAddr(const void *p) {
	return (uint32_t) __cvta_generic_to_shared(p);
}

// mbarrier + TMA bulk copy (global -> shared), the Blackwell/Hopper async-proxy path
__device__ __forceinline__ void mbarInit(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkCopyG2S(void *dst, const void *src, uint32_t bytes, uint64_t *bar, unsigned long long policy) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
		::"r"(smemAddr(dst)), "l"(src), "r"(bytes), "r"(smemAddr(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t *bar, uint32_t parity) {
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"WAIT_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra DONE_%=;\n\t"
		"bra WAIT_%=;\n\t"
		"DONE_%=:\n\t}" ::"r"(smemAddr(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ uint32_t ldsU32(uint32_t addr) {
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}

// reverse complement of a 2-bit packed h-mer held right-aligned (first base most significant):
// complement, reverse the 32 two-bit groups, drop the unused groups (query.cpp:447-450 on codes)
__device__ __forceinline__ unsigned long long revcompKey(unsigned long long x, uint32_t h) {
	x = __brevll(~x);
	x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
	return x >> (64 - 2 * h);
}

struct WarpState {
	uint32_t hits[32][kHitSeg];    // table<<31 | leaf id, per read of the warp's tile
	uint16_t queue[kQueueCap];     // slot<<10 | strands<<8 | forward window position
	uint16_t hit_cnt[32];
	uint16_t strip_base[34];       // exclusive prefix sum of the reads' strip counts, [32] = total
	uint8_t rl[32];
	uint32_t set_cnt[2];           // cooperative dedup: distinct U / D leaves of the read in work
	uint64_t bar[2];               // mbarriers of the warp's two tile buffers
};

// base j of a packed read (16 bases per word, first base in the top bits)
__device__ __forceinline__ uint32_t packedBase(uint32_t pk_addr, uint32_t j) {
	return (ldsU32(pk_addr + ((j >> 4) << 2)) >> (30u - 2u * (j & 15u))) & 3u;
}

// the h-mer starting at base i of a packed read: bits [2i, 2i+2h) of the read's bit stream
__device__ __forceinline__ unsigned long long packedWindow(uint32_t pk_addr, uint32_t i, uint32_t h) {
	const uint32_t a = pk_addr + ((i >> 4) << 2), sh = 2u * (i & 15u);
	const uint32_t w0 = ldsU32(a), w1 = ldsU32(a + 4), w2 = ldsU32(a + 8);
	const uint32_t hi = __funnelshift_l(w1, w0, sh), lo = __funnelshift_l(w2, w1, sh);
	return (((unsigned long long) hi << 32) | lo) >> (64 - 2 * h);
}

// the n-base window (n <= 32) starting at base i of a packed read, right-aligned
__device__ __forceinline__ unsigned long long packedBases(uint32_t pk_addr, uint32_t i, uint32_t n) {
	const uint32_t a = pk_addr + ((i >> 4) << 2), sh = 2u * (i & 15u);
	const uint32_t w0 = ldsU32(a), w1 = ldsU32(a + 4), w2 = ldsU32(a + 8);
	const uint32_t hi = __funnelshift_l(w1, w0, sh), lo = __funnelshift_l(w2, w1, sh);
	return (((unsigned long long) hi << 32) | lo) >> (64 - 2 * n);
}

// Walk the path-compressed trie below a bucket root: find64_p's loop (hashtrie.cpp:356-366).  The
// forward strand consumes the bases right of the window, the reverse-complement strand the
// complemented bases left of it; `next` = first base to consume / one
// past it.  A chain node stands for a run of single-child nodes: all its bases must be there and
// match (an internal node is never an answer), checked with one comparison.
template <bool REVERSE>
__device__ __forceinline__ uint32_t descend(uint32_t ref, const uint32_t *__restrict__ nodes, uint32_t pk_addr, uint32_t rl, uint32_t next) {
	while (ref != kRefNone && !(ref & kRefLeafTag)) {
		uint4 nd;
#ifdef CAMMIQ_STREAM_HINTS
		asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
			: "=r"(nd.x), "=r"(nd.y), "=r"(nd.z), "=r"(nd.w) : "l"(nodes + 4 * (size_t) (ref - 1)), "l"(policyEvictFirst()));
#else
		asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
			: "=r"(nd.x), "=r"(nd.y), "=r"(nd.z), "=r"(nd.w) : "l"(nodes + 4 * (size_t) (ref - 1)));
#endif
		if ((nd.x & kChainTag) == kChainTag) {
			const uint32_t n = nd.x & 63u;
			const unsigned long long want = ((unsigned long long) nd.y << 32) | nd.z;
			unsigned long long have;
			if (REVERSE) {
				if (next < n)
					return kRefNone;
				next -= n;
				have = revcompKey(packedBases(pk_addr, next, n), n);
			} else {
				if (next + n > rl)
					return kRefNone;
				have = packedBases(pk_addr, next, n);
				next += n;
			}
			if (have != want)
				return kRefNone;
			ref = nd.w;
		} else {
			uint32_t code;
			if (REVERSE) {
				if (next == 0)
					return kRefNone;
				next--;
				code = 3u - packedBase(pk_addr, next);
			} else {
				if (next >= rl)
					return kRefNone;
				code = packedBase(pk_addr, next);
				next++;
			}
			ref = code == 0 ? nd.x : code == 1 ? nd.y : code == 2 ? nd.z : nd.w;
		}
	}
	return ref;
}

// Leaves under a bucket root reached by one strand of a read go to the read's hit list.
template <bool REVERSE, int MODE>
__device__ __forceinline__ void collectLeaves(const ScanParams &p, WarpState &ws, uint32_t *warp_spill, uint32_t pk_addr,
		uint32_t slot, uint32_t rl, uint32_t next, unsigned long long refs, uint32_t &n_leaf_hits) {
	uint32_t leaf[2];
	leaf[0] = descend<REVERSE>((uint32_t) refs, p.nodes_u, pk_addr, rl, next);
	leaf[1] = descend<REVERSE>((uint32_t) (refs >> 32), p.nodes_d, pk_addr, rl, next);
#pragma unroll
	for (int t = 0; t < 2; t++) {
		if (leaf[t] == kRefNone)
			continue;
		n_leaf_hits++;
		// (requesting what phase 3 will touch for this leaf -- genome ids, read counter -- from HBM
		// here, like requesting a candidate's bucket when phase 1 finds it, was measured and cost
		// more than it saved: 4.47 ms with the three prefetches, 4.05-4.22 ms with any one of them
		// off; the extra lines push filter words out of L2.  Experiment builds can switch them on.)
		const uint32_t lid = leaf[t] & ~kRefLeafTag;
#ifdef CAMMIQ_LEAF_PREFETCH
		if (t) prefetchL2(&p.leaf_d_ref[lid]);
		else prefetchL2(&p.leaf_u_ref[lid]);
#endif
#ifdef CAMMIQ_RCOUNT_PREFETCH
		if (MODE == CQ_MODE_P)
			prefetchL2(t ? &p.rcount_d[lid] : &p.rcount_u[lid]);
#endif
		uint32_t e = (leaf[t] & ~kRefLeafTag) | (t ? kRefLeafTag : 0u);
		// 16-bit shared counter bumped through its containing 32-bit word
		uint32_t *word = reinterpret_cast<uint32_t *>(&ws.hit_cnt[slot & ~1u]);
		uint32_t old = atomicAdd(word, (slot & 1u) ? 0x10000u : 1u);
		uint32_t at = (slot & 1u) ? (old >> 16) : (old & 0xFFFFu);
		if (at < (uint32_t) kHitSeg) ws.hits[slot][at] = e;
		else if (at - kHitSeg < p.spill_stride) warp_spill[(size_t) slot * p.spill_stride + at - kHitSeg] = e;
	}
}

// Phase 2: the warp drains its candidate queue.  One candidate per lane: re-extract the h-mer of
// the window, probe the prefix table (HBM; keys are placed by their canonical h-mer, so both
// strands of a window share the probe sequence and ONE bucket load serves both), descend, append
// leaves to the read's hit list.  A palindromic h-mer (its own reverse complement) is found by
// both strands: the two tags are equal and both descents run.
template <int MODE>
__device__ __forceinline__ void drainQueue(const ScanParams &p, WarpState &ws, uint32_t pk_warp, uint32_t *warp_spill,
		int lane, uint32_t nq, uint32_t &n_leaf_hits, uint32_t &n_chained) {
	__syncwarp();
	const uint32_t h = p.h;
	for (uint32_t base = 0; base < nq; base += 32) {
		const uint32_t k = base + lane;
		if (k < nq) {
			const uint32_t item = ws.queue[k];
			const uint32_t slot = item >> 10, i = item & 0xFFu;
			const uint32_t strands = (item >> 8) & 3u;
			const uint32_t pk_addr = pk_warp + slot * p.words_per_read * 4u;
			const uint32_t rl = ws.rl[slot];
			const unsigned long long hf = packedWindow(pk_addr, i, h), hr = revcompKey(hf, h);
			const unsigned long long tag_f = hf | kKeyOccupied, tag_r = hr | kKeyOccupied;
			uint32_t A, B;
			filterHash(hf < hr ? hf : hr, A, B);
			uint64_t b = tableBucket(B, p.table_shift); // requested from HBM when phase 1 found the candidate
			unsigned long long refs_f = 0, refs_r = 0;
			bool found_f = !(strands & 1u), found_r = !(strands & 2u); // "nothing left to find" per strand
			bool hit_f = false, hit_r = false;
			for (;;) {
				unsigned long long k0, k1, r0, r1;
				loadBucket(p.table + b, k0, k1, r0, r1);
				const unsigned long long k0m = k0 & ~kBucketOverflow;
				if (!found_f) {
					if (k0m == tag_f) { refs_f = r0; found_f = hit_f = true; }
					else if (k1 == tag_f) { refs_f = r1; found_f = hit_f = true; }
				}
				if (!found_r) {
					if (k0m == tag_r) { refs_r = r0; found_r = hit_r = true; }
					else if (k1 == tag_r) { refs_r = r1; found_r = hit_r = true; }
				}
				if ((found_f && found_r) || !(k0 & kBucketOverflow))
					break;
				b = (b + 1) & p.table_mask; // a key passed this full bucket: look further
				n_chained++;
			}
			if (hit_f)
				collectLeaves<false, MODE>(p, ws, warp_spill, pk_addr, slot, rl, i + h, refs_f, n_leaf_hits);
			if (hit_r)
				collectLeaves<true, MODE>(p, ws, warp_spill, pk_addr, slot, rl, i, refs_r, n_leaf_hits);
		}
	}
	__syncwarp();
}

// FILTER: 0 = no filter (phase 1 loads every position's bucket keys), 1 = selective L2 filter,
// 2 = L2 filter as a sieve in front of the bucket-key loads of phase 1 (index too large to be selective)
template <int MODE, int FILTER>
__global__ void __launch_bounds__(kScanThreads, CAMMIQ_MIN_BLOCKS) scan_reads_kernel(ScanParams p) {
	// [8 warps][kTileBufs buffers][32 * words_per_read words: a packed tile] | [2*(G+1) u32 genome counters]
	extern __shared__ __align__(128) uint8_t dyn_smem[];
	__shared__ __align__(16) WarpState warp_state[kWarpsPerBlock];
	__shared__ unsigned long long block_tot[2]; // nundet, nconf

	const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
	const uint32_t tile_bytes = 32u * p.words_per_read * 4u;
	uint32_t *smem_counts = reinterpret_cast<uint32_t *>(dyn_smem + (size_t) kWarpsPerBlock * kTileBufs * tile_bytes);
	const uint32_t G1 = p.n_genomes + 1, ncnt = 2 * G1;
	if (p.smem_counters)
		for (uint32_t i = tid; i < ncnt; i += blockDim.x)
			smem_counts[i] = 0;
	if (tid < 2)
		block_tot[tid] = 0;
	WarpState &ws = warp_state[wib];
	if (lane == 0) {
		mbarInit(&ws.bar[0], 1);
		mbarInit(&ws.bar[1], 1);
	}
	__syncthreads();

	const uint32_t h = p.h;
	const unsigned long long pol_keep = policyEvictLast(), pol_stream = policyEvictFirst();
	const unsigned long long kmask = ~0ull >> (64 - 2 * h);
	const uint32_t top_shift = 2 * h - 2;
	const uint32_t lt_mask = (1u << lane) - 1u;
	uint8_t *tiles = dyn_smem + (size_t) wib * kTileBufs * tile_bytes;
	const uint32_t tiles_addr = smemAddr(tiles);
	const size_t warp_global = (size_t) blockIdx.x * kWarpsPerBlock + wib;
	uint32_t *warp_spill = p.hit_spill + warp_global * 32 * p.spill_stride;
	uint32_t *warp_set = p.dedup_sets + warp_global * p.dedup_slots;
	uint32_t n_undet = 0, n_conf = 0, n_invalid = 0, n_probes = 0; // per lane (a lane sees < 2^32 / 512 reads per launch)
	uint32_t n_cand = 0, n_leaf_hits = 0, n_chained = 0, n_sieve = 0; // n_cand warp-uniform, the others per lane
	uint32_t parity = 0; // bit b: phase of buffer b's mbarrier

	const uint32_t n_sub = (uint32_t) ((p.n_reads + 31) / 32); // the host keeps a launch below 2^32 tiles
	const uint32_t warp_stride = gridDim.x * kWarpsPerBlock;
	uint32_t sub = blockIdx.x * kWarpsPerBlock + wib;

	// one bulk copy per tile, straight into the layout the phases read
	auto issueCopy = [&](uint32_t tile, uint32_t buf) {
		if (lane == 0) {
			mbarExpectTx(&ws.bar[buf], tile_bytes);
			bulkCopyG2S(tiles + buf * tile_bytes, p.words + (size_t) tile * 32 * p.words_per_read, tile_bytes, &ws.bar[buf], pol_stream);
		}
	};

	if (sub < n_sub && kTileBufs == 2)
		issueCopy(sub, 0);
#if !defined(CAMMIQ_STATIC_TILES) && CAMMIQ_TILE_BUFS == 1
	// Tiles are handed out dynamically: a warp's first tile is its own index, every further one comes
	// from a global counter (tile = number of warps + the counter's old value), fetched at the START of
	// the tile in work so that the atomic's round trip hides behind the tile.  A CTA that becomes
	// resident late (a collective or another stream's kernel holds its slot for a while) then simply
	// takes fewer tiles instead of finishing its fixed share late.  Every processed tile takes exactly
	// one number, so the n_sub-th taker is the last one of the launch and leaves the counter at 0.
	uint32_t taken = 0;
	for (uint32_t it = 0; sub < n_sub; it++) {
		if (lane == 0) {
			taken = atomicAdd(p.tile_counter, 1u);
			if (taken == n_sub - 1)
				atomicExch(p.tile_counter, 0u);
		}
#else
	for (uint32_t it = 0; sub < n_sub; sub += warp_stride, it++) {
#endif
		const uint32_t buf = kTileBufs == 2 ? (it & 1u) : 0u;
		// the other buffer is free (its tile was finished one iteration ago): the next tile's words
		// arrive while this tile is scanned
		if (kTileBufs == 2) {
			if (sub + warp_stride < n_sub)
				issueCopy(sub + warp_stride, buf ^ 1u);
		} else
			issueCopy(sub, 0);
		const uint64_t r = (uint64_t) sub * 32 + lane;
		const bool have = r < p.n_reads;
		const uint32_t rl = have ? p.lengths[r] : 0u;
		mbarWait(&ws.bar[buf], (parity >> buf) & 1u);
		parity ^= 1u << buf;
		const uint32_t pk_warp = tiles_addr + buf * tile_bytes;

		const bool valid = have && rl >= h; // invalid reads arrive with length 0
		const uint32_t npos = valid ? rl - h + 1 : 0;
		// strips of the tile, flattened over its reads
		const uint32_t my_strips = (npos + kStrip - 1) / kStrip;
		uint32_t incl = my_strips;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += up;
		}
		ws.strip_base[lane] = (uint16_t) (incl - my_strips);
		if (lane == 31)
			ws.strip_base[32] = (uint16_t) incl;
		ws.rl[lane] = (uint8_t) rl;
		ws.hit_cnt[lane] = 0;
		const uint32_t total_strips = __shfl_sync(0xffffffffu, incl, 31);
		const uint32_t first_strips = __shfl_sync(0xffffffffu, my_strips, 0);
		const bool uniform_strips = __all_sync(0xffffffffu, my_strips == first_strips) && first_strips > 0;
		const uint32_t strip_inv = uniform_strips ? (65536u + my_strips - 1) / my_strips : 0u;
		n_probes += 2 * npos; // both strands of every window
		__syncwarp();

		// ---- phase 1: a strip of kStrip positions per lane, all probes in flight before any test ---
		uint32_t nq = 0; // queue fill, warp-uniform
		// a strip's positions are tested half a strip late (the second half in the next round): every
		// filter word has at least half a strip of hashing between its load and its first use
		uint2 pend_ff[kStrip / 2];
		uint32_t pend_bsel[kStrip / 2], pend_item = 0;
#pragma unroll
		for (int t = 0; t < kStrip / 2; t++) {
			pend_ff[t] = make_uint2(0u, 0u); // an all-zero word fails the test: nothing is pending yet
			pend_bsel[t] = 0;
		}
		// one position's filter word -> candidate queue (warp-wide: every lane calls it)
		auto testAndQueue = [&](uint2 word, uint32_t sel_bits, uint32_t item) {
			const bool cand = filterTest(word.x, word.y, sel_bits, p.filter_sel);
			const uint32_t m = __ballot_sync(0xffffffffu, cand);
			if (m) { // warp-uniform; the queue slot of a lane is its rank among the lanes with a candidate
				if (cand) {
					ws.queue[nq + __popc(m & lt_mask)] = (uint16_t) (item | (3u << 8)); // either strand may hold the key
#ifdef CAMMIQ_BUCKET_PREFETCH
					prefetchL2(p.table + tableBucket(sel_bits, p.table_shift));
#endif
				}
				nq += __popc(m);
			}
		};
		for (uint32_t k0 = 0; k0 < total_strips; k0 += 32) {
			const uint32_t k = k0 + lane;
			const bool live = k < total_strips;
			// the read this strip belongs to: a division when all reads of the tile have the same
			// number of strips (the usual case), else the last read whose first strip is <= k
			uint32_t slot = 0, i0 = 0;
			if (uniform_strips) {
				slot = (k * strip_inv) >> 16; // exact: k < 2^10, strips per read <= 32
				i0 = (k - slot * my_strips) * kStrip;
				if (!live) slot = 0, i0 = 0;
			} else {
#pragma unroll
				for (int step = 16; step > 0; step >>= 1)
					if (ws.strip_base[slot + step] <= k) slot += step;
				i0 = live ? (k - ws.strip_base[slot]) * kStrip : 0u;
			}
			const uint32_t s_rl = ws.rl[slot];
			const uint32_t count = live ? min((uint32_t) kStrip, s_rl - h + 1 - i0) : 0u;
			// bases i0 .. i0+h+kStrip-2 of the read lie in three packed words (i0 is a multiple of kStrip)
			static_assert(kStrip == 4 || kStrip == 8, "strips must not straddle the three-word window");
			const uint32_t a = pk_warp + slot * p.words_per_read * 4u + ((i0 >> 4) << 2), sh = (i0 & 15u) * 2u;
			const uint32_t w0 = ldsU32(a), w1 = ldsU32(a + 4), w2 = ldsU32(a + 8);
			const uint32_t hi = __funnelshift_l(w1, w0, sh), lo = __funnelshift_l(w2, w1, sh), lo2 = w2 << sh;
			unsigned long long hf = (((unsigned long long) hi << 32) | lo) >> (64 - 2 * h);
			unsigned long long hr = revcompKey(hf, h);
			// the bases that enter the window at the strip's later positions, first one in the top bits
			const uint32_t enter = 2 * h >= 32 ? __funnelshift_l(lo2, lo, 2 * h - 32) : __funnelshift_l(lo, hi, 2 * h);

#ifndef CAMMIQ_NO_PIPELINE
			if (FILTER == 1) {
				uint2 ff[kStrip];
				uint32_t bsel[kStrip];
				auto probe = [&](int t) {
					if (t > 0) {
						const uint32_t c = (enter >> (32 - 2 * t)) & 3u;
						hf = ((hf << 2) | c) & kmask;                                     // query.cpp:493-494
						hr = (hr >> 2) | ((unsigned long long) (3u - c) << top_shift);    // the window's reverse complement
					}
					// hr is the reverse complement of the window hf covers: ONE probe with the canonical
					// h-mer answers both strands
					uint32_t A;
					filterHash(hf <= hr ? hf : hr, A, bsel[t]);
					ff[t] = make_uint2(0u, 0u);
					if ((uint32_t) t < count)
						ff[t] = loadFilterWord(p.filter + filterWordIndex(A, p.filter_words), pol_keep);
				};
				const uint32_t item = (slot << 10) | i0;
#pragma unroll
				for (int t = 0; t < kStrip / 2; t++)
					probe(t);
#pragma unroll
				for (int t = 0; t < kStrip / 2; t++) // the previous round's second half
					testAndQueue(pend_ff[t], pend_bsel[t], pend_item + t);
#pragma unroll
				for (int t = kStrip / 2; t < kStrip; t++)
					probe(t);
#pragma unroll
				for (int t = 0; t < kStrip / 2; t++) // this round's first half
					testAndQueue(ff[t], bsel[t], item + t);
#pragma unroll
				for (int t = 0; t < kStrip / 2; t++) {
					pend_ff[t] = ff[kStrip / 2 + t];
					pend_bsel[t] = bsel[kStrip / 2 + t];
				}
				pend_item = item + kStrip / 2;
				// a round adds at most 32 * kStrip entries: drain above half
				if (nq > (uint32_t) (kQueueCap - 32 * kStrip)) {
					n_cand += nq;
					drainQueue<MODE>(p, ws, pk_warp, warp_spill, lane, nq, n_leaf_hits, n_chained);
					nq = 0;
				}
				continue;
			}
#endif
			uint2 ff[kStrip];                     // FILTER: filter words
			uint32_t bsel[kStrip];
			unsigned long long kf[kStrip];        // !FILTER: forward key, the two keys of the strands' shared bucket
			unsigned long long bk0[kStrip], bk1[kStrip];
#pragma unroll
			for (int t = 0; t < kStrip; t++) {
				if (t > 0) {
					const uint32_t c = (enter >> (32 - 2 * t)) & 3u;
					hf = ((hf << 2) | c) & kmask;                                     // query.cpp:493-494
					hr = (hr >> 2) | ((unsigned long long) (3u - c) << top_shift);    // the window's reverse complement
				}
				if (FILTER) {
					// hr is the reverse complement of the window hf covers: ONE probe with the canonical
					// h-mer answers both strands
					uint32_t A;
					filterHash(hf <= hr ? hf : hr, A, bsel[t]);
					ff[t] = make_uint2(0u, 0u);
					if ((uint32_t) t < count)
						ff[t] = loadFilterWord(p.filter + filterWordIndex(A, p.filter_words), pol_keep);
					if (FILTER == 2)
						kf[t] = hf;
				} else {
					kf[t] = hf;
					bk0[t] = bk1[t] = 0ull;
					// a key and its reverse complement share their home bucket: one sector per position
					if ((uint32_t) t < count) {
						uint32_t A;
						filterHash(hf < hr ? hf : hr, A, bsel[t]);
						loadBucketKeys(p.table + tableBucket(bsel[t], p.table_shift), bk0[t], bk1[t]);
					}
				}
			}
			if (FILTER == 2) {
				// sieve: only the positions whose filter word passes (about half at 1.5 bits per key) go
				// to HBM for their bucket's keys; an all-zero word (past the strip's end) never passes
#pragma unroll
				for (int t = 0; t < kStrip; t++) {
					bk0[t] = bk1[t] = 0ull;
					if (filterTest(ff[t].x, ff[t].y, bsel[t], p.filter_sel)) {
						loadBucketKeys(p.table + tableBucket(bsel[t], p.table_shift), bk0[t], bk1[t]);
						n_sieve++;
					}
				}
			}
#ifdef CAMMIQ_INTERLEAVE
			// the candidates of the PREVIOUS round are resolved while this round's probes are in
			// flight: their buckets were requested a full round ago, and the filter words of this
			// round get the time of a phase-2 pass to arrive
			if (FILTER == 1 && nq > 0) {
				n_cand += nq;
				drainQueue<MODE>(p, ws, pk_warp, warp_spill, lane, nq, n_leaf_hits, n_chained);
				nq = 0;
			}
#endif
#pragma unroll
			for (int t = 0; t < kStrip; t++) {
				bool cand_f, cand_r;
				if (FILTER == 1) {
					// the canonical h-mer of the window may be a key, in either orientation: phase 2 looks
					// for both.  An all-zero word (position past the strip's end) fails the test.
					cand_f = cand_r = filterTest(ff[t].x, ff[t].y, bsel[t], p.filter_sel);
				} else {
					// candidate = the bucket holds the key, or a key spilled past this bucket; the reverse
					// strand's key is recomputed rather than kept live across the loads
					const unsigned long long tag_f = kf[t] | kKeyOccupied, tag_r = revcompKey(kf[t], h) | kKeyOccupied;
					const unsigned long long k0m = bk0[t] & ~kBucketOverflow;
					const bool over = (bk0[t] & kBucketOverflow) != 0;
					cand_f = k0m == tag_f || bk1[t] == tag_f || over;
					cand_r = k0m == tag_r || bk1[t] == tag_r || over;
				}
				const uint32_t strands = (cand_f ? 1u : 0u) | (cand_r ? 2u : 0u);
				const uint32_t m = __ballot_sync(0xffffffffu, strands != 0);
				if (m) { // warp-uniform; the queue slot of a lane is its rank among the lanes with a candidate
					if (strands) {
						ws.queue[nq + __popc(m & lt_mask)] = (uint16_t) ((slot << 10) | (strands << 8) | (i0 + t));
#ifdef CAMMIQ_BUCKET_PREFETCH
						if (FILTER == 1)
							prefetchL2(p.table + tableBucket(bsel[t], p.table_shift));
#endif
					}
					nq += __popc(m);
				}
			}
			// a round adds at most 32 * kStrip entries: drain above half
			if (nq > (uint32_t) (kQueueCap - 32 * kStrip)) {
				n_cand += nq;
				drainQueue<MODE>(p, ws, pk_warp, warp_spill, lane, nq, n_leaf_hits, n_chained);
				nq = 0;
			}
		}
#ifndef CAMMIQ_NO_PIPELINE
		if (FILTER == 1) {
#pragma unroll
			for (int t = 0; t < kStrip / 2; t++) // the last round's second half
				testAndQueue(pend_ff[t], pend_bsel[t], pend_item + t);
		}
#endif
		n_cand += nq;
		drainQueue<MODE>(p, ws, pk_warp, warp_spill, lane, nq, n_leaf_hits, n_chained);

		// ---- phase 3: lane r: leaf set of read r -> decision (query.cpp:529-636) -------------------
		uint32_t cls = CQ_CLASS_UNLABELED, rid_a = 0, rid_b = 0, distinct_u = 0, distinct_d = 0;
		const uint32_t nh = valid ? min((uint32_t) ws.hit_cnt[lane], (uint32_t) kHitSeg + p.spill_stride) : 0;
		const uint32_t *my_spill = warp_spill + (size_t) lane * p.spill_stride;
		if (nh > 0) {
			uint32_t min_r = 0xFFFFFFFFu, max_r = 0;
			unsigned long long min_p = ~0ull, max_p = 0;
			for (uint32_t i = 0; i < nh; i++) {
				uint32_t e = i < (uint32_t) kHitSeg ? ws.hits[lane][i] : my_spill[i - kHitSeg];
				if (e & kRefLeafTag) {
					uint2 ab = loadStreamU32x2(&p.leaf_d_ref[e & ~kRefLeafTag]);
					uint32_t l = min(ab.x, ab.y), g = max(ab.x, ab.y);
					unsigned long long key = ((unsigned long long) l << 32) | g;
					min_p = key < min_p ? key : min_p;
					max_p = key > max_p ? key : max_p;
				} else {
					uint32_t rid = loadStreamU32(&p.leaf_u_ref[e]);
					min_r = min(min_r, rid);
					max_r = max(max_r, rid);
				}
			}
			const int nr = (min_r == 0xFFFFFFFFu) ? 0 : (min_r == max_r ? 1 : 2);
			const int np = (min_p == ~0ull) ? 0 : (min_p == max_p ? 1 : 2);
			const uint32_t a0 = (uint32_t) (min_p >> 32), b0 = (uint32_t) min_p;
			if (np == 0) {
				if (nr == 1) { cls = CQ_CLASS_U; rid_a = min_r; }
				else cls = CQ_CLASS_CONFLICT;
			} else if (np == 1) {
				if (nr == 0) { cls = CQ_CLASS_D_PAIR; rid_a = a0; rid_b = b0; }
				else if (nr == 2) cls = CQ_CLASS_CONFLICT;
				else if (a0 != min_r && b0 != min_r) cls = CQ_CLASS_CONFLICT;
				else { cls = CQ_CLASS_UD; rid_a = min_r; }
			} else if (nr == 2) {
				cls = CQ_CLASS_CONFLICT;
			} else {
				// |P| >= 2: does every pair contain r (|R| == 1), or which of a0 / b0 lies in
				// every pair (|R| == 0, the intersection of query.cpp:604-633)
				bool all_a = true, all_b = true;
				const uint32_t ta = nr == 1 ? min_r : a0, tb = nr == 1 ? min_r : b0;
				for (uint32_t i = 0; i < nh; i++) {
					uint32_t e = i < (uint32_t) kHitSeg ? ws.hits[lane][i] : my_spill[i - kHitSeg];
					if (e & kRefLeafTag) {
						uint2 ab = loadStreamU32x2(&p.leaf_d_ref[e & ~kRefLeafTag]);
						all_a &= (ab.x == ta || ab.y == ta);
						all_b &= (ab.x == tb || ab.y == tb);
					}
				}
				if (nr == 1) {
					if (all_a) { cls = CQ_CLASS_UD; rid_a = min_r; }
					else cls = CQ_CLASS_CONFLICT;
				} else {
					int ni = (all_a ? 1 : 0) + ((b0 != a0 && all_b) ? 1 : 0);
					if (ni == 1) { cls = CQ_CLASS_D_INTER; rid_a = all_a ? a0 : b0; }
					else cls = CQ_CLASS_CONFLICT;
				}
			}
		}
		// distinct leaves: rcount += 1 per distinct leaf of an accepted read (query.cpp:550-551; the
		// reference collects leaf pointers in a std::set, so a leaf hit at several positions counts once)
		const bool accepted = cls >= CQ_CLASS_U;
		const bool want_sets = p.read_nleaf_u != NULL;
		const bool count_leaves = nh > 0 && ((MODE == CQ_MODE_P && accepted) || want_sets);
		if (count_leaves && nh <= p.light_hits) {
			for (uint32_t i = 0; i < nh; i++) {
				uint32_t e = i < (uint32_t) kHitSeg ? ws.hits[lane][i] : my_spill[i - kHitSeg];
				bool first = true;
				for (uint32_t q = 0; q < i && first; q++)
					first = (q < (uint32_t) kHitSeg ? ws.hits[lane][q] : my_spill[q - kHitSeg]) != e;
				if (!first)
					continue;
				const bool is_d = (e & kRefLeafTag) != 0;
				const uint32_t leaf = e & ~kRefLeafTag;
				if (MODE == CQ_MODE_P && accepted)
					redAddStream(is_d ? &p.rcount_d[leaf] : &p.rcount_u[leaf], pol_stream);
				if (want_sets) {
					uint32_t at = is_d ? distinct_d : distinct_u;
					if (at < p.leaf_cap)
						(is_d ? p.read_leaf_d : p.read_leaf_u)[r * p.leaf_cap + at] = leaf;
				}
				if (is_d) distinct_d++;
				else distinct_u++;
			}
		}
		// long hit lists (near-duplicate strains, dense indices): the whole warp deduplicates one
		// read at a time through a hash set in global scratch -- linear instead of quadratic work
		uint32_t heavy = __ballot_sync(0xffffffffu, count_leaves && nh > p.light_hits);
		while (heavy) {
			const int src = __ffs(heavy) - 1;
			heavy &= heavy - 1;
			const uint32_t s_nh = __shfl_sync(0xffffffffu, nh, src);
			const bool s_acc = __shfl_sync(0xffffffffu, (MODE == CQ_MODE_P && accepted) ? 1 : 0, src) != 0;
			const uint64_t s_r = (uint64_t) sub * 32 + src;
			const uint32_t set_mask = p.dedup_slots - 1;
			for (uint32_t q = lane; q < p.dedup_slots; q += 32)
				warp_set[q] = kSetEmpty;
			if (lane < 2)
				ws.set_cnt[lane] = 0;
			__syncwarp();
			const uint32_t *s_spill = warp_spill + (size_t) src * p.spill_stride;
			for (uint32_t i = lane; i < s_nh; i += 32) {
				const uint32_t e = i < (uint32_t) kHitSeg ? ws.hits[src][i] : s_spill[i - kHitSeg];
				uint32_t q = (e * 0x9E3779B1u) >> 7 & set_mask;
				bool first;
				for (;;) {
					const uint32_t seen = atomicCAS(&warp_set[q], kSetEmpty, e);
					if (seen == kSetEmpty || seen == e) {
						first = seen == kSetEmpty;
						break;
					}
					q = (q + 1) & set_mask;
				}
				if (first) {
					const bool is_d = (e & kRefLeafTag) != 0;
					const uint32_t leaf = e & ~kRefLeafTag;
					if (s_acc)
						redAddStream(is_d ? &p.rcount_d[leaf] : &p.rcount_u[leaf], pol_stream);
					const uint32_t at = atomicAdd(&ws.set_cnt[is_d ? 1 : 0], 1u);
					if (want_sets && at < p.leaf_cap)
						(is_d ? p.read_leaf_d : p.read_leaf_u)[s_r * p.leaf_cap + at] = leaf;
				}
			}
			__syncwarp();
			if (lane == src) {
				distinct_u = ws.set_cnt[0];
				distinct_d = ws.set_cnt[1];
			}
			__syncwarp();
		}

		// ---- counters (the effects of query.cpp:542-636) ---------------------------------------------
		const bool inc_u = have && (cls == CQ_CLASS_U || cls == CQ_CLASS_UD || (cls == CQ_CLASS_D_INTER && MODE == CQ_MODE_SC));
		const bool inc_d = have && cls >= CQ_CLASS_D_PAIR;
		const bool inc_b = have && cls == CQ_CLASS_D_PAIR;
		if (have) {
			if (!valid) n_invalid++;
			if (cls == CQ_CLASS_UNLABELED) n_undet++;
			else if (cls == CQ_CLASS_CONFLICT) n_conf++;
		}
		// lanes that bump the same genome combine first (match_any): one add per distinct genome
		// and warp, however skewed the sample is
#pragma unroll
		for (int which = 0; which < 3; which++) {
			const bool inc = which == 0 ? inc_u : which == 1 ? inc_d : inc_b;
			const uint32_t at = (which == 0 ? 0u : G1) + (which == 2 ? rid_b : rid_a);
			const uint32_t m = __ballot_sync(0xffffffffu, inc);
			if (inc) {
				const uint32_t peers = __match_any_sync(m, at);
				if (lane == __ffs(peers) - 1) {
					if (p.smem_counters) atomicAdd(&smem_counts[at], (uint32_t) __popc(peers));
					else atomicAdd(&p.counts[at], (unsigned long long) __popc(peers));
				}
			}
		}
		if (MODE == CQ_MODE_SC) {
			// one pair record per D_PAIR read; the warp reserves its records with one atomic
			const uint32_t m = __ballot_sync(0xffffffffu, inc_b);
			if (m) {
				unsigned long long base = 0;
				if (lane == __ffs(m) - 1)
					base = atomicAdd(p.pair_count, (unsigned long long) __popc(m));
				base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
				if (inc_b)
					p.pair_records[base + __popc(m & lt_mask)] = ((unsigned long long) rid_a << 32) | rid_b;
			}
		}
		if (have) {
			if (p.read_class) {
				p.read_class[r] = (uint8_t) cls;
				p.read_rid_a[r] = rid_a;
				p.read_rid_b[r] = rid_b;
			}
			if (p.read_nleaf_u) {
				p.read_nleaf_u[r] = distinct_u;
				p.read_nleaf_d[r] = distinct_d;
			}
		}
		__syncwarp(); // every lane is done with the packed tile and the warp state
#if !defined(CAMMIQ_STATIC_TILES) && CAMMIQ_TILE_BUFS == 1
		{
			const unsigned long long nx = (unsigned long long) warp_stride + __shfl_sync(0xffffffffu, taken, 0);
			sub = nx < n_sub ? (uint32_t) nx : n_sub;
		}
#endif
	}

	// ---- block totals ---------------------------------------------------------------------------------
	unsigned long long undet64 = n_undet, conf64 = n_conf, invalid64 = n_invalid, probes64 = n_probes, leaf64 = n_leaf_hits,
		chain64 = n_chained, sieve64 = n_sieve;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		undet64 += __shfl_xor_sync(0xffffffffu, undet64, o);
		conf64 += __shfl_xor_sync(0xffffffffu, conf64, o);
		invalid64 += __shfl_xor_sync(0xffffffffu, invalid64, o);
		probes64 += __shfl_xor_sync(0xffffffffu, probes64, o);
		leaf64 += __shfl_xor_sync(0xffffffffu, leaf64, o);
		chain64 += __shfl_xor_sync(0xffffffffu, chain64, o);
		sieve64 += __shfl_xor_sync(0xffffffffu, sieve64, o);
	}
	if (lane == 0) {
		if (undet64) atomicAdd(&block_tot[0], undet64);
		if (conf64) atomicAdd(&block_tot[1], conf64);
		if (invalid64) atomicAdd(&p.counts[ncnt + 2], invalid64);
		if (probes64) atomicAdd(&p.probe_count[0], probes64);
		if (n_cand) atomicAdd(&p.probe_count[1], (unsigned long long) n_cand);
		if (leaf64) atomicAdd(&p.probe_count[2], leaf64);
		if (chain64) atomicAdd(&p.probe_count[3], chain64);
		if (sieve64) atomicAdd(&p.probe_count[4], sieve64);
	}
	__syncthreads();
	if (p.smem_counters) {
		uint32_t *dst = p.partials + (size_t) blockIdx.x * ncnt;
		for (uint32_t i = tid; i < ncnt; i += blockDim.x)
			dst[i] = smem_counts[i];
	}
	if (tid < 2 && block_tot[tid])
		atomicAdd(&p.counts[ncnt + tid], block_tot[tid]);
}

// ------------------------------------------------------------------- read packing (A1)
//
// One thread per (read, word): 16 bases -> one 32-bit word of the scan's tile layout.  ASCII input
// is what query64_* consumes (one byte per base, any alignment): decoded with the codes of
// query.cpp:1860-1883, validated (a byte outside ACGTacgt zeroes the read's length: the read is
// then counted unlabeled and invalid), packed first base most significant.  PACKED_IN: the bytes
// the host packer / FASTQ reader produce (four bases per byte, validated there) are regrouped into
// words.  Reads past n_reads up to a whole tile are written as zeros.
struct PackParams {
	const uint8_t *bases;
	const uint64_t *offsets;   // NULL: read i starts at (read_base + i) * stride
	const uint32_t *offsets32; // PACKED_IN only: 32-bit offsets, or NULL
	uint64_t stride, read_base;
	const uint8_t *lengths_in;
	uint64_t n_reads, n_padded; // n_padded = n_reads rounded up to 32
	uint32_t words_per_read;
	uint32_t base_words;       // ceil(longest / 16): the words of a read that can hold bases
	uint32_t reads_per_block;  // floor(256 / base_words)
	uint32_t inv_words;        // ceil(2^16 / base_words)
	uint32_t *words;           // [n_padded][words_per_read]
	uint8_t *lengths_out;      // ASCII input: a copy of lengths_in in which invalid reads are zeroed
};

// prmt without the selector masking __byte_perm adds (the callers' selector nibbles are 0..7 by construction)
__device__ __forceinline__ uint32_t prmt(uint32_t x, uint32_t y, uint32_t sel) {
	uint32_t v;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(v) : "r"(x), "r"(y), "r"(sel));
	return v;
}

__device__ __forceinline__ uint32_t ldgU32(const uint8_t *a) {
	uint32_t v;
	asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(a));
	return v;
}

template <bool PACKED_IN>
__global__ void __launch_bounds__(256) pack_tiles_kernel(PackParams q) {
	// a block pass takes floor(256 / base_words) whole reads, one thread per word that can hold bases
	// (the two or three slack words of a read are zeroed by the thread of its last word): read and word
	// of a thread come from 32-bit arithmetic with a precomputed reciprocal (exact for thread ids below
	// 2^16 / base_words).  The grid is persistent (a few blocks per SM striding over the passes), and a
	// thread keeps its word column: input and output addresses advance by constants.  The kernel is
	// bound by the integer pipe (about 130 instructions per 16 bases; ncu: ALU 74 %, DRAM 41 %), not by HBM: everything below is written to
	// keep that count down.
	const uint32_t k = (threadIdx.x * q.inv_words) >> 16;
	const uint32_t c = threadIdx.x - k * q.base_words;
	if (k >= q.reads_per_block)
		return;
	const bool last_word = c + 1 == q.base_words;
	const bool slack3 = q.words_per_read - q.base_words == 3; // (base_words + 2) | 1: two or three slack words
	const bool strided = !q.offsets && !(PACKED_IN && q.offsets32);
	const uint64_t step = (uint64_t) gridDim.x * q.reads_per_block;
	uint64_t r = (uint64_t) blockIdx.x * q.reads_per_block + k;
	uint32_t *out = q.words + r * q.words_per_read + c;
	const uint64_t out_step = step * q.words_per_read;
	const uint32_t col = (PACKED_IN ? 4u : 16u) * c; // byte of this thread's word within a read
	const uint8_t *in = q.bases + (q.read_base + r) * q.stride + col;
	const uint64_t in_step = step * q.stride;
	const int first_base = (int) (16u * c);
	for (; r < q.n_padded; r += step, out += out_step, in += in_step) {
		uint32_t word = 0;
		const int n = r < q.n_reads ? (int) q.lengths_in[r] - first_base : 0; // bases of this read in word c
		if (n > 0) {
			const uint8_t *a = in;
			if (!strided)
				a = q.bases + col + (PACKED_IN && q.offsets32 ? (uint64_t) q.offsets32[r] : q.offsets[r]);
			const uint8_t *al = (const uint8_t *) ((uintptr_t) a & ~(uintptr_t) 3);
			const uint32_t sh = (uint32_t) ((uintptr_t) a & 3u) * 8u;
			const uint32_t n16 = (uint32_t) min(n, 16);
			if (PACKED_IN) {
				// four bytes of the host-packed read = this word, big-endian; bytes past the read are masked
				const uint32_t w0 = ldgU32(al), w1 = sh ? ldgU32(al + 4) : 0u; // an aligned word never needs the next one
				word = __byte_perm(__funnelshift_r(w0, w1, sh), 0u, 0x0123);
			} else {
				// sixteen ASCII bytes at any alignment: up to five aligned words, funnel-shifted
				const uint32_t nw = ((sh >> 3) + n16 + 3u) >> 2; // aligned words the bases touch
				uint32_t w[5];
#pragma unroll
				for (int t = 0; t < 5; t++)
					w[t] = (uint32_t) t < nw ? ldgU32(al + 4 * t) : 0u;
				uint32_t bad = 0, g[4];
#pragma unroll
				for (int t = 0; t < 4; t++) {
					const uint32_t d = __funnelshift_r(w[t], w[t + 1], sh);
					// codes A/a=0 C/c=1 G/g=2 T/t=3 from bits 1 and 2 of each byte (on anything else the code is
					// arbitrary and the comparison below fails); validity: fold case and compare with the letter
					// each code stands for (the four codes gathered into one permute selector)
					const uint32_t code4 = ((d >> 1) ^ (d >> 2)) & 0x03030303u;
					const uint32_t nib = code4 | (code4 >> 4);
					const uint32_t expect4 = prmt(0x54474341u, 0u, prmt(nib, 0u, 0x0020u));
					// bytes past the read do not count: all-ones shifted left by 8 x (bases left), the shift
					// clamped to [0, 32], marks them
					const int left = 8 * (int) n16 - 32 * t;
					const uint32_t dead = __funnelshift_lc(0u, 0xFFFFFFFFu, (uint32_t) (t == 0 ? left : max(left, 0)));
					bad |= ((d & 0xDFDFDFDFu) ^ expect4) & ~dead;
					// four 2-bit codes -> the top byte, first base in the top bits
					g[t] = code4 * 0x40100401u;
				}
				if (bad)
					q.lengths_out[r] = 0; // every thread of an invalid read that sees a bad byte stores the same 0
				word = prmt(prmt(g[3], g[2], 0x0073u), prmt(g[1], g[0], 0x7300u), 0x7610u);
			}
			word &= __funnelshift_lc(0u, 0xFFFFFFFFu, 32u - 2u * n16);
		}
		out[0] = word;
		if (last_word) {
			out[1] = 0u;
			out[2] = 0u;
			if (slack3)
				out[3] = 0u;
		}
	}
}

// counts[i] += sum over blocks of partials[b][i]; one thread per counter, coalesced over i.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const uint32_t *__restrict__ partials,
		uint32_t n_blocks, uint32_t ncnt, unsigned long long *counts) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= ncnt)
		return;
	unsigned long long sum = 0;
	for (uint32_t b = 0; b < n_blocks; b++)
		sum += partials[(size_t) b * ncnt + i];
	counts[i] += sum;
}

// read_cnts_b of query64_sc (query.cpp:994-997): the scan leaves one (a<<32|b) record per
// D_PAIR read; these two kernels fold them into (pair, count) entries on the device so that
// only the distinct pairs cross PCIe.  Open-addressing table, capacity a power of two >= 2x
// the number of records (cannot fill up); lanes of a warp holding the same pair combine first,
// so a sample dominated by one pair does not serialise on one address.
struct PairSlot {
	unsigned long long key;   // a<<32 | b, kEmptyKey = free
	unsigned long long count;
};

__global__ void __launch_bounds__(256) init_pairs_kernel(PairSlot *table, uint64_t n_slots) {
	const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n_slots) {
		table[i].key = kEmptyKey;
		table[i].count = 0;
	}
}

__global__ void __launch_bounds__(256) aggregate_pairs_kernel(const unsigned long long *__restrict__ records, uint64_t n_records,
		PairSlot *table, uint64_t mask) {
	const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	const bool have = i < n_records;
	const unsigned long long key = have ? records[i] : kEmptyKey;
	const unsigned active = __ballot_sync(0xffffffffu, have);
	if (!have)
		return;
	const unsigned peers = __match_any_sync(active, key);
	if ((int) (threadIdx.x & 31u) != __ffs(peers) - 1)
		return; // another lane of the warp adds for this pair
	const unsigned long long add = (unsigned long long) __popc(peers);
	uint64_t b = mixKey(key) & mask;
	for (;;) {
		const unsigned long long seen = atomicCAS(&table[b].key, kEmptyKey, key);
		if (seen == kEmptyKey || seen == key) {
			atomicAdd(&table[b].count, add);
			return;
		}
		b = (b + 1) & mask;
	}
}

__global__ void __launch_bounds__(256) compact_pairs_kernel(const PairSlot *__restrict__ table, uint64_t n_slots, PairSlot *out,
		unsigned long long *n_out) {
	const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_slots || table[i].key == kEmptyKey)
		return;
	out[atomicAdd(n_out, 1ull)] = table[i];
}

// ------------------------------------------------------------------- ILP input assembly (8f.3)
//
// What runILP_* computes per leaf before it builds the model (query.cpp:1154-1181, 1508-1535):
//   U leaf l of genome i:  wcov  = ucount1 * (rl - depth) * 1.0 / rl * pow(1 - erate, depth)
//   D leaf l of (i, j):    wcov1 / wcov2 likewise from ucount1 / ucount2
// with rl a uint32_t (so ucount * (rl - depth) is 32-bit unsigned arithmetic, kept as such), and
// per genome the sum of its coverages over map_sp[i] -- the coefficient of COV[i] in the
// constraints EXP1 / EXP2 (query.cpp:1196-1230) -- and the sum of its leaves' rcount.  One thread
// per leaf; the per-genome sums are double-precision atomics, so their last bits depend on the
// order of addition (the reference sums in map_sp order) and pow() is CUDA's, not glibc's:
// callers compare with a relative tolerance.
struct IlpParams {
	const uint32_t *ref1, *ref2;   // genome ids of leaf l at [l * ref_stride] (ref2 = NULL for the unique table)
	uint32_t ref_stride;
	const uint16_t *ucount1, *ucount2;
	const uint8_t *depth;
	const uint32_t *rcount;        // [n] pleafNode::rcount
	uint64_t n;
	uint32_t n_genomes;
	uint32_t rl;
	double one_minus_e;
	double *wcov1, *wcov2;         // [n] out (wcov2 = NULL for the unique table)
	double *genome_wcov;           // [G+1] out, += coverage of the genome's leaves
	unsigned long long *genome_rcount; // [G+1] out, += rcount of the genome's leaves
};

__global__ void __launch_bounds__(256) ilp_inputs_kernel(IlpParams q) {
	const uint64_t l = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (l >= q.n)
		return;
	const uint32_t depth = q.depth[l];
	const double decay = pow(q.one_minus_e, (double) depth);
	const double w1 = (double) ((uint32_t) q.ucount1[l] * (q.rl - depth)) * 1.0 / (double) q.rl * decay;
	q.wcov1[l] = w1;
	const uint32_t a = q.ref1[l * q.ref_stride], rc = q.rcount[l];
	if (a >= 1 && a <= q.n_genomes) {
		atomicAdd(&q.genome_wcov[a], w1);
		if (rc) atomicAdd(&q.genome_rcount[a], (unsigned long long) rc);
	}
	if (q.ref2 != NULL) {
		const double w2 = (double) ((uint32_t) q.ucount2[l] * (q.rl - depth)) * 1.0 / (double) q.rl * decay;
		q.wcov2[l] = w2;
		const uint32_t b = q.ref2[l * q.ref_stride];
		if (b >= 1 && b <= q.n_genomes) {
			atomicAdd(&q.genome_wcov[b], w2);
			if (rc) atomicAdd(&q.genome_rcount[b], (unsigned long long) rc);
		}
	}
}

// ------------------------------------------------------------------- lookup roofline probes

__global__ void __launch_bounds__(256) random_sector_kernel(const TableBucket *table, uint64_t mask,
		uint64_t n_probes, uint64_t seed, unsigned long long *sink) {
	const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t n_threads = (uint64_t) gridDim.x * blockDim.x;
	unsigned long long acc = 0;
	for (uint64_t i = tid; i < n_probes; i += n_threads * kProbeUnroll) {
		unsigned long long k0[kProbeUnroll], r0[kProbeUnroll], k1[kProbeUnroll], r1[kProbeUnroll];
#pragma unroll
		for (int u = 0; u < kProbeUnroll; u++) {
			uint64_t j = i + (uint64_t) u * n_threads;
			if (j < n_probes)
				loadBucket(table + (mixKey(j + seed) & mask), k0[u], r0[u], k1[u], r1[u]);
		}
#pragma unroll
		for (int u = 0; u < kProbeUnroll; u++) {
			uint64_t j = i + (uint64_t) u * n_threads;
			if (j < n_probes)
				acc += k0[u] ^ r0[u] ^ k1[u] ^ r1[u];
		}
	}
	if (acc == 0x123456789ull)
		*sink = acc;
}

template <int BYTES>
__global__ void __launch_bounds__(256) random_gather_kernel(const uint8_t *region, uint64_t mask,
		uint64_t n_probes, uint64_t seed, unsigned long long *sink) {
	const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t n_threads = (uint64_t) gridDim.x * blockDim.x;
	unsigned long long acc = 0;
	for (uint64_t i = tid; i < n_probes; i += n_threads * kProbeUnroll) {
		unsigned long long v[kProbeUnroll][4];
#pragma unroll
		for (int u = 0; u < kProbeUnroll; u++) {
			uint64_t j = i + (uint64_t) u * n_threads;
			v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0;
			if (j < n_probes) {
				const uint8_t *a = region + ((mixKey(j + seed) * BYTES) & mask);
				if (BYTES == 4) { unsigned int t; asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(t) : "l"(a)); v[u][0] = t; }
				else if (BYTES == 8) asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v[u][0]) : "l"(a));
				else if (BYTES == 16) asm volatile("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(v[u][0]), "=l"(v[u][1]) : "l"(a));
				else asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[u][0]), "=l"(v[u][1]), "=l"(v[u][2]), "=l"(v[u][3]) : "l"(a));
			}
		}
#pragma unroll
		for (int u = 0; u < kProbeUnroll; u++)
			acc += v[u][0] ^ v[u][1] ^ v[u][2] ^ v[u][3];
	}
	if (acc == 0x123456789ull)
		*sink = acc;
}

} // namespace cammiq
#endif
