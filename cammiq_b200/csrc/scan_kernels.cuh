// sm_100a kernels of the read-matching path.
//
//   scan_reads_kernel      A1-A8 in one launch, persistent grid: every WARP pulls sub-tiles of
//                          32 reads (ASCII, or 2-bit packed by the host) into its own shared-
//                          memory buffer with one TMA bulk copy (cp.async.bulk + mbarrier), then
//                            phase 1 (thread per read): decode each base to its 2-bit code,
//                              roll the h-base prefix hash of BOTH strands in registers and
//                              test the canonical h-mer of every position against the L2-
//                              resident membership filter (or the prefix table itself when the
//                              index is too large for a filter); positives go to a per-warp
//                              queue,
//                            phase 2 (warp-cooperative): queued candidates probe the prefix
//                              table in HBM (one 32-byte sector = one bucket) and descend the
//                              CSR trie; leaves land in per-read hit lists,
//                            phase 3 (thread per read): leaf set -> decision -> counters.
//   reduce_partials_kernel per-block genome-count partials -> 64-bit totals (no atomics)
//   init_pairs_kernel, aggregate_pairs_kernel, compact_pairs_kernel
//                          query64_sc's pair map: per-read pair records -> (pair, count) entries
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

namespace cammiq {

static const int kScanThreads = 256;          // 8 warps, each streaming its own 32-read sub-tiles
static const int kWarpsPerBlock = kScanThreads / 32;
static const int kQueueCap = 512;             // per-warp candidate queue (drained when > 256 used)
static const int kHitSeg = 8;                 // per-read hit slots in shared memory
static const int kHitSpill = 1024;            // per-read overflow in global memory (2 tables x 2 strands x 251)
static const int kStepUnroll = 4;             // bases per thread between filter tests (8 loads in flight)
static const int kProbeUnroll = 4;            // micro-benchmark unroll
static const int kMaxBlocksPerSM = 4;
static const uint32_t kMaxSmemGenomes = 8191; // 2*(G+1) u32 block counters must fit 64 KB

struct ScanParams {
	// index
	const TableSlot *table;
	uint64_t table_mask;
	const uint2 *filter;      // NULL: no filter, phase 1 probes the table
	uint32_t filter_words;    // number of 64-bit filter words
	const uint32_t *nodes_u, *nodes_d;
	const uint32_t *leaf_u_ref;
	const uint2 *leaf_d_ref;
	uint32_t h;
	uint32_t n_genomes;
	// reads in device memory.  ASCII (exactly the state query64_* consumes), or PACKED: 2 bits per
	// base, base j of a read in byte j/4 at bits 7-2*(j%4)..6-2*(j%4) (first base most
	// significant), ceil(len/4) bytes per read, validated on the host (an invalid read arrives
	// with length 0)
	const uint8_t *bases;
	const uint64_t *offsets;  // NULL: read i starts at (read_base + i)*stride
	const uint32_t *offsets32; // PACKED only: 32-bit offsets (batch-relative), or NULL
	uint64_t stride;
	uint64_t read_base;       // caller's index of this launch's first read (chunked submission)
	const uint8_t *lengths;
	uint64_t n_reads;
	uint32_t tile_cap;        // bytes of one staging buffer (32 reads); 2 per warp
	// outputs
	int smem_counters;        // 1: block-private genome counters + partials, 0: global atomics
	uint32_t *partials;       // [gridDim.x][2*(G+1)]
	unsigned long long *counts; // [2*(G+1)+4]: cnt_u | cnt_d | nundet nconf n_invalid n_pair_records
	uint32_t *rcount_u, *rcount_d;
	unsigned long long *pair_records; // SC: (a<<32|b) per D_PAIR read
	uint32_t *hit_spill;      // [total warps][32][kHitSpill]
	unsigned long long *probe_count; // [4]: probes, candidates, leaf hits, chained bucket loads
	// optional per-read outputs
	uint8_t *read_class;
	uint32_t *read_rid_a, *read_rid_b;
	uint32_t leaf_cap;
	uint32_t *read_nleaf_u, *read_nleaf_d, *read_leaf_u, *read_leaf_d;
};

// ------------------------------------------------------------------------------ helpers

// A/a=0 C/c=1 G/g=2 T/t=3 (query.cpp:1860-1883); `bad` is raised for any other byte.
__device__ __forceinline__ uint32_t decodeBase(uint32_t c, bool &bad) {
	uint32_t u = (c & 0xDFu) - 0x41u; // fold case; 'A' -> 0, 'C' -> 2, 'G' -> 6, 'T' -> 19
	bad |= (u > 19u) || !((0x80045u >> u) & 1u);
	uint32_t t = (c >> 1) & 3u;
	return t ^ (t >> 1);
}

// 32 bytes = one sector = one prefix-table bucket, fetched with a single 256-bit load
// (LDG.E.256 on sm_100a), read-only path, no L1 allocation (every probe is a fresh sector).
__device__ __forceinline__ void loadBucket(const TableSlot *b, unsigned long long &k0,
		unsigned long long &r0, unsigned long long &k1, unsigned long long &r1) {
	asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
		: "=l"(k0), "=l"(r0), "=l"(k1), "=l"(r1) : "l"(b));
}

// L2 residency is the whole point of the filter: its words are loaded with an evict_last
// policy while every streaming access of the kernel (table sectors, leaf ids, rcount
// updates, the read tiles) is issued evict_first, so the 64 MB filter is what the L2 keeps.
__device__ __forceinline__ unsigned long long policyEvictLast() {
	unsigned long long p;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
	return p;
}
__device__ __forceinline__ unsigned long long policyEvictFirst() {
	unsigned long long p;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
	return p;
}
__device__ __forceinline__ uint2 loadFilterWord(const uint2 *f, unsigned long long policy) {
	uint2 v;
	asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(f), "l"(policy));
	return v;
}
__device__ __forceinline__ uint32_t loadStreamU32(const uint32_t *a) {
	uint32_t v;
	asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(a));
	return v;
}
__device__ __forceinline__ uint2 loadStreamU32x2(const uint2 *a) {
	uint2 v;
	asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(a));
	return v;
}
__device__ __forceinline__ void redAddStream(uint32_t *a, unsigned long long policy) {
	asm volatile("red.global.add.L2::cache_hint.u32 [%0], 1, %1;" ::"l"(a), "l"(policy) : "memory");
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

struct WarpState {
	uint32_t hits[32][kHitSeg];    // table<<31 | leaf id, per read of the warp's sub-tile
	uint16_t queue[kQueueCap];     // slot<<9 | strand<<8 | position
	uint32_t soff[32];             // byte offset of the slot's read inside the staged sub-tile
	unsigned long long goff[32];   // its offset in the caller's base buffer (fallback path)
	uint16_t hit_cnt[32];
	uint8_t rl[32];
	uint32_t q_count;
	uint32_t staged;               // sub-tile is in shared memory (else: read from global)
	uint64_t bar;                  // mbarrier of the warp's staging buffer
};

// first base of a slot's read, generic address space (phase 2 / trie descent only)
__device__ __forceinline__ const uint8_t *slotBases(const ScanParams &p, const WarpState &ws, const uint8_t *buf, uint32_t slot) {
	return ws.staged ? buf + ws.soff[slot] : p.bases + ws.goff[slot];
}

// base `j` of strand `strand` of a read (strand 1 = reverse complement, query.cpp:447-450)
template <bool PACKED>
__device__ __forceinline__ uint32_t strandBase(const uint8_t *s, uint32_t rl, uint32_t strand, uint32_t j) {
	const uint32_t at = strand ? rl - 1 - j : j;
	uint32_t c;
	if (PACKED) {
		c = ((uint32_t) s[at >> 2] >> (6u - 2u * (at & 3u))) & 3u;
	} else {
		bool bad = false;
		c = decodeBase(s[at], bad);
	}
	return strand ? 3u - c : c;
}

// Walk the CSR trie below a bucket root: find64_p's loop (hashtrie.cpp:356-366).
template <bool PACKED>
__device__ __forceinline__ uint32_t descend(uint32_t ref, const uint32_t *__restrict__ nodes,
		const uint8_t *s, uint32_t rl, uint32_t strand, uint32_t next) {
	while (ref != kRefNone && !(ref & kRefLeafTag)) {
		if (next >= rl)
			return kRefNone;
		uint32_t code = strandBase<PACKED>(s, rl, strand, next);
		ref = loadStreamU32(&nodes[4 * (size_t) (ref - 1) + code]);
		next++;
	}
	return ref;
}

__device__ __forceinline__ uint32_t ldsWord(uint32_t addr) {
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}

// reverse the 32 two-bit groups of x
__device__ __forceinline__ unsigned long long reverseGroups(unsigned long long x) {
	x = __brevll(x);
	return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
}

// 2-bit packing of bases [j, j+n) of strand `strand` of a read (n <= 32, first base most
// significant, right-aligned).  The reverse-complement strand's window is the reverse
// complement of the forward window [rl-j-n, rl-j); staged reads are decoded four bases per
// 32-bit shared load (the bytes were validated by phase 1).
template <bool PACKED>
__device__ __forceinline__ unsigned long long strandWindow(const uint8_t *s, bool in_smem, uint32_t rl, uint32_t strand,
		uint32_t j, uint32_t n) {
	const uint32_t i = strand ? rl - j - n : j;
	unsigned long long hv = 0;
	if (PACKED) {
		// the window is bits [2i, 2i+2n) of the read's big-endian bit stream
		if (in_smem) {
			// three aligned words cover the <= 9 bytes the window touches (the staging buffer has slack)
			const uint32_t a = smemAddr(s) + (i >> 2);
			const uint32_t w0 = __byte_perm(ldsWord(a & ~3u), 0u, 0x0123), w1 = __byte_perm(ldsWord((a & ~3u) + 4u), 0u, 0x0123),
				w2 = __byte_perm(ldsWord((a & ~3u) + 8u), 0u, 0x0123);
			const uint32_t sh = (a & 3u) * 8u + 2u * (i & 3u); // 0..30
			const uint32_t hi = __funnelshift_l(w1, w0, sh), lo = __funnelshift_l(w2, w1, sh);
			hv = (((unsigned long long) hi << 32) | lo) >> (64 - 2 * n);
		} else {
			for (uint32_t t = 0; t < n; t++)
				hv = (hv << 2) | (((uint32_t) s[(i + t) >> 2] >> (6u - 2u * ((i + t) & 3u))) & 3u);
		}
	} else if (in_smem) {
		const uint32_t a0 = smemAddr(s) + i;
		for (uint32_t t = 0; t < n; t += 4) {
			const uint32_t a = a0 + t;
			const uint32_t w4 = __funnelshift_r(ldsWord(a & ~3u), ldsWord((a & ~3u) + 4u), (a & 3u) * 8u);
			const uint32_t t4 = (w4 >> 1) & 0x03030303u;
			const uint32_t code4 = t4 ^ ((t4 >> 1) & 0x01010101u);
#pragma unroll
			for (int u = 0; u < 4; u++)
				if (t + u < n) hv = (hv << 2) | ((code4 >> (8 * u)) & 3u);
		}
	} else {
		bool bad = false;
		for (uint32_t t = 0; t < n; t++)
			hv = (hv << 2) | decodeBase(s[i + t], bad);
	}
	return strand ? reverseGroups(~hv) >> (64 - 2 * n) : hv;
}

// Leaves under a bucket root reached by one strand of a read go to the read's hit list.
template <bool PACKED>
__device__ __forceinline__ void collectLeaves(const ScanParams &p, WarpState &ws, uint32_t *warp_spill, const uint8_t *s,
		uint32_t slot, uint32_t rl, uint32_t strand, uint32_t next, unsigned long long refs, uint32_t &n_leaf_hits) {
	uint32_t leaf[2];
	leaf[0] = descend<PACKED>((uint32_t) refs, p.nodes_u, s, rl, strand, next);
	leaf[1] = descend<PACKED>((uint32_t) (refs >> 32), p.nodes_d, s, rl, strand, next);
#pragma unroll
	for (int t = 0; t < 2; t++) {
		if (leaf[t] == kRefNone)
			continue;
		n_leaf_hits++;
		uint32_t e = (leaf[t] & ~kRefLeafTag) | (t ? kRefLeafTag : 0u);
		// 16-bit shared counter bumped through its containing 32-bit word
		uint32_t *word = reinterpret_cast<uint32_t *>(&ws.hit_cnt[slot & ~1u]);
		uint32_t old = atomicAdd(word, (slot & 1u) ? 0x10000u : 1u);
		uint32_t at = (slot & 1u) ? (old >> 16) : (old & 0xFFFFu);
		if (at < (uint32_t) kHitSeg) ws.hits[slot][at] = e;
		else if (at < (uint32_t) (kHitSeg + kHitSpill)) warp_spill[(size_t) slot * kHitSpill + at - kHitSeg] = e;
	}
}

// Phase 2: the warp drains its candidate queue.  One candidate per lane: recompute the h-mer,
// probe the prefix table (HBM), descend, append leaves to the owning read's hit list.
// FILTER: phase 1 does not look for palindromic h-mers (equal to their own reverse complement);
// for those the forward candidate also serves the reverse strand here and a reverse candidate
// (a false positive of the filter's other pattern) is dropped.
template <bool PACKED, bool FILTER>
__device__ __forceinline__ void drainQueue(const ScanParams &p, WarpState &ws, const uint8_t *buf, uint32_t *warp_spill,
		int lane, uint32_t &n_leaf_hits, uint32_t &n_chained) {
	__syncwarp();
	const uint32_t nq = ws.q_count;
	const uint32_t h = p.h;
	for (uint32_t base = 0; base < nq; base += 32) {
		const uint32_t k = base + lane;
		if (k < nq) {
			const uint32_t item = ws.queue[k];
			const uint32_t slot = item >> 9, strand = (item >> 8) & 1u, pos = item & 0xFFu;
			const uint8_t *s = slotBases(p, ws, buf, slot);
			const uint32_t rl = ws.rl[slot];
			const unsigned long long hv = strandWindow<PACKED>(s, ws.staged != 0, rl, strand, pos, h);
			// keys are placed by their canonical h-mer (flat_index.hpp, homeBucketHost)
			const unsigned long long hv_rc = reverseGroups(~hv) >> (64 - 2 * h);
			const bool palin = FILTER && hv == hv_rc;
			if (palin && strand)
				continue;
			uint64_t b = mixKey(hv < hv_rc ? hv : hv_rc) & p.table_mask;
			unsigned long long k0, r0, k1, r1, refs = 0;
			bool found = false;
			for (;;) {
				loadBucket(p.table + 2 * b, k0, r0, k1, r1);
				if (k0 == hv) { refs = r0; found = true; }
				else if (k1 == hv) { refs = r1; found = true; }
				if (found || k0 == kEmptyKey || k1 == kEmptyKey)
					break;
				b = (b + 1) & p.table_mask; // full bucket: the key may have spilled to the next one
				n_chained++;
			}
			if (found) {
				collectLeaves<PACKED>(p, ws, warp_spill, s, slot, rl, strand, pos + h, refs, n_leaf_hits);
				if (palin) // the reverse strand holds the same h-mer at position rl-h-pos
					collectLeaves<PACKED>(p, ws, warp_spill, s, slot, rl, 1u, rl - pos, refs, n_leaf_hits);
			}
		}
	}
	__syncwarp();
	if (lane == 0)
		ws.q_count = 0;
	__syncwarp();
}

__device__ __forceinline__ uint32_t ldsU8(uint32_t addr) {
	uint32_t v;
	asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}

__device__ __forceinline__ uint32_t ldsU32(uint32_t addr) {
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}

template <int MODE, bool FILTER, bool PACKED>
__global__ void __launch_bounds__(kScanThreads, 3) scan_reads_kernel(ScanParams p) {
	// [8 warps][tile_cap bytes of ASCII] | [2*(G+1) u32 genome counters]
	extern __shared__ __align__(128) uint8_t dyn_smem[];
	__shared__ __align__(16) WarpState warp_state[kWarpsPerBlock];
	__shared__ unsigned long long block_tot[2]; // nundet, nconf

	const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
	uint32_t *smem_counts = reinterpret_cast<uint32_t *>(dyn_smem + (size_t) kWarpsPerBlock * p.tile_cap);
	const uint32_t G1 = p.n_genomes + 1, ncnt = 2 * G1;
	if (p.smem_counters)
		for (uint32_t i = tid; i < ncnt; i += blockDim.x)
			smem_counts[i] = 0;
	if (tid < 2)
		block_tot[tid] = 0;
	WarpState &ws = warp_state[wib];
	if (lane == 0) {
		ws.q_count = 0;
		mbarInit(&ws.bar, 1);
	}
	__syncthreads();

	const uint32_t h = p.h;
	const unsigned long long pol_keep = policyEvictLast(), pol_stream = policyEvictFirst();
	const unsigned long long kmask = ~0ull >> (64 - 2 * h);
	const uint32_t top_shift = 2 * h - 2;
	uint8_t *wbuf = dyn_smem + (size_t) wib * p.tile_cap;
	uint32_t *warp_spill = p.hit_spill + ((size_t) blockIdx.x * kWarpsPerBlock + wib) * 32 * kHitSpill;
	unsigned long long n_undet = 0, n_conf = 0, n_invalid = 0;
	uint32_t n_probes = 0, n_cand = 0, n_leaf_hits = 0, n_chained = 0; // per lane
	uint32_t parity = 0;

	// Every warp owns sub-tiles of 32 reads (one read per lane) and streams them through its own
	// staging buffer with its own mbarrier: no block-wide barrier in the loop, the other warps
	// of the SM cover the (short) bulk-copy latency.
	const uint64_t n_sub = (p.n_reads + 31) / 32;
	const uint64_t warp_stride = (uint64_t) gridDim.x * kWarpsPerBlock;
	uint64_t sub = (uint64_t) blockIdx.x * kWarpsPerBlock + wib;

	struct SubTile {
		unsigned long long off, start; // this lane's read offset; 16-byte aligned start of the span
		uint32_t rl, bytes;
		bool have, staged;
	};
	// offsets / lengths of a sub-tile's reads and the byte span they cover (warp-uniform)
	auto describe = [&](uint64_t sub_idx) -> SubTile {
		SubTile t;
		const uint64_t r = sub_idx * 32 + lane;
		t.have = r < p.n_reads;
		if (PACKED)
			t.off = t.have ? (p.offsets32 ? (unsigned long long) p.offsets32[r] : p.offsets ? p.offsets[r] : (p.read_base + r) * p.stride) : ~0ull;
		else
			t.off = t.have ? (p.offsets ? p.offsets[r] : (p.read_base + r) * p.stride) : ~0ull;
		t.rl = t.have ? p.lengths[r] : 0;
		unsigned long long lo = t.off, hi = t.have ? t.off + (PACKED ? (t.rl + 3u) >> 2 : t.rl) : 0ull;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			unsigned long long tl = __shfl_xor_sync(0xffffffffu, lo, o), th = __shfl_xor_sync(0xffffffffu, hi, o);
			lo = tl < lo ? tl : lo;
			hi = th > hi ? th : hi;
		}
		t.start = lo & ~15ull;
		t.bytes = hi > t.start ? (uint32_t) min((hi - t.start + 15ull) & ~15ull, 0xFFFFFFF0ull) : 0u;
		t.staged = t.bytes > 0 && t.bytes <= p.tile_cap; // else: the reads are fetched from global directly
		return t;
	};

	SubTile cur;
	cur.have = false; cur.staged = false; cur.rl = 0; cur.off = 0; cur.start = 0; cur.bytes = 0;
	if (sub < n_sub)
		cur = describe(sub);
	for (; sub < n_sub; sub += warp_stride) {
		// one TMA bulk copy brings the sub-tile's ASCII bytes into the warp's buffer
		if (cur.staged) {
			if (lane == 0) {
				mbarExpectTx(&ws.bar, cur.bytes);
				bulkCopyG2S(wbuf, p.bases + cur.start, cur.bytes, &ws.bar, pol_stream);
			}
		}
		// the next sub-tile's offsets / lengths are fetched while this one is in flight
		SubTile nxt;
		nxt.have = false; nxt.staged = false; nxt.rl = 0; nxt.off = 0; nxt.start = 0; nxt.bytes = 0;
		if (sub + warp_stride < n_sub)
			nxt = describe(sub + warp_stride);
		if (cur.staged) {
			mbarWait(&ws.bar, parity);
			parity ^= 1u;
		}
		const uint64_t r = sub * 32 + lane;
		const bool have = cur.have, staged = cur.staged;
		const uint32_t rl = cur.rl;
		const uint8_t *tbuf = wbuf;
		const uint32_t soff = staged && have ? (uint32_t) (cur.off - cur.start) : 0u;
		ws.soff[lane] = soff;
		ws.goff[lane] = have ? cur.off : 0ull;
		ws.rl[lane] = (uint8_t) rl;
		ws.hit_cnt[lane] = 0;
		if (lane == 0)
			ws.staged = staged ? 1u : 0u;
		__syncwarp();

		// ---- phase 1: thread per read, rolling hashes of both strands, filter / table test ------
		// Four bases per iteration: one (unaligned) 32-bit shared load, SIMD-in-register decode
		// and validation, then four roll steps; positions with a full h-base window are probed.
		const uint32_t sbase = smemAddr(tbuf) + soff;
		const uint8_t *gbase = p.bases + (have ? cur.off : 0ull);
		uint32_t bad4 = 0;
		unsigned long long hf = 0, hr = 0;
		const uint32_t wmax = __reduce_max_sync(0xffffffffu, rl);
		for (uint32_t j0 = 0; j0 < wmax; j0 += kStepUnroll) {
			// bytes j0..j0+3 of the read (garbage past rl is masked below)
			uint32_t w4 = 0;
			if (PACKED) {
				// one byte = the four codes of this iteration (validated and zero-padded by the host)
				if (j0 < rl)
					w4 = staged ? ldsU8(sbase + (j0 >> 2)) : (uint32_t) gbase[j0 >> 2];
			} else if (j0 < rl) {
				if (staged) {
					const uint32_t a = sbase + j0;
					w4 = __funnelshift_r(ldsU32(a & ~3u), ldsU32((a & ~3u) + 4u), (a & 3u) * 8u);
				} else {
#pragma unroll
					for (int u = 0; u < kStepUnroll; u++)
						if (j0 + u < rl) w4 |= (uint32_t) gbase[j0 + u] << (8 * u);
				}
			}
			// codes: A/a=0 C/c=1 G/g=2 T/t=3 in each byte; validity: fold case and compare with the
			// letter the code stands for (one byte permute)
			uint32_t code4;
			if (PACKED) {
				code4 = w4;
			} else {
				const uint32_t t4 = (w4 >> 1) & 0x03030303u;
				code4 = t4 ^ ((t4 >> 1) & 0x01010101u);
				const uint32_t nib = code4 | (code4 >> 4);
				const uint32_t expect4 = __byte_perm(0x54474341u, 0u, (nib & 0xFFu) | ((nib >> 8) & 0xFF00u));
				const uint32_t left = rl > j0 ? rl - j0 : 0u;
				const uint32_t live = left >= 4u ? 0xFFFFFFFFu : ((1u << (8u * left)) - 1u);
				bad4 |= ((w4 & 0xDFDFDFDFu) ^ expect4) & live;
			}

			uint2 ff[kStepUnroll];                                // FILTER: filter words
			uint32_t bsel[kStepUnroll];
			unsigned long long kf[kStepUnroll];                   // !FILTER: the forward key and the strands' shared bucket
			unsigned long long bk[kStepUnroll][4];
#pragma unroll
			for (int u = 0; u < kStepUnroll; u++) {
				const uint32_t j = j0 + u;
				const uint32_t c = PACKED ? (code4 >> (6 - 2 * u)) & 3u : (code4 >> (8 * u)) & 3u;
				hf = ((hf << 2) | c) & kmask;
				hr = (hr >> 2) | ((unsigned long long) (3u - c) << top_shift);
				if (j + 1 >= h && j < rl) {
					if (FILTER) {
						// hr is the reverse complement of the window hf covers: ONE probe with the
						// canonical h-mer answers both strands
						uint32_t a;
						const bool fwd_is_canon = hf <= hr;
						filterHash(fwd_is_canon ? hf : hr, a, bsel[u]);
						// bit 31 of B is unused by the selectors: remember the orientation
						bsel[u] = (bsel[u] & 0x7FFFFFFFu) | (fwd_is_canon ? 0x80000000u : 0u);
						ff[u] = loadFilterWord(p.filter + filterWordIndex(a, p.filter_words), pol_keep);
					} else {
						kf[u] = hf;
						// a key and its reverse complement share their home bucket: one sector per position
						loadBucket(p.table + 2 * (mixKey(hf < hr ? hf : hr) & p.table_mask), bk[u][0], bk[u][1], bk[u][2], bk[u][3]);
					}
				}
			}
			if (j0 + kStepUnroll >= h) { // warp-uniform: some step of this iteration has a full window
#pragma unroll
				for (int u = 0; u < kStepUnroll; u++) {
					const uint32_t j = j0 + u;
					bool cand_f = false, cand_r = false;
					if (j + 1 >= h && j < rl) {
						if (FILTER) {
							// pattern of B: the canonical orientation is a key; other pattern: its reverse
							// complement is.  fwd_canon says which strand holds the canonical orientation.
							const bool same = filterTest(ff[u].x, ff[u].y, bsel[u]);
							const bool other = filterTest(ff[u].x, ff[u].y, filterOtherPattern(bsel[u]));
							// (a palindromic h-mer is its own reverse complement: phase 2 serves its reverse
							// strand from the forward candidate)
							const bool fwd_canon = (bsel[u] & 0x80000000u) != 0;
							cand_f = fwd_canon ? same : other;
							cand_r = fwd_canon ? other : same;
						} else {
							// candidate = the bucket holds the key, or is full and the key may have spilled
							// the reverse strand's key is recomputed rather than kept live across the loads
							const unsigned long long kr = reverseGroups(~kf[u]) >> (64 - 2 * h);
							const bool full = bk[u][0] != kEmptyKey && bk[u][2] != kEmptyKey;
							cand_f = bk[u][0] == kf[u] || bk[u][2] == kf[u] || full;
							cand_r = bk[u][0] == kr || bk[u][2] == kr || full;
						}
					}
					if (cand_f | cand_r) { // rare: one branch on the common path
						if (cand_f) {
							// forward strand, position i = j-h+1
							uint32_t at = atomicAdd(&ws.q_count, 1u);
							ws.queue[at] = (uint16_t) (((uint32_t) lane << 9) | (j + 1 - h));
							n_cand++;
						}
						if (cand_r) {
							// reverse-complement strand: this window is rc position rl-1-j
							uint32_t at = atomicAdd(&ws.q_count, 1u);
							ws.queue[at] = (uint16_t) (((uint32_t) lane << 9) | 0x100u | (rl - 1 - j));
							n_cand++;
						}
					}
				}
				__syncwarp();
				// at most 2*kStepUnroll*32 = 256 candidates arrive per iteration: drain above half
				if (ws.q_count > (uint32_t) (kQueueCap - 2 * kStepUnroll * 32))
					drainQueue<PACKED, FILTER>(p, ws, tbuf, warp_spill, lane, n_leaf_hits, n_chained);
			}
		}
		const bool bad = bad4 != 0;
		if (have && rl >= h)
			n_probes += 2 * (rl - h + 1); // both strands of every window
		drainQueue<PACKED, FILTER>(p, ws, tbuf, warp_spill, lane, n_leaf_hits, n_chained);

		// ---- phase 3: thread per read: leaf set -> decision (query.cpp:529-636) ------------------
		uint32_t cls = CQ_CLASS_UNLABELED, rid_a = 0, rid_b = 0, distinct_u = 0, distinct_d = 0;
		const bool valid = have && !bad && rl >= h;
		const uint32_t nh = valid ? min((uint32_t) ws.hit_cnt[lane], (uint32_t) (kHitSeg + kHitSpill)) : 0;
		const uint32_t *my_spill = warp_spill + (size_t) lane * kHitSpill;
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
			// distinct leaves: rcount += 1 per distinct leaf of an accepted read (query.cpp:550-551)
			const bool accepted = cls >= CQ_CLASS_U;
			const bool want_sets = p.read_nleaf_u != NULL;
			if ((MODE == CQ_MODE_P && accepted) || want_sets) {
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
		}

		// ---- counters (the effects of query.cpp:542-636) ---------------------------------------------
		if (have) {
			const bool inc_u = cls == CQ_CLASS_U || cls == CQ_CLASS_UD || (cls == CQ_CLASS_D_INTER && MODE == CQ_MODE_SC);
			const bool inc_d = cls >= CQ_CLASS_D_PAIR;
			if (!valid) n_invalid++;
			if (cls == CQ_CLASS_UNLABELED) n_undet++;
			else if (cls == CQ_CLASS_CONFLICT) n_conf++;
			else if (p.smem_counters) {
				if (inc_u) atomicAdd(&smem_counts[rid_a], 1u);
				if (inc_d) atomicAdd(&smem_counts[G1 + rid_a], 1u);
				if (cls == CQ_CLASS_D_PAIR) atomicAdd(&smem_counts[G1 + rid_b], 1u);
			} else {
				if (inc_u) atomicAdd(&p.counts[rid_a], 1ull);
				if (inc_d) atomicAdd(&p.counts[G1 + rid_a], 1ull);
				if (cls == CQ_CLASS_D_PAIR) atomicAdd(&p.counts[G1 + rid_b], 1ull);
			}
			if (MODE == CQ_MODE_SC && cls == CQ_CLASS_D_PAIR) {
				unsigned long long at = atomicAdd(&p.counts[2 * G1 + 3], 1ull);
				p.pair_records[at] = ((unsigned long long) rid_a << 32) | rid_b;
			}
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
		__syncwarp(); // every lane is done with this staging buffer and the warp state
		cur = nxt;
	}

	// ---- block totals ---------------------------------------------------------------------------------
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		n_undet += __shfl_xor_sync(0xffffffffu, n_undet, o);
		n_conf += __shfl_xor_sync(0xffffffffu, n_conf, o);
		n_invalid += __shfl_xor_sync(0xffffffffu, n_invalid, o);
	}
	unsigned long long probes64 = n_probes, cand64 = n_cand, leaf64 = n_leaf_hits, chain64 = n_chained;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		probes64 += __shfl_xor_sync(0xffffffffu, probes64, o);
		cand64 += __shfl_xor_sync(0xffffffffu, cand64, o);
		leaf64 += __shfl_xor_sync(0xffffffffu, leaf64, o);
		chain64 += __shfl_xor_sync(0xffffffffu, chain64, o);
	}
	if (lane == 0) {
		if (n_undet) atomicAdd(&block_tot[0], n_undet);
		if (n_conf) atomicAdd(&block_tot[1], n_conf);
		if (n_invalid) atomicAdd(&p.counts[ncnt + 2], n_invalid);
		if (probes64) atomicAdd(&p.probe_count[0], probes64);
		if (cand64) atomicAdd(&p.probe_count[1], cand64);
		if (leaf64) atomicAdd(&p.probe_count[2], leaf64);
		if (chain64) atomicAdd(&p.probe_count[3], chain64);
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

// ------------------------------------------------------------------- lookup roofline probes

__global__ void __launch_bounds__(256) random_sector_kernel(const TableSlot *table, uint64_t mask,
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
				loadBucket(table + 2 * (mixKey(j + seed) & mask), k0[u], r0[u], k1[u], r1[u]);
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
