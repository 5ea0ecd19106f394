// sm_100a kernels of the read-matching path.
//
//   scan_reads_kernel      A1-A8 in one launch: a CTA pulls a tile of 256 ASCII reads into
//                          shared memory with one TMA bulk copy (cp.async.bulk + mbarrier),
//                          then
//                            phase 1 (thread per read): decode each base to its 2-bit code,
//                              roll the h-base prefix hash of BOTH strands in registers and
//                              test every position against the L2-resident membership filter
//                              (or the prefix table itself when the index is too large for a
//                              filter); positives go to a per-warp queue,
//                            phase 2 (warp-cooperative): queued candidates probe the prefix
//                              table in HBM (one 32-byte sector = one bucket) and descend the
//                              CSR trie; leaves land in per-read hit lists,
//                            phase 3 (thread per read): leaf set -> decision -> counters.
//   reduce_partials_kernel per-block genome-count partials -> 64-bit totals (no atomics)
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

static const int kScanThreads = 256;          // = reads per tile
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
	uint64_t filter_mask;     // words - 1
	const uint32_t *nodes_u, *nodes_d;
	const uint32_t *leaf_u_ref;
	const uint2 *leaf_d_ref;
	uint32_t h;
	uint32_t n_genomes;
	// reads: ASCII in device memory, exactly the state query64_* consumes
	const uint8_t *bases;
	const uint64_t *offsets;  // NULL: read i starts at (read_base + i)*stride
	uint64_t stride;
	uint64_t read_base;       // caller's index of this launch's first read (chunked submission)
	const uint8_t *lengths;
	uint64_t n_reads;
	uint32_t tile_cap;        // bytes of shared memory reserved for the ASCII tile
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
	uint32_t queue[kQueueCap];     // slot<<16 | strand<<15 | position
	uint32_t hits[32][kHitSeg];    // table<<31 | leaf id
	const uint8_t *sptr[32];       // first base of the slot's read (shared or global)
	uint32_t hit_cnt[32];
	uint32_t rl[32];
	uint32_t q_count;
};

// base `j` of strand `strand` of a read (strand 1 = reverse complement, query.cpp:447-450)
__device__ __forceinline__ uint32_t strandBase(const uint8_t *s, uint32_t rl, uint32_t strand, uint32_t j) {
	bool bad = false;
	uint32_t c = strand ? 3u - decodeBase(s[rl - 1 - j], bad) : decodeBase(s[j], bad);
	return c;
}

// Walk the CSR trie below a bucket root: find64_p's loop (hashtrie.cpp:356-366).
__device__ __forceinline__ uint32_t descend(uint32_t ref, const uint32_t *__restrict__ nodes,
		const uint8_t *s, uint32_t rl, uint32_t strand, uint32_t next) {
	while (ref != kRefNone && !(ref & kRefLeafTag)) {
		if (next >= rl)
			return kRefNone;
		uint32_t code = strandBase(s, rl, strand, next);
		ref = loadStreamU32(&nodes[4 * (size_t) (ref - 1) + code]);
		next++;
	}
	return ref;
}

// Phase 2: the warp drains its candidate queue.  One candidate per lane: recompute the h-mer,
// probe the prefix table (HBM), descend, append leaves to the owning read's hit list.
__device__ __forceinline__ void drainQueue(const ScanParams &p, WarpState &ws, uint32_t *warp_spill, int lane,
		uint32_t &n_leaf_hits, uint32_t &n_chained) {
	__syncwarp();
	const uint32_t nq = ws.q_count;
	const uint32_t h = p.h;
	for (uint32_t base = 0; base < nq; base += 32) {
		const uint32_t k = base + lane;
		if (k < nq) {
			const uint32_t item = ws.queue[k];
			const uint32_t slot = item >> 16, strand = (item >> 15) & 1u, pos = item & 0x7FFFu;
			const uint8_t *s = ws.sptr[slot];
			const uint32_t rl = ws.rl[slot];
			unsigned long long hv = 0;
			for (uint32_t t = 0; t < h; t++)
				hv = (hv << 2) | strandBase(s, rl, strand, pos + t);
			uint64_t b = mixKey(hv) & p.table_mask;
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
				uint32_t leaf[2];
				leaf[0] = descend((uint32_t) refs, p.nodes_u, s, rl, strand, pos + h);
				leaf[1] = descend((uint32_t) (refs >> 32), p.nodes_d, s, rl, strand, pos + h);
#pragma unroll
				for (int t = 0; t < 2; t++) {
					if (leaf[t] == kRefNone)
						continue;
					n_leaf_hits++;
					uint32_t e = (leaf[t] & ~kRefLeafTag) | (t ? kRefLeafTag : 0u);
					uint32_t at = atomicAdd(&ws.hit_cnt[slot], 1u);
					if (at < (uint32_t) kHitSeg) ws.hits[slot][at] = e;
					else if (at < (uint32_t) (kHitSeg + kHitSpill)) warp_spill[(size_t) slot * kHitSpill + at - kHitSeg] = e;
				}
			}
		}
	}
	__syncwarp();
	if (lane == 0)
		ws.q_count = 0;
	__syncwarp();
}

template <int MODE, bool FILTER>
__global__ void __launch_bounds__(kScanThreads, 2) scan_reads_kernel(ScanParams p) {
	extern __shared__ __align__(128) uint8_t dyn_smem[]; // [tile_cap ASCII tile][2*(G+1) u32 counters]
	__shared__ WarpState warp_state[kWarpsPerBlock];
	__shared__ __align__(8) uint64_t tile_bar;
	__shared__ unsigned long long block_tot[2]; // nundet, nconf
	__shared__ unsigned long long span_lo[kWarpsPerBlock], span_hi[kWarpsPerBlock];

	uint8_t *tile = dyn_smem;
	uint32_t *smem_counts = reinterpret_cast<uint32_t *>(dyn_smem + p.tile_cap);
	const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
	const uint32_t G1 = p.n_genomes + 1, ncnt = 2 * G1;
	if (p.smem_counters)
		for (uint32_t i = tid; i < ncnt; i += blockDim.x)
			smem_counts[i] = 0;
	if (tid < 2)
		block_tot[tid] = 0;
	if (tid == 0)
		mbarInit(&tile_bar, 1);
	WarpState &ws = warp_state[wib];
	if (lane == 0)
		ws.q_count = 0;
	__syncthreads();

	const uint32_t h = p.h;
	const unsigned long long pol_keep = policyEvictLast(), pol_stream = policyEvictFirst();
	const unsigned long long kmask = ~0ull >> (64 - 2 * h);
	const uint32_t top_shift = 2 * h - 2;
	uint32_t *warp_spill = p.hit_spill + ((size_t) blockIdx.x * kWarpsPerBlock + wib) * 32 * kHitSpill;
	unsigned long long n_undet = 0, n_conf = 0, n_invalid = 0;
	uint32_t n_probes = 0, n_cand = 0, n_leaf_hits = 0, n_chained = 0; // per lane
	uint32_t bar_parity = 0;
	const uint64_t n_tiles = (p.n_reads + kScanThreads - 1) / kScanThreads;

	for (uint64_t tile_idx = blockIdx.x; tile_idx < n_tiles; tile_idx += gridDim.x) {
		const uint64_t r = tile_idx * kScanThreads + tid;
		const bool have = r < p.n_reads;
		const uint64_t off = have ? (p.offsets ? p.offsets[r] : (p.read_base + r) * p.stride) : ~0ull;
		uint32_t rl = have ? p.lengths[r] : 0;

		// ---- stage the tile: one TMA bulk copy of the byte range the 256 reads span -------------
		unsigned long long lo = have ? off : ~0ull, hi = have ? off + rl : 0ull;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			unsigned long long tl = __shfl_xor_sync(0xffffffffu, lo, o), th = __shfl_xor_sync(0xffffffffu, hi, o);
			lo = tl < lo ? tl : lo;
			hi = th > hi ? th : hi;
		}
		if (lane == 0) {
			span_lo[wib] = lo;
			span_hi[wib] = hi;
		}
		__syncthreads(); // also: every thread is done with the previous tile
		lo = span_lo[0];
		hi = span_hi[0];
#pragma unroll
		for (int w = 1; w < kWarpsPerBlock; w++) {
			lo = span_lo[w] < lo ? span_lo[w] : lo;
			hi = span_hi[w] > hi ? span_hi[w] : hi;
		}
		const unsigned long long start = lo & ~15ull;
		const unsigned long long bytes = hi > start ? ((hi - start + 15ull) & ~15ull) : 0ull;
		const bool staged = bytes > 0 && bytes <= p.tile_cap; // else: reads are fetched from global directly
		if (staged) {
			if (tid == 0) {
				mbarExpectTx(&tile_bar, (uint32_t) bytes);
				bulkCopyG2S(tile, p.bases + start, (uint32_t) bytes, &tile_bar, pol_stream);
			}
			mbarWait(&tile_bar, bar_parity);
			bar_parity ^= 1u;
		}
		__syncthreads(); // span_lo/hi may be rewritten by the next iteration only after all read them
		const uint8_t *s = have ? (staged ? tile + (off - start) : p.bases + off) : tile;
		ws.sptr[lane] = s;
		ws.rl[lane] = rl;
		ws.hit_cnt[lane] = 0;
		__syncwarp();

		// ---- phase 1: thread per read, rolling hashes of both strands, filter / table test ------
		bool bad = false;
		unsigned long long hf = 0, hr = 0;
		const uint32_t wmax = __reduce_max_sync(0xffffffffu, rl);
		for (uint32_t j = 0; j + 1 < h; j++) {
			if (j < rl) {
				uint32_t c = decodeBase(s[j], bad);
				hf = (hf << 2) | c;
				hr = (hr >> 2) | ((unsigned long long) (3u - c) << top_shift);
			}
		}
		for (uint32_t j0 = h - 1; j0 < wmax; j0 += kStepUnroll) {
			unsigned long long kf[kStepUnroll], kr[kStepUnroll];
			uint2 ff[kStepUnroll], fr[kStepUnroll];              // FILTER: filter words
			unsigned long long bf[kStepUnroll][4], br[kStepUnroll][4]; // !FILTER: table buckets
			uint32_t mf[kStepUnroll][2], mr[kStepUnroll][2];
#pragma unroll
			for (int u = 0; u < kStepUnroll; u++) {
				const uint32_t j = j0 + u;
				if (j < rl) {
					uint32_t c = decodeBase(s[j], bad);
					hf = ((hf << 2) | c) & kmask;
					hr = (hr >> 2) | ((unsigned long long) (3u - c) << top_shift);
					kf[u] = hf;
					kr[u] = hr;
					const unsigned long long xf = mixKey(hf), xr = mixKey(hr);
					if (FILTER) {
						uint64_t wf, wr;
						filterProbe(xf, p.filter_mask, wf, mf[u][0], mf[u][1]);
						filterProbe(xr, p.filter_mask, wr, mr[u][0], mr[u][1]);
						ff[u] = loadFilterWord(p.filter + wf, pol_keep);
						fr[u] = loadFilterWord(p.filter + wr, pol_keep);
					} else {
						loadBucket(p.table + 2 * (xf & p.table_mask), bf[u][0], bf[u][1], bf[u][2], bf[u][3]);
						loadBucket(p.table + 2 * (xr & p.table_mask), br[u][0], br[u][1], br[u][2], br[u][3]);
					}
				}
			}
#pragma unroll
			for (int u = 0; u < kStepUnroll; u++) {
				const uint32_t j = j0 + u;
				bool cand_f = false, cand_r = false;
				if (j < rl) {
					n_probes += 2;
					if (FILTER) {
						cand_f = ((ff[u].x & mf[u][0]) == mf[u][0]) && ((ff[u].y & mf[u][1]) == mf[u][1]);
						cand_r = ((fr[u].x & mr[u][0]) == mr[u][0]) && ((fr[u].y & mr[u][1]) == mr[u][1]);
					} else {
						// candidate = the bucket holds the key, or is full and the key may have spilled
						cand_f = bf[u][0] == kf[u] || bf[u][2] == kf[u] || (bf[u][0] != kEmptyKey && bf[u][2] != kEmptyKey);
						cand_r = br[u][0] == kr[u] || br[u][2] == kr[u] || (br[u][0] != kEmptyKey && br[u][2] != kEmptyKey);
					}
				}
				if (cand_f) {
					// forward strand, position i = j-h+1
					uint32_t at = atomicAdd(&ws.q_count, 1u);
					ws.queue[at] = ((uint32_t) lane << 16) | (j + 1 - h);
					n_cand++;
				}
				if (cand_r) {
					// reverse-complement strand: this window is rc position rl-1-j
					uint32_t at = atomicAdd(&ws.q_count, 1u);
					ws.queue[at] = ((uint32_t) lane << 16) | 0x8000u | (rl - 1 - j);
					n_cand++;
				}
			}
			__syncwarp();
			// at most 2*kStepUnroll*32 = 256 candidates arrive per iteration: drain above half
			if (ws.q_count > (uint32_t) (kQueueCap - 2 * kStepUnroll * 32))
				drainQueue(p, ws, warp_spill, lane, n_leaf_hits, n_chained);
		}
		drainQueue(p, ws, warp_spill, lane, n_leaf_hits, n_chained);

		// ---- phase 3: thread per read: leaf set -> decision (query.cpp:529-636) ------------------
		uint32_t cls = CQ_CLASS_UNLABELED, rid_a = 0, rid_b = 0, distinct_u = 0, distinct_d = 0;
		const bool valid = have && !bad && rl >= h;
		const uint32_t nh = valid ? min(ws.hit_cnt[lane], (uint32_t) (kHitSeg + kHitSpill)) : 0;
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
		__syncwarp();
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
