// sm_100a kernels of the read-matching path.
//
//   pack_reads_kernel     A1: ASCII -> 2-bit codes, MSB-first 64-bit words, validity check
//   scan_reads_kernel     A2-A8: both-strand prefix probes of the merged table, trie descent,
//                         per-read leaf set -> decision -> counters
//   reduce_partials_kernel per-block genome-count partials -> 64-bit totals (no atomics)
//   random_sector_kernel  the measured lookup roofline (random 32-byte sector gather)
//
// Reference behaviour each piece reproduces: query.cpp:480-527 (scan), hashtrie.cpp:350-369
// (find64_p), query.cpp:529-636 / 964-1067 (decision), query.cpp:447-450 (reverse complement).
#ifndef CAMMIQ_SCAN_KERNELS_CUH
#define CAMMIQ_SCAN_KERNELS_CUH

#include <cstdint>
#include <cuda_runtime.h>

#include "flat_index.hpp"

namespace cammiq {

static const int kScanThreads = 256;
static const int kWarpsPerBlock = kScanThreads / 32;
static const int kMaxWordsPerRead = 8;   // 256 bases
static const int kHitCap = 64;           // per-warp shared-memory hit list; overflow spills to global
static const int kSpillCap = 1024;       // 2 tables x 2 strands x <=251 positions
static const int kProbeUnroll = 4;       // independent sector loads in flight per lane
static const uint32_t kMaxSmemGenomes = 8191; // 2*(G+1) u32 block counters must fit 64 KB

struct ScanParams {
	// index
	const TableSlot *table;
	uint64_t table_mask;
	const uint32_t *nodes_u, *nodes_d;
	const uint32_t *leaf_u_ref;
	const uint2 *leaf_d_ref;
	uint32_t h;
	uint32_t n_genomes;
	// reads
	const uint64_t *packed;
	const uint8_t *len;
	uint32_t words_per_read;
	uint64_t n_reads;
	// outputs
	int mode;
	int smem_counters;        // 1: block-private genome counters + partials, 0: global atomics
	uint32_t *partials;       // [gridDim.x][2*(G+1)]
	unsigned long long *counts; // [2*(G+1)+4]: cnt_u | cnt_d | nundet nconf n_invalid n_pair_records
	uint32_t *rcount_u, *rcount_d;
	unsigned long long *pair_records; // SC: (a<<32|b) per D_PAIR read
	uint32_t *spill;          // [total warps][kSpillCap]
	unsigned long long *probe_count; // [4]: probes, bucket hits, leaf hits, extra (chained) bucket loads
	// optional per-read outputs
	uint8_t *read_class;
	uint32_t *read_rid_a, *read_rid_b;
	uint32_t leaf_cap;
	uint32_t *read_nleaf_u, *read_nleaf_d, *read_leaf_u, *read_leaf_d;
};

struct PackParams {
	const uint8_t *bases;
	const uint64_t *offsets; // NULL: fixed stride
	uint64_t stride;
	const uint8_t *lengths;
	uint64_t n_reads;
	uint32_t words_per_read;
	uint32_t h;
	uint64_t *packed;
	uint8_t *len_out;
	unsigned long long *n_invalid;
};

// ------------------------------------------------------------------------------------ pack

__device__ __forceinline__ int baseCodeDev(uint32_t c) {
	// A/a=0 C/c=1 G/g=2 T/t=3, else -1 (query.cpp:1860-1883)
	uint32_t u = c & 0xDFu; // fold case
	int code = (u == 'A') ? 0 : (u == 'C') ? 1 : (u == 'G') ? 2 : (u == 'T') ? 3 : -1;
	return code;
}

__global__ void __launch_bounds__(256) pack_reads_kernel(PackParams p) {
	const int lane = threadIdx.x & 31;
	const uint64_t warp = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint64_t n_warps = ((uint64_t) gridDim.x * blockDim.x) >> 5;
	unsigned long long invalid = 0;
	for (uint64_t r = warp; r < p.n_reads; r += n_warps) {
		const uint8_t *src = p.bases + (p.offsets ? p.offsets[r] : r * p.stride);
		const uint32_t rl = p.lengths[r];
		uint32_t v = 0;
		bool bad = false;
#pragma unroll
		for (int k = 0; k < 8; k++) {
			uint32_t j = lane * 8 + k;
			int code = 0;
			if (j < rl) {
				code = baseCodeDev(src[j]);
				bad |= code < 0;
			}
			v = (v << 2) | (uint32_t) (code & 3);
		}
		// four lanes make one 64-bit word, first base most significant
		unsigned long long w = (unsigned long long) v << (16 * (3 - (lane & 3)));
		w |= __shfl_xor_sync(0xffffffffu, w, 1);
		w |= __shfl_xor_sync(0xffffffffu, w, 2);
		if ((lane & 3) == 0 && (uint32_t) (lane >> 2) < p.words_per_read)
			p.packed[r * p.words_per_read + (lane >> 2)] = w;
		bool ok = !__any_sync(0xffffffffu, bad) && rl >= p.h;
		if (lane == 0) {
			p.len_out[r] = ok ? (uint8_t) rl : 0;
			invalid += ok ? 0 : 1;
		}
	}
	if (lane == 0 && invalid)
		atomicAdd(p.n_invalid, invalid);
}

// ------------------------------------------------------------------------------------ scan

// 32 bytes = one sector = one prefix-table bucket, fetched with a single 256-bit load
// (LDG.E.256 on sm_100a), read-only path, no L1 allocation (every probe is a fresh sector).
__device__ __forceinline__ void loadBucket(const TableSlot *b, unsigned long long &k0,
		unsigned long long &r0, unsigned long long &k1, unsigned long long &r1) {
	asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
		: "=l"(k0), "=l"(r0), "=l"(k1), "=l"(r1) : "l"(b));
}

// bases [i, i+32) of a packed strand as one word, base i most significant
__device__ __forceinline__ unsigned long long window64(const unsigned long long *w, uint32_t i) {
	uint32_t q = i >> 5, r = (i & 31) * 2;
	unsigned long long hi = w[q], lo = w[q + 1];
	return r ? ((hi << r) | (lo >> (64 - r))) : hi;
}

__device__ __forceinline__ uint32_t baseAt(const unsigned long long *w, uint32_t j) {
	return (uint32_t) (w[j >> 5] >> (62 - 2 * (j & 31))) & 3u;
}

// reverse the 32 two-bit groups of x
__device__ __forceinline__ unsigned long long reverseGroups(unsigned long long x) {
	x = __brevll(x);
	return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
}

// Walk the CSR trie below a bucket root: find64_p's loop (hashtrie.cpp:356-366).
// ref = bucket root, the read continues at base index `next` of strand `w`, `remaining`
// bases are left.  Returns the leaf ref or kRefNone.
__device__ __forceinline__ uint32_t descend(uint32_t ref, const uint32_t *__restrict__ nodes,
		const unsigned long long *w, uint32_t next, uint32_t remaining) {
	while (ref != kRefNone && !(ref & kRefLeafTag)) {
		if (remaining == 0)
			return kRefNone;
		uint32_t code = baseAt(w, next);
		ref = __ldg(&nodes[4 * (size_t) (ref - 1) + code]);
		next++;
		remaining--;
	}
	return ref;
}

__device__ __forceinline__ unsigned long long warpMin64(unsigned long long v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
		v = t < v ? t : v;
	}
	return v;
}
__device__ __forceinline__ unsigned long long warpMax64(unsigned long long v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
		v = t > v ? t : v;
	}
	return v;
}

struct WarpScratch {
	unsigned long long fwd[kMaxWordsPerRead + 1];
	unsigned long long rev[kMaxWordsPerRead + 1];
	uint32_t hits[kHitCap];
};

__device__ __forceinline__ uint32_t hitAt(const WarpScratch &s, const uint32_t *spill, uint32_t i) {
	return i < (uint32_t) kHitCap ? s.hits[i] : spill[i - kHitCap];
}

template <int MODE>
__global__ void __launch_bounds__(kScanThreads) scan_reads_kernel(ScanParams p) {
	extern __shared__ uint32_t smem_counts[]; // [2*(G+1)] when p.smem_counters
	__shared__ WarpScratch scratch[kWarpsPerBlock];
	__shared__ unsigned long long block_tot[2]; // nundet, nconf

	const int lane = threadIdx.x & 31;
	const int wib = threadIdx.x >> 5;
	const uint32_t ncnt = 2 * (p.n_genomes + 1);
	if (p.smem_counters)
		for (uint32_t i = threadIdx.x; i < ncnt; i += blockDim.x)
			smem_counts[i] = 0;
	if (threadIdx.x < 2)
		block_tot[threadIdx.x] = 0;
	__syncthreads();

	WarpScratch &s = scratch[wib];
	const uint64_t warp = (uint64_t) blockIdx.x * kWarpsPerBlock + wib;
	const uint64_t n_warps = (uint64_t) gridDim.x * kWarpsPerBlock;
	uint32_t *spill = p.spill + warp * kSpillCap;
	const uint32_t h = p.h, wpr = p.words_per_read;
	const uint32_t lt_mask = (1u << lane) - 1;
	unsigned long long n_undet = 0, n_conf = 0, n_probes = 0;
	uint32_t n_bucket_hits = 0, n_leaf_hits = 0, n_chained = 0; // per lane

	for (uint64_t r = warp; r < p.n_reads; r += n_warps) {
		const uint32_t rl = p.len[r];
		uint32_t cls = CQ_CLASS_UNLABELED, rid_a = 0, rid_b = 0, nhits = 0;
		uint32_t distinct_u = 0, distinct_d = 0;

		if (rl >= h && rl > 0) {
			// ---- stage both strands of the read in shared memory ------------------------
			if (lane <= (int) wpr)
				s.fwd[lane] = lane < (int) wpr ? p.packed[r * wpr + lane] : 0ull;
			__syncwarp();
			if (lane <= (int) wpr) {
				// reverse-complement word `lane` = bases rl-32*lane-1 down to rl-32*lane-32
				unsigned long long x = 0;
				int start = (int) rl - 32 * (lane + 1);
				if (start >= 0)
					x = window64(s.fwd, (uint32_t) start);
				else if (start > -32)
					x = s.fwd[0] >> (2 * (uint32_t) (-start));
				x = reverseGroups(~x);
				// bases past the end of the strand read as 0
				int valid = (int) rl - 32 * lane;
				if (valid <= 0) x = 0;
				else if (valid < 32) x &= ~0ull << (64 - 2 * valid);
				s.rev[lane] = x;
			}
			__syncwarp();

			// ---- probes: every position of both strands (query.cpp:486-501, 512-527) ------
			const uint32_t npos = rl - h + 1, total = 2 * npos;
			n_probes += total;
			for (uint32_t base = 0; base < total; base += 32 * kProbeUnroll) {
				unsigned long long hv[kProbeUnroll], k0[kProbeUnroll], r0[kProbeUnroll],
					k1[kProbeUnroll], r1[kProbeUnroll];
				uint64_t bidx[kProbeUnroll];
#pragma unroll
				for (int u = 0; u < kProbeUnroll; u++) {
					uint32_t q = base + u * 32 + lane;
					if (q < total) {
						const unsigned long long *w = q >= npos ? s.rev : s.fwd;
						uint32_t pos = q >= npos ? q - npos : q;
						hv[u] = window64(w, pos) >> (64 - 2 * h);
						bidx[u] = mixKey(hv[u]) & p.table_mask;
						loadBucket(p.table + 2 * bidx[u], k0[u], r0[u], k1[u], r1[u]);
					}
				}
#pragma unroll
				for (int u = 0; u < kProbeUnroll; u++) {
					uint32_t q = base + u * 32 + lane;
					uint32_t leaf_u = kRefNone, leaf_d = kRefNone;
					if (base + u * 32 < total) { // warp-uniform
						if (q < total) {
							unsigned long long refs = 0;
							bool found = false;
							for (;;) {
								if (k0[u] == hv[u]) { refs = r0[u]; found = true; }
								else if (k1[u] == hv[u]) { refs = r1[u]; found = true; }
								if (found || k0[u] == kEmptyKey || k1[u] == kEmptyKey)
									break;
								bidx[u] = (bidx[u] + 1) & p.table_mask; // full bucket: next one
								n_chained++;
								loadBucket(p.table + 2 * bidx[u], k0[u], r0[u], k1[u], r1[u]);
							}
							if (found) {
								n_bucket_hits++;
								const unsigned long long *w = q >= npos ? s.rev : s.fwd;
								uint32_t pos = q >= npos ? q - npos : q;
								uint32_t next = pos + h, remaining = rl - h - pos;
								leaf_u = descend((uint32_t) refs, p.nodes_u, w, next, remaining);
								leaf_d = descend((uint32_t) (refs >> 32), p.nodes_d, w, next, remaining);
								n_leaf_hits += (leaf_u != kRefNone) + (leaf_d != kRefNone);
							}
						}
						// append hits (U entries keep bit 31 clear, D entries set it)
						unsigned mu = __ballot_sync(0xffffffffu, leaf_u != kRefNone);
						unsigned md = __ballot_sync(0xffffffffu, leaf_d != kRefNone);
						if (mu | md) {
							if (leaf_u != kRefNone) {
								uint32_t at = nhits + __popc(mu & lt_mask);
								uint32_t e = leaf_u & ~kRefLeafTag;
								if (at < (uint32_t) kHitCap) s.hits[at] = e;
								else if (at < (uint32_t) (kHitCap + kSpillCap)) spill[at - kHitCap] = e;
							}
							nhits += __popc(mu);
							if (leaf_d != kRefNone) {
								uint32_t at = nhits + __popc(md & lt_mask);
								uint32_t e = leaf_d | kRefLeafTag;
								if (at < (uint32_t) kHitCap) s.hits[at] = e;
								else if (at < (uint32_t) (kHitCap + kSpillCap)) spill[at - kHitCap] = e;
							}
							nhits += __popc(md);
						}
					}
				}
			}
			__syncwarp();
			if (nhits > (uint32_t) kHitCap)
				__threadfence_block(); // spilled entries are re-read by other lanes below

			// ---- decision from reductions over the hit list (query.cpp:529-636) ----------
			if (nhits > 0) {
				uint32_t min_r = 0xFFFFFFFFu, max_r = 0;
				unsigned long long min_p = ~0ull, max_p = 0;
				for (uint32_t c = 0; c < nhits; c += 32) {
					if (c + lane < nhits) {
						uint32_t e = hitAt(s, spill, c + lane);
						if (e & kRefLeafTag) {
							uint2 ab = __ldg(&p.leaf_d_ref[e & ~kRefLeafTag]);
							uint32_t lo = min(ab.x, ab.y), hi = max(ab.x, ab.y);
							unsigned long long key = ((unsigned long long) lo << 32) | hi;
							min_p = key < min_p ? key : min_p;
							max_p = key > max_p ? key : max_p;
						} else {
							uint32_t rid = __ldg(&p.leaf_u_ref[e]);
							min_r = min(min_r, rid);
							max_r = max(max_r, rid);
						}
					}
				}
				min_r = __reduce_min_sync(0xffffffffu, min_r);
				max_r = __reduce_max_sync(0xffffffffu, max_r);
				min_p = warpMin64(min_p);
				max_p = warpMax64(max_p);
				const int nr = (min_r == 0xFFFFFFFFu) ? 0 : (min_r == max_r ? 1 : 2);
				const int np = (min_p == ~0ull) ? 0 : (min_p == max_p ? 1 : 2);
				const uint32_t a0 = (uint32_t) (min_p >> 32), b0 = (uint32_t) min_p;
				if (np == 0) {
					if (nr == 1) { cls = CQ_CLASS_U; rid_a = min_r; }
					else cls = CQ_CLASS_CONFLICT; // nr >= 2 (nr == 0 impossible with hits)
				} else if (np == 1) {
					if (nr == 0) { cls = CQ_CLASS_D_PAIR; rid_a = a0; rid_b = b0; }
					else if (nr == 2) cls = CQ_CLASS_CONFLICT;
					else if (a0 != min_r && b0 != min_r) cls = CQ_CLASS_CONFLICT;
					else { cls = CQ_CLASS_UD; rid_a = min_r; }
				} else if (nr == 2) {
					cls = CQ_CLASS_CONFLICT;
				} else {
					// second pass over the pairs: does every pair contain r (nr == 1), or which
					// of a0 / b0 lies in every pair (nr == 0)
					bool all_a = true, all_b = true;
					const uint32_t ta = nr == 1 ? min_r : a0, tb = nr == 1 ? min_r : b0;
					for (uint32_t c = 0; c < nhits; c += 32) {
						if (c + lane < nhits) {
							uint32_t e = hitAt(s, spill, c + lane);
							if (e & kRefLeafTag) {
								uint2 ab = __ldg(&p.leaf_d_ref[e & ~kRefLeafTag]);
								all_a &= (ab.x == ta || ab.y == ta);
								all_b &= (ab.x == tb || ab.y == tb);
							}
						}
					}
					all_a = __all_sync(0xffffffffu, all_a);
					all_b = __all_sync(0xffffffffu, all_b);
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

			// ---- distinct leaves: rcount (+1 per distinct leaf of an accepted read) --------
			const bool accepted = cls >= CQ_CLASS_U;
			const bool want_sets = p.read_nleaf_u != NULL;
			if (nhits > 0 && ((MODE == CQ_MODE_P && accepted) || want_sets)) {
				for (uint32_t c = 0; c < nhits; c += 32) {
					const bool have = c + lane < nhits;
					uint32_t e = have ? hitAt(s, spill, c + lane) : 0;
					unsigned act = __ballot_sync(0xffffffffu, have);
					bool leader = false;
					if (have) {
						unsigned grp = __match_any_sync(act, e);
						leader = (__ffs(grp) - 1) == lane;
						// seen in an earlier chunk?
						for (uint32_t j = 0; j < c && leader; j++)
							if (hitAt(s, spill, j) == e) leader = false;
					}
					const bool is_d = (e & kRefLeafTag) != 0;
					const uint32_t leaf = e & ~kRefLeafTag;
					if (MODE == CQ_MODE_P && accepted && leader)
						atomicAdd(is_d ? &p.rcount_d[leaf] : &p.rcount_u[leaf], 1u);
					unsigned mu = __ballot_sync(0xffffffffu, leader && !is_d);
					unsigned md = __ballot_sync(0xffffffffu, leader && is_d);
					if (want_sets && leader) {
						uint32_t at = is_d ? distinct_d + __popc(md & lt_mask) : distinct_u + __popc(mu & lt_mask);
						if (at < p.leaf_cap)
							(is_d ? p.read_leaf_d : p.read_leaf_u)[r * p.leaf_cap + at] = leaf;
					}
					distinct_u += __popc(mu);
					distinct_d += __popc(md);
				}
			}
		}

		// ---- counters (query.cpp:542-636 effects) --------------------------------------------
		if (lane == 0) {
			const uint32_t G1 = p.n_genomes + 1;
			const bool inc_u = cls == CQ_CLASS_U || cls == CQ_CLASS_UD || (cls == CQ_CLASS_D_INTER && MODE == CQ_MODE_SC);
			const bool inc_d = cls >= CQ_CLASS_D_PAIR;
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

	n_bucket_hits = __reduce_add_sync(0xffffffffu, n_bucket_hits);
	n_leaf_hits = __reduce_add_sync(0xffffffffu, n_leaf_hits);
	n_chained = __reduce_add_sync(0xffffffffu, n_chained);
	if (lane == 0) {
		if (n_undet) atomicAdd(&block_tot[0], n_undet);
		if (n_conf) atomicAdd(&block_tot[1], n_conf);
		if (n_probes) atomicAdd(&p.probe_count[0], n_probes);
		if (n_bucket_hits) atomicAdd(&p.probe_count[1], (unsigned long long) n_bucket_hits);
		if (n_leaf_hits) atomicAdd(&p.probe_count[2], (unsigned long long) n_leaf_hits);
		if (n_chained) atomicAdd(&p.probe_count[3], (unsigned long long) n_chained);
	}
	__syncthreads();
	if (p.smem_counters) {
		uint32_t *dst = p.partials + (size_t) blockIdx.x * ncnt;
		for (uint32_t i = threadIdx.x; i < ncnt; i += blockDim.x)
			dst[i] = smem_counts[i];
	}
	if (threadIdx.x < 2 && block_tot[threadIdx.x])
		atomicAdd(&p.counts[ncnt + threadIdx.x], block_tot[threadIdx.x]);
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

// ------------------------------------------------------------------- lookup roofline probe

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
