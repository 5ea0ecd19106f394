// Flattened, device-ready form of the two CAMMiQ indices.
//
//   prefix table   open addressing over 32-byte buckets (= one DRAM/L2 sector), two slots per
//                  bucket, keys first: {key0, key1, (u_ref0, d_ref0), (u_ref1, d_ref1)}.  U and D
//                  share the hash length (query.cpp:460), so ONE probe answers both tables, and
//                  the 16 bytes of keys alone decide whether a read position is a candidate
//                  (what phase 1 of the scan loads when no filter fronts the table).  Linear
//                  probing by bucket; bit 63 of key0 is the bucket's OVERFLOW flag: some key whose
//                  probe sequence passed this bucket lies further on.  A lookup stops at the first
//                  bucket that holds the key or whose flag is clear, so a miss costs one sector
//                  unless a key really spilled past the bucket (2 % of buckets at load 0.30).
//   trie nodes     per table, 16 bytes per node; only buckets whose root is not already a leaf
//                  have any (rare when h == k).  A node with several children holds its 4 child
//                  refs; a run of single-child nodes -- what almost every key longer than h is --
//                  is ONE chain node {tag | length, up to 32 bases, next ref}: a lookup compares
//                  the read's next bases with it in one step instead of walking one dependent
//                  load per base (Hash::find64_p's loop, hashtrie.cpp:356-366, on compressed paths).
//   leaf refs      per table, the genome id(s) the classification needs: u32 for U,
//                  {u32,u32} for D.  The remaining leaf fields (ucount, depth) stay on the
//                  host for the ILP set-up.
#ifndef CAMMIQ_FLAT_INDEX_HPP
#define CAMMIQ_FLAT_INDEX_HPP

#include <cstdint>
#include <cstdlib>
#include <sys/mman.h>
#include <string>
#include <vector>

#include "index_codec.hpp"

namespace cammiq {

struct TableBucket {
	uint64_t key[2];    // 0 = free slot, else kKeyOccupied | h-mer (h <= 31: 62 bits); key[0] bit 63 = overflow flag
	uint32_t ref[2][2]; // [slot][0 = U root, 1 = D root]
};
static const uint64_t kKeyOccupied = 1ull << 62;
static const uint64_t kBucketOverflow = 1ull << 63;
static const uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull; // free entry of the SC pair table (scan_kernels.cuh)
static const int kSlotsPerBucket = 2;

// Bucket index of a key; the same function runs on the host (flatten, find_host) and in the
// scan kernel.
#if defined(__CUDACC__)
__host__ __device__
#endif
inline uint64_t mixKey(uint64_t x) {
	x ^= x >> 31;
	x *= 0x9E3779B97F4A7C15ull;
	x ^= x >> 29;
	x *= 0xBF58476D1CE4E5B9ull;
	x ^= x >> 32;
	return x;
}

// L2-resident membership filter in front of the table: register-blocked Bloom filter, one
// 64-bit word per key, four bits of it set.  ~98-99% of probes miss; a
// filter that fits the B200's L2 (random gathers over <= 64 MB run at ~285 G/s versus ~45 G
// sectors/s from HBM, profiles/r01_microbench_gather.json) answers them without touching DRAM.
// The filter is keyed by the CANONICAL h-mer, min(key, reverse complement of key): the scan holds
// both strands' hashes of a window anyway (hf, hr), so ONE probe and ONE test per read position
// answer both strands; which orientation is the key is sorted out by the table probe of the few
// positions that pass (the bucket of a key and of its reverse complement is the same one).
// The hash runs once per read position and is deliberately lean: two independent multiply-add
// hashes of the key's 32-bit halves (4 IMAD), the word index from the high bits of the first,
// the bit selectors from the second.
// Two regimes.  With at least kFilterMinBitsPerKey bits per key the filter is selective (a few
// false positives per read) and a position that passes goes straight to the table probe of
// phase 2.  An index too large for that (cfg4: 3.5e8 keys) still gets a 64 MB filter, used as a
// SIEVE: fewer bits per key are set (filter_sel_mask switches selectors off) and a position
// that passes -- about half of them at 1.5 bits per key -- loads its table bucket's keys right
// in phase 1.  The sieve halves the HBM accesses of the regime that is bound by them.
// 48 MB, not the 64 MB that a lone gather kernel still finds L2-resident: next to the scan's other
// traffic a 64 MB filter misses L2 on more than half of its probes (24 GB of DRAM traffic per 10M
// reads against 15 GB at 40-48 MB, profiles/r02_filter_size_sweep.json); the extra false positives
// of the smaller filter (11 instead of 15 bits per key at cfg2) cost less than the misses.
static const uint64_t kFilterMaxBytesDefault = 48ull << 20;
static const uint32_t kFilterMinBitsPerKey = 8;

// reverse complement of a 2-bit packed h-mer (first base most significant): complement, then
// reverse the 32 two-bit groups (bytes, nibbles, pairs) and drop the unused low groups
inline uint64_t revcompKeyHost(uint64_t key, uint32_t h) {
	uint64_t x = __builtin_bswap64(~key);
	x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
	x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
	return x >> (64 - 2 * h);
}
inline uint64_t canonicalKeyHost(uint64_t key, uint32_t h) {
	uint64_t rc = revcompKeyHost(key, h);
	return key < rc ? key : rc;
}


#if defined(__CUDACC__)
__host__ __device__
#endif
inline void filterHash(uint64_t key, uint32_t &A, uint32_t &B) {
	const uint32_t lo = (uint32_t) key, hi = (uint32_t) (key >> 32);
	A = lo * 0x9E3779B1u + hi * 0x85EBCA77u;
	B = lo * 0xC2B2AE3Du + hi * 0x27D4EB2Fu;
}
// Home bucket of a key.  Keys are stored as they are, but PLACED by their canonical h-mer: a key
// and its reverse complement share one probe sequence, so the scan -- which holds both strands'
// hashes of a window -- finds either with ONE bucket load per read position.  The bucket is a
// multiplicative hash of the filter hash B the scan has computed for the position anyway: the
// moment a position passes the filter its bucket address costs two more instructions, cheap
// enough to request the sector from HBM right there (prefetch), long before phase 2 needs it.
// table_shift = 32 - log2(number of buckets); the table has between 2^6 and 2^32 buckets.
#if defined(__CUDACC__)
__host__ __device__
#endif
inline uint64_t tableBucket(uint32_t B, uint32_t table_shift) { return (uint64_t) ((B * 0x9E3779B1u) >> table_shift); }
inline uint64_t homeBucketHost(uint64_t key, uint32_t h, uint32_t table_shift) {
	uint32_t A, B;
	filterHash(canonicalKeyHost(key, h), A, B);
	return tableBucket(B, table_shift);
}
// word index for a filter of `words` words (any count below 2^32): the high half of A * words
#if defined(__CUDACC__)
__host__ __device__
#endif
inline uint32_t filterWordIndex(uint32_t A, uint32_t words) {
#if defined(__CUDA_ARCH__)
	return __umulhi(A, words);
#else
	return (uint32_t) (((uint64_t) A * words) >> 32);
#endif
}
// B selects four bits of the 64-bit word as (byte, bit) pairs: byte indices = the four nibbles
// of B & 0x7777, bits inside those bytes = the four nibbles of (B >> 16) & 0x7777.  On the device
// that is two byte permutes: one gathers the four selected bytes of the word, one builds the four
// one-bit masks from a constant table -- no variable shifts.
// `sel` = 0x77777777 uses all four pairs; a sieve with fewer bits per key zeroes nibbles of it
// (0x00770077: two pairs, 0x00070007: one), which turns the unused pairs into (byte 0, bit 0) -- a bit
// the builder then sets in every word.
static const uint32_t kFilterSelAll = 0x77777777u;
inline uint64_t filterMask(uint32_t B, uint32_t sel = kFilterSelAll) {
	uint64_t m = 0;
	B &= sel;
	for (int i = 0; i < 4; i++)
		m |= 1ull << (8 * ((B >> (4 * i)) & 7u) + ((B >> (16 + 4 * i)) & 7u));
	return m;
}
#if defined(__CUDACC__)
__host__ __device__
#endif
inline bool filterTest(uint32_t x, uint32_t y, uint32_t B, uint32_t sel = kFilterSelAll) {
#if defined(__CUDA_ARCH__)
	const uint32_t bytes = __byte_perm(x, y, B & sel & 0xFFFFu);
	const uint32_t bits = __byte_perm(0x08040201u, 0x80402010u, (B & sel) >> 16);
	return (~bytes & bits) == 0u;
#else
	const uint64_t m = filterMask(B, sel), w = (uint64_t) x | ((uint64_t) y << 32);
	return (w & m) == m;
#endif
}

// Uninitialised, owning array (the multi-GB table is first-touched by the build threads
// instead of being value-initialised by one)
template <typename T>
struct RawArray {
	T *p = NULL;
	size_t n = 0;
	RawArray() {}
	RawArray(const RawArray &) = delete;
	RawArray &operator=(const RawArray &) = delete;
	~RawArray() { free(p); }
	bool alloc(size_t count) {
		free(p);
		p = NULL;
		n = 0;
		// 2 MB alignment + MADV_HUGEPAGE: a multi-GB table probed at random is TLB-bound with
		// 4 KB pages, and first-touching it takes one fault per page
		const size_t bytes = sizeof(T) * (count ? count : 1), huge = (size_t) 2 << 20;
		void *q = NULL;
		if (bytes >= 4 * huge) {
			if (posix_memalign(&q, huge, (bytes + huge - 1) / huge * huge) != 0)
				q = NULL;
#ifdef MADV_HUGEPAGE
			if (q != NULL)
				madvise(q, (bytes + huge - 1) / huge * huge, MADV_HUGEPAGE);
#endif
		} else
			q = malloc(bytes);
		if (q == NULL)
			return false;
		p = (T *) q;
		n = count;
		return true;
	}
	size_t size() const { return n; }
	T *data() { return p; }
	const T *data() const { return p; }
	T &operator[](size_t i) { return p[i]; }
	const T &operator[](size_t i) const { return p[i]; }
};

// Chain node: word 0 = kChainTag | number of bases (1..32), words 1-2 = the bases (2-bit codes,
// first base most significant, right-aligned in 64 bits: word 1 = high half), word 3 = the ref
// that follows the chain.  Leaf refs carry 0x80000000 | id, so the tag needs ids below 2^30.
static const uint32_t kChainTag = 0xC0000000u;
static const uint32_t kChainMaxBases = 32;
static const uint64_t kMaxLeavesPerTable = 1ull << 30;

struct FlatIndex {
	uint32_t hash_len = 0;
	uint64_t n_table_buckets = 0; // power of two, 2^6 .. 2^32
	uint32_t table_shift = 26;    // 32 - log2(n_table_buckets)
	uint64_t n_keys = 0;
	RawArray<TableBucket> table;  // n_table_buckets
	std::vector<uint64_t> filter; // any multiple of 128 words, empty = no filter (index too large for L2)
	uint32_t filter_words = 0;    // = filter.size()
	uint32_t filter_sel_mask = kFilterSelAll; // selector pairs in use (see filterTest)
	bool filter_sieve = false;    // too few bits per key to be selective: passing positions probe the table in phase 1
	DecodedIndex u, d;            // leaves (file order) + trie nodes + buckets of each table
	// path-compressed tries (what the device and flatFind walk); the table refs point into these
	FlatVec<uint32_t>::type cnodes_u, cnodes_d;
	double decode_ms = 0, flatten_ms = 0;

	uint64_t deviceBytes() const;
};

// Build the merged prefix table from the two decoded indices (moved into out).
int flattenIndices(DecodedIndex &u, DecodedIndex &d, double load_factor, FlatIndex &out, std::string &err);

// (Re)build the membership filter with at most max_bytes (0 = no filter).
void buildFilter(FlatIndex &fi, uint64_t max_bytes);

// Hash::find64_p on the flattened layout (host; layout verification only).
uint64_t flatFind(const FlatIndex &fi, int table, uint64_t bucket, const uint8_t *cand, size_t len);

} // namespace cammiq
#endif
