#include "flat_index.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <new>
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "../../include/cammiq_gpu.h"

namespace cammiq {

uint64_t FlatIndex::deviceBytes() const {
	return table.size() * sizeof(TableBucket) + filter.size() * 8 + (cnodes_u.size() + cnodes_d.size()) * 4 +
		u.numLeaves() * 4 + d.numLeaves() * 8 + (u.numLeaves() + d.numLeaves()) * 4;
}

namespace {

unsigned flattenThreads() {
	unsigned n = std::thread::hardware_concurrency();
	return std::max(1u, std::min(n, 16u));
}

// A bucket slot: where a key's two root refs live.
struct SlotRef {
	TableBucket *b;
	int k;
};

// Find or place `key` in bucket b.  Returns the slot index, or -1 when the bucket is full of
// other keys (the caller raises the bucket's overflow flag and moves on).
inline int bucketInsert(TableBucket &b, uint64_t key, bool &fresh) {
	const uint64_t tag = key | kKeyOccupied;
	for (int k = 0; k < kSlotsPerBucket; k++) {
		const uint64_t have = b.key[k] & ~kBucketOverflow;
		if (have == tag) {
			fresh = false;
			return k;
		}
		if (have == 0) {
			b.key[k] |= tag;
			fresh = true;
			return k;
		}
	}
	return -1;
}

// Insert (or find) key; returns the slot.  Linear probing by bucket; every full bucket passed
// on the way gets its overflow flag, so lookups know to look further.
inline SlotRef probeInsert(RawArray<TableBucket> &t, uint64_t mask, uint32_t shift, uint32_t h, uint64_t key, bool &fresh) {
	uint64_t b = homeBucketHost(key, h, shift);
	for (;;) {
		const int k = bucketInsert(t[b], key, fresh);
		if (k >= 0) {
			SlotRef r = {&t[b], k};
			return r;
		}
		t[b].key[0] |= kBucketOverflow;
		b = (b + 1) & mask;
	}
}

// ---- path compression of the decoded tries ------------------------------------------------------
// Runs of single-child nodes become chain nodes (flat_index.hpp).  Buckets are independent, so the
// bucket range is cut into one piece per thread; a counting pass sizes every piece's share of the
// output, a second pass writes it.
struct Compressor {
	const FlatVec<uint32_t>::type &nodes; // decoded: 4 child refs per node
	uint32_t *out;                        // compressed nodes (NULL: count only)
	uint64_t next;                        // next free compressed node

	uint32_t run(uint32_t ref) {
		if (ref == kRefNone || refIsLeaf(ref))
			return ref;
		// follow single-child nodes
		uint64_t bases = 0;
		uint32_t len = 0, cur = ref;
		while (cur != kRefNone && !refIsLeaf(cur) && len < kChainMaxBases) {
			const uint32_t *ch = &nodes[4 * (size_t) refNodeId(cur)];
			int only = -1, n = 0;
			for (int c = 0; c < 4; c++)
				if (ch[c] != kRefNone) {
					only = c;
					n++;
				}
			if (n != 1)
				break;
			bases = (bases << 2) | (uint64_t) only;
			len++;
			cur = ch[only];
		}
		const uint64_t at = next++;
		if (len > 0) {
			const uint32_t follow = run(cur);
			if (out != NULL) {
				out[4 * at + 0] = kChainTag | len;
				out[4 * at + 1] = (uint32_t) (bases >> 32);
				out[4 * at + 2] = (uint32_t) bases;
				out[4 * at + 3] = follow;
			}
		} else {
			const uint32_t *ch = &nodes[4 * (size_t) refNodeId(ref)];
			uint32_t kids[4];
			for (int c = 0; c < 4; c++)
				kids[c] = run(ch[c]);
			if (out != NULL)
				for (int c = 0; c < 4; c++)
					out[4 * at + c] = kids[c];
		}
		return (uint32_t) at + 1;
	}
};

// x.nodes / x.bucket_root -> cnodes + compressed roots (one per bucket, file order)
int compressTries(const DecodedIndex &x, FlatVec<uint32_t>::type &cnodes, FlatVec<uint32_t>::type &croot, std::string &err) {
	if (x.numLeaves() > kMaxLeavesPerTable) {
		err = "Index too large: more than 2^30 leaves in one table.";
		return CQ_ENOMEM;
	}
	const size_t nb = x.bucket_root.size();
	croot.resize(nb);
	const unsigned T = flattenThreads();
	std::vector<uint64_t> need(T + 1, 0);
	{
		std::vector<std::thread> pool;
		for (unsigned p = 0; p < T; p++)
			pool.emplace_back([&, p]() {
				Compressor c = {x.nodes, NULL, 0};
				for (size_t i = nb * p / T; i < nb * (p + 1) / T; i++)
					c.run(x.bucket_root[i]);
				need[p + 1] = c.next;
			});
		for (auto &th : pool) th.join();
	}
	for (unsigned p = 0; p < T; p++)
		need[p + 1] += need[p];
	if (need[T] >= 0x7FFFFFFFull) {
		err = "Index too large: more than 2^31 trie nodes in one table.";
		return CQ_ENOMEM;
	}
	cnodes.resize(4 * (size_t) need[T]);
	{
		std::vector<std::thread> pool;
		for (unsigned p = 0; p < T; p++)
			pool.emplace_back([&, p]() {
				Compressor c = {x.nodes, cnodes.data(), need[p]};
				for (size_t i = nb * p / T; i < nb * (p + 1) / T; i++)
					croot[i] = c.run(x.bucket_root[i]);
			});
		for (auto &th : pool) th.join();
	}
	return CQ_OK;
}

} // namespace

int flattenIndices(DecodedIndex &u, DecodedIndex &d, double load_factor, FlatIndex &out, std::string &err) {
	auto t0 = std::chrono::high_resolution_clock::now();
	if (u.doubly_unique || !d.doubly_unique) {
		err = "Index flags: expected a unique (.bin1) and a doubly-unique (.bin2) index.";
		return CQ_EFORMAT;
	}
	if (u.hash_len != d.hash_len) {
		err = "Hash lengths of the unique and doubly-unique index differ.";
		return CQ_EFORMAT;
	}
	if (load_factor <= 0.0 || load_factor > 1.0)
		load_factor = 0.30;
	out.hash_len = u.hash_len;
	FlatVec<uint32_t>::type croot[2];
	int rc_c;
	if ((rc_c = compressTries(u, out.cnodes_u, croot[0], err)) != 0) return rc_c;
	if ((rc_c = compressTries(d, out.cnodes_d, croot[1], err)) != 0) return rc_c;
	if (getenv("CAMMIQ_VERBOSE")) fprintf(stderr, "[flatten] tries compressed %.0f ms: %zu -> %zu and %zu -> %zu nodes\n", std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count(), u.nodes.size() / 4, out.cnodes_u.size() / 4, d.nodes.size() / 4, out.cnodes_d.size() / 4);
	uint64_t upper = u.bucket_key.size() + d.bucket_key.size();
	uint64_t want = (uint64_t) ((double) upper / (load_factor * kSlotsPerBucket)) + 1;
	uint64_t nb = 64;
	uint32_t shift = 26;
	while (nb < want && shift > 0) {
		nb <<= 1;
		shift--;
	}
	if (nb < want) {
		err = "Index too large: the prefix table would need more than 2^32 buckets.";
		return CQ_ENOMEM;
	}
	out.n_table_buckets = nb;
	out.table_shift = shift;
	// the table is a few GB: allocate it raw and first-touch it from all threads
	if (!out.table.alloc(nb))
		throw std::bad_alloc();
	{
		const TableBucket empty = {{0, 0}, {{kRefNone, kRefNone}, {kRefNone, kRefNone}}};
		const unsigned Ti = flattenThreads();
		std::vector<std::thread> pool;
		for (unsigned p = 0; p < Ti; p++)
			pool.emplace_back([&, p]() {
				const size_t lo = out.table.size() * p / Ti, hi = out.table.size() * (p + 1) / Ti;
				for (size_t i = lo; i < hi; i++)
					out.table[i] = empty;
			});
		for (auto &th : pool) th.join();
	}
	if (getenv("CAMMIQ_VERBOSE")) fprintf(stderr, "[flatten] table init %.0f ms\n", std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count());
	const uint64_t mask = nb - 1;
	const uint64_t key_limit = (u.hash_len >= 32) ? UINT64_MAX : (1ull << (2 * u.hash_len));
	// Parallel build: thread p owns the home buckets [nb*p/T, nb*(p+1)/T) and inserts, in file
	// order, exactly the keys whose home bucket lies there, probing inside its own range; the
	// few keys whose probe sequence would cross the range end are inserted afterwards by one
	// thread.  Every key stored past its home bucket still has only full buckets before it, so
	// lookups (stop at the first bucket with a free slot) are unaffected.
	const unsigned T = flattenThreads();
	// pre-pass: which thread owns each key (its home bucket's range), computed once in parallel
	// instead of by every insert thread
	std::vector<uint8_t> owner[2];
	for (int t = 0; t < 2; t++) {
		const DecodedIndex &x = t == 0 ? u : d;
		owner[t].resize(x.bucket_key.size());
		std::vector<std::thread> pool;
		for (unsigned q = 0; q < T; q++)
			pool.emplace_back([&, q, t]() {
				const size_t a = x.bucket_key.size() * q / T, e = x.bucket_key.size() * (q + 1) / T;
				for (size_t i = a; i < e; i++) {
					const uint64_t b = homeBucketHost(x.bucket_key[i], u.hash_len, shift);
					unsigned p = (unsigned) (((unsigned __int128) b * T) / nb);
					while (p + 1 < T && b >= nb * (p + 1) / T) p++;
					while (p > 0 && b < nb * p / T) p--;
					owner[t][i] = (uint8_t) p;
				}
			});
		for (auto &th : pool) th.join();
	}
	struct Deferred { uint8_t table; uint64_t index; };
	std::vector<std::vector<Deferred>> deferred(T);
	std::vector<uint64_t> fresh_keys(T, 0);
	std::atomic<bool> bad_key(false);
	{
		std::vector<std::thread> pool;
		for (unsigned p = 0; p < T; p++)
			pool.emplace_back([&, p]() {
				const uint64_t hi = nb * (p + 1) / T;
				uint64_t fresh = 0;
				// table accesses are cache misses: the home bucket of a key is prefetched kAhead
				// owned keys before it is inserted
				static const unsigned kAhead = 24;
				struct Pending { uint64_t i, b; };
				Pending ring[kAhead];
				for (int t = 0; t < 2; t++) {
					const DecodedIndex &x = t == 0 ? u : d;
					const std::vector<uint8_t> &own = owner[t];
					auto insert = [&](const Pending &e) {
						const uint64_t key = x.bucket_key[e.i];
						if (key >= key_limit) {
							bad_key = true;
							return;
						}
						// probe inside the owned range; a bucket that is left behind full gets its
						// overflow flag (also when the key ends up deferred: it will lie further on)
						SlotRef hit = {NULL, 0};
						for (uint64_t b = e.b; b < hi && hit.b == NULL; b++) {
							bool is_new = false;
							const int k = bucketInsert(out.table[b], key, is_new);
							if (k >= 0) {
								hit.b = &out.table[b];
								hit.k = k;
								fresh += is_new ? 1 : 0;
							} else
								out.table[b].key[0] |= kBucketOverflow;
						}
						if (hit.b == NULL) {
							Deferred df = {(uint8_t) t, e.i};
							deferred[p].push_back(df);
							return;
						}
						// a repeated key inside one file: the later bucket replaces the earlier one, as
						// map64[bucket] = root does (hashtrie.cpp:500)
						hit.b->ref[hit.k][t] = croot[t][e.i];
					};
					uint64_t queued = 0;
					for (size_t i = 0; i < x.bucket_key.size(); i++) {
						if (own[i] != (uint8_t) p)
							continue;
						Pending e = {(uint64_t) i, homeBucketHost(x.bucket_key[i], u.hash_len, shift)};
						__builtin_prefetch(&out.table[e.b * kSlotsPerBucket], 1);
						if (queued >= kAhead)
							insert(ring[queued % kAhead]); // the oldest entry: file order is kept
						ring[queued % kAhead] = e;
						queued++;
					}
					for (uint64_t q = queued > kAhead ? queued - kAhead : 0; q < queued; q++)
						insert(ring[q % kAhead]);
				}
				fresh_keys[p] = fresh;
			});
		for (auto &th : pool) th.join();
	}
	if (getenv("CAMMIQ_VERBOSE")) fprintf(stderr, "[flatten] parallel insert done %.0f ms\n", std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count());
	if (bad_key) {
		err = "Bucket key does not fit 2*hash_len bits.";
		return CQ_EFORMAT;
	}
	uint64_t n_keys = 0;
	for (unsigned p = 0; p < T; p++) {
		n_keys += fresh_keys[p];
		for (const Deferred &df : deferred[p]) {
			const DecodedIndex &x = df.table == 0 ? u : d;
			bool fresh;
			// the flags of the owner's range are already raised; from the range end on, the key
			// wraps into buckets of another (finished) range
			SlotRef s = probeInsert(out.table, mask, shift, u.hash_len, x.bucket_key[df.index], fresh);
			n_keys += fresh ? 1 : 0;
			s.b->ref[s.k][df.table] = croot[df.table][df.index];
		}
	}
	out.n_keys = n_keys;
	out.u = std::move(u);
	out.d = std::move(d);
	if (getenv("CAMMIQ_VERBOSE")) fprintf(stderr, "[flatten] before filter %.0f ms\n", std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count());
	buildFilter(out, kFilterMaxBytesDefault);
	out.flatten_ms = std::chrono::duration<double, std::milli>(
		std::chrono::high_resolution_clock::now() - t0).count();
	return CQ_OK;
}

void buildFilter(FlatIndex &fi, uint64_t max_bytes) {
	fi.filter.clear();
	fi.filter.shrink_to_fit();
	fi.filter_words = 0;
	if (max_bytes < 8192)
		return;
	// 16 bits per key when the budget allows, never fewer than kFilterMinBitsPerKey
	uint64_t words = std::max<uint64_t>(1024, (fi.n_keys * 16 + 63) / 64);
	words = std::min<uint64_t>(words, max_bytes / 8);
	words = std::min<uint64_t>(words, (1ull << 31));
	words &= ~127ull; // whole 1 KB blocks
	fi.filter_sel_mask = kFilterSelAll;
	fi.filter_sieve = false;
	if (words * 64 < fi.n_keys * kFilterMinBitsPerKey) {
		// not selective at this size: use it as a sieve in front of the table probe, with the number
		// of bits per key that minimises the share of positions passing (about ln 2 * bits per key)
		if (words * 64 < fi.n_keys)
			return; // below one bit per key nothing is gained
		const double bpk = (double) (words * 64) / (double) fi.n_keys;
		fi.filter_sel_mask = bpk >= 5.0 ? kFilterSelAll : bpk >= 2.5 ? 0x00770077u : 0x00070007u;
		fi.filter_sieve = true;
	}
	// test hook: CAMMIQ_FILTER_FORCE_SIEVE = 1, 2 or 4 selector pairs puts any index in the sieve regime
	if (const char *force = getenv("CAMMIQ_FILTER_FORCE_SIEVE")) {
		const int k = atoi(force);
		fi.filter_sel_mask = k >= 4 ? kFilterSelAll : k >= 2 ? 0x00770077u : 0x00070007u;
		fi.filter_sieve = true;
	}
	// with selector pairs switched off the test looks at bit 0 of the word in their place
	fi.filter.assign(words, fi.filter_sel_mask == kFilterSelAll ? 0ull : 1ull);
	fi.filter_words = (uint32_t) words;
	const unsigned T = flattenThreads();
	std::vector<std::thread> pool;
	for (unsigned p = 0; p < T; p++)
		pool.emplace_back([&, p]() {
			const size_t lo = fi.table.size() * p / T, hi = fi.table.size() * (p + 1) / T;
			for (size_t i = lo; i < hi; i++)
				for (int k = 0; k < kSlotsPerBucket; k++) {
					const uint64_t stored = fi.table[i].key[k] & ~kBucketOverflow;
					if (stored == 0)
						continue;
					uint32_t A, B;
					const uint64_t key = stored & ~kKeyOccupied, canon = canonicalKeyHost(key, fi.hash_len);
					filterHash(canon, A, B);
					__atomic_fetch_or(&fi.filter[filterWordIndex(A, fi.filter_words)], filterMask(B, fi.filter_sel_mask), __ATOMIC_RELAXED);
				}
		});
	for (auto &th : pool) th.join();
}

uint64_t flatFind(const FlatIndex &fi, int table, uint64_t bucket, const uint8_t *cand, size_t len) {
	uint64_t mask = fi.n_table_buckets - 1;
	if (!fi.filter.empty()) {
		// same gate as phase 1 of the scan kernel: a filter miss ends the lookup
		uint32_t A, B;
		const uint64_t canon = canonicalKeyHost(bucket, fi.hash_len);
		filterHash(canon, A, B);
		uint64_t w = fi.filter[filterWordIndex(A, fi.filter_words)];
		if (!filterTest((uint32_t) w, (uint32_t) (w >> 32), B, fi.filter_sel_mask))
			return UINT64_MAX;
	}
	uint64_t b = homeBucketHost(bucket, fi.hash_len, fi.table_shift);
	uint32_t ref = kRefNone;
	const uint64_t tag = bucket | kKeyOccupied;
	for (;;) {
		const TableBucket &tb = fi.table[b];
		bool hit = false;
		for (int k = 0; k < kSlotsPerBucket; k++)
			if ((tb.key[k] & ~kBucketOverflow) == tag) {
				ref = tb.ref[k][table == CQ_TABLE_U ? 0 : 1];
				hit = true;
			}
		if (hit || !(tb.key[0] & kBucketOverflow))
			break;
		b = (b + 1) & mask;
	}
	const FlatVec<uint32_t>::type &cn = table == CQ_TABLE_U ? fi.cnodes_u : fi.cnodes_d;
	size_t i = 0;
	while (ref != kRefNone) {
		if (refIsLeaf(ref))
			return refLeafId(ref);
		const uint32_t *nd = &cn[4 * (size_t) refNodeId(ref)];
		if ((nd[0] & kChainTag) == kChainTag) {
			// a run of single-child nodes: every base of it must be there and match
			const uint32_t n = nd[0] & 63u;
			const uint64_t want = ((uint64_t) nd[1] << 32) | nd[2];
			if (i + n > len)
				return UINT64_MAX;
			uint64_t have = 0;
			for (uint32_t t = 0; t < n; t++) {
				int code = baseCode(cand[i + t]);
				if (code < 0)
					return UINT64_MAX;
				have = (have << 2) | (uint64_t) code;
			}
			if (have != want)
				return UINT64_MAX;
			i += n;
			ref = nd[3];
		} else {
			if (i >= len)
				return UINT64_MAX;
			int code = baseCode(cand[i++]);
			if (code < 0)
				return UINT64_MAX;
			ref = nd[code];
		}
	}
	return UINT64_MAX;
}

} // namespace cammiq
