// Synthetic workload generator (TOOLING for bench.py and tests; not on the query path).
//
// Procedural, seeded, block-structured strain genomes -- no genome is ever materialised:
//   genome g (cluster c = g / cluster_size) is a sequence of blocks of `block_len` bases;
//   each (cluster, block) is private (content unique to g), pair-shared (identical in g and
//   its partner strain g^1) or cluster-shared (identical in every strain of the cluster).
// k-mers inside private blocks are unique to g, k-mers inside pair-shared blocks are doubly
// unique, cluster-shared blocks contribute nothing -- the structure the reference builder
// extracts from real strain collections, at the leaf densities
// measured there (about 17 unique and 10 doubly-unique substrings per kbp, 95% of length
// k = h, the rest spread up to Lmax).
//
//   cqs_write_index   emits format-exact index_u.bin1 / index_d.bin2 (+.aux) through the
//                     product's own encoder plus genome_map.out and the three meta files
//                     that `cammiq --query` opens (query.cpp:158-205)
//   cqs_make_reads    simulated reads (CAMMiQ-simulate's model: uniform start, reverse
//                     complement w.p. 0.5, i.i.d. substitutions) into a caller buffer
//   cqs_write_fastq   the same reads as a FASTQ file (for the reference CLI / harness)
//
// Everything is a pure function of (seed, indices), independent of the thread count.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cammiq_gpu.h"
#include "index_codec.hpp"

using namespace cammiq;

extern "C" {

typedef struct {
	uint64_t seed;
	uint32_t n_genomes;
	uint32_t genome_len;   // bases per genome (multiple of block_len is used)
	uint32_t cluster_size; // strains per cluster
	uint32_t block_len;
	uint32_t permille_private, permille_pair; // remainder = cluster-shared
	uint32_t u_per_block, d_per_block;        // keys per private / pair-shared block
	uint32_t k;            // minimum key length = hash length h
	uint32_t lmax;         // maximum key length (lmax - k <= 32)
	uint32_t permille_deep;// keys longer than k
	uint32_t threads;
} cqs_params;

typedef struct {
	uint64_t n_leaves_u, n_leaves_d, n_dropped;
	double gen_ms, sort_ms, write_ms;
} cqs_index_stats;

int cqs_write_index(const cqs_params *p, const char *dir, cqs_index_stats *stats);
int cqs_make_reads(const cqs_params *p, uint64_t first, uint64_t n, uint32_t read_len, double erate,
		uint8_t *bases, uint32_t *src_genome);
int cqs_write_fastq(const cqs_params *p, uint64_t first, uint64_t n, uint32_t read_len, double erate,
		const char *path);
}

namespace {

inline uint64_t mix(uint64_t x) {
	x += 0x9E3779B97F4A7C15ull;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
	return x ^ (x >> 31);
}
inline uint64_t H(uint64_t seed, uint64_t a, uint64_t b, uint64_t c) {
	return mix(mix(mix(seed ^ (a * 0xD6E8FEB86659FD93ull)) ^ b) ^ (c * 0xA24BAED4963EE407ull));
}

struct Model {
	cqs_params p;
	uint32_t n_blocks;
	explicit Model(const cqs_params &q) : p(q) { n_blocks = p.genome_len / p.block_len; }
	// 0 private, 1 pair-shared, 2 cluster-shared
	int blockType(uint32_t g, uint32_t b, uint32_t &partner) const {
		uint32_t c = g / p.cluster_size;
		uint32_t t = (uint32_t) (H(p.seed, 4, c, b) % 1000);
		partner = g;
		if (t < p.permille_private)
			return 0;
		if (t < p.permille_private + p.permille_pair) {
			uint32_t in_cluster = g % p.cluster_size, mate = in_cluster ^ 1u;
			uint32_t pg = c * p.cluster_size + mate;
			if (mate >= p.cluster_size || pg >= p.n_genomes)
				return 0; // odd strain out: its copy of the block is private
			partner = pg;
			return 1;
		}
		return 2;
	}
	uint64_t blockSeed(uint32_t g, uint32_t b) const {
		uint32_t partner;
		int t = blockType(g, b, partner);
		if (t == 0) {
			// the odd strain of a pair block must not collide with a true private block
			return H(p.seed, 1, g, b);
		}
		if (t == 1)
			return H(p.seed, 2, std::min(g, partner), b);
		return H(p.seed, 3, g / p.cluster_size, b);
	}
	// 2-bit code of base `off` of a block
	static inline uint32_t blockBase(uint64_t bseed, uint32_t off) {
		uint64_t w = mix(bseed + (uint64_t) (off >> 5) * 0x9E3779B97F4A7C15ull);
		return (uint32_t) (w >> (2 * (off & 31))) & 3u;
	}
	uint32_t genomeBase(uint32_t g, uint64_t pos, uint32_t &cache_b, uint64_t &cache_seed) const {
		uint32_t b = (uint32_t) (pos / p.block_len);
		if (b != cache_b) {
			cache_b = b;
			cache_seed = blockSeed(g, b);
		}
		return blockBase(cache_seed, (uint32_t) (pos % p.block_len));
	}
};

struct KeyRec {
	uint64_t bucket; // 2-bit hash of the first k bases
	uint64_t extra;  // remaining bases, first base most significant, left-aligned in 64 bits
	uint32_t rid1, rid2;
	uint16_t ucount1, ucount2;
	uint8_t extra_len;
	uint8_t table; // 0 = U, 1 = D
};

inline bool recLess(const KeyRec &a, const KeyRec &b) {
	if (a.bucket != b.bucket) return a.bucket < b.bucket;
	if (a.extra_len != b.extra_len) return a.extra_len < b.extra_len;
	if (a.extra != b.extra) return a.extra < b.extra;
	if (a.rid1 != b.rid1) return a.rid1 < b.rid1;
	return a.rid2 < b.rid2;
}

void generateKeys(const Model &m, uint32_t g_lo, uint32_t g_hi, std::vector<KeyRec> &out_u, std::vector<KeyRec> &out_d) {
	const cqs_params &p = m.p;
	for (uint32_t g = g_lo; g < g_hi; g++) {
		for (uint32_t b = 0; b < m.n_blocks; b++) {
			uint32_t partner;
			int t = m.blockType(g, b, partner);
			if (t == 2)
				continue;
			if (t == 1 && partner < g)
				continue; // the pair block is emitted once, by its lower strain
			uint64_t bseed = m.blockSeed(g, b);
			uint32_t nkeys = t == 0 ? p.u_per_block : p.d_per_block;
			for (uint32_t j = 0; j < nkeys; j++) {
				uint64_t r = H(bseed, 100, j, t);
				uint32_t len = p.k;
				if (p.lmax > p.k && (r >> 40) % 1000 < p.permille_deep)
					len = p.k + 1 + (uint32_t) ((r >> 20) % (p.lmax - p.k));
				uint32_t pos = (uint32_t) (r % (p.block_len - p.lmax + 1));
				bool rc = (r >> 60) & 1;
				KeyRec k;
				memset(&k, 0, sizeof(k));
				uint64_t bucket = 0, extra = 0;
				for (uint32_t i = 0; i < len; i++) {
					uint32_t code = rc ? 3u - Model::blockBase(bseed, pos + len - 1 - i) : Model::blockBase(bseed, pos + i);
					if (i < p.k)
						bucket = (bucket << 2) | code;
					else
						extra |= (uint64_t) code << (62 - 2 * (i - p.k));
				}
				k.bucket = bucket;
				k.extra = extra;
				k.extra_len = (uint8_t) (len - p.k);
				k.ucount1 = (uint16_t) (1 + (r >> 8) % 3);
				k.table = (uint8_t) t;
				if (t == 0) {
					k.rid1 = g + 1;
					out_u.push_back(k);
				} else {
					k.rid1 = g + 1;
					k.rid2 = partner + 1;
					k.ucount2 = (uint16_t) (1 + (r >> 12) % 3);
					out_d.push_back(k);
				}
			}
		}
	}
}

// Sort records by bucket in parallel (partition on the top bits of the bucket key).
void parallelSort(std::vector<KeyRec> &v, uint32_t h, uint32_t threads) {
	const int PB = 8;
	const uint32_t shift = 2 * h > (uint32_t) PB ? 2 * h - PB : 0;
	std::vector<size_t> cnt((1 << PB) + 1, 0);
	for (auto &r : v) cnt[(r.bucket >> shift) + 1]++;
	for (int i = 0; i < (1 << PB); i++) cnt[i + 1] += cnt[i];
	std::vector<KeyRec> tmp(v.size());
	std::vector<size_t> at(cnt.begin(), cnt.end() - 1);
	for (auto &r : v) tmp[at[r.bucket >> shift]++] = r;
	v.swap(tmp);
	tmp.clear();
	tmp.shrink_to_fit();
	std::vector<std::thread> pool;
	for (uint32_t t = 0; t < threads; t++)
		pool.emplace_back([&, t]() {
			for (int part = (int) t; part < (1 << PB); part += (int) threads)
				std::sort(v.begin() + cnt[part], v.begin() + cnt[part + 1], recLess);
		});
	for (auto &th : pool) th.join();
}

// a is a prefix of b (or equal) given both share the bucket
inline bool extraPrefix(const KeyRec &a, const KeyRec &b) {
	if (a.extra_len > b.extra_len) return false;
	if (a.extra_len == 0) return true;
	uint64_t mask = ~0ull << (64 - 2 * a.extra_len);
	return (a.extra & mask) == (b.extra & mask);
}

// Records sorted by bucket -> DecodedIndex (buckets in sorted order, prefix-free).
uint64_t assemble(const std::vector<KeyRec> &v, uint32_t h, bool doubly, DecodedIndex &out) {
	out = DecodedIndex();
	out.doubly_unique = doubly;
	out.hash_len = h;
	out.bucket_key.reserve(v.size());
	out.bucket_root.reserve(v.size());
	out.ref_id1.reserve(v.size());
	out.ref_id2.reserve(v.size());
	out.ucount1.reserve(v.size());
	out.ucount2.reserve(v.size());
	out.depth.reserve(v.size());
	uint64_t dropped = 0;
	std::vector<const KeyRec *> keep;
	auto addLeaf = [&](const KeyRec &r) -> uint32_t {
		uint32_t id = (uint32_t) out.ref_id1.size();
		out.ref_id1.push_back(r.rid1);
		out.ref_id2.push_back(doubly ? r.rid2 : 0);
		out.ucount1.push_back(r.ucount1);
		out.ucount2.push_back(doubly ? r.ucount2 : 0);
		out.depth.push_back((uint8_t) (h + r.extra_len));
		out.max_ref_id = std::max(out.max_ref_id, std::max(r.rid1, r.rid2));
		return kRefLeafTag | id;
	};
	for (size_t i = 0; i < v.size();) {
		size_t j = i;
		while (j < v.size() && v[j].bucket == v[i].bucket) j++;
		// keep a prefix-free subset (sorted by length first, so shorter keys win)
		keep.clear();
		for (size_t a = i; a < j; a++) {
			bool ok = true;
			for (const KeyRec *k : keep)
				if (extraPrefix(*k, v[a])) { ok = false; break; }
			if (ok) keep.push_back(&v[a]);
			else dropped++;
		}
		uint32_t root;
		if (keep.size() == 1 && keep[0]->extra_len == 0)
			root = addLeaf(*keep[0]);
		else {
			// leaves must be numbered in pre-order (A<C<G<T): order the kept keys lexicographically
			std::sort(keep.begin(), keep.end(), [](const KeyRec *a, const KeyRec *b) { return a->extra < b->extra; });
			out.nodes.insert(out.nodes.end(), 4, kRefNone);
			uint32_t root_node = (uint32_t) (out.nodes.size() / 4 - 1);
			root = root_node + 1;
			for (const KeyRec *k : keep) {
				uint32_t cur = root_node;
				for (uint32_t d = 0; d < k->extra_len; d++) {
					uint32_t code = (uint32_t) (k->extra >> (62 - 2 * d)) & 3u;
					size_t slot = 4 * (size_t) cur + code;
					if (d + 1 == k->extra_len)
						out.nodes[slot] = addLeaf(*k);
					else {
						if (out.nodes[slot] == kRefNone) {
							out.nodes.insert(out.nodes.end(), 4, kRefNone);
							out.nodes[slot] = (uint32_t) (out.nodes.size() / 4 - 1) + 1;
						}
						cur = refNodeId(out.nodes[slot]);
					}
				}
			}
		}
		out.bucket_key.push_back(v[i].bucket);
		out.bucket_root.push_back(root);
		i = j;
	}
	return dropped;
}

double nowMs() {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

bool validParams(const cqs_params *p) {
	return p != NULL && p->n_genomes >= 1 && p->cluster_size >= 1 && p->block_len >= 64 &&
		p->genome_len >= p->block_len && p->k >= 5 && p->k <= 31 && p->lmax >= p->k &&
		p->lmax - p->k <= 32 && p->lmax < p->block_len && p->permille_private + p->permille_pair <= 1000;
}

void makeRead(const Model &m, uint64_t idx, uint32_t rl, double erate, uint8_t *dst, uint32_t *src) {
	static const char ALPHA[4] = {'A', 'C', 'G', 'T'};
	const cqs_params &p = m.p;
	uint64_t r = H(p.seed, 7, idx, rl);
	uint32_t g = (uint32_t) (r % p.n_genomes);
	uint64_t glen = (uint64_t) m.n_blocks * p.block_len;
	uint64_t pos = mix(r) % (glen - rl + 1);
	bool rc = (mix(r ^ 0x5555) >> 33) & 1;
	uint32_t cb = 0xFFFFFFFFu;
	uint64_t cs = 0;
	uint64_t e = mix(r ^ 0xABCDEF);
	const uint64_t thresh = erate <= 0 ? 0 : (uint64_t) (erate * 4294967296.0);
	for (uint32_t i = 0; i < rl; i++) {
		uint32_t code = rc ? 3u - m.genomeBase(g, pos + rl - 1 - i, cb, cs) : m.genomeBase(g, pos + i, cb, cs);
		if (thresh) {
			e = mix(e);
			if ((e & 0xFFFFFFFFull) < thresh)
				code = (code + 1 + (uint32_t) ((e >> 40) % 3)) & 3u; // substitution, CAMMiQ-simulate:119-141
		}
		dst[i] = (uint8_t) ALPHA[code];
	}
	if (src) *src = g + 1;
}

} // namespace

extern "C" int cqs_write_index(const cqs_params *p, const char *dir, cqs_index_stats *stats) {
	if (!validParams(p) || dir == NULL)
		return CQ_EINVAL;
	Model m(*p);
	const uint32_t T = std::max<uint32_t>(1, p->threads);
	double t0 = nowMs();
	std::vector<std::vector<KeyRec>> pu(T), pd(T);
	{
		std::vector<std::thread> pool;
		for (uint32_t t = 0; t < T; t++)
			pool.emplace_back([&, t]() {
				uint32_t lo = (uint32_t) ((uint64_t) p->n_genomes * t / T), hi = (uint32_t) ((uint64_t) p->n_genomes * (t + 1) / T);
				generateKeys(m, lo, hi, pu[t], pd[t]);
			});
		for (auto &th : pool) th.join();
	}
	std::vector<KeyRec> vu, vd;
	for (uint32_t t = 0; t < T; t++) {
		vu.insert(vu.end(), pu[t].begin(), pu[t].end());
		vd.insert(vd.end(), pd[t].begin(), pd[t].end());
		std::vector<KeyRec>().swap(pu[t]);
		std::vector<KeyRec>().swap(pd[t]);
	}
	double t1 = nowMs();
	parallelSort(vu, p->k, T);
	parallelSort(vd, p->k, T);
	double t2 = nowMs();
	std::string d(dir);
	if (!d.empty() && d.back() != '/') d += '/';
	uint64_t dropped = 0, nu = 0, nd = 0;
	std::string err_u, err_d;
	int rc_u = 0, rc_d = 0;
	{
		// assemble + encode the two tables on two threads
		std::thread tu([&]() {
			DecodedIndex x;
			uint64_t dr = assemble(vu, p->k, false, x);
			std::vector<KeyRec>().swap(vu);
			nu = x.numLeaves();
			rc_u = encodeIndexFile(d + "index_u.bin1", x, err_u);
			__atomic_fetch_add(&dropped, dr, __ATOMIC_RELAXED);
		});
		std::thread td([&]() {
			DecodedIndex x;
			uint64_t dr = assemble(vd, p->k, true, x);
			std::vector<KeyRec>().swap(vd);
			nd = x.numLeaves();
			rc_d = encodeIndexFile(d + "index_d.bin2", x, err_d);
			__atomic_fetch_add(&dropped, dr, __ATOMIC_RELAXED);
		});
		tu.join();
		td.join();
	}
	if (rc_u != 0 || rc_d != 0)
		return rc_u != 0 ? rc_u : rc_d;
	// map + meta files (formats: SURVEY.md section 5.9)
	FILE *fm = fopen((d + "genome_map.out").c_str(), "w");
	FILE *fl = fopen((d + "genome_lengths.out").c_str(), "w");
	FILE *fu = fopen((d + "unique_lmer_count_u.out").c_str(), "w");
	FILE *fd = fopen((d + "unique_lmer_count_d.out").c_str(), "w");
	if (!fm || !fl || !fu || !fd)
		return CQ_EIO;
	for (uint32_t g = 1; g <= p->n_genomes; g++) {
		fprintf(fm, "genome_%06u.fna\t%u\t%u\tSynthetic strain %u of cluster %u\n", g, g, 100000 + g, g,
			(g - 1) / p->cluster_size + 1);
		fprintf(fl, "%u\t%u\n", g, m.n_blocks * p->block_len);
		fprintf(fu, "%u\t%u\n", g, 0u);
		fprintf(fd, "%u\t%u\n", g, 0u);
	}
	fclose(fm); fclose(fl); fclose(fu); fclose(fd);
	double t3 = nowMs();
	if (stats) {
		stats->n_leaves_u = nu;
		stats->n_leaves_d = nd;
		stats->n_dropped = dropped;
		stats->gen_ms = t1 - t0;
		stats->sort_ms = t2 - t1;
		stats->write_ms = t3 - t2;
	}
	return CQ_OK;
}

extern "C" int cqs_make_reads(const cqs_params *p, uint64_t first, uint64_t n, uint32_t read_len, double erate,
		uint8_t *bases, uint32_t *src_genome) {
	if (!validParams(p) || bases == NULL || read_len == 0 || read_len > 255 ||
		(uint64_t) read_len > (uint64_t) (p->genome_len / p->block_len) * p->block_len)
		return CQ_EINVAL;
	Model m(*p);
	const uint32_t T = std::max<uint32_t>(1, p->threads);
	std::vector<std::thread> pool;
	for (uint32_t t = 0; t < T; t++)
		pool.emplace_back([&, t]() {
			uint64_t lo = n * t / T, hi = n * (t + 1) / T;
			for (uint64_t i = lo; i < hi; i++)
				makeRead(m, first + i, read_len, erate, bases + i * read_len, src_genome ? src_genome + i : NULL);
		});
	for (auto &th : pool) th.join();
	return CQ_OK;
}

extern "C" int cqs_write_fastq(const cqs_params *p, uint64_t first, uint64_t n, uint32_t read_len, double erate,
		const char *path) {
	if (path == NULL)
		return CQ_EINVAL;
	std::vector<uint8_t> buf((size_t) n * read_len);
	int rc = cqs_make_reads(p, first, n, read_len, erate, buf.data(), NULL);
	if (rc != 0)
		return rc;
	FILE *f = fopen(path, "w");
	if (f == NULL)
		return CQ_EIO;
	std::string qual(read_len, 'I');
	for (uint64_t i = 0; i < n; i++) {
		fprintf(f, "@r%llu\n", (unsigned long long) (first + i));
		fwrite(buf.data() + i * read_len, 1, read_len, f);
		fprintf(f, "\n+\n%s\n", qual.c_str());
	}
	fclose(f);
	return CQ_OK;
}
