/*
 * TEST INFRASTRUCTURE ONLY (oracle/).  Driver around the UNMODIFIED reference query code
 * (/root/reference/src/{query,hashtrie,binaryio}.cpp), compiled by oracle/Makefile into
 * oracle/_ref/ref_harness.  It only calls reference functions and reads reference state:
 *
 *   FqReader::loadIdx_p / loadSmap / prepallFastq / readallFastq      (query.cpp:109-156, 427-445)
 *   FqReader::query64_p / query64mt_p / query64_sc                    (query.cpp:458-1080)
 *   Hash::find64_p                                                    (hashtrie.cpp:350-369)
 *   Genome::read_cnts_u/d, FqReader::nundet/nconf/read_cnts_b,
 *   Hash::map_sp, pleafNode::{rcount,depth,ucount1,ucount2}           (query.hpp, hashtrie.hpp)
 *
 * and writes them to a plain-text dump that tests/ and tests/golden/ consume.  Private
 * members are reached with the `#define private public` trick; no reference source is
 * edited or copied.
 *
 * Usage:
 *   ref_harness dump <idx_u> <idx_d> <map> <p|mt|sc> <nthreads> <per_read_n> <out> <fastq>...
 *   ref_harness time <idx_u> <idx_d> <map> <p|mt|sc> <nthreads> <fastq> [reps]
 *       prints one JSON line per repetition (the reference's own scan, timed around the
 *       query64_* call like its "Time for query" bracket) and leaves with _exit so that the
 *       pointer-trie teardown (tens of seconds at 3e7 leaves) is not waited for.  The line
 *       carries the full per-genome vectors (cu, cd), the pair map (sc) and, for the per-leaf
 *       counts, sum and a position-weighted digest over Hash::map_sp in its own order:
 *       sum over g, k of rcount(map_sp[g][k]) * mix64(g << 32 | k)  (mod 2^64)
 *   ref_harness ilp <idx_u> <idx_d> <map> <erate> <out> <fastq>
 *       query64_p, then what runILP_* derives per leaf before it builds the model, evaluated
 *       with the expressions of query.cpp:1087, 1157-1160, 1171-1175 on the reference's own
 *       nodes (the solver code itself is compiled out without CPLEX / Gurobi): one line
 *       "<U|D> <genome> <k> <rcount> <wcov1 %.17g> <wcov2 %.17g>" per map_sp entry
 *
 *   ref_harness readdump <fastq> <min_len>
 *       what FqReader::readFastq (query.cpp:371-425) holds after reading the file: one line
 *       "<length> <bases>" per read on stdout ((uint8_t) length, the first `length` bytes as stored).
 *       The reference seeds rand() from the clock; this binary interposes srand() so that
 *       CAMMIQ_SEED (when set) is the seed instead, which makes the N substitution reproducible.
 *
 * Canonical leaf ids in the dump = rank of the leaf's full key (h-base bucket prefix +
 * trie path) in lexicographic order within its table; both sides can compute it
 * independently of the bucket order in the file.
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>
#include <fcntl.h>
#include <unistd.h>
#include <omp.h>

#define private public
#include "query.hpp"
#undef private

/* The reference calls srand(clock) at the top of readFastq; defined here, this srand wins over
   libc's for every caller in the executable.  glibc's rand() shares its state with random(). */
extern "C" void srand(unsigned int seed) {
	const char *s = getenv("CAMMIQ_SEED");
	srandom(s != NULL ? (unsigned int) atoi(s) : seed);
}

typedef std::unordered_map<const trieNode*, uint64_t> LeafIds;

static void numberLeaves(const trieNode *node, LeafIds &ids, uint64_t &next) {
	if (node == NULL)
		return;
	if (node->isEnd) {
		ids[node] = next++;
		return;
	}
	for (int c = 0; c < 4; c++)
		numberLeaves(node->children[c], ids, next);
}

static void canonicalIds(Hash *ht, LeafIds &ids) {
	std::vector<std::pair<uint64_t, trieNode*>> buckets;
	for (auto &it : ht->getmap())
		buckets.push_back(std::make_pair(it.first, it.second));
	std::sort(buckets.begin(), buckets.end());
	uint64_t next = 0;
	for (auto &b : buckets)
		numberLeaves(b.second, ids, next);
}

static void dumpLeafTable(FILE *out, const char *tag, Hash *ht, size_t G, LeafIds &ids) {
	for (size_t rid = 1; rid <= G; rid++) {
		auto it = ht->map_sp.find((uint32_t) rid);
		if (it == ht->map_sp.end()) {
			fprintf(out, "%s %zu 0\n", tag, rid);
			continue;
		}
		fprintf(out, "%s %zu %zu", tag, rid, it->second.size());
		for (auto pn : it->second)
			fprintf(out, " %lu:%u:%u:%u:%u:%u", (unsigned long) ids[pn], pn->refID1, pn->refID2,
				(unsigned) pn->depth, (unsigned) pn->ucount1, (unsigned) pn->ucount2);
		fprintf(out, "\n");
	}
}

static void dumpRcounts(FILE *out, const char *tag, Hash *ht, size_t G) {
	for (size_t rid = 1; rid <= G; rid++) {
		auto it = ht->map_sp.find((uint32_t) rid);
		if (it == ht->map_sp.end()) {
			fprintf(out, "%s %zu 0\n", tag, rid);
			continue;
		}
		fprintf(out, "%s %zu %zu", tag, rid, it->second.size());
		for (auto pn : it->second)
			fprintf(out, " %u", pn->rcount);
		fprintf(out, "\n");
	}
}

/* Per-read distinct leaf sets through the reference's public lookup primitive, walking
   positions exactly as query.cpp:482-527 does (forward strand, then reverse complement). */
static void readLeafSets(FqReader &fq, uint8_t *read, size_t rl, std::set<pleafNode*> &pn) {
	uint32_t h = fq.hash_len_u;
	uint8_t rc[256];
	fq.getRC(rc, read, rl);
	uint8_t *strands[2] = {read, rc};
	for (int s = 0; s < 2; s++) {
		uint8_t *r = strands[s];
		for (size_t i = 0; i + h <= rl; i++) {
			uint64_t hv = fq.ht_u->computeHashVal64(r + i);
			pleafNode *p = fq.ht_u->find64_p(hv, r + i + h, rl - h - i);
			if (p != NULL) pn.insert(p);
			p = fq.ht_d->find64_p(hv, r + i + h, rl - h - i);
			if (p != NULL) pn.insert(p);
		}
	}
}

static uint64_t mix64(uint64_t x) {
	x += 0x9E3779B97F4A7C15ull;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
	return x ^ (x >> 31);
}

static void rcountDigest(Hash *ht, size_t G, uint64_t &sum, uint64_t &digest) {
	sum = digest = 0;
	for (size_t rid = 1; rid <= G; rid++) {
		auto it = ht->map_sp.find((uint32_t) rid);
		if (it == ht->map_sp.end())
			continue;
		for (size_t k = 0; k < it->second.size(); k++) {
			sum += it->second[k]->rcount;
			digest += (uint64_t) it->second[k]->rcount * mix64(((uint64_t) rid << 32) | (uint64_t) k);
		}
	}
}

static void runQuery(FqReader &fq, const std::string &mode, size_t fi) {
	if (mode == "p")
		fq.query64_p(fi);
	else if (mode == "mt")
		fq.query64mt_p(fi);
	else
		fq.query64_sc(fi);
}

static int readDump(const std::string &fastq, size_t min_len) {
	std::string none;
	FqReader fq(none, none, none, none, 0.01f, false);
	fq.qfilenames.push_back(fastq);
	fq.prepallFastq();
	if (min_len == 0)
		fq.readallFastq();
	else
		fq.readallFastq(min_len);
	for (size_t r = 0; r < fq.reads[0].size(); r++) {
		printf("%u ", (unsigned) fq.rlengths[0][r]);
		fwrite(fq.reads[0][r], 1, fq.rlengths[0][r], stdout);
		printf("\n");
	}
	fflush(stdout);
	_exit(0);
}

static void ilpLeaves(FILE *out, const char *tag, Hash *ht, size_t G, uint32_t rl, double erate) {
	for (size_t i = 0; i < G; i++) {
		auto it = ht->map_sp.find((uint32_t) (i + 1));
		if (it == ht->map_sp.end())
			continue;
		size_t k = 0;
		for (auto pn : it->second) {
			/* the expressions of query.cpp:1157-1160 (unique) and 1171-1175 (doubly unique), verbatim */
			double wcov1 = (pn->ucount1 * (rl - pn->depth) * 1.0 / rl);
			double wcov2 = (pn->ucount2 * (rl - pn->depth) * 1.0 / rl);
			wcov1 = wcov1 * pow(1 - erate, pn->depth);
			wcov2 = wcov2 * pow(1 - erate, pn->depth);
			fprintf(out, "%s %zu %zu %u %.17g %.17g\n", tag, i + 1, k++, pn->rcount, wcov1, wcov2);
		}
	}
}

static int ilpDump(int argc, char **argv) {
	std::string idx_u = argv[2], idx_d = argv[3], map_fn = argv[4], out_fn = argv[6], fastq = argv[7], empty;
	float erate_f = (float) atof(argv[5]); /* FqReader keeps -e as a float (query.hpp:62) */
	(void) argc;
	FqReader fq(idx_u, idx_d, map_fn, empty, erate_f, false);
	fq.loadIdx_p();
	fq.loadSmap();
	fq.qfilenames.push_back(fastq);
	fq.prepallFastq();
	fq.readallFastq();
	fq.getFqnameWithoutDir(0);
	fq.query64_p(0);
	FILE *out = fopen(out_fn.c_str(), "w");
	if (out == NULL)
		return 1;
	size_t G = fq.genomes.size() - 1;
	uint32_t rl = fq.tlengths[0] / fq.reads[0].size(); /* query.cpp:1087 */
	double erate = fq.erate_;                          /* the double parameter of runILP_* (query.cpp:252) */
	fprintf(out, "RL %u\n", rl);
	ilpLeaves(out, "U", fq.ht_u, G, rl, erate);
	ilpLeaves(out, "D", fq.ht_d, G, rl, erate);
	fclose(out);
	return 0;
}

int main(int argc, char **argv) {
	if (argc == 4 && std::string(argv[1]) == "readdump")
		return readDump(argv[2], (size_t) atol(argv[3]));
	if (argc == 8 && std::string(argv[1]) == "ilp")
		return ilpDump(argc, argv);
	if (argc < 8) {
		fprintf(stderr, "usage: see header of oracle/ref_harness.cpp\n");
		return 2;
	}
	std::string cmd = argv[1], idx_u = argv[2], idx_d = argv[3], map_fn = argv[4], mode = argv[5];
	int nthreads = atoi(argv[6]);
	size_t per_read_n = 0;
	std::string out_fn, empty;
	std::vector<std::string> fastqs;
	int a = 7;
	if (cmd == "dump") {
		per_read_n = (size_t) atol(argv[a++]);
		out_fn = argv[a++];
	}
	for (; a < argc; a++)
		fastqs.push_back(argv[a]);
	int reps = 1;
	if (cmd == "time" && fastqs.size() > 1) {
		reps = atoi(fastqs.back().c_str());
		fastqs.pop_back();
		if (reps < 1) reps = 1;
	}
	if (fastqs.empty() || (mode != "p" && mode != "mt" && mode != "sc")) {
		fprintf(stderr, "bad arguments\n");
		return 2;
	}
	omp_set_num_threads(nthreads);

	auto t0 = std::chrono::high_resolution_clock::now();
	FqReader fq(idx_u, idx_d, map_fn, empty, 0.01f, false);
	fq.loadIdx_p();
	fq.loadSmap();
	fq.nthreads = nthreads;
	fq.qfilenames = fastqs;
	fq.prepallFastq();
	fq.readallFastq();
	double load_ms = std::chrono::duration<double, std::milli>(
		std::chrono::high_resolution_clock::now() - t0).count();
	size_t G = fq.genomes.size() - 1;

	if (cmd == "time") {
		fq.getFqnameWithoutDir(0);
		for (int rep = 0; rep < reps; rep++) {
			auto t1 = std::chrono::high_resolution_clock::now();
			runQuery(fq, mode, 0);
			double q_ms = std::chrono::duration<double, std::milli>(
				std::chrono::high_resolution_clock::now() - t1).count();
			uint64_t su = 0, sd = 0;
			for (size_t i = 1; i <= G; i++) {
				su += fq.genomes[i]->read_cnts_u;
				sd += fq.genomes[i]->read_cnts_d;
			}
			printf("{\"rep\": %d, \"reads\": %zu, \"query_ms\": %.3f, \"load_ms\": %.3f, \"threads\": %d, "
				"\"mode\": \"%s\", \"nundet\": %zu, \"nconf\": %zu, \"sum_u\": %lu, \"sum_d\": %lu",
				rep, fq.reads[0].size(), q_ms, load_ms, nthreads, mode.c_str(), fq.nundet, fq.nconf,
				(unsigned long) su, (unsigned long) sd);
			printf(", \"cu\": [");
			for (size_t i = 1; i <= G; i++)
				printf(i < G ? "%lu, " : "%lu", (unsigned long) fq.genomes[i]->read_cnts_u);
			printf("], \"cd\": [");
			for (size_t i = 1; i <= G; i++)
				printf(i < G ? "%lu, " : "%lu", (unsigned long) fq.genomes[i]->read_cnts_d);
			printf("]");
			if (mode != "sc") {
				uint64_t s1, d1, s2, d2;
				rcountDigest(fq.ht_u, G, s1, d1);
				rcountDigest(fq.ht_d, G, s2, d2);
				printf(", \"rcu_sum\": %lu, \"rcu_digest\": \"%016lx\", \"rcd_sum\": %lu, \"rcd_digest\": \"%016lx\"",
					(unsigned long) s1, (unsigned long) d1, (unsigned long) s2, (unsigned long) d2);
			} else {
				printf(", \"pairs\": [");
				size_t k = 0;
				for (auto &it : fq.read_cnts_b)
					printf(k++ ? ", [%u, %u, %lu]" : "[%u, %u, %lu]", it.first.first, it.first.second, (unsigned long) it.second);
				printf("]");
			}
			printf("}\n");
			fflush(stdout);
			if (rep + 1 < reps) {
				if (mode == "sc") fq.resetCounters_sc();
				else fq.resetCounters();
			}
		}
		fflush(stderr);
		_exit(0);
	}

	FILE *out = fopen(out_fn.c_str(), "w");
	if (out == NULL) {
		fprintf(stderr, "cannot open %s\n", out_fn.c_str());
		return 1;
	}
	LeafIds ids_u, ids_d;
	canonicalIds(fq.ht_u, ids_u);
	canonicalIds(fq.ht_d, ids_d);
	fprintf(out, "H %u %u\nNU %lu\nND %lu\nG %zu\nMODE %s\n", fq.hash_len_u, fq.hash_len_d,
		(unsigned long) fq.ht_u->leaf_cnt, (unsigned long) fq.ht_d->leaf_cnt, G, mode.c_str());
	fprintf(out, "TAXID");
	for (size_t i = 1; i <= G; i++)
		fprintf(out, " %u", fq.genomes[i]->taxID);
	fprintf(out, "\n");
	dumpLeafTable(out, "LEAFU", fq.ht_u, G, ids_u);
	dumpLeafTable(out, "LEAFD", fq.ht_d, G, ids_d);

	for (size_t fi = 0; fi < fastqs.size(); fi++) {
		fq.getFqnameWithoutDir(fi);
		runQuery(fq, mode, fi);
		fprintf(out, "FILE %s %zu\nNUNDET %zu\nNCONF %zu\n", fq.current_filename.c_str(),
			fq.reads[fi].size(), fq.nundet, fq.nconf);
		fprintf(out, "CU");
		for (size_t i = 1; i <= G; i++)
			fprintf(out, " %lu", (unsigned long) fq.genomes[i]->read_cnts_u);
		fprintf(out, "\nCD");
		for (size_t i = 1; i <= G; i++)
			fprintf(out, " %lu", (unsigned long) fq.genomes[i]->read_cnts_d);
		fprintf(out, "\n");
		if (mode != "sc") {
			dumpRcounts(out, "RCU", fq.ht_u, G);
			dumpRcounts(out, "RCD", fq.ht_d, G);
		} else {
			fprintf(out, "PAIRS %zu", fq.read_cnts_b.size());
			for (auto &it : fq.read_cnts_b)
				fprintf(out, " %u:%u:%lu", it.first.first, it.first.second, (unsigned long) it.second);
			fprintf(out, "\n");
		}
		if (mode == "sc")
			fq.resetCounters_sc();
		else
			fq.resetCounters();
	}

	/* Optional per-read records for the first per_read_n reads of the first file: the
	   reference's own decision for that read alone (query64_* on a one-read file slot)
	   plus its distinct leaf set. */
	if (per_read_n > 0) {
		size_t n = std::min(per_read_n, fq.reads[0].size());
		size_t slot = fq.reads.size();
		fq.reads.push_back(std::vector<uint8_t*>(1, (uint8_t*) NULL));
		fq.rlengths.push_back(std::vector<uint8_t>(1, 0));
		fflush(stderr);
		int saved = dup(2), devnull = open("/dev/null", O_WRONLY);
		dup2(devnull, 2);
		for (size_t r = 0; r < n; r++) {
			fq.reads[slot][0] = fq.reads[0][r];
			fq.rlengths[slot][0] = fq.rlengths[0][r];
			fq.nundet = 0;
			fq.nconf = 0;
			for (size_t i = 1; i <= G; i++) {
				fq.genomes[i]->read_cnts_u = 0;
				fq.genomes[i]->read_cnts_d = 0;
			}
			fq.read_cnts_b.clear();
			runQuery(fq, mode, slot);
			fprintf(out, "READ %zu %zu %zu U", r, fq.nundet, fq.nconf);
			for (size_t i = 1; i <= G; i++)
				for (uint64_t k = 0; k < fq.genomes[i]->read_cnts_u; k++)
					fprintf(out, " %zu", i);
			fprintf(out, " D");
			for (size_t i = 1; i <= G; i++)
				for (uint64_t k = 0; k < fq.genomes[i]->read_cnts_d; k++)
					fprintf(out, " %zu", i);
			std::set<pleafNode*> pn;
			readLeafSets(fq, fq.reads[0][r], fq.rlengths[0][r], pn);
			std::vector<uint64_t> lu, ld;
			for (auto p : pn) {
				if (p->refID2 == 0) lu.push_back(ids_u[p]);
				else ld.push_back(ids_d[p]);
			}
			std::sort(lu.begin(), lu.end());
			std::sort(ld.begin(), ld.end());
			fprintf(out, " LU");
			for (auto v : lu) fprintf(out, " %lu", (unsigned long) v);
			fprintf(out, " LD");
			for (auto v : ld) fprintf(out, " %lu", (unsigned long) v);
			fprintf(out, "\n");
		}
		fflush(stderr);
		dup2(saved, 2);
		close(devnull);
		close(saved);
		fq.reads[slot][0] = NULL;
	}
	fclose(out);
	return 0;
}
