// The reference-side binding of libcammiq_gpu.so: the three member functions that ARE the hot
// path of the reference,
//
//     void FqReader::query64_p(size_t file_idx)     query.cpp:458   standard mode, 1 thread
//     void FqReader::query64mt_p(size_t file_idx)   query.cpp:650   standard mode, OpenMP
//     void FqReader::query64_sc(size_t file_idx)    query.cpp:891   --read_cnts mode
//
// (declared at query.hpp:113-115, called from queryFastq_p / queryFastq_sc, query.cpp:247-250,
// 286-289, 319, 356), re-defined over the C ABI of include/cammiq_gpu.h.  This file is compiled
// against the reference's OWN headers and linked with the reference's OWN main.cpp, query.cpp
// (loaders, FASTQ reader, ILP set-up, outputUniqueCnts), hashtrie.cpp, binaryio.cpp, build.cpp and
// gsa.cpp, none of them edited: oracle/Makefile weakens the three symbols in the reference's
// query.o (objcopy --weaken-symbol), so the definitions below win at link time and every call
// site of the reference reaches the GPU.  Everything else of `cammiq --query` stays the
// reference's code.
//
// State the reference functions read (query.hpp:30-63): reads[file_idx] (one heap block per read),
// rlengths[file_idx], ht_u / ht_d, genomes, current_filename.  State they write: Genome::
// read_cnts_u / read_cnts_d, nundet, nconf, pleafNode::rcount (P modes), read_cnts_b (SC mode),
// and the stderr lines of query.cpp:472, 637-647.  All of it is reproduced here; counters are
// ADDED to the reference's fields, as the reference's loops do, so resetCounters(_sc) keeps its
// meaning.
//
// The reference class cannot be given new members without editing query.hpp, so the GPU handles
// of an FqReader live in a side table keyed by the object's address.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "query.hpp" // the reference's header, from where it lies

#include "../../../include/cammiq_gpu.h"

namespace {

struct GpuState {
	cq_index *index;
	cq_ctx *ctx;
	size_t n_genomes;
	std::vector<pleafNode *> leaf_u, leaf_d; // file-order leaf id -> the reference's node
};

std::map<const void *, GpuState> g_states;

void die(const char *what) {
	fprintf(stderr, "%s: %s\n", what, cq_last_error());
	abort();
}

// Hash::map_sp[rid] lists the leaves of genome rid in file order (hashtrie.cpp:452-453, 476);
// cq_index_map_sp returns the library's file-order leaf ids of rid in the same order.  Zipping
// the two gives id -> pleafNode* without touching the reference's loader.
void mapLeaves(const cq_index *index, int table, Hash *ht, size_t G, uint64_t n_leaves, std::vector<pleafNode *> &out) {
	std::vector<uint64_t> off(G + 2, 0);
	uint64_t total = 0;
	if (cq_index_map_sp(index, table, (uint32_t) G, off.data(), NULL, &total) != 0)
		die("cq_index_map_sp");
	std::vector<uint64_t> ids(total ? total : 1);
	if (cq_index_map_sp(index, table, (uint32_t) G, off.data(), ids.data(), &total) != 0)
		die("cq_index_map_sp");
	out.assign(n_leaves, (pleafNode *) NULL);
	for (size_t rid = 1; rid <= G; rid++) {
		auto it = ht->map_sp.find((uint32_t) rid);
		const size_t have = it == ht->map_sp.end() ? 0 : it->second.size();
		if (have != off[rid + 1] - off[rid]) {
			fprintf(stderr, "GPU binding: genome %zu has %zu leaves in the reference's index and %lu in the library's.\n",
				rid, have, (unsigned long) (off[rid + 1] - off[rid]));
			abort();
		}
		for (size_t k = 0; k < have; k++)
			out[ids[off[rid] + k]] = it->second[k];
	}
	for (uint64_t l = 0; l < n_leaves; l++)
		if (out[l] == NULL) {
			fprintf(stderr, "GPU binding: leaf %lu of table %d belongs to no genome of the map file.\n", (unsigned long) l, table);
			abort();
		}
}

// Everything the three members hand over: pointers to the (private) FqReader state they may touch.
struct Call {
	const void *self;
	const std::vector<uint8_t *> *reads;
	const std::vector<uint8_t> *rlengths;
	std::vector<Genome *> *genomes;
	Hash *ht_u, *ht_d;
	const std::string *idx_u, *idx_d, *filename;
	size_t *nundet, *nconf;
	std::map<std::pair<uint32_t, uint32_t>, uint64_t> *read_cnts_b;
};

GpuState &stateOf(const Call &c) {
	auto it = g_states.find(c.self);
	if (it != g_states.end())
		return it->second;
	// first query of this FqReader: the index goes to the device once (the reference has already
	// loaded its own copy, which the ILP set-up keeps using)
	GpuState st;
	st.index = NULL;
	st.ctx = NULL;
	st.n_genomes = c.genomes->size() - 1; // genomes[0] is NULL (query.cpp:126)
	if (cq_index_load(c.idx_u->c_str(), c.idx_d->c_str(), 0.0, &st.index) != 0)
		die("cq_index_load");
	const char *dev = getenv("CAMMIQ_DEVICE");
	if (cq_ctx_create(dev ? atoi(dev) : 0, NULL, &st.ctx) != 0)
		die("cq_ctx_create");
	if (cq_index_upload(st.ctx, st.index, (uint32_t) st.n_genomes) != 0)
		die("cq_index_upload");
	cq_index_info info;
	if (cq_index_get_info(st.index, &info) != 0)
		die("cq_index_get_info");
	mapLeaves(st.index, CQ_TABLE_U, c.ht_u, st.n_genomes, info.n_leaves_u, st.leaf_u);
	mapLeaves(st.index, CQ_TABLE_D, c.ht_d, st.n_genomes, info.n_leaves_d, st.leaf_d);
	return g_states.insert(std::make_pair(c.self, st)).first->second;
}

void gpuQuery(const Call &c, int mode) {
	auto start = std::chrono::high_resolution_clock::now();
	GpuState &st = stateOf(c);
	fprintf(stderr, "Querying %s.\n", c.filename->c_str());
	// the reference keeps one heap block per read (query.cpp:384); the ABI takes one buffer
	const size_t n = c.reads->size(), G = st.n_genomes;
	std::vector<uint64_t> offsets(n ? n : 1);
	size_t total = 0;
	for (size_t i = 0; i < n; i++) {
		offsets[i] = total;
		total += (*c.rlengths)[i];
	}
	std::vector<uint8_t> bases(total ? total : 1);
	for (size_t i = 0; i < n; i++)
		memcpy(&bases[offsets[i]], (*c.reads)[i], (*c.rlengths)[i]);
	std::vector<uint64_t> cu(G + 1, 0), cd(G + 1, 0);
	std::vector<uint32_t> ru, rd;
	std::vector<cq_pair_count> pairs;
	cq_result r;
	memset(&r, 0, sizeof(r));
	r.cnt_u = cu.data();
	r.cnt_d = cd.data();
	if (mode == CQ_MODE_P) {
		ru.assign(st.leaf_u.size() ? st.leaf_u.size() : 1, 0);
		rd.assign(st.leaf_d.size() ? st.leaf_d.size() : 1, 0);
		r.rcount_u = ru.data();
		r.rcount_d = rd.data();
	} else {
		pairs.resize(n + 1);
		r.pairs = pairs.data();
		r.pairs_cap = pairs.size();
	}
	// the device accumulators hold this call alone; its sums are added to the reference's fields
	if (cq_reset(st.ctx) != 0)
		die("cq_reset");
	if (cq_query(st.ctx, mode, bases.data(), offsets.data(), 0, c.rlengths->data(), n, &r) != 0)
		die("cq_query");
	for (size_t g = 1; g <= G; g++) {
		(*c.genomes)[g]->read_cnts_u += cu[g];
		(*c.genomes)[g]->read_cnts_d += cd[g];
	}
	*c.nundet += r.nundet;
	*c.nconf += r.nconf;
	if (mode == CQ_MODE_P) {
		for (size_t l = 0; l < st.leaf_u.size(); l++)
			st.leaf_u[l]->rcount += ru[l];
		for (size_t l = 0; l < st.leaf_d.size(); l++)
			st.leaf_d[l]->rcount += rd[l];
	} else {
		for (uint64_t i = 0; i < r.n_pairs; i++)
			(*c.read_cnts_b)[std::make_pair(pairs[i].a, pairs[i].b)] += pairs[i].count;
	}
	// the reference's progress line appears before every 100 000th read (query.cpp:637-638)
	for (size_t nrd = 0; nrd < n; nrd += 100000)
		fprintf(stderr, "Processed %lu reads.\r", (unsigned long) (nrd + 1));
	fprintf(stderr, "\nNumber of unlabeled reads: %lu.\n", (unsigned long) *c.nundet);
	fprintf(stderr, "Number of reads with conflict labels: %lu.\n", (unsigned long) *c.nconf);
	fprintf(stderr, "Completed query %s.\n", c.filename->c_str());
	auto duration = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::high_resolution_clock::now() - start).count();
	fprintf(stderr, "Time for query: %lu ms.\n", (unsigned long) duration);
}

} // namespace

#define CAMMIQ_GPU_CALL() \
	{ this, &reads[file_idx], &rlengths[file_idx], &genomes, ht_u, ht_d, &IDXFILEU, &IDXFILED, &current_filename, &nundet, &nconf, &read_cnts_b }

void FqReader::query64_p(size_t file_idx) {
	assert(hash_len_u == hash_len_d);
	const Call c = CAMMIQ_GPU_CALL();
	gpuQuery(c, CQ_MODE_P);
}

// the OpenMP variant computes the same counters (query.cpp:650-889); one GPU path serves both
void FqReader::query64mt_p(size_t file_idx) {
	assert(hash_len_u == hash_len_d);
	const Call c = CAMMIQ_GPU_CALL();
	gpuQuery(c, CQ_MODE_P);
}

void FqReader::query64_sc(size_t file_idx) {
	assert(hash_len_u == hash_len_d);
	const Call c = CAMMIQ_GPU_CALL();
	gpuQuery(c, CQ_MODE_SC);
}
