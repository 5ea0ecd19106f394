#include "query_driver.hpp"

#include <dirent.h>
#include <math.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <set>
#include <sstream>
#include <thread>

namespace cammiq {

static uint64_t nowMs() {
	return (uint64_t) std::chrono::duration_cast<std::chrono::milliseconds>(
		std::chrono::high_resolution_clock::now().time_since_epoch()).count();
}

void FqReader::die(const char *what) {
	fprintf(stderr, "%s: %s\n", what, cq_last_error());
	abort(); // the reference aborts on every error (query.cpp:151-153, binaryio.cpp:190-197)
}

FqReader::FqReader(uint32_t hl_u, const std::string &idx_u, uint32_t hl_d, const std::string &idx_d,
		const std::string &map_fn, const std::string &output_fn, float erate, bool debug) {
	hash_len_u = hl_u;
	hash_len_d = hl_d;
	IDXFILEU = idx_u;
	IDXFILED = idx_d;
	IDXDIR = "./";
	size_t found = IDXFILEU.find_last_of("/");
	if (found != IDXFILEU.npos)
		IDXDIR = IDXFILEU.substr(0, found + 1);
	MAPFILE = map_fn;
	OUTPUTFILE = output_fn;
	erate_ = erate;
	debug_display = debug;
}

FqReader::~FqReader() {
	if (ctx_thread.joinable())
		ctx_thread.join();
	cq_multi_destroy(multi);
	if (index != NULL)
		cq_index_free(index);
	for (auto g : genomes)
		delete g;
	for (auto r : reads)
		delete r;
}

// FqReader::loadIdx_p (query.cpp:109-123): both index files are decoded on two host threads
// inside cq_index_load and flattened into the device layout.
void FqReader::startContexts() {
	ctx_rc = 0;
	ctx_thread = std::thread([this]() {
		// one context per GPU (and their NCCL communicator when there are several)
		if ((ctx_rc = cq_multi_create(n_gpus, NULL, &multi)) != 0)
			ctx_err = cq_last_error(); // the error text is per thread: kept for the report
	});
}

void FqReader::loadIdx_p() {
	uint64_t start = nowMs();
	startContexts(); // CUDA start-up overlaps the index decode
	if (cq_index_load(IDXFILEU.c_str(), IDXFILED.c_str(), 0.0, &index) != 0) {
		fprintf(stderr, "%s\n", cq_last_error());
		abort();
	}
	cq_index_info info;
	cq_index_get_info(index, &info);
	fprintf(stderr, "Index: %s\nHash Length: %d\n", IDXFILEU.c_str(), (int) info.hash_len);
	fprintf(stderr, "Index: %s\nHash Length: %d\n", IDXFILED.c_str(), (int) info.hash_len);
	if ((hash_len_u > 0 && hash_len_u != info.hash_len) || (hash_len_d > 0 && hash_len_d != info.hash_len)) {
		fprintf(stderr, "The hash length given with -h does not match the one encoded in the index.\n");
		abort(); // assert(ht->getHashLength() == hash_len), query.hpp:78-90
	}
	hash_len_u = hash_len_d = info.hash_len;
	rcount_u.assign((size_t) info.n_leaves_u, 0);
	rcount_d.assign((size_t) info.n_leaves_d, 0);
	fprintf(stderr, "Loaded index files into memory.\n");
	fprintf(stderr, "Time for loading index: %lu ms.\n", (unsigned long) (nowMs() - start));
}

// FqReader::loadSmap (query.cpp:125-156): fasta_name \t genome_id \t taxid \t name; one Genome
// per first appearance of a taxid, a repeated taxid only extends the name of genomes[id].
void FqReader::loadSmap() {
	genomes.push_back(NULL);
	std::string line, gname, id, taxid;
	std::ifstream in(MAPFILE.c_str());
	std::set<uint32_t> taxids;
	if (!in.is_open()) {
		fprintf(stderr, "Can not open map file %s.\n", MAPFILE.c_str());
		abort();
	}
	while (std::getline(in, line)) {
		std::istringstream ls(line);
		while (std::getline(ls, gname, '\t')) {
			std::getline(ls, id, '\t');
			std::getline(ls, taxid, '\t');
			std::getline(ls, gname, '\t');
		}
		uint32_t taxid_ = (uint32_t) atoi(taxid.c_str());
		if (taxids.count(taxid_) != 0) {
			size_t gi = (size_t) atoi(id.c_str());
			if (gi < genomes.size() && genomes[gi] != NULL)
				genomes[gi]->name += ('/' + gname);
		} else {
			Genome *g = new Genome();
			g->taxID = taxid_;
			g->name = gname;
			genomes.push_back(g);
			taxids.insert(taxid_);
		}
	}
	fprintf(stderr, "Loaded genome map file.\n");

	// The index becomes resident here: the counters are sized by the number of genomes.
	const uint32_t G = (uint32_t) genomes.size() - 1;
	const uint64_t t_ctx = nowMs();
	if (ctx_thread.joinable())
		ctx_thread.join();
	else
		startContexts(), ctx_thread.join();
	const uint64_t t_up = nowMs();
	if (ctx_rc != 0 || multi == NULL) {
		fprintf(stderr, "Cannot create the GPU context: %s\n", ctx_err.c_str());
		abort();
	}
	if (cq_multi_upload(multi, index, G) != 0)
		die("Cannot place the index on the GPU");
	if (getenv("CAMMIQ_VERBOSE"))
		fprintf(stderr, "[cammiq] waited %lu ms for the GPU context(s), index upload %lu ms\n",
			(unsigned long) (t_up - t_ctx), (unsigned long) (nowMs() - t_up));
}

// FqReader::loadGenomeLength (query.cpp:158-205)
void FqReader::loadGenomeLength() {
	struct { const char *fn; const char *err; int field; } files[3] = {
		{"genome_lengths.out", "Can not open genome length file.\n", 0},
		{"unique_lmer_count_u.out", "Can not open unique count file.\n", 1},
		{"unique_lmer_count_d.out", "Can not open doubly-unique count file.\n", 2}};
	for (int k = 0; k < 3; k++) {
		std::ifstream in((IDXDIR + files[k].fn).c_str());
		if (!in.is_open()) {
			fprintf(stderr, "%s", files[k].err);
			abort();
		}
		std::string line, id, val;
		while (std::getline(in, line)) {
			std::istringstream ls(line);
			ls >> id;
			ls >> val;
			size_t gi = (size_t) atoi(id.c_str());
			if (gi == 0 || gi >= genomes.size())
				continue; // the reference indexes genomes[] unchecked
			uint32_t v = (uint32_t) atoi(val.c_str());
			if (files[k].field == 0) genomes[gi]->glength = v;
			else if (files[k].field == 1) genomes[gi]->nus = v;
			else genomes[gi]->nds = v;
		}
	}
	fprintf(stderr, "Loaded genome length file.\n");
}

// FqReader::getFqList (query.cpp:207-229)
void FqReader::getFqList(const std::string &INDIR) {
	if (!qfilenames.empty()) {
		qfilenames.clear();
		fprintf(stderr, "Flushed existing file names.\n");
	}
	DIR *dir = opendir(INDIR.c_str());
	if (dir == NULL) {
		fprintf(stderr, "Input directory not exists.\n");
		abort();
	}
	struct dirent *ent;
	while ((ent = readdir(dir)) != NULL) {
		std::string filename = ent->d_name;
		if (filename.length() >= 7) {
			std::string fext = filename.substr(filename.find_last_of(".") + 1);
			if (fext == "fq" || fext == "fastq")
				qfilenames.push_back(INDIR + filename);
		}
	}
	closedir(dir);
}

void FqReader::readAll(size_t min_l) {
	// the reference seeds rand() from the clock for every file (query.cpp:375-376);
	// CAMMIQ_SEED makes the N substitution reproducible
	for (size_t i = 0; i < qfilenames.size(); i++) {
		const char *seed = getenv("CAMMIQ_SEED");
		srand(seed ? (unsigned) atoi(seed) : (unsigned) std::chrono::high_resolution_clock::now().time_since_epoch().count());
		ReadSet *rs = new ReadSet();
		readFastq(qfilenames[i], min_l, *rs); // a missing file yields an empty read set, as in the reference
		reads.push_back(rs);
		if (getenv("CAMMIQ_VERBOSE"))
			fprintf(stderr, "[cammiq] %s: %lu reads parsed and packed in %.1f ms on %d threads (%lu with N)\n",
				qfilenames[i].c_str(), (unsigned long) rs->size(), rs->parse_ms, rs->threads, (unsigned long) rs->n_with_n);
		fprintf(stderr, "Loaded query file %s.\n", qfilenames[i].c_str());
	}
}

void FqReader::getFqnameWithoutDir(size_t file_idx) {
	std::stringstream fn_stream(qfilenames[file_idx]);
	while (fn_stream.good())
		getline(fn_stream, current_filename, '/');
}

// query64_p / query64mt_p / query64_sc (query.cpp:458-1080): the whole scan is cq_query.
void FqReader::queryGpu(size_t file_idx, int mode) {
	uint64_t start = nowMs();
	if (hash_len_u != hash_len_d) {
		fprintf(stderr, "Hash lengths of the two indices differ.\n");
		abort(); // assert(hash_len_u == hash_len_d), query.cpp:460
	}
	ReadSet &rs = *reads[file_idx];
	const uint32_t G = (uint32_t) genomes.size() - 1;
	const uint64_t n = rs.size();
	fprintf(stderr, "Querying %s.\n", current_filename.c_str());
	std::vector<uint64_t> cu(G + 1, 0), cd(G + 1, 0);
	std::vector<cq_pair_count> pairs(mode == CQ_MODE_SC ? std::max<uint64_t>(n, 1) : 1);
	cq_result res;
	memset(&res, 0, sizeof(res));
	res.cnt_u = cu.data();
	res.cnt_d = cd.data();
	if (mode == CQ_MODE_P) {
		res.rcount_u = rcount_u.data();
		res.rcount_d = rcount_d.data();
	} else {
		res.pairs = pairs.data();
		res.pairs_cap = pairs.size();
	}
	// reads sharded over the GPUs (index replicated), counters combined by one NCCL reduce: all
	// of it behind the ABI; with one GPU this is cq_query_packed
	if (cq_multi_query_packed(multi, mode, rs.bases, rs.offsets.data(), 0, rs.lengths.data(), n, &res) != 0)
		die("GPU query failed");
	for (uint32_t g = 1; g <= G; g++) {
		genomes[g]->read_cnts_u = cu[g];
		genomes[g]->read_cnts_d = cd[g];
	}
	nundet = res.nundet;
	nconf = res.nconf;
	ninvalid = res.n_invalid;
	if (mode == CQ_MODE_SC) {
		read_cnts_b.clear();
		for (uint64_t i = 0; i < res.n_pairs; i++)
			read_cnts_b[std::make_pair(pairs[i].a, pairs[i].b)] = pairs[i].count;
	}
	// the reference's progress line appears before every 100 000th read (query.cpp:637-638)
	for (uint64_t nrd = 0; nrd < n; nrd += 100000)
		fprintf(stderr, "Processed %lu reads.\r", (unsigned long) (nrd + 1));
	fprintf(stderr, "\nNumber of unlabeled reads: %lu.\n", (unsigned long) nundet);
	fprintf(stderr, "Number of reads with conflict labels: %lu.\n", (unsigned long) nconf);
	if (ninvalid > 0)
		fprintf(stderr, "Number of reads shorter than the hash length or with non-ACGT bases (counted as unlabeled): %lu.\n",
			(unsigned long) ninvalid);
	fprintf(stderr, "Completed query %s.\n", current_filename.c_str());
	fprintf(stderr, "Time for query: %lu ms.\n", (unsigned long) (nowMs() - start));
}

// FqReader::resetCounters / resetCounters_sc (query.cpp:1820-1858)
void FqReader::resetCounters(bool sc) {
	uint64_t start = nowMs();
	nconf = nundet = ninvalid = 0;
	for (size_t i = 1; i < genomes.size(); i++)
		genomes[i]->read_cnts_u = genomes[i]->read_cnts_d = 0;
	if (sc)
		read_cnts_b.clear();
	else {
		std::fill(rcount_u.begin(), rcount_u.end(), 0);
		std::fill(rcount_d.begin(), rcount_d.end(), 0);
	}
	if (cq_multi_reset(multi) != 0) die("cq_reset failed");
	fprintf(stderr, "Time for resetting counters: %lu ms.\n", (unsigned long) (nowMs() - start));
}

// FqReader::outputUniqueCnts (query.cpp:1786-1818)
void FqReader::outputUniqueCnts(size_t file_idx) {
	size_t n_species = genomes.size() - 1;
	FILE *fout = fopen(OUTPUTFILE.c_str(), file_idx == 0 ? "w" : "a");
	if (fout == NULL) {
		fprintf(stderr, "Can not open output file %s.\n", OUTPUTFILE.c_str());
		abort();
	}
	if (file_idx == 0) {
		fprintf(fout, "QUERY/TAXID\t");
		for (size_t i = 1; i <= n_species; i++)
			fprintf(fout, i < n_species ? "%u\t" : "%u\n", genomes[i]->taxID);
	}
	fprintf(fout, "%s\t", current_filename.c_str());
	for (size_t i = 1; i <= n_species; i++)
		fprintf(fout, i < n_species ? "%lu\t" : "%lu\n", (unsigned long) genomes[i]->read_cnts_u);
	fclose(fout);
}

// Extension (--dump_ilp_inputs <file>): everything runILP_* reads from the scan, in the order
// it reads it (query.cpp:1100-1181): per genome the counters, then per map_sp leaf (file
// order) refIDs, ucounts, depth, rcount and the weighted coverage coefficient(s) wcov.
void FqReader::dumpIlpInputs(size_t file_idx) {
	FILE *f = fopen(ilp_dump.c_str(), file_idx == 0 ? "w" : "a");
	if (f == NULL) {
		fprintf(stderr, "Can not open output file %s.\n", ilp_dump.c_str());
		abort();
	}
	const uint32_t G = (uint32_t) genomes.size() - 1;
	ReadSet &rs = *reads[file_idx];
	uint32_t rl = rs.size() ? (uint32_t) (rs.total_length / rs.size()) : 0; // query.cpp:1087
	fprintf(f, "QUERY\t%s\treads\t%lu\trl\t%u\terate\t%g\tnundet\t%lu\tnconf\t%lu\n", current_filename.c_str(),
		(unsigned long) rs.size(), rl, erate_, (unsigned long) nundet, (unsigned long) nconf);
	for (uint32_t g = 1; g <= G; g++)
		fprintf(f, "GENOME\t%u\t%u\t%lu\t%lu\t%u\t%u\t%u\n", g, genomes[g]->taxID, (unsigned long) genomes[g]->read_cnts_u,
			(unsigned long) genomes[g]->read_cnts_d, genomes[g]->glength, genomes[g]->nus, genomes[g]->nds);
	for (int table = 0; table < 2; table++) {
		cq_leaf_view lv;
		cq_index_leaves(index, table, &lv);
		std::vector<uint64_t> off(G + 2), ids;
		uint64_t total = 0;
		cq_index_map_sp(index, table, G, off.data(), NULL, &total);
		ids.resize(total ? total : 1);
		cq_index_map_sp(index, table, G, off.data(), ids.data(), &total);
		const std::vector<uint32_t> &rc = table == 0 ? rcount_u : rcount_d;
		for (uint32_t g = 1; g <= G; g++)
			for (uint64_t k = off[g]; k < off[g + 1]; k++) {
				uint64_t l = ids[k];
				// runILP_* receives erate as a double converted from the float option (query.cpp:252)
				const double er = erate_;
				const uint32_t d = lv.depth[l];
				double w1 = rl ? (lv.ucount1[l] * (rl - d) * 1.0 / rl) * pow(1 - er, d) : 0.0;
				double w2 = rl ? (lv.ucount2[l] * (rl - d) * 1.0 / rl) * pow(1 - er, d) : 0.0;
				fprintf(f, "%s\t%u\t%lu\t%u\t%u\t%u\t%u\t%u\t%u\t%.17g\t%.17g\n", table == 0 ? "LEAFU" : "LEAFD", g,
					(unsigned long) l, lv.ref_id1[l], lv.ref_id2[l], lv.ucount1[l], lv.ucount2[l], lv.depth[l], rc[l], w1, w2);
			}
	}
	fclose(f);
}

// FqReader::queryFastq_p (query.cpp:231-301).  The ILP (runILP_cplex / runILP_gurobi) is only
// compiled into the reference when a solver is present; this build, like a reference build
// without CPLEX/GUROBI, stops after the scan.
void FqReader::queryFastq_p(const std::vector<std::string> &files, size_t min_l) {
	if (!qfilenames.empty()) {
		qfilenames.clear();
		fprintf(stderr, "Flushed existing file names.\n");
	}
	qfilenames = files;
	readAll(min_l);
	loadGenomeLength();
	for (size_t fq_idx = 0; fq_idx < qfilenames.size(); fq_idx++) {
		getFqnameWithoutDir(fq_idx);
		queryGpu(fq_idx, CQ_MODE_P);
		if (!ilp_dump.empty())
			dumpIlpInputs(fq_idx);
		if (fq_idx + 1 < qfilenames.size())
			resetCounters(false);
	}
}

void FqReader::queryFastq_p(const std::string &dir, size_t min_l) {
	getFqList(dir);
	std::vector<std::string> files = qfilenames;
	qfilenames.clear();
	queryFastq_p(files, min_l);
}

// FqReader::queryFastq_sc (query.cpp:303-369)
void FqReader::queryFastq_sc(int id_mode, const std::vector<std::string> &files, size_t min_l, bool load_lengths) {
	if (!qfilenames.empty()) {
		qfilenames.clear();
		fprintf(stderr, "Flushed existing file names.\n");
	}
	qfilenames = files;
	readAll(min_l);
	if (load_lengths)
		loadGenomeLength();
	for (size_t fq_idx = 0; fq_idx < qfilenames.size(); fq_idx++) {
		getFqnameWithoutDir(fq_idx);
		if (nthreads > 1)
			fprintf(stderr, "Single cell queries only support one thread.\n");
		queryGpu(fq_idx, CQ_MODE_SC);
		if (id_mode <= 1)
			outputUniqueCnts(fq_idx);
		if (fq_idx + 1 < qfilenames.size())
			resetCounters(true);
	}
}

void FqReader::queryFastq_sc(int id_mode, const std::string &dir, size_t min_l) {
	getFqList(dir);
	std::vector<std::string> files = qfilenames;
	qfilenames.clear();
	queryFastq_sc(id_mode, files, min_l, false); // the directory variant does not load the lengths (query.cpp:303-332)
}

} // namespace cammiq
