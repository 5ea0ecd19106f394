// Host-side query orchestration: the GPU-backed stand-in for class FqReader
// (reference: query.hpp:26-139, query.cpp:24-456, 1786-1858).  Same constructor arguments,
// same call sequence from main (loadIdx_p, loadSmap, queryFastq_p / queryFastq_sc), same
// stderr lines and output files; query64_p / query64mt_p / query64_sc are one call into the
// C ABI (cq_query) instead of the CPU loops.
#ifndef CAMMIQ_QUERY_DRIVER_HPP
#define CAMMIQ_QUERY_DRIVER_HPP

#include <cstdint>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/cammiq_gpu.h"
#include "fastq_reader.hpp"

namespace cammiq {

struct Genome { // query.hpp:12-24
	uint64_t read_cnts_u = 0, read_cnts_d = 0;
	uint32_t glength = 0, nus = 0, nds = 0, taxID = 0;
	std::string name;
};

class FqReader {
public:
	int nthreads = 1;  // -t: kept for CLI compatibility; the scan runs on the GPU(s)
	int n_gpus = 1;    // --gpus extension: reads sharded over devices 0..n_gpus-1
	std::string ilp_dump; // --dump_ilp_inputs extension

	// hash lengths of 0 mean "use the one encoded in the index" (query.cpp:34-84)
	FqReader(uint32_t hl_u, const std::string &idx_u, uint32_t hl_d, const std::string &idx_d,
			const std::string &map_fn, const std::string &output_fn, float erate, bool debug);
	~FqReader();

	void loadIdx_p();
	void loadSmap();
	void loadGenomeLength();
	void getFqList(const std::string &dir);
	void queryFastq_p(const std::vector<std::string> &files, size_t min_l);
	void queryFastq_p(const std::string &dir, size_t min_l);
	void queryFastq_sc(int id_mode, const std::vector<std::string> &files, size_t min_l, bool load_lengths);
	void queryFastq_sc(int id_mode, const std::string &dir, size_t min_l);

private:
	void readAll(size_t min_l);
	void getFqnameWithoutDir(size_t file_idx);
	void queryGpu(size_t file_idx, int mode); // query64_p / query64mt_p / query64_sc
	void outputUniqueCnts(size_t file_idx);
	void resetCounters(bool sc);
	void dumpIlpInputs(size_t file_idx);
	void die(const char *what);

	std::vector<std::string> qfilenames;
	std::string current_filename;
	std::vector<ReadSet *> reads;
	size_t nconf = 0, nundet = 0, ninvalid = 0;
	std::string MAPFILE, IDXFILEU, IDXFILED, IDXDIR, OUTPUTFILE;
	std::vector<Genome *> genomes; // index 0 unused (NULL), query.cpp:126
	std::map<std::pair<uint32_t, uint32_t>, uint64_t> read_cnts_b;
	uint32_t hash_len_u, hash_len_d;
	float erate_;
	bool debug_display;

	cq_index *index = NULL;
	cq_multi *multi = NULL; // the GPU(s): cq_multi_* shards reads and reduces the counters
	// the GPU contexts come up on a thread of their own while the host decodes the index
	std::thread ctx_thread;
	int ctx_rc = 0;
	std::string ctx_err; // cq_last_error() is per thread: kept for the report
	void startContexts();
	std::vector<uint32_t> rcount_u, rcount_d; // pleafNode::rcount in file order
};

} // namespace cammiq
#endif
