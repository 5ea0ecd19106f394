// `cammiq` -- drop-in command line for the query side of CAMMiQ, backed by libcammiq_gpu.
// Mirrors the reference's hand-rolled argv loop for `--query` (main.cpp:74-446, 519-549):
// same options, same validation messages, same dispatch to queryFastq_p / queryFastq_sc.
// `--build` is not part of this path: indices are produced by the reference's builder, whose
// files this program reads unchanged.
// Extensions (not in the reference): --gpus <n>, --dump_ilp_inputs <file>.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <unistd.h>

#include "fastq_reader.hpp"
#include "../pack_reads.hpp"
#include "query_driver.hpp"

static bool validFile(const char *fn) {
	FILE *fd = fopen(fn, "r");
	if (fd == NULL)
		return false;
	fclose(fd);
	return true;
}

static void printUsage() {
	fprintf(stderr, "\n");
	fprintf(stderr, "CAMMiQ (GPU query path): Metagenomic microbial abundance quantification.\n\n");
	fprintf(stderr, "Usage: ./cammiq --query (--<options_for_query>) parameters\n\n");
	fprintf(stderr, "options_for_query = read_cnts | doubly_unique.\n");
	fprintf(stderr, "-h <integers>,\t hash length(s): integer, default is the one encoded in the index.\n");
	fprintf(stderr, "-f <strings>,\t map file name.\n");
	fprintf(stderr, "-Q <string>,\t directoty containing the fastq files.\n");
	fprintf(stderr, "-q <strings>,\t querying file names separated by space.\n");
	fprintf(stderr, "-i <strings>,\t indexing file names separated by space.\n");
	fprintf(stderr, "-o <string>,\t output file name.\n");
	fprintf(stderr, "-e <float>,\t expected sequencing error probability in queries.\n");
	fprintf(stderr, "-t <integer>,\t number of threads (accepted for compatibility; the scan runs on the GPU).\n");
	fprintf(stderr, "--gpus <integer>,\t number of GPUs to shard the reads over (default 1).\n");
	fprintf(stderr, "--dump_ilp_inputs <string>,\t write the counters and per-leaf coefficients the ILP reads.\n\n");
	fprintf(stderr, "--build is served by the reference's builder; its index files are read unchanged.\n");
	fprintf(stderr, "--help,\t to print user options.\n");
}

#define NEED_VALUE(msg)                      \
	if (++i >= argc) {                       \
		fprintf(stderr, msg);                \
		exit(EXIT_FAILURE);                  \
	}

// --dump_reads <fastq> [min_len] (diagnostic, not in the reference): what readFastq hands the
// scan, one line per read: "<length> <bases>", '!' for a read the kernel will treat as invalid.
static int dumpReads(const char *path, size_t min_len) {
	const char *seed = getenv("CAMMIQ_SEED");
	srand(seed ? (unsigned) atoi(seed) : 1u);
	cammiq::ReadSet rs;
	if (!cammiq::readFastq(path, min_len, rs)) {
		fprintf(stderr, "Cannot open %s.\n", path);
		return 1;
	}
	std::vector<uint8_t> line(256);
	for (size_t i = 0; i < rs.size(); i++) {
		cammiq::unpackRead(rs.bases + rs.offsets[i], rs.lengths[i], line.data());
		printf("%u %.*s\n", (unsigned) rs.lengths[i], (int) rs.lengths[i], (const char *) line.data());
	}
	fprintf(stderr, "%lu reads, total length %lu, %lu with N, %d threads, %.1f ms\n", (unsigned long) rs.size(),
		(unsigned long) rs.total_length, (unsigned long) rs.n_with_n, rs.threads, rs.parse_ms);
	return 0;
}

int main(int argc, char **argv) {
	if (argc >= 3 && std::string(argv[1]) == "--dump_reads")
		return dumpReads(argv[2], argc >= 4 ? (size_t) atoi(argv[3]) : 0);
	if (argc == 2) {
		std::string val(argv[1]);
		if (val == "--help" || val == "--HELP") {
			printUsage();
			return 0;
		}
		exit(EXIT_FAILURE);
	}
	int t = 1, h = -1, h1 = -1, h2 = -1, n_gpus = 1;
	std::string fm_name = "", fi_name1 = "./index_u.bin1", fi_name2 = "./index_d.bin2", fq_dir = "";
	std::vector<std::string> fq_names;
	int mode = -1, id_mode = 0;
	bool debug_info = 0;
	std::string output, ilp_dump;
	float erate = 0.01;
	size_t min_rl = 0;

	for (int i = 1; i < argc; i++) {
		std::string val(argv[i]);
		if (val == "--build") {
			fprintf(stderr, "Index construction is not part of the GPU query path: build the index with the reference's cammiq --build.\n");
			exit(EXIT_FAILURE);
		}
		if (val == "--query") {
			mode = 1;
			continue;
		}
		if (val == "--unique" || val == "--both") {
			fprintf(stderr, "Option --unique is only valid in mode BUILD.\n");
			exit(EXIT_FAILURE);
		}
		if (val == "--doubly_unique") {
			if (id_mode == 0) {
				fprintf(stderr, "Option --doubly_unique is only valid in --read_cnts queries.\n");
				exit(EXIT_FAILURE);
			}
			id_mode = 2;
			continue;
		}
		if (val == "--read_cnts") {
			if (mode <= 0) {
				fprintf(stderr, "Option --read_cnts is only valid in mode QUERY.\n");
				exit(EXIT_FAILURE);
			}
			id_mode = 1;
			continue;
		}
		if (val == "--enable_ilp_display") {
			if (mode <= 0) {
				fprintf(stderr, "Option --enable_ilp_display is only valid in mode QUERY.\n");
				exit(EXIT_FAILURE);
			}
			debug_info = 1;
			continue;
		}
		if (val == "--read_length_filter") {
			if (mode <= 0) {
				fprintf(stderr, "Option --read_length_filter is only valid in mode QUERY.\n");
				exit(EXIT_FAILURE);
			}
			NEED_VALUE("Please specify a parameter value for --read_length_filter.\n");
			min_rl = (size_t) atoi(argv[i]);
			continue;
		}
		// fine-grained ILP parameters: parsed for compatibility, consumed by the solver stage only
		if (val == "--read_cnt_thres" || val == "--easy_to_identify_thres" || val == "--ilp_epsilon" ||
			val == "--ilp_alpha" || val == "--max_depth") {
			if (id_mode == 1) {
				fprintf(stderr, "Option %s is not valid in READ_CNTS queries.\n", val.c_str());
				exit(EXIT_FAILURE);
			}
			if (++i >= argc) {
				fprintf(stderr, "Please specify a parameter value for %s.\n", val.c_str());
				exit(EXIT_FAILURE);
			}
			continue;
		}
		if (val == "--unique_read_cnt_thres" || val == "--doubly_unique_read_cnt_thres") {
			if (id_mode == 0) {
				fprintf(stderr, "Option %s is only valid in READ_CNTS queries.\n", val.c_str());
				exit(EXIT_FAILURE);
			}
			if (++i >= argc) {
				fprintf(stderr, "Please specify a parameter value for %s.\n", val.c_str());
				exit(EXIT_FAILURE);
			}
			continue;
		}
		if (val == "-k" || val == "-L" || val == "-Lmax") {
			fprintf(stderr, "Parameter %s is only valid in mode BUILD.\n", val.c_str() + 1);
			exit(EXIT_FAILURE);
		}
		if (val == "-i") {
			NEED_VALUE("Please specify index file names.\n");
			while (i < argc && argv[i][0] != '-') {
				std::string filename = argv[i++];
				std::string ext = filename.substr(filename.find_last_of(".") + 1);
				if (ext == "idx1" || ext == "bin1")
					fi_name1 = filename;
				if (ext == "idx2" || ext == "bin2")
					fi_name2 = filename;
			}
			i--;
			continue;
		}
		if (val == "-o") {
			if (mode <= 0) {
				fprintf(stderr, "Parameter o is only valid in mode QUERY.\n");
				exit(EXIT_FAILURE);
			}
			NEED_VALUE("Please specify an output file name.\n");
			output = argv[i];
			continue;
		}
		if (val == "-e") {
			if (mode <= 0) {
				fprintf(stderr, "Parameter e is only valid in mode QUERY.\n");
				exit(EXIT_FAILURE);
			}
			NEED_VALUE("Please specify the expected sequencing error rate.\n");
			erate = (float) atof(argv[i]);
			if (erate < 0.0 || erate > 0.2) {
				fprintf(stderr, "The error rate should be in range [0, 0.2].\n");
				exit(EXIT_FAILURE);
			}
			continue;
		}
		if (val == "-h") {
			NEED_VALUE("Please specify hash length as an integer.\n");
			if (i + 1 < argc && argv[i + 1][0] != '-') {
				h1 = atoi(argv[i++]);
				h2 = atoi(argv[i]);
			} else
				h = atoi(argv[i]);
			if ((h != -1 && (h <= 4 || h >= 32)) || (h1 != -1 && (h1 <= 4 || h1 >= 32)) || (h2 != -1 && (h2 <= 4 || h2 >= 32))) {
				fprintf(stderr, "The hash length should be in range [5, 31].\n");
				exit(EXIT_FAILURE);
			}
			continue;
		}
		if (val == "-t") {
			NEED_VALUE("Please specify the worker threads number.\n");
			t = atoi(argv[i]);
			continue;
		}
		if (val == "--gpus") {
			NEED_VALUE("Please specify the number of GPUs.\n");
			n_gpus = atoi(argv[i]);
			if (n_gpus < 1 || n_gpus > 64) {
				fprintf(stderr, "The number of GPUs should be in range [1, 64].\n");
				exit(EXIT_FAILURE);
			}
			continue;
		}
		if (val == "--dump_ilp_inputs") {
			NEED_VALUE("Please specify a file name for --dump_ilp_inputs.\n");
			ilp_dump = argv[i];
			continue;
		}
		if (val == "-f") {
			NEED_VALUE("Please specify file names.\n");
			while (i < argc && argv[i][0] != '-') {
				std::string filename = argv[i++];
				std::string ext = filename.substr(filename.find_last_of(".") + 1);
				if (ext == "out" || ext == "map")
					fm_name = filename;
			}
			i--;
			continue;
		}
		if (val == "-q") {
			NEED_VALUE("Please specify query file names.\n");
			while (i < argc && argv[i][0] != '-') {
				std::string filename = argv[i++];
				std::string ext = filename.substr(filename.find_last_of(".") + 1);
				if (ext == "fq" || ext == "fastq") {
					if (validFile(filename.c_str()))
						fq_names.push_back(filename);
					else {
						fprintf(stderr, "Failed to find input file %s.\n", filename.c_str());
						exit(EXIT_FAILURE);
					}
				}
			}
			i--;
			continue;
		}
		if (val == "-Q") {
			NEED_VALUE("Please specify the directory containing fastq files.\n");
			fq_dir = argv[i];
			if (!validFile(fq_dir.c_str())) {
				fprintf(stderr, "Failed to find input directory %s.\n", fq_dir.c_str());
				exit(EXIT_FAILURE);
			}
			continue;
		}
		if (val == "-D") {
			NEED_VALUE("Please specify the directory containing fasta files.\n");
			continue;
		}
		fprintf(stderr, "Failed to recognize option: %s. \n", val.c_str());
		exit(EXIT_FAILURE);
	}

	if (mode != 1)
		return 0;
	cammiq::FqReader *fqr;
	if (h == -1) {
		if (h1 == -1 || h2 == -1) {
			fprintf(stderr, "Warning: Hash length not specified, using that encoded in the index.\n");
			fqr = new cammiq::FqReader(0, fi_name1, 0, fi_name2, fm_name, output, erate, debug_info);
		} else
			fqr = new cammiq::FqReader((uint32_t) h1, fi_name1, (uint32_t) h2, fi_name2, fm_name, output, erate, debug_info);
	} else
		fqr = new cammiq::FqReader((uint32_t) h, fi_name1, (uint32_t) h, fi_name2, fm_name, output, erate, debug_info);
	fqr->n_gpus = n_gpus;
	fqr->ilp_dump = ilp_dump;
	fqr->loadIdx_p();
	fqr->loadSmap();
	fqr->nthreads = t;
	if (!fq_names.empty()) {
		if (id_mode == 0)
			fqr->queryFastq_p(fq_names, min_rl);
		else
			fqr->queryFastq_sc(id_mode, fq_names, min_rl, true);
	} else if (fq_dir != "") {
		if (id_mode == 0)
			fqr->queryFastq_p(fq_dir, min_rl);
		else
			fqr->queryFastq_sc(id_mode, fq_dir, min_rl);
	} else {
		fprintf(stderr, "Please specify at least one query file or directory.\n");
		exit(EXIT_FAILURE);
	}
	// Everything is written and closed: leave without tearing down the CUDA context and the
	// multi-GB host arrays one by one (CAMMIQ_CLEAN_EXIT=1 keeps the orderly path for leak checks).
	if (getenv("CAMMIQ_CLEAN_EXIT") == NULL) {
		fflush(NULL);
		_exit(0);
	}
	delete fqr;
	return 0;
}
