// FASTQ ingest for the query path -- the host-side state query64_* consumes
// (reference: FqReader::readFastq, query.cpp:371-425).
//
// The reference reads a file with one getline per line on one thread and keeps one heap block
// of ASCII per read.  Here the file is mapped, split into one slice per worker thread, and
// parsed in parallel; every read goes straight to the 2-bit packed layout the scan kernel
// consumes (include/cammiq_gpu.h, cq_query_packed), back to back in ONE page-locked buffer with
// per-read offsets and uint8 lengths.
#ifndef CAMMIQ_FASTQ_READER_HPP
#define CAMMIQ_FASTQ_READER_HPP

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace cammiq {

struct ReadSet {
	uint8_t *bases = NULL;   // packed reads; pinned (cq_host_alloc) when a GPU is present, else malloc
	bool pinned = false;
	size_t cap_bases = 0;
	std::vector<uint64_t> offsets; // first byte of every read in `bases`
	std::vector<uint8_t> lengths;  // (uint8_t) line length, as the reference stores it (query.cpp:387);
	                               // 0 for a read holding a byte outside ACGTacgt after N substitution
	uint64_t total_length = 0;     // sum of the FULL line lengths (FqReader::tlengths)
	uint64_t n_bases = 0;          // bytes used in `bases`
	uint64_t n_with_n = 0;         // reads in which an 'N' was substituted
	double parse_ms = 0;
	int threads = 0;
	~ReadSet();
	void clear();
	size_t size() const { return lengths.size(); }
};

// 4-line records located by line number alone (line 4k+1 holds the bases), exactly as the
// reference's getline loop sees them.  Every 'N' of a read is replaced by ONE random base drawn
// per read (alphabet[rand() & 3], query.cpp:383); rand() is consumed once per accepted read, in
// file order, whether or not it holds an N, like the reference.  Reads whose line is shorter
// than min_len are skipped (query.cpp:410).  threads <= 0: CAMMIQ_IO_THREADS, else
// min(16, hardware threads).  Returns false when the file cannot be opened.
bool readFastq(const std::string &path, size_t min_len, ReadSet &out, int threads = 0);

} // namespace cammiq
#endif
