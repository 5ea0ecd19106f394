// FASTQ ingest for the query path -- the host-side state query64_* consumes
// (reference: FqReader::readFastq, query.cpp:371-425).  Reads of one file are stored back to
// back in ONE page-locked buffer (bases) with per-read offsets and uint8 lengths, which is the
// layout cq_query streams to the GPU; the reference keeps one heap block per read.
#ifndef CAMMIQ_FASTQ_READER_HPP
#define CAMMIQ_FASTQ_READER_HPP

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace cammiq {

struct ReadSet {
	uint8_t *bases = NULL;   // pinned (cq_host_alloc) when a GPU is present, else malloc
	bool pinned = false;
	size_t cap_bases = 0;
	std::vector<uint64_t> offsets;
	std::vector<uint8_t> lengths; // (uint8_t) line length, as the reference stores it (query.cpp:387)
	uint64_t total_length = 0;    // sum of the FULL line lengths (FqReader::tlengths)
	uint64_t n_bases = 0;
	~ReadSet();
	void clear();
	size_t size() const { return lengths.size(); }
};

// 4-line records, bases = line 2.  Every 'N' of a read is replaced by ONE random base drawn
// per read (alphabet[rand() & 3], query.cpp:383); rand() is consumed once per read whether or
// not it holds an N, like the reference.  Reads whose line is shorter than min_len are skipped
// (query.cpp:410).  Returns false when the file cannot be opened.
bool readFastq(const std::string &path, size_t min_len, ReadSet &out);

} // namespace cammiq
#endif
