#include "fastq_reader.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../../include/cammiq_gpu.h"

namespace cammiq {

ReadSet::~ReadSet() { clear(); }

void ReadSet::clear() {
	if (bases != NULL) {
		if (pinned)
			cq_host_free(bases);
		else
			free(bases);
	}
	bases = NULL;
	cap_bases = 0;
	offsets.clear();
	lengths.clear();
	total_length = 0;
	n_bases = 0;
}

namespace {

bool slurp(const std::string &fn, std::vector<char> &buf) {
	FILE *f = fopen(fn.c_str(), "rb");
	if (f == NULL)
		return false;
	fseek(f, 0, SEEK_END);
	long n = ftell(f);
	fseek(f, 0, SEEK_SET);
	buf.resize((size_t) n);
	bool ok = n == 0 || fread(buf.data(), 1, (size_t) n, f) == (size_t) n;
	fclose(f);
	return ok;
}

} // namespace

bool readFastq(const std::string &path, size_t min_len, ReadSet &out) {
	static const char alphabet[4] = {'A', 'C', 'G', 'T'};
	out.clear();
	std::vector<char> file;
	if (!slurp(path, file))
		return false;
	// one buffer for all bases: the file size bounds them
	void *p = NULL;
	size_t cap = file.size() + 64;
	if (cq_host_alloc(cap, &p) == 0) {
		out.bases = (uint8_t *) p;
		out.pinned = true;
	} else {
		out.bases = (uint8_t *) malloc(cap);
		out.pinned = false;
	}
	out.cap_bases = cap;
	const char *cur = file.data(), *end = cur + file.size();
	auto nextLine = [&](const char *&b, const char *&e) -> bool {
		if (cur >= end)
			return false;
		b = cur;
		const char *nl = (const char *) memchr(cur, '\n', (size_t) (end - cur));
		e = nl ? nl : end;
		cur = nl ? nl + 1 : end;
		return true;
	};
	const char *b, *e;
	uint64_t at = 0;
	while (nextLine(b, e)) {        // header line (std::getline loop condition, query.cpp:380)
		if (!nextLine(b, e)) {      // bases; a missing line reads as empty like a failed getline
			b = e = end;
		}
		size_t rl = (size_t) (e - b);
		if (rl >= min_len) {
			const char sub = alphabet[rand() & 3];
			uint8_t *dst = out.bases + at;
			for (size_t i = 0; i < rl; i++)
				dst[i] = (uint8_t) (b[i] == 'N' ? sub : b[i]);
			out.offsets.push_back(at);
			out.lengths.push_back((uint8_t) rl);
			out.total_length += rl;
			at += rl;
		}
		const char *b2, *e2;
		nextLine(b2, e2); // '+'
		nextLine(b2, e2); // qualities
	}
	out.n_bases = at;
	return true;
}

} // namespace cammiq
