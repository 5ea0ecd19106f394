#include "fastq_reader.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../../include/cammiq_gpu.h"
#include "../pack_reads.hpp"

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
	n_with_n = 0;
}

namespace {

struct LineRef {
	const char *b; // first base
	uint64_t rl;   // full line length
};

struct Slice {
	std::vector<LineRef> reads;   // accepted reads whose bases line starts in this slice
	std::vector<uint64_t> with_n; // (global) indices of reads the packer rejected
	uint64_t newlines = 0, packed_bytes = 0, total_length = 0;
};

uint64_t countNewlines(const char *p, const char *end) {
	uint64_t n = 0;
	while (p < end) {
		const char *nl = (const char *) memchr(p, '\n', (size_t) (end - p));
		if (nl == NULL)
			break;
		n++;
		p = nl + 1;
	}
	return n;
}

int ioThreads(int requested, uint64_t file_bytes) {
	int t = requested;
	if (t <= 0) {
		const char *env = getenv("CAMMIQ_IO_THREADS");
		t = env ? atoi(env) : (int) std::min(std::thread::hardware_concurrency(), 16u);
	}
	t = std::max(1, std::min(t, 64));
	// a slice below a megabyte is not worth a thread
	return (int) std::max<uint64_t>(1, std::min<uint64_t>((uint64_t) t, file_bytes >> 20));
}

} // namespace

bool readFastq(const std::string &path, size_t min_len, ReadSet &out, int threads) {
	static const char alphabet[4] = {'A', 'C', 'G', 'T'};
	auto t_start = std::chrono::high_resolution_clock::now();
	out.clear();
	const int fd = open(path.c_str(), O_RDONLY);
	if (fd < 0)
		return false;
	struct stat st;
	if (fstat(fd, &st) != 0) {
		close(fd);
		return false;
	}
	const uint64_t size = (uint64_t) st.st_size;
	if (size == 0) {
		close(fd);
		return true;
	}
	// map the file; a stream that cannot be mapped is read into memory instead
	std::vector<char> fallback;
	const char *data = (const char *) mmap(NULL, size, PROT_READ, MAP_PRIVATE, fd, 0);
	const bool mapped = data != MAP_FAILED;
	if (mapped) {
		madvise((void *) data, size, MADV_WILLNEED);
	} else {
		fallback.resize(size);
		uint64_t got = 0;
		while (got < size) {
			ssize_t r = read(fd, fallback.data() + got, size - got);
			if (r <= 0)
				break;
			got += (uint64_t) r;
		}
		if (got != size) {
			close(fd);
			return false;
		}
		data = fallback.data();
	}
	close(fd);

	const int T = ioThreads(threads, size);
	out.threads = T;
	WorkerPool pool(T);
	std::vector<Slice> slices((size_t) T);
	std::vector<uint64_t> cut((size_t) T + 1);
	for (int t = 0; t <= T; t++)
		cut[(size_t) t] = size / (uint64_t) T * (uint64_t) t;
	cut[(size_t) T] = size;

	// pass 1: newlines per slice -> the line number every slice starts at
	pool.run([&](int t) { slices[(size_t) t].newlines = countNewlines(data + cut[(size_t) t], data + cut[(size_t) t + 1]); });
	std::vector<uint64_t> before((size_t) T + 1, 0);
	for (int t = 0; t < T; t++)
		before[(size_t) t + 1] = before[(size_t) t] + slices[(size_t) t].newlines;

	// pass 2: every slice walks the lines that START inside it; line 4k+1 is a read
	pool.run([&](int t) {
		Slice &s = slices[(size_t) t];
		const uint64_t lo = cut[(size_t) t], hi = cut[(size_t) t + 1];
		uint64_t p, li;
		if (lo == 0) {
			p = 0;
			li = 0;
		} else {
			const char *nl = (const char *) memchr(data + lo - 1, '\n', (size_t) (size - (lo - 1)));
			if (nl == NULL)
				return;
			p = (uint64_t) (nl - data) + 1;
			li = before[(size_t) t] + (p - 1 >= lo ? 1 : 0);
		}
		while (p < hi) {
			const char *nl = (const char *) memchr(data + p, '\n', (size_t) (size - p));
			const uint64_t end = nl ? (uint64_t) (nl - data) : size;
			if ((li & 3) == 1) {
				const uint64_t rl = end - p;
				if (rl >= min_len) {
					LineRef r = {data + p, rl};
					s.reads.push_back(r);
					s.packed_bytes += packedBytes((uint8_t) rl);
					s.total_length += rl;
				}
			}
			p = end + 1;
			li++;
		}
	});
	// A last record cut off after its header (query.cpp:380-381).  With a final newline the second
	// getline extracts nothing and leaves an EMPTY bases line; without one the first getline has
	// already hit end-of-file, the second does not touch its string, and the HEADER TEXT is what
	// the reference stores as the read.
	const uint64_t n_lines = before[(size_t) T] + (data[size - 1] != '\n' ? 1 : 0);
	if ((n_lines & 3) == 1) {
		uint64_t start = size;
		if (data[size - 1] != '\n') {
			start = size - 1;
			while (start > 0 && data[start - 1] != '\n')
				start--;
		}
		const uint64_t rl = size - start;
		if (rl >= min_len) {
			Slice &s = slices[(size_t) T - 1];
			LineRef r = {data + start, rl};
			s.reads.push_back(r);
			s.packed_bytes += packedBytes((uint8_t) rl);
			s.total_length += rl;
		}
	}

	std::vector<uint64_t> read_base((size_t) T + 1, 0), byte_base((size_t) T + 1, 0);
	for (int t = 0; t < T; t++) {
		read_base[(size_t) t + 1] = read_base[(size_t) t] + slices[(size_t) t].reads.size();
		byte_base[(size_t) t + 1] = byte_base[(size_t) t] + slices[(size_t) t].packed_bytes;
		out.total_length += slices[(size_t) t].total_length;
	}
	const uint64_t n = read_base[(size_t) T], total = byte_base[(size_t) T];
	void *pin = NULL;
	const size_t cap = (size_t) total + 64;
	if (cq_host_alloc(cap, &pin) == 0) {
		out.bases = (uint8_t *) pin;
		out.pinned = true;
	} else {
		out.bases = (uint8_t *) malloc(cap);
		out.pinned = false;
	}
	out.cap_bases = cap;
	out.offsets.resize((size_t) n);
	out.lengths.resize((size_t) n);
	out.n_bases = total;

	// the reference draws rand() once per accepted read, in file order: produce that sequence on
	// a thread of its own while the workers pack
	std::vector<uint8_t> subs((size_t) n);
	std::thread draw([&]() {
		for (uint64_t i = 0; i < n; i++)
			subs[(size_t) i] = (uint8_t) (rand() & 3);
	});
	pool.run([&](int t) {
		Slice &s = slices[(size_t) t];
		uint64_t at = byte_base[(size_t) t], idx = read_base[(size_t) t];
		for (size_t k = 0; k < s.reads.size(); k++, idx++) {
			const uint32_t len = (uint8_t) s.reads[k].rl;
			const bool ok = packRead((const uint8_t *) s.reads[k].b, len, out.bases + at);
			out.offsets[(size_t) idx] = at;
			out.lengths[(size_t) idx] = ok ? (uint8_t) len : 0;
			if (!ok)
				s.with_n.push_back(idx);
			at += packedBytes(len);
		}
	});
	draw.join();
	// reads the packer rejected: substitute N as the reference does and try again (rare)
	uint8_t tmp[256];
	for (int t = 0; t < T; t++) {
		Slice &s = slices[(size_t) t];
		for (size_t k = 0; k < s.with_n.size(); k++) {
			const uint64_t idx = s.with_n[k];
			const LineRef &r = s.reads[(size_t) (idx - read_base[(size_t) t])];
			const uint32_t len = (uint8_t) r.rl;
			bool had_n = false;
			for (uint32_t i = 0; i < len; i++) {
				had_n |= r.b[i] == 'N';
				tmp[i] = (uint8_t) (r.b[i] == 'N' ? alphabet[subs[(size_t) idx]] : r.b[i]);
			}
			out.n_with_n += had_n ? 1 : 0;
			const bool ok = packRead(tmp, len, out.bases + out.offsets[(size_t) idx]);
			out.lengths[(size_t) idx] = ok ? (uint8_t) len : 0;
		}
	}
	if (mapped)
		munmap((void *) data, size);
	out.parse_ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t_start).count();
	return true;
}

} // namespace cammiq
