// Host side of the read pipeline: ASCII reads -> 2-bit codes
// before they cross PCIe, four bases per byte, validated on the way.
//
// Packed layout (what scan_reads_kernel<.., PACKED=true> consumes): base j of a read sits in
// byte j/4 at bits 7-2*(j%4) .. 6-2*(j%4), i.e. the first base is most significant and the
// read is a big-endian bit stream; a read of `len` bases owns ceil(len/4) bytes and the unused
// low bits of its last byte are zero.  Codes follow the reference's hash alphabet
// (query.cpp:1860-1883): A/a=0 C/c=1 G/g=2 T/t=3.  A read holding any other byte is invalid:
// the reference's scan would abort on it, the ASCII kernel counts it as unclassified, and the
// packer reports it so the caller can hand the kernel a zero length for it (same outcome).
#ifndef CAMMIQ_PACK_READS_HPP
#define CAMMIQ_PACK_READS_HPP

#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace cammiq {

inline size_t packedBytes(uint32_t len) { return (len + 3u) >> 2; }

// Packs one read; returns false when a byte outside ACGTacgt was met (the output is then
// unspecified but stays within ceil(len/4) bytes).  AVX-512 / AVX2 / scalar, chosen at run time.
bool packRead(const uint8_t *ascii, uint32_t len, uint8_t *dst);

// Inverse (tests, diagnostics): 2-bit codes -> "ACGT".
void unpackRead(const uint8_t *packed, uint32_t len, uint8_t *ascii);

// "avx512" | "avx2" | "scalar": what packRead dispatches to on this machine.
const char *packIsaName();

// A fixed set of worker threads that run the same callable on disjoint slices; the calling
// thread takes slice 0, so WorkerPool(1) spawns nothing.
class WorkerPool {
public:
	explicit WorkerPool(int n_threads);
	~WorkerPool();
	int size() const { return n_; }
	// runs fn(slice) for slice in [0, size()) and returns when all are done
	void run(const std::function<void(int)> &fn);

private:
	void loop(int id);
	int n_;
	std::vector<std::thread> threads_;
	std::mutex mu_;
	std::condition_variable cv_start_, cv_done_;
	const std::function<void(int)> *job_;
	uint64_t generation_;
	int pending_;
	bool stop_;
};

// One batch of ASCII reads as the C ABI describes them (offsets == NULL: fixed stride).
struct AsciiReads {
	const uint8_t *bases;
	const uint64_t *offsets;
	uint64_t stride;
	const uint8_t *lengths;
};

// Where the packed reads of one batch go: back to back (`dense`, per-read 32-bit offsets) or at
// a fixed stride of ceil(max_len/4) bytes.
struct PackedLayout {
	bool dense;
	uint64_t stride;      // !dense
	uint64_t total_bytes; // bytes the batch occupies
	uint32_t max_len;     // longest read of the batch (bases)
	std::vector<uint64_t> slice_start; // dense: first byte of every worker's slice
};

// Pass 1 over the lengths of reads [first, first+n): sizes the batch.
PackedLayout planBatch(WorkerPool &pool, const uint8_t *lengths, uint64_t first, uint64_t n, bool dense);

// Pass 2: packs reads [first, first+n) of `in` into `out` as `layout` says.
// out_lengths[k] = lengths[first+k], or 0 for an invalid read; out_offsets[k] (dense only) =
// the read's first byte in `out`.  Returns the number of invalid reads.
uint64_t packBatch(WorkerPool &pool, const AsciiReads &in, uint64_t first, uint64_t n, const PackedLayout &layout,
		uint8_t *out, uint32_t *out_offsets, uint8_t *out_lengths);

} // namespace cammiq

#endif
