// ASCII -> 2-bit read packer and its worker pool; see pack_reads.hpp for the layout.
#include "pack_reads.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>

#if defined(__x86_64__)
#include <immintrin.h>
#define CAMMIQ_X86 1
#endif

namespace cammiq {

namespace {

struct CodeTable {
	uint8_t v[256];
	CodeTable() {
		memset(v, 0xFF, sizeof(v));
		v['A'] = v['a'] = 0;
		v['C'] = v['c'] = 1;
		v['G'] = v['g'] = 2;
		v['T'] = v['t'] = 3;
	}
};
const CodeTable kCodes;

inline bool packScalar(const uint8_t *s, uint32_t len, uint8_t *dst) {
	uint32_t bad = 0, j = 0;
	for (; j + 4 <= len; j += 4) {
		const uint32_t c0 = kCodes.v[s[j]], c1 = kCodes.v[s[j + 1]], c2 = kCodes.v[s[j + 2]], c3 = kCodes.v[s[j + 3]];
		bad |= (c0 | c1 | c2 | c3) & 0x80u;
		dst[j >> 2] = (uint8_t) ((c0 << 6) | ((c1 & 3u) << 4) | ((c2 & 3u) << 2) | (c3 & 3u));
	}
	if (j < len) {
		uint32_t b = 0;
		for (uint32_t u = 0; j + u < len; u++) {
			const uint32_t c = kCodes.v[s[j + u]];
			bad |= c & 0x80u;
			b |= (c & 3u) << (6 - 2 * u);
		}
		dst[j >> 2] = (uint8_t) b;
	}
	return bad == 0;
}

#ifdef CAMMIQ_X86
// 64 bases per step.  Masked loads keep the tail inside the read; zeroed lanes decode to code 0
// and therefore leave the padding bits of the last byte clear.
__attribute__((target("avx512f,avx512bw,avx512vl"))) inline bool packAvx512(const uint8_t *s, uint32_t len, uint8_t *dst) {
	const __m512i three = _mm512_set1_epi8(3), one = _mm512_set1_epi8(1), fold = _mm512_set1_epi8((char) 0xDF);
	const __m512i letters = _mm512_broadcast_i32x4(_mm_setr_epi8('A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0));
	const __m512i pair = _mm512_set1_epi16(0x0104), quad = _mm512_set1_epi32(0x00010010);
	__mmask64 bad = 0;
	for (uint32_t j = 0; j < len; j += 64) {
		const uint32_t nb = std::min<uint32_t>(64u, len - j);
		const __mmask64 live = nb == 64 ? ~(__mmask64) 0 : (((__mmask64) 1 << nb) - 1);
		const __m512i w = _mm512_maskz_loadu_epi8(live, s + j);
		const __m512i t = _mm512_and_si512(_mm512_srli_epi16(w, 1), three);
		const __m512i code = _mm512_xor_si512(t, _mm512_and_si512(_mm512_srli_epi16(t, 1), one));
		bad |= _mm512_mask_cmpneq_epi8_mask(live, _mm512_and_si512(w, fold), _mm512_shuffle_epi8(letters, code));
		// c0*4+c1 per 16-bit lane, then (..)*16+(..) per 32-bit lane: one packed byte per dword
		const __m512i p32 = _mm512_madd_epi16(_mm512_maddubs_epi16(code, pair), quad);
		const __m128i bytes = _mm512_cvtepi32_epi8(p32);
		const uint32_t nout = (nb + 3) >> 2;
		_mm_mask_storeu_epi8(dst + (j >> 2), (__mmask16) ((1u << nout) - 1u), bytes);
	}
	return bad == 0;
}

// 32 bases per step; the tail goes through a zeroed bounce buffer.
__attribute__((target("avx2"))) inline bool packAvx2(const uint8_t *s, uint32_t len, uint8_t *dst) {
	const __m256i three = _mm256_set1_epi8(3), one = _mm256_set1_epi8(1), fold = _mm256_set1_epi8((char) 0xDF);
	const __m256i letters = _mm256_broadcastsi128_si256(_mm_setr_epi8('A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0));
	const __m256i pair = _mm256_set1_epi16(0x0104), quad = _mm256_set1_epi32(0x00010010);
	const __m256i gather = _mm256_broadcastsi128_si256(_mm_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1));
	uint32_t bad = 0;
	for (uint32_t j = 0; j < len; j += 32) {
		const uint32_t nb = std::min<uint32_t>(32u, len - j);
		__m256i w;
		if (nb == 32) {
			w = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + j));
		} else {
			uint8_t tmp[32];
			memset(tmp, 'A', sizeof(tmp)); // 'A' = code 0: valid and leaves the padding bits clear
			memcpy(tmp, s + j, nb);
			w = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(tmp));
		}
		const __m256i t = _mm256_and_si256(_mm256_srli_epi16(w, 1), three);
		const __m256i code = _mm256_xor_si256(t, _mm256_and_si256(_mm256_srli_epi16(t, 1), one));
		const __m256i ok = _mm256_cmpeq_epi8(_mm256_and_si256(w, fold), _mm256_shuffle_epi8(letters, code));
		bad |= ~(uint32_t) _mm256_movemask_epi8(ok);
		const __m256i p32 = _mm256_madd_epi16(_mm256_maddubs_epi16(code, pair), quad);
		const __m256i g = _mm256_shuffle_epi8(p32, gather);
		uint8_t out8[8];
		const uint32_t lo = (uint32_t) _mm256_extract_epi32(g, 0), hi = (uint32_t) _mm256_extract_epi32(g, 4);
		memcpy(out8, &lo, 4);
		memcpy(out8 + 4, &hi, 4);
		memcpy(dst + (j >> 2), out8, (nb + 3) >> 2);
	}
	return bad == 0;
}
#endif

typedef bool (*PackFn)(const uint8_t *, uint32_t, uint8_t *);

// One worker's share of a batch: reads [a, b) of the batch that starts at caller index `first`.
struct SliceArgs {
	const AsciiReads *in;
	uint64_t first, a, b;
	bool dense;
	uint64_t stride, at; // fixed stride, or the slice's first byte when dense
	uint8_t *out;
	uint32_t *out_offsets;
	uint8_t *out_lengths;
};
typedef uint64_t (*SliceFn)(const SliceArgs &);

// A core streams ~10 GB/s on its own; fetching a few reads ahead keeps more lines in flight.
#ifdef CAMMIQ_X86
#define CAMMIQ_PREFETCH(p) _mm_prefetch(reinterpret_cast<const char *>(p), _MM_HINT_T0)
#else
#define CAMMIQ_PREFETCH(p) __builtin_prefetch(p)
#endif
static size_t prefetchAhead() {
	const char *e = getenv("CAMMIQ_PACK_PREFETCH");
	return e ? (size_t) atol(e) : 4096;
}
static const size_t kPrefetchAhead = prefetchAhead();
// CAMMIQ_PACK_STREAM=0 (experiments): back-to-back reads of one length go through the per-read loop too
static const bool kStreamPath = !(getenv("CAMMIQ_PACK_STREAM") && atoi(getenv("CAMMIQ_PACK_STREAM")) == 0);

// the loop is stamped out per ISA so that the packer inlines into it
#define CAMMIQ_SLICE_LOOP(NAME, TARGET, PACK)                                                     \
	TARGET uint64_t NAME(const SliceArgs &x) {                                                    \
		uint64_t bad = 0, at = x.at;                                                              \
		for (uint64_t k = x.a; k < x.b; k++) {                                                    \
			const uint64_t i = x.first + k;                                                       \
			const uint32_t len = x.in->lengths[i];                                                \
			const uint8_t *src = x.in->bases + (x.in->offsets ? x.in->offsets[i] : i * x.in->stride); \
			uint8_t *dst = x.out + (x.dense ? at : k * x.stride);                                 \
			CAMMIQ_PREFETCH(src + kPrefetchAhead);                                                \
			CAMMIQ_PREFETCH(src + kPrefetchAhead + 64);                                           \
			const bool ok = PACK(src, len, dst);                                                  \
			x.out_lengths[k] = ok ? (uint8_t) len : 0;                                            \
			bad += ok ? 0 : 1;                                                                    \
			if (x.dense) {                                                                        \
				x.out_offsets[k] = (uint32_t) at;                                                 \
				at += packedBytes(len);                                                           \
			}                                                                                     \
		}                                                                                         \
		return bad;                                                                               \
	}
CAMMIQ_SLICE_LOOP(sliceScalar, , packScalar)
#ifdef CAMMIQ_X86
// Reads of one length L (a multiple of 4) that lie back to back (stride == L, packed stride == L/4) are
// ONE stream of bases: 16 reads = L/4 whole 64-byte blocks, converted without a mask, a tail or a per-read
// loop; the bytes land exactly where the per-read path would put them.  Returns false when a byte outside
// ACGTacgt was met (the caller then repeats these 16 reads one by one to find the invalid ones).
template <bool NT>
__attribute__((target("avx512f,avx512bw,avx512vl"))) inline bool packStream16(const uint8_t *s, uint32_t n_bytes, uint8_t *dst,
		size_t prefetch_ahead) {
	const __m512i three = _mm512_set1_epi8(3), fold = _mm512_set1_epi8((char) 0xDF);
	const __m512i letters = _mm512_broadcast_i32x4(_mm_setr_epi8('A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0));
	const __m512i pair = _mm512_set1_epi16(0x0104), quad = _mm512_set1_epi32(0x00010010);
	__mmask64 bad = 0;
	for (uint32_t j = 0; j < n_bytes; j += 64) {
		CAMMIQ_PREFETCH(s + j + prefetch_ahead);
		const __m512i w = _mm512_loadu_si512(reinterpret_cast<const void *>(s + j));
		// code = bits 1 and 2 of the letter, Gray-decoded: ((w >> 1) ^ (w >> 2)) & 3 (the 16-bit shifts leak
		// a neighbour's bits only into bits 6-7, which the mask drops)
		const __m512i code = _mm512_ternarylogic_epi32(_mm512_srli_epi16(w, 1), _mm512_srli_epi16(w, 2), three, 0x28);
		bad |= _mm512_cmpneq_epi8_mask(_mm512_and_si512(w, fold), _mm512_shuffle_epi8(letters, code));
		const __m512i p32 = _mm512_madd_epi16(_mm512_maddubs_epi16(code, pair), quad);
		// the packed bytes are read next by the DMA engine, not by this core: with a 16-byte aligned
		// destination they bypass the cache (no read-for-ownership of the output lines)
		if (NT)
			_mm_stream_si128(reinterpret_cast<__m128i *>(dst + (j >> 2)), _mm512_cvtepi32_epi8(p32));
		else
			_mm_storeu_si128(reinterpret_cast<__m128i *>(dst + (j >> 2)), _mm512_cvtepi32_epi8(p32));
	}
	return bad == 0;
}

__attribute__((target("avx512f,avx512bw,avx512vl"))) uint64_t sliceAvx512(const SliceArgs &x) {
	uint64_t bad = 0, at = x.at;
	const uint64_t L = x.in->stride;
	const bool stream = kStreamPath && !x.dense && x.in->offsets == NULL && L >= 4 && L <= 255 && (L & 3) == 0 && x.stride == L / 4;
	const __m128i want = _mm_set1_epi8((char) L);
	bool streamed = false;
	for (uint64_t k = x.a; k < x.b;) {
		uint64_t upto = k + 1;
		if (stream && k + 16 <= x.b &&
			_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128(reinterpret_cast<const __m128i *>(x.in->lengths + x.first + k)), want)) == 0xFFFF) {
			const uint8_t *src = x.in->bases + (x.first + k) * L;
			uint8_t *dst = x.out + k * x.stride;
			const bool ok = ((uintptr_t) dst & 15) == 0 ? packStream16<true>(src, (uint32_t) (16 * L), dst, kPrefetchAhead)
													   : packStream16<false>(src, (uint32_t) (16 * L), dst, kPrefetchAhead);
			if (ok) {
				_mm_storeu_si128(reinterpret_cast<__m128i *>(x.out_lengths + k), want);
				streamed = true;
				k += 16;
				continue;
			}
			upto = k + 16; // an invalid byte among these 16 reads: one by one
		}
		for (; k < upto; k++) {
			const uint64_t i = x.first + k;
			const uint32_t len = x.in->lengths[i];
			const uint8_t *src = x.in->bases + (x.in->offsets ? x.in->offsets[i] : i * x.in->stride);
			uint8_t *dst = x.out + (x.dense ? at : k * x.stride);
			CAMMIQ_PREFETCH(src + kPrefetchAhead);
			CAMMIQ_PREFETCH(src + kPrefetchAhead + 64);
			const bool ok = packAvx512(src, len, dst);
			x.out_lengths[k] = ok ? (uint8_t) len : 0;
			bad += ok ? 0 : 1;
			if (x.dense) {
				x.out_offsets[k] = (uint32_t) at;
				at += packedBytes(len);
			}
		}
	}
	if (streamed)
		_mm_sfence(); // the streaming stores are visible before the worker reports its slice done
	return bad;
}
CAMMIQ_SLICE_LOOP(sliceAvx2, __attribute__((target("avx2"))), packAvx2)
#endif

struct Dispatch {
	PackFn fn;
	SliceFn slice;
	const char *name;
	Dispatch() : fn(packScalar), slice(sliceScalar), name("scalar") {
#ifdef CAMMIQ_X86
		__builtin_cpu_init();
		if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl")) {
			fn = packAvx512;
			slice = sliceAvx512;
			name = "avx512";
		} else if (__builtin_cpu_supports("avx2")) {
			fn = packAvx2;
			slice = sliceAvx2;
			name = "avx2";
		}
#endif
		const char *force = getenv("CAMMIQ_PACK_ISA"); // tests: exercise the slower paths too
		if (force) {
			if (!strcmp(force, "scalar")) { fn = packScalar; slice = sliceScalar; name = "scalar"; }
#ifdef CAMMIQ_X86
			else if (!strcmp(force, "avx2") && __builtin_cpu_supports("avx2")) { fn = packAvx2; slice = sliceAvx2; name = "avx2"; }
#endif
		}
	}
};

const Dispatch &dispatch() {
	static const Dispatch d;
	return d;
}

} // namespace

bool packRead(const uint8_t *ascii, uint32_t len, uint8_t *dst) { return dispatch().fn(ascii, len, dst); }

const char *packIsaName() { return dispatch().name; }

void unpackRead(const uint8_t *packed, uint32_t len, uint8_t *ascii) {
	for (uint32_t j = 0; j < len; j++)
		ascii[j] = (uint8_t) "ACGT"[(packed[j >> 2] >> (6 - 2 * (j & 3))) & 3];
}

// ------------------------------------------------------------------------------ worker pool

WorkerPool::WorkerPool(int n_threads) : n_(std::max(n_threads, 1)), job_(NULL), generation_(0), pending_(0), stop_(false) {
	for (int i = 1; i < n_; i++)
		threads_.push_back(std::thread(&WorkerPool::loop, this, i));
}

WorkerPool::~WorkerPool() {
	{
		std::lock_guard<std::mutex> lk(mu_);
		stop_ = true;
	}
	cv_start_.notify_all();
	for (size_t i = 0; i < threads_.size(); i++)
		threads_[i].join();
}

void WorkerPool::loop(int id) {
	uint64_t seen = 0;
	for (;;) {
		const std::function<void(int)> *job;
		{
			std::unique_lock<std::mutex> lk(mu_);
			while (!stop_ && generation_ == seen)
				cv_start_.wait(lk);
			if (stop_)
				return;
			seen = generation_;
			job = job_;
		}
		(*job)(id);
		{
			std::lock_guard<std::mutex> lk(mu_);
			if (--pending_ == 0)
				cv_done_.notify_one();
		}
	}
}

void WorkerPool::run(const std::function<void(int)> &fn) {
	if (n_ == 1) {
		fn(0);
		return;
	}
	{
		std::lock_guard<std::mutex> lk(mu_);
		job_ = &fn;
		pending_ = n_ - 1;
		generation_++;
	}
	cv_start_.notify_all();
	fn(0);
	std::unique_lock<std::mutex> lk(mu_);
	while (pending_ != 0)
		cv_done_.wait(lk);
}

// ------------------------------------------------------------------------------ batches

namespace {
// slices are multiples of 64 reads so that two workers never share a cache line of the outputs
inline void sliceOf(uint64_t n, int T, int t, uint64_t &a, uint64_t &b) {
	const uint64_t per = ((n + T - 1) / T + 63) & ~63ull;
	a = std::min<uint64_t>(n, per * (uint64_t) t);
	b = std::min<uint64_t>(n, a + per);
}
} // namespace

PackedLayout planBatch(WorkerPool &pool, const uint8_t *lengths, uint64_t first, uint64_t n, bool dense) {
	const int T = pool.size();
	std::vector<uint64_t> bytes((size_t) T, 0);
	std::vector<uint32_t> longest((size_t) T, 0);
	pool.run([&](int t) {
		uint64_t a, b, sum = 0;
		uint32_t mx = 0;
		sliceOf(n, T, t, a, b);
		for (uint64_t k = a; k < b; k++) {
			const uint32_t len = lengths[first + k];
			sum += packedBytes(len);
			mx = std::max(mx, len);
		}
		bytes[(size_t) t] = sum;
		longest[(size_t) t] = mx;
	});
	PackedLayout L;
	L.dense = dense;
	L.max_len = 0;
	L.slice_start.assign((size_t) T, 0);
	uint64_t at = 0;
	for (int t = 0; t < T; t++) {
		L.slice_start[(size_t) t] = at;
		at += bytes[(size_t) t];
		L.max_len = std::max(L.max_len, longest[(size_t) t]);
	}
	L.stride = packedBytes(L.max_len);
	L.total_bytes = dense ? at : n * L.stride;
	return L;
}

uint64_t packBatch(WorkerPool &pool, const AsciiReads &in, uint64_t first, uint64_t n, const PackedLayout &layout,
		uint8_t *out, uint32_t *out_offsets, uint8_t *out_lengths) {
	const int T = pool.size();
	std::vector<uint64_t> invalid((size_t) T, 0);
	const SliceFn slice = dispatch().slice;
	pool.run([&](int t) {
		SliceArgs x;
		x.in = &in;
		x.first = first;
		sliceOf(n, T, t, x.a, x.b);
		x.dense = layout.dense;
		x.stride = layout.stride;
		x.at = layout.dense ? layout.slice_start[(size_t) t] : 0;
		x.out = out;
		x.out_offsets = out_offsets;
		x.out_lengths = out_lengths;
		invalid[(size_t) t] = slice(x);
	});
	uint64_t total = 0;
	for (int t = 0; t < T; t++)
		total += invalid[(size_t) t];
	return total;
}

} // namespace cammiq
