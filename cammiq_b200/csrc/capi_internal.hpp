// Internals shared by the translation units of libcammiq_gpu.so (capi.cu, multi_gpu.cpp).
// Not part of the ABI: include/cammiq_gpu.h is.
#ifndef CAMMIQ_CAPI_INTERNAL_HPP
#define CAMMIQ_CAPI_INTERNAL_HPP

#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/cammiq_gpu.h"
#include "flat_index.hpp"
#include "pack_reads.hpp"

namespace cammiq {
struct PairSlot; // scan_kernels.cuh
}
using cammiq::FlatIndex;
using cammiq::PairSlot;
using cammiq::TableBucket;
using cammiq::WorkerPool;

// records `msg` as the calling thread's cq_last_error() and returns `code`
int cqFail(int code, const std::string &msg);

#define CQ_CUDA(call)                                                                      \
	do {                                                                                   \
		cudaError_t e_ = (call);                                                           \
		if (e_ != cudaSuccess)                                                             \
			return cqFail(CQ_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));   \
	} while (0)

struct cq_index {
	FlatIndex flat;
};

struct cq_ctx {
	int device = 0;
	int n_sms = 0;
	cudaStream_t stream = NULL;
	bool own_stream = false;
	cudaEvent_t ev[2] = {NULL, NULL}; // H2D bracket of the last cq_reads_stage
	struct StepEvents { cudaEvent_t e[4]; }; // pack start | scan start | scan end | reduce end
	std::vector<StepEvents> steps;
	size_t steps_used = 0;
	// resident index
	bool has_index = false;
	uint32_t h = 0, n_genomes = 0;
	uint64_t n_leaves_u = 0, n_leaves_d = 0, table_mask = 0;
	uint32_t table_shift = 26;
	TableBucket *d_table = NULL;
	uint32_t *d_nodes_u = NULL, *d_nodes_d = NULL, *d_leaf_u_ref = NULL;
	uint2 *d_leaf_d_ref = NULL;
	// accumulators
	unsigned long long *d_counts = NULL; // 2*(G+1)+4
	uint32_t *d_rcount_u = NULL, *d_rcount_d = NULL;
	// the other accumulator set of cq_swap_accumulators (NULL until its first call)
	unsigned long long *d_counts_alt = NULL;
	uint32_t *d_rcount_u_alt = NULL, *d_rcount_d_alt = NULL;
	uint32_t *d_partials = NULL;
	uint32_t *d_spill = NULL; // per-read hit overflow, [max grid warps][32][spill_stride]; grows with the longest read seen
	uint32_t *d_dedup = NULL; // hash-set scratch of the cooperative leaf dedup, [max grid warps][dedup_slots]
	size_t cap_spill = 0, cap_dedup = 0;
	uint2 *d_filter = NULL;
	uint32_t filter_words = 0, filter_sel = 0x77777777u;
	bool filter_sieve = false;
	int max_grid = 0;
	unsigned long long *d_probe_count = NULL;
	int grid = 0;
	bool smem_counters = true;
	size_t smem_bytes = 0;
	// staged reads
	uint8_t *d_bases = NULL;
	uint64_t *d_offsets = NULL;
	uint8_t *d_lengths = NULL;
	size_t cap_bases = 0, cap_reads_off = 0, cap_reads_len = 0;
	uint64_t staged_reads = 0, staged_stride = 0, staged_bytes = 0;
	bool staged_has_offsets = false;
	uint32_t staged_max_len = 0;
	uint64_t staged_shift = 0; // offset of d_bases[0] in the caller's base buffer
	bool staged_packed = false;
	size_t last_dyn_smem[6] = {(size_t) -1, (size_t) -1, (size_t) -1, (size_t) -1, (size_t) -1, (size_t) -1}; // per kernel variant
	int last_per_sm[6] = {0, 0, 0, 0, 0, 0};
	int last_carveout[6] = {0, 0, 0, 0, 0, 0};
	// host->device pipeline of cq_query: kStages chunk buffers rotate through copy and scan
	static const int kStages = 3;
	cudaStream_t copy_stream = NULL;
	cudaEvent_t ev_copied[kStages] = {NULL, NULL, NULL}, ev_free[kStages] = {NULL, NULL, NULL};
	uint8_t *d_cbases[kStages] = {NULL, NULL, NULL};
	uint64_t *d_coffsets[kStages] = {NULL, NULL, NULL};
	uint32_t *d_coffsets32[kStages] = {NULL, NULL, NULL};
	uint8_t *d_clengths[kStages] = {NULL, NULL, NULL};
	size_t cap_cbases[kStages] = {0, 0, 0}, cap_coffsets[kStages] = {0, 0, 0}, cap_coffsets32[kStages] = {0, 0, 0},
		cap_clengths[kStages] = {0, 0, 0};
	// reads in the scan's tile layout (pack_tiles_kernel): one buffer per pipeline stage + one
	// for the staged reads (index kStages); the length copies hold 0 for reads found invalid
	uint32_t *d_words[kStages + 1] = {NULL, NULL, NULL, NULL};
	uint8_t *d_len2[kStages + 1] = {NULL, NULL, NULL, NULL};
	size_t cap_words[kStages + 1] = {0, 0, 0, 0}, cap_len2[kStages + 1] = {0, 0, 0, 0};
	// host packing (cq_ctx_set_host_packing): worker pool + pinned staging of the packed chunks
	int pack_threads = 0;
	WorkerPool *pool = NULL;
	uint8_t *h_pbases[kStages] = {NULL, NULL, NULL};
	uint8_t *h_plengths[kStages] = {NULL, NULL, NULL};
	uint32_t *h_poffsets[kStages] = {NULL, NULL, NULL};
	size_t cap_h_pbases[kStages] = {0, 0, 0}, cap_h_plengths[kStages] = {0, 0, 0}, cap_h_poffsets[kStages] = {0, 0, 0};
	// SC pair records (device, grows)
	unsigned long long *d_pairs = NULL;
	size_t cap_pairs = 0;
	uint64_t sc_reads_since_reset = 0; // upper bound of the pair records held
	PairSlot *d_pair_table = NULL, *d_pair_out = NULL; // aggregation scratch of cq_fetch
	size_t cap_pair_table = 0, cap_pair_out = 0;
	// per-read outputs (device, sized per call)
	uint8_t *d_read_class = NULL;
	uint32_t *d_read_rid_a = NULL, *d_read_rid_b = NULL, *d_nleaf_u = NULL, *d_nleaf_d = NULL,
		*d_leaf_u = NULL, *d_leaf_d = NULL;
	size_t cap_read_class = 0, cap_read_rid_a = 0, cap_read_rid_b = 0, cap_nleaf_u = 0, cap_nleaf_d = 0,
		cap_leaf_u = 0, cap_leaf_d = 0;
	uint32_t leaf_cap = 0;
	bool want_per_read = false, want_sets = false;
	cq_timing timing;
};


// One pass of the hot path over host reads, results left on the device (cq_query = this + cqFetch).
// per_read != NULL: optional per-read outputs of this call go to its buffers.
int cqSubmitHost(cq_ctx *c, int mode, bool packed, const uint8_t *bases, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t n_reads, cq_result *per_read, const char *who);
// Totals to host buffers.  counts / rcount_u / rcount_d: device arrays on c's device holding the
// totals to report (the context's own accumulators, or sums over several devices).
int cqFetchFrom(cq_ctx *c, int mode, const unsigned long long *d_counts, const uint32_t *d_rcount_u, const uint32_t *d_rcount_d,
		cq_result *out, bool with_pairs);
// the (pair, count) entries of this context's pair records, unsorted
int cqCollectPairs(cq_ctx *c, std::vector<cq_pair_count> &out);

#endif
