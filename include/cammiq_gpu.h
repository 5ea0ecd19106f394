/*
 * cammiq_gpu.h -- C ABI of libcammiq_gpu.so: the B200 (sm_100a) implementation of CAMMiQ's
 * query-time read-matching path.  Plain C types only; every buffer in a signature is a
 * caller-owned HOST buffer unless its name starts with d_ (device pointer).
 *
 * The reference has no FFI layer; the seam this ABI replaces is the body of three member
 * functions of class FqReader plus the index load that feeds them (citations are
 * /root/reference/src/<file>:<lines>):
 *
 *   cq_index_load / cq_index_free     Hash::loadIdx64_p + decodeTrie_p       hashtrie.cpp:425-507
 *                                     FqReader::loadIdx_p (2 loader threads)  query.cpp:109-123
 *   cq_index_leaves / cq_index_map_sp pleafNode fields, Hash::map_sp          hashtrie.hpp:37-47,60
 *   cq_ctx_create / cq_ctx_destroy    FqReader ctor / dtor                    query.cpp:34-107
 *   cq_index_upload                   end of loadIdx_p (index becomes resident)
 *   cq_query                          FqReader::query64_p / query64mt_p (CQ_MODE_P)
 *                                     FqReader::query64_sc (CQ_MODE_SC)       query.cpp:458-1080
 *   cq_query_packed / cq_pack_reads   the read storage readFastq fills        query.cpp:371-425
 *   cq_ctx_set_host_packing           (ASCII `reads` vector -> 2-bit codes before PCIe)
 *   cq_reset                          resetCounters / resetCounters_sc        query.cpp:1820-1858
 *   cq_get_timing                     the "Time for query" bracket            query.cpp:645-647
 *   cq_multi_*                        the same calls over the GPUs of one box: reads sharded,
 *                                     index replicated, ONE NCCL sum-reduce of the counters
 *                                     (the reference's own data
 *                                     parallelism is the OpenMP loop of query64mt_p, query.cpp:664)
 *   cq_ilp_inputs                     ILP set-up coefficients over the leaf arrays
 *                                     (runILP_*: query.cpp:1154-1181, 1508-1535)
 *
 * All functions return 0 on success and a negative CQ_E* code on failure;
 * cq_last_error() then holds a message for the calling thread.  No exceptions cross the
 * ABI.  There is NO CPU fallback: every cq_ctx_* / cq_query* call fails with CQ_ENODEV
 * when no CUDA device is usable.
 */
#ifndef CAMMIQ_GPU_H
#define CAMMIQ_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CQ_ABI_VERSION 3

enum {
	CQ_OK = 0,
	CQ_EINVAL = -1,  /* bad argument */
	CQ_EIO = -2,     /* cannot open / read an index file */
	CQ_EFORMAT = -3, /* index file violates the format */
	CQ_ENOMEM = -4,
	CQ_ENODEV = -5,  /* no usable CUDA device / library built without one */
	CQ_ECUDA = -6,   /* CUDA runtime error, text in cq_last_error() */
	CQ_ESTATE = -7   /* call order violated (e.g. query before upload) */
};

/* Which reference function cq_query stands in for. */
enum {
	CQ_MODE_P = 0, /* query64_p / query64mt_p: per-genome counts + per-leaf rcount */
	CQ_MODE_SC = 1 /* query64_sc: per-genome counts + pair map read_cnts_b, no rcount */
};

enum { CQ_TABLE_U = 0, CQ_TABLE_D = 1 };

/* Per-read decision (rows of the table in SURVEY.md section 8a; query.cpp:542-636, 977-1067). */
enum {
	CQ_CLASS_UNLABELED = 0, /* nundet++ */
	CQ_CLASS_CONFLICT = 1,  /* nconf++ */
	CQ_CLASS_U = 2,         /* cnt_u[rid_a]++ */
	CQ_CLASS_D_PAIR = 3,    /* cnt_d[rid_a]++, cnt_d[rid_b]++ (SC: pairs[(rid_a,rid_b)]++) */
	CQ_CLASS_UD = 4,        /* cnt_u[rid_a]++, cnt_d[rid_a]++ */
	CQ_CLASS_D_INTER = 5    /* P: cnt_d[rid_a]++ ; SC: cnt_u[rid_a]++ and cnt_d[rid_a]++ */
};

typedef struct cq_index cq_index; /* host-side decoded + flattened index pair (U and D) */
typedef struct cq_ctx cq_ctx;     /* one CUDA device: resident index, counters, streams */

const char *cq_last_error(void);
int cq_abi_version(void);

/* ---------------------------------------------------------------- index (host side) -- */

typedef struct {
	uint32_t hash_len;       /* h, equal for both tables (assert at query.cpp:460) */
	uint64_t n_leaves_u;     /* Hash::leaf_cnt of index_u.bin1 */
	uint64_t n_leaves_d;     /* Hash::leaf_cnt of index_d.bin2 */
	uint64_t n_buckets_u;    /* distinct h-mer buckets per table */
	uint64_t n_buckets_d;
	uint64_t n_keys;         /* distinct h-mers over both tables = occupied table slots */
	uint64_t n_table_buckets;/* 32-byte buckets in the merged prefix table (power of two) */
	uint64_t n_nodes_u;      /* trie nodes of the index file below non-leaf bucket roots */
	uint64_t n_nodes_d;
	uint64_t n_cnodes_u;     /* device trie nodes after path compression (runs of single-child */
	uint64_t n_cnodes_d;     /* nodes become one chain node of up to 32 bases) */
	uint32_t max_ref_id;     /* largest genome id stored in a leaf */
	uint64_t filter_bytes;   /* membership filter size, 0 = none */
	uint64_t device_bytes;   /* bytes cq_index_upload will place on the device */
	double decode_ms, flatten_ms;
} cq_index_info;

/*
 * Decode <path_u>(+.aux) and <path_d>(+.aux) on two host threads and flatten them into the
 * device layout.  load_factor in (0,1]: fraction of table slots occupied (0 = default).
 * Fails with CQ_EFORMAT when the two hash lengths differ or a stream is malformed.
 */
int cq_index_load(const char *path_u, const char *path_d, double load_factor, cq_index **out);
void cq_index_free(cq_index *idx);
int cq_index_get_info(const cq_index *idx, cq_index_info *info);
/*
 * (Re)build the L2-resident membership filter that fronts the prefix table, using at most
 * max_bytes (whole KB; 0 removes the filter, so every position probes the table in HBM).
 * cq_index_load builds it with a 48 MB budget -- what stays L2-resident next to the scan's
 * other traffic on a B200.  With fewer than 8 bits per key the filter is no longer selective
 * and is used as a sieve instead (1 or 2 bits per key; positions that pass load their bucket's
 * keys in the same phase); below 1 bit per key it is dropped.
 * Takes effect at the next cq_index_upload.
 */
int cq_index_set_filter_budget(cq_index *idx, uint64_t max_bytes);

/* Leaf fields in FILE order (leaf id = order of appearance in the index file). */
typedef struct {
	uint64_t n;
	const uint32_t *ref_id1;
	const uint32_t *ref_id2; /* all zero for CQ_TABLE_U */
	const uint16_t *ucount1;
	const uint16_t *ucount2;
	const uint8_t *depth;    /* hash_len + trie depth, uint8 arithmetic as in the reference */
} cq_leaf_view;
int cq_index_leaves(const cq_index *idx, int table, cq_leaf_view *view);

/*
 * Hash::map_sp as CSR: for rid in 1..n_genomes the file-order leaf ids that carry rid, in
 * file order (D leaves appear under both ids).  offsets has n_genomes+2 entries; the ids
 * of rid are ids[offsets[rid] .. offsets[rid+1]).  ids may be NULL to size the array;
 * *total receives the entry count.
 */
int cq_index_map_sp(const cq_index *idx, int table, uint32_t n_genomes, uint64_t *offsets,
		uint64_t *ids, uint64_t *total);

/*
 * Layout verification accessor (host): one Hash::find64_p on the FLATTENED layout.
 * bucket = 2-bit hash of the h-mer, cand/len = the bases that follow it.  *leaf receives
 * the file-order leaf id or UINT64_MAX.  Used by the flattening tests; it is not a query
 * path and cq_query never calls it.
 */
int cq_index_find_host(const cq_index *idx, int table, uint64_t bucket, const uint8_t *cand,
		size_t len, uint64_t *leaf);

/* ------------------------------------------------------------------- device context -- */

/* device = CUDA ordinal.  stream = a cudaStream_t the caller wants the work enqueued on
   (e.g. torch's current stream), or NULL to let the context create its own.
   A context is not thread-safe: use it from one thread at a time (one context per GPU, each
   driven by its own thread or process, is the multi-GPU pattern).  A cq_index is immutable
   after cq_index_load / cq_index_set_filter_budget and may be uploaded to any number of
   contexts concurrently. */
int cq_ctx_create(int device, void *stream, cq_ctx **out);
void cq_ctx_destroy(cq_ctx *ctx);

/* Copy the flattened index to the device and size the counters for genome ids 1..n_genomes.
   Fails with CQ_EINVAL when a leaf carries an id outside 1..n_genomes. */
int cq_index_upload(cq_ctx *ctx, const cq_index *idx, uint32_t n_genomes);

typedef struct {
	uint32_t a, b; /* a <= b */
	uint64_t count;
} cq_pair_count;

typedef struct {
	/* in: capacities of the caller's buffers (0 / NULL = not wanted) */
	uint64_t *cnt_u;        /* [n_genomes+1], index 0 unused: Genome::read_cnts_u */
	uint64_t *cnt_d;        /* [n_genomes+1]: Genome::read_cnts_d */
	uint32_t *rcount_u;     /* [n_leaves_u] pleafNode::rcount, file order (CQ_MODE_P) */
	uint32_t *rcount_d;     /* [n_leaves_d] */
	cq_pair_count *pairs;   /* [pairs_cap] read_cnts_b sorted by (a,b) (CQ_MODE_SC) */
	uint64_t pairs_cap;
	/* optional per-read records, [n_reads] each */
	uint8_t *read_class;
	uint32_t *read_rid_a;
	uint32_t *read_rid_b;
	/* optional per-read distinct leaf sets: up to leaf_cap file-order ids per table per read,
	   ascending; read_nleaf_* hold the true set sizes */
	uint32_t leaf_cap;
	uint32_t *read_nleaf_u; /* [n_reads] */
	uint32_t *read_nleaf_d;
	uint32_t *read_leaf_u;  /* [n_reads * leaf_cap] */
	uint32_t *read_leaf_d;
	/* out */
	uint64_t nundet;        /* FqReader::nundet */
	uint64_t nconf;         /* FqReader::nconf */
	uint64_t n_invalid;     /* reads the reference cannot process (see cq_query) */
	uint64_t n_pairs;       /* distinct pairs (may exceed pairs_cap: then CQ_EINVAL) */
} cq_result;

/*
 * One pass of the hot path over n_reads reads held in HOST memory as ASCII, exactly the
 * state query64_* consumes: read i = bases[offsets[i] .. offsets[i]+lengths[i]).
 * offsets may be NULL for fixed-stride storage: read i starts at i*stride.
 * Synchronous.  Device counters ACCUMULATE across calls until cq_reset (the reference's
 * counters live until resetCounters); on return `out` holds the accumulated totals.
 * Per-read outputs cover this call's reads only.
 *
 * Defined behaviour where the reference has none: a read
 * shorter than hash_len (query.cpp:486 underflows) or holding a byte outside ACGTacgt
 * (symbolIdx = -1) is counted as unlabeled and in n_invalid.  'N' must already have been
 * substituted by the FASTQ reader, as in the reference (query.cpp:383).
 */
int cq_query(cq_ctx *ctx, int mode, const uint8_t *bases, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t n_reads, cq_result *out);

/* ------------------------------------------- 2-bit packed reads -- */
/*
 * The reference keeps reads as ASCII (`std::vector<uint8_t*> reads`, query.cpp:371-425) and so
 * does cq_query's interface; ASCII is 4x the bytes PCIe has to move.  With host packing on,
 * cq_query converts each chunk of reads to 2-bit codes with `threads` host threads (AVX-512 /
 * AVX2 / scalar), validates them there, and overlaps pack -> copy -> scan chunk by chunk.
 * Results are identical either way.  threads = 0 turns packing off (ASCII crosses PCIe, the
 * kernel decodes); threads < 0 restores the default: CAMMIQ_PACK_THREADS if set, else off when
 * LOCAL_WORLD_SIZE > 1 (several ranks share the host's memory bandwidth, which is what packing
 * spends), else min(16, hardware threads) when the host has at least 4, else off.
 */
int cq_ctx_set_host_packing(cq_ctx *ctx, int threads);

/*
 * Packed layout: base j of a read sits in byte j/4 at bits 7-2*(j%4)..6-2*(j%4) (A=0 C=1 G=2
 * T=3, hash alphabet of query.cpp:1860-1883; first base most significant), ceil(len/4) bytes
 * per read, unused low bits of the last byte zero.
 *
 * cq_pack_reads: host-only helper for callers that ingest reads themselves (a FASTQ parser
 * writing straight into pinned buffers).  Read i goes to packed + i*packed_stride
 * (packed_stride >= ceil(max length/4)); packed_lengths[i] = lengths[i], or 0 when the read
 * holds a byte outside ACGTacgt (cq_query's rule for such reads then applies);
 * *n_invalid (may be NULL) receives how many did.  Works without a GPU.
 */
int cq_pack_reads(const uint8_t *bases, const uint64_t *offsets, uint64_t stride, const uint8_t *lengths,
		uint64_t n_reads, int threads, uint8_t *packed, uint64_t packed_stride, uint8_t *packed_lengths,
		uint64_t *n_invalid);
/* "avx512", "avx2" or "scalar": the packer this host runs. */
const char *cq_pack_isa(void);

/*
 * cq_query over reads the caller already holds PACKED in host memory: read i =
 * ceil(lengths[i]/4) bytes at packed[offsets[i]] (offsets NULL: at i*stride, stride in
 * bytes).  lengths as produced by cq_pack_reads (0 = invalid read).  Otherwise as cq_query.
 */
int cq_query_packed(cq_ctx *ctx, int mode, const uint8_t *packed, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t n_reads, cq_result *out);

/* Zero the accumulated counters (cnt_u/d, rcount, nundet, nconf, n_invalid, pairs). */
int cq_reset(cq_ctx *ctx);

/* Page-locked host memory for read and result buffers: cq_query's host<->device copies run
   asynchronously (and overlap the scan) only from pinned memory. */
int cq_host_alloc(size_t bytes, void **out);
void cq_host_free(void *p);

/* --------------------------------------------------- several GPUs of one box (8b / 8e) -- */
/*
 * cq_multi is cq_ctx over n devices: the index is replicated (cq_multi_upload), cq_multi_query
 * cuts the reads into n contiguous shards of (nearly) equal base counts, one host thread per device
 * runs the ordinary pipeline on its shard, and ONE grouped NCCL sum-reduce brings the counter
 * block (and, in CQ_MODE_P, the per-leaf rcount arrays) to device 0 and from there to `out`.
 * On return `out` holds the totals over all devices, exactly as cq_query would have produced
 * them on one (integer sums: independent of n).  Counters accumulate across calls until
 * cq_multi_reset, like cq_query's.  Per-read outputs are written by each device into its slice of
 * the caller's buffers.  devices = NULL selects ordinals 0..n_gpus-1.
 * NCCL is loaded at run time (libnccl.so.2; a copy the process already holds, e.g. torch's, is
 * reused); with n_gpus = 1 no NCCL is needed.  Not thread-safe; one cq_multi per process is the
 * intended use (it owns one host thread per device while a query runs).
 */
typedef struct cq_multi cq_multi;
int cq_multi_create(int n_gpus, const int *devices, cq_multi **out);
void cq_multi_destroy(cq_multi *m);
int cq_multi_n_gpus(const cq_multi *m);
int cq_multi_upload(cq_multi *m, const cq_index *idx, uint32_t n_genomes);
int cq_multi_query(cq_multi *m, int mode, const uint8_t *bases, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t n_reads, cq_result *out);
int cq_multi_query_packed(cq_multi *m, int mode, const uint8_t *packed, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t n_reads, cq_result *out);
int cq_multi_reset(cq_multi *m);
/* The i-th device's context (owned by m): host-packing setting, timing, device info. */
int cq_multi_ctx(cq_multi *m, int i, cq_ctx **out);
typedef struct {
	int n_gpus;
	int nccl_version;        /* ncclGetVersion of the library in use, 0 when n_gpus = 1 */
	int devices[8];
	uint64_t shard_reads[8]; /* reads each device scanned in the last query */
	double reduce_ms;        /* CUDA-event time of the last NCCL reduce on device 0 */
} cq_multi_info;
int cq_multi_get_info(const cq_multi *m, cq_multi_info *out);

/* ----------------------------------------------------- ILP input assembly (8f.3) -- */
/*
 * What runILP_cplex / runILP_gurobi derive from the scan before they build the model
 * (query.cpp:1154-1181, 1196-1230; 1508-1535): per leaf
 *     wcov = ucount * (rl - depth) * 1.0 / rl * pow(1 - erate, depth)
 * (rl = the uint32_t average read length of query.cpp:1087, so ucount * (rl - depth) is 32-bit
 * unsigned arithmetic as in the reference), per genome the sum of its leaves' coverages over
 * Hash::map_sp[g] (the coefficient of COV[g] in the constraints EXP1 / EXP2) and the sum of
 * their rcount.  Computed on the device from the context's accumulated rcount arrays, one thread
 * per leaf; leaf order = file order (cq_index_leaves, cq_index_map_sp).  Any output may be NULL.
 * Per-leaf values agree with the host formula to the last bits of pow(); the per-genome sums are
 * accumulated with atomics (compare with a relative tolerance, 1e-12 is ample).
 */
typedef struct {
	double erate;           /* -e (the float the reference holds, widened) */
	uint32_t read_length;   /* rl */
	double *wcov_u;         /* [n_leaves_u] */
	double *wcov_d1;        /* [n_leaves_d] coverage w.r.t. refID1 */
	double *wcov_d2;        /* [n_leaves_d] coverage w.r.t. refID2 */
	double *genome_wcov_u;  /* [n_genomes+1], index 0 unused */
	double *genome_wcov_d;  /* [n_genomes+1] */
	uint64_t *genome_rcount_u; /* [n_genomes+1] */
	uint64_t *genome_rcount_d; /* [n_genomes+1] */
} cq_ilp_args;
int cq_ilp_inputs(cq_ctx *ctx, const cq_index *idx, cq_ilp_args *io);

/* ------------------------------------------- device-resident entry points (plumbing) -- */
/*
 * For callers that keep reads in HBM and combine counters themselves (bench.py, the
 * multi-GPU launcher: one process per GPU, NCCL reduce of the counter block).
 * cq_reads_stage copies ASCII reads host->device once; cq_query_staged runs pack + scan +
 * count reduction on the staged reads asynchronously on the context's stream.
 */
int cq_reads_stage(cq_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t n_reads);
/* The same for reads the caller holds packed (layout and arguments of cq_query_packed). */
int cq_reads_stage_packed(cq_ctx *ctx, const uint8_t *packed, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t n_reads);
int cq_query_staged(cq_ctx *ctx, int mode);
/* cq_query / cq_query_packed without the copy of the totals to the host: the reads flow through the
   same pack -> copy -> scan pipeline from HOST buffers and the results stay in the device
   accumulators (for launchers that combine the accumulators of several processes with their own
   collective first and fetch the reduced totals once, on one rank).  Returns when every chunk
   has been submitted; the caller's read buffers are free again after cq_sync. */
int cq_query_submit(cq_ctx *ctx, int mode, const uint8_t *bases, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t n_reads);
int cq_query_submit_packed(cq_ctx *ctx, int mode, const uint8_t *packed, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t n_reads);
int cq_sync(cq_ctx *ctx);
/* Copy the accumulated totals to host buffers (same semantics as cq_query's out). */
int cq_fetch(cq_ctx *ctx, int mode, cq_result *out);

typedef struct {
	void *d_counts;        /* uint64[2*(n_genomes+1)+4]: cnt_u | cnt_d | nundet nconf n_invalid (reserved); summing this
	                          block over devices gives the totals -- nothing device-local lives in it */
	uint64_t n_counts;
	void *d_rcount_u;      /* uint32[n_leaves_u] */
	uint64_t n_rcount_u;
	void *d_rcount_d;      /* uint32[n_leaves_d] */
	uint64_t n_rcount_d;
} cq_device_counters;
int cq_get_device_counters(cq_ctx *ctx, cq_device_counters *out);
/* Double-buffered accumulators, for a caller that overlaps ITS reduction of one batch's counters (the
 * collective of SURVEY.md section 8e) with the scan of the next batch.  The context owns two accumulator
 * sets (counter block + both rcount arrays; the second is allocated and zeroed at the first call).  The
 * call makes the other set current -- everything enqueued afterwards (cq_reset, scans, cq_fetch,
 * cq_get_device_counters) uses it -- and reports the set that WAS current in *previous (may be NULL).
 * Host-side pointer exchange only: work already enqueued keeps the set it was launched with.  The
 * caller orders its reads of `previous` after the context's stream (an event) and finishes them before
 * it swaps back and writes that set again.  The pair records of mode SC are not part of a set (they are
 * gathered per call, not reduced): swap in mode P, or fetch the pairs before swapping. */
int cq_swap_accumulators(cq_ctx *ctx, cq_device_counters *previous);
/* The cudaStream_t every call of this context enqueues on (for the caller's collectives). */
int cq_get_stream(cq_ctx *ctx, void **stream);

typedef struct {
	double h2d_ms, pack_ms, scan_ms, reduce_ms, d2h_ms, total_ms; /* last cq_query / cq_query_staged */
	uint64_t scan_launches, kernel_launches;                      /* since ctx creation */
	uint64_t probes;  /* prefix-table probes issued by the last scan (2*(rl-h+1) per valid read) */
	uint64_t bucket_hits, leaf_hits; /* last scan: probes whose h-mer is a key / leaves reached */
	uint64_t chained_loads;          /* last scan: extra bucket loads after a full bucket */
	/* CUDA-event durations summed over every step since cq_timing_reset (or ctx creation) */
	double pack_ms_sum, scan_ms_sum, reduce_ms_sum;
	uint64_t steps;
	/* geometry of the last scan launch */
	uint32_t grid_blocks, blocks_per_sm, dyn_smem_bytes, regs_per_thread;
	/* host packing of the last cq_query: wall time spent in the packer, threads used (0 = off) */
	double host_pack_ms;
	uint32_t host_pack_threads;
	uint32_t smem_carveout_pct; /* shared-memory carve-out the last scan launch asked for (percent of 228 KB) */
	uint64_t h2d_bytes;   /* bytes the last cq_query / cq_query_packed copied host->device */
	uint64_t sieve_loads; /* last scan, sieve regime: positions that passed the sieve and loaded their bucket's keys */
} cq_timing;
/* Synchronises the stream, folds the per-step CUDA events into the sums and returns them. */
int cq_get_timing(cq_ctx *ctx, cq_timing *out);
int cq_timing_reset(cq_ctx *ctx);

/* Random 32-byte-sector gather micro-benchmark over the resident prefix table: the
   measured "lookup roofline" of SURVEY.md section 8d.  n_probes random bucket reads;
   *gsectors_per_s receives 1e-9 * sectors/s (CUDA-event timed). */
int cq_bench_random_sectors(cq_ctx *ctx, uint64_t n_probes, int iters, double *gsectors_per_s);

/* Same measurement over a scratch region of region_bytes (power of two) read with
   access_bytes = 4, 8, 16 or 32 per probe: maps out the L2 / HBM random-access curve that
   sizes the resident filter.  persist != 0 pins the region with an L2 access-policy window. */
int cq_bench_random_gather(cq_ctx *ctx, uint64_t region_bytes, int access_bytes, uint64_t n_probes,
		int iters, int persist, double *gaccesses_per_s);

typedef struct {
	char name[128];
	int sm_count, cc_major, cc_minor;
	uint64_t l2_bytes, persisting_l2_max_bytes, access_policy_max_window_bytes, global_mem_bytes;
	int sm_clock_khz, mem_clock_khz, mem_bus_bits;
} cq_device_info;
int cq_get_device_info(cq_ctx *ctx, cq_device_info *out);

#ifdef __cplusplus
}
#endif
#endif
