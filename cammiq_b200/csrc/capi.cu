// C ABI of libcammiq_gpu.so (include/cammiq_gpu.h): host index load/flatten, device context,
// and the query entry points that stand in for FqReader::query64_p / query64mt_p /
// query64_sc (query.cpp:458-1080).  No CPU fallback anywhere: without a CUDA device every
// context / query call fails with CQ_ENODEV.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "capi_internal.hpp"
#include "index_codec.hpp"
#include "scan_kernels.cuh"

using namespace cammiq;

static thread_local std::string g_err;

int cqFail(int code, const std::string &msg) {
	g_err = msg;
	return code;
}
static int fail(int code, const std::string &msg) { return cqFail(code, msg); }

extern "C" const char *cq_last_error(void) { return g_err.c_str(); }
extern "C" int cq_abi_version(void) { return CQ_ABI_VERSION; }

// ------------------------------------------------------------------------------- index

extern "C" int cq_index_load(const char *path_u, const char *path_d, double load_factor, cq_index **out) {
	if (path_u == NULL || path_d == NULL || out == NULL)
		return fail(CQ_EINVAL, "cq_index_load: NULL argument.");
	*out = NULL;
	auto t0 = std::chrono::high_resolution_clock::now();
	DecodedIndex u, d;
	std::string err_u, err_d;
	int rc_u = 0, rc_d = 0;
	// two loader threads, one per index file, like FqReader::loadIdx_p (query.cpp:109-123)
	std::thread tu([&]() { rc_u = decodeIndexFile(path_u, u, err_u); });
	std::thread td([&]() { rc_d = decodeIndexFile(path_d, d, err_d); });
	tu.join();
	td.join();
	if (rc_u != 0)
		return fail(rc_u, err_u);
	if (rc_d != 0)
		return fail(rc_d, err_d);
	double decode_ms = std::chrono::duration<double, std::milli>(
		std::chrono::high_resolution_clock::now() - t0).count();
	cq_index *idx = new (std::nothrow) cq_index();
	if (idx == NULL)
		return fail(CQ_ENOMEM, "cq_index_load: out of memory.");
	std::string err;
	int rc;
	try {
		rc = flattenIndices(u, d, load_factor, idx->flat, err);
	} catch (const std::bad_alloc &) {
		delete idx;
		return fail(CQ_ENOMEM, "cq_index_load: out of memory while flattening.");
	}
	if (rc != 0) {
		delete idx;
		return fail(rc, err);
	}
	idx->flat.decode_ms = decode_ms;
	*out = idx;
	return CQ_OK;
}

extern "C" void cq_index_free(cq_index *idx) { delete idx; }

extern "C" int cq_index_get_info(const cq_index *idx, cq_index_info *info) {
	if (idx == NULL || info == NULL)
		return fail(CQ_EINVAL, "cq_index_get_info: NULL argument.");
	const FlatIndex &f = idx->flat;
	memset(info, 0, sizeof(*info));
	info->hash_len = f.hash_len;
	info->n_leaves_u = f.u.numLeaves();
	info->n_leaves_d = f.d.numLeaves();
	info->n_buckets_u = f.u.bucket_key.size();
	info->n_buckets_d = f.d.bucket_key.size();
	info->n_keys = f.n_keys;
	info->n_table_buckets = f.n_table_buckets;
	info->n_nodes_u = f.u.numNodes();
	info->n_nodes_d = f.d.numNodes();
	info->n_cnodes_u = f.cnodes_u.size() / 4;
	info->n_cnodes_d = f.cnodes_d.size() / 4;
	info->max_ref_id = std::max(f.u.max_ref_id, f.d.max_ref_id);
	info->filter_bytes = f.filter.size() * 8;
	info->device_bytes = f.deviceBytes();
	info->decode_ms = f.decode_ms;
	info->flatten_ms = f.flatten_ms;
	return CQ_OK;
}

extern "C" int cq_index_set_filter_budget(cq_index *idx, uint64_t max_bytes) {
	if (idx == NULL)
		return fail(CQ_EINVAL, "cq_index_set_filter_budget: NULL index.");
	try {
		buildFilter(idx->flat, max_bytes);
	} catch (const std::bad_alloc &) {
		return fail(CQ_ENOMEM, "cq_index_set_filter_budget: out of memory.");
	}
	return CQ_OK;
}

extern "C" int cq_index_leaves(const cq_index *idx, int table, cq_leaf_view *view) {
	if (idx == NULL || view == NULL || (table != CQ_TABLE_U && table != CQ_TABLE_D))
		return fail(CQ_EINVAL, "cq_index_leaves: bad argument.");
	const DecodedIndex &x = table == CQ_TABLE_U ? idx->flat.u : idx->flat.d;
	view->n = x.numLeaves();
	view->ref_id1 = x.ref_id1.data();
	view->ref_id2 = x.ref_id2.data();
	view->ucount1 = x.ucount1.data();
	view->ucount2 = x.ucount2.data();
	view->depth = x.depth.data();
	return CQ_OK;
}

extern "C" int cq_index_map_sp(const cq_index *idx, int table, uint32_t n_genomes, uint64_t *offsets,
		uint64_t *ids, uint64_t *total) {
	if (idx == NULL || offsets == NULL || (table != CQ_TABLE_U && table != CQ_TABLE_D))
		return fail(CQ_EINVAL, "cq_index_map_sp: bad argument.");
	const DecodedIndex &x = table == CQ_TABLE_U ? idx->flat.u : idx->flat.d;
	const uint64_t n = x.numLeaves();
	std::vector<uint64_t> cnt((size_t) n_genomes + 2, 0);
	auto ok = [&](uint32_t r) { return r >= 1 && r <= n_genomes; };
	for (uint64_t l = 0; l < n; l++) {
		if (ok(x.ref_id1[l])) cnt[x.ref_id1[l]]++;
		if (x.doubly_unique && ok(x.ref_id2[l])) cnt[x.ref_id2[l]]++;
	}
	uint64_t sum = 0;
	for (uint32_t r = 0; r <= n_genomes; r++) {
		offsets[r] = sum;
		sum += cnt[r];
	}
	offsets[n_genomes + 1] = sum;
	if (total) *total = sum;
	if (ids != NULL) {
		std::vector<uint64_t> fill((size_t) n_genomes + 2, 0);
		for (uint64_t l = 0; l < n; l++) {
			uint32_t a = x.ref_id1[l], b = x.ref_id2[l];
			if (ok(a)) ids[offsets[a] + fill[a]++] = l;
			if (x.doubly_unique && ok(b)) ids[offsets[b] + fill[b]++] = l;
		}
	}
	return CQ_OK;
}

extern "C" int cq_index_find_host(const cq_index *idx, int table, uint64_t bucket, const uint8_t *cand,
		size_t len, uint64_t *leaf) {
	if (idx == NULL || leaf == NULL || (table != CQ_TABLE_U && table != CQ_TABLE_D) || (len > 0 && cand == NULL))
		return fail(CQ_EINVAL, "cq_index_find_host: bad argument.");
	*leaf = flatFind(idx->flat, table, bucket, cand, len);
	return CQ_OK;
}

// ----------------------------------------------------------------------------- context

static void freeDevice(cq_ctx *c) {
	cudaFree(c->d_table); cudaFree(c->d_nodes_u); cudaFree(c->d_nodes_d);
	cudaFree(c->d_leaf_u_ref); cudaFree(c->d_leaf_d_ref); cudaFree(c->d_counts);
	cudaFree(c->d_rcount_u); cudaFree(c->d_rcount_d); cudaFree(c->d_partials);
	cudaFree(c->d_counts_alt); cudaFree(c->d_rcount_u_alt); cudaFree(c->d_rcount_d_alt);
	c->d_counts_alt = NULL; c->d_rcount_u_alt = c->d_rcount_d_alt = NULL;
	cudaFree(c->d_spill); cudaFree(c->d_dedup); cudaFree(c->d_probe_count); cudaFree(c->d_filter);
	c->d_filter = NULL;
	c->d_dedup = NULL;
	c->cap_spill = c->cap_dedup = 0;
	c->d_table = NULL; c->d_nodes_u = c->d_nodes_d = c->d_leaf_u_ref = NULL; c->d_leaf_d_ref = NULL;
	c->d_counts = NULL; c->d_rcount_u = c->d_rcount_d = c->d_partials = c->d_spill = NULL;
	c->d_probe_count = NULL;
	c->has_index = false;
}

// CAMMIQ_PACK_THREADS if set, else min(16, hardware threads) on hosts with at least 4, else 0.
// Packing trades PCIe bytes for host memory traffic; with several ranks on one host
// (LOCAL_WORLD_SIZE > 1, one process per GPU) the host's memory system is the shared bottleneck
// and plain DMA of the ASCII reads is faster (measured at 2 ranks), so the default is off there.
static int defaultPackThreads() {
	const char *env = getenv("CAMMIQ_PACK_THREADS");
	if (env != NULL)
		return std::max(0, atoi(env));
	const char *lws = getenv("LOCAL_WORLD_SIZE");
	if (lws != NULL && atoi(lws) > 1)
		return 0;
	const unsigned hw = std::thread::hardware_concurrency();
	return hw >= 4 ? (int) std::min(hw, 16u) : 0;
}

extern "C" int cq_ctx_create(int device, void *stream, cq_ctx **out) {
	if (out == NULL)
		return fail(CQ_EINVAL, "cq_ctx_create: NULL argument.");
	*out = NULL;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0)
		return fail(CQ_ENODEV, std::string("no usable CUDA device (") +
			(e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") + "); there is no CPU fallback.");
	if (device < 0 || device >= n)
		return fail(CQ_EINVAL, "cq_ctx_create: device ordinal out of range.");
	CQ_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop;
	CQ_CUDA(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10)
		return fail(CQ_ENODEV, std::string("device ") + prop.name + " is not sm_100; this library is built for sm_100a only.");
	cq_ctx *c = new (std::nothrow) cq_ctx();
	if (c == NULL)
		return fail(CQ_ENOMEM, "cq_ctx_create: out of memory.");
	memset(&c->timing, 0, sizeof(c->timing));
	c->device = device;
	c->n_sms = prop.multiProcessorCount;
	if (stream != NULL)
		c->stream = (cudaStream_t) stream;
	else {
		cudaError_t e2 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
		if (e2 != cudaSuccess) {
			delete c;
			return fail(CQ_ECUDA, cudaGetErrorString(e2));
		}
		c->own_stream = true;
	}
	for (int i = 0; i < 2; i++)
		cudaEventCreate(&c->ev[i]);
	for (int i = 0; i < cq_ctx::kStages; i++) {
		cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming);
		cudaEventCreateWithFlags(&c->ev_free[i], cudaEventDisableTiming);
	}
	cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
	c->pack_threads = defaultPackThreads();
	*out = c;
	return CQ_OK;
}

extern "C" void cq_ctx_destroy(cq_ctx *c) {
	if (c == NULL)
		return;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	freeDevice(c);
	cudaFree(c->d_bases); cudaFree(c->d_offsets); cudaFree(c->d_lengths);
	cudaFree(c->d_pairs); cudaFree(c->d_pair_table); cudaFree(c->d_pair_out); cudaFree(c->d_read_class); cudaFree(c->d_read_rid_a);
	cudaFree(c->d_read_rid_b); cudaFree(c->d_nleaf_u); cudaFree(c->d_nleaf_d); cudaFree(c->d_leaf_u);
	cudaFree(c->d_leaf_d);
	for (int i = 0; i < 2; i++)
		if (c->ev[i]) cudaEventDestroy(c->ev[i]);
	for (int i = 0; i < cq_ctx::kStages; i++) {
		if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]);
		if (c->ev_free[i]) cudaEventDestroy(c->ev_free[i]);
		cudaFree(c->d_cbases[i]); cudaFree(c->d_coffsets[i]); cudaFree(c->d_coffsets32[i]); cudaFree(c->d_clengths[i]);
		cudaFree(c->d_words[i]); cudaFree(c->d_len2[i]);
		cudaFreeHost(c->h_pbases[i]); cudaFreeHost(c->h_plengths[i]); cudaFreeHost(c->h_poffsets[i]);
	}
	cudaFree(c->d_words[cq_ctx::kStages]); cudaFree(c->d_len2[cq_ctx::kStages]);
	delete c->pool;
	if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
	for (auto &se : c->steps)
		for (int i = 0; i < 4; i++)
			cudaEventDestroy(se.e[i]);
	if (c->own_stream)
		cudaStreamDestroy(c->stream);
	delete c;
}

// The index arrays live in pageable memory; a plain cudaMemcpy of them is staged by the driver
// on one thread (~5 GB/s).  Large arrays go through two pinned buffers filled by a few host
// threads instead, so the copy engine sees pinned memory and runs near PCIe speed.
struct Uploader {
	static const size_t kChunk = 32u << 20;
	cudaStream_t st;
	WorkerPool pool;
	uint8_t *stage[2] = {NULL, NULL};
	cudaEvent_t ev[2] = {NULL, NULL};
	uint64_t k = 0;
	explicit Uploader(cudaStream_t s) : st(s), pool((int) std::max(1u, std::min(std::thread::hardware_concurrency(), 8u))) {}
	~Uploader() {
		cudaStreamSynchronize(st);
		for (int i = 0; i < 2; i++) {
			if (stage[i]) cudaFreeHost(stage[i]);
			if (ev[i]) cudaEventDestroy(ev[i]);
		}
	}
	int copy(void *dst, const void *src, size_t bytes) {
		if (bytes < (4u << 20)) {
			CQ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
			return CQ_OK;
		}
		for (int i = 0; i < 2; i++)
			if (stage[i] == NULL) {
				CQ_CUDA(cudaHostAlloc((void **) &stage[i], kChunk, cudaHostAllocDefault));
				CQ_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
			}
		const int T = pool.size();
		for (size_t off = 0; off < bytes; off += kChunk, k++) {
			const int b = (int) (k & 1);
			const size_t n = std::min(kChunk, bytes - off);
			if (k >= 2)
				CQ_CUDA(cudaEventSynchronize(ev[b]));
			const uint8_t *from = (const uint8_t *) src + off;
			uint8_t *to = stage[b];
			pool.run([&](int t) {
				const size_t lo = n * (size_t) t / (size_t) T, hi = n * ((size_t) t + 1) / (size_t) T;
				memcpy(to + lo, from + lo, hi - lo);
			});
			CQ_CUDA(cudaMemcpyAsync((uint8_t *) dst + off, to, n, cudaMemcpyHostToDevice, st));
			CQ_CUDA(cudaEventRecord(ev[b], st));
		}
		return CQ_OK;
	}
};

template <typename T>
static int uploadArray(T **dst, const T *src, size_t n, Uploader &up) {
	size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
	CQ_CUDA(cudaMalloc((void **) dst, bytes));
	if (n > 0)
		return up.copy(*dst, src, n * sizeof(T));
	return CQ_OK;
}

extern "C" int cq_index_upload(cq_ctx *c, const cq_index *idx, uint32_t n_genomes) {
	if (c == NULL || idx == NULL || n_genomes == 0)
		return fail(CQ_EINVAL, "cq_index_upload: bad argument.");
	const FlatIndex &f = idx->flat;
	if (std::max(f.u.max_ref_id, f.d.max_ref_id) > n_genomes)
		return fail(CQ_EINVAL, "cq_index_upload: a leaf carries a genome id larger than n_genomes.");
	for (size_t i = 0; i < f.u.ref_id1.size(); i++)
		if (f.u.ref_id1[i] == 0)
			return fail(CQ_EINVAL, "cq_index_upload: a unique leaf carries genome id 0.");
	CQ_CUDA(cudaSetDevice(c->device));
	freeDevice(c);
	int rc;
	const bool verbose = getenv("CAMMIQ_VERBOSE") != NULL;
	auto t0 = std::chrono::high_resolution_clock::now();
	auto lap = [&](const char *what) {
		if (!verbose) return;
		cudaStreamSynchronize(c->stream);
		auto t1 = std::chrono::high_resolution_clock::now();
		fprintf(stderr, "[upload] %s %.0f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
		t0 = t1;
	};
	Uploader up(c->stream);
	if ((rc = uploadArray(&c->d_table, f.table.data(), f.table.size(), up)) != 0) return rc;
	lap("prefix table");
	if ((rc = uploadArray(&c->d_nodes_u, f.cnodes_u.data(), f.cnodes_u.size(), up)) != 0) return rc;
	if ((rc = uploadArray(&c->d_nodes_d, f.cnodes_d.data(), f.cnodes_d.size(), up)) != 0) return rc;
	if ((rc = uploadArray(&c->d_leaf_u_ref, f.u.ref_id1.data(), f.u.ref_id1.size(), up)) != 0) return rc;
	FlatVec<uint2>::type dref(f.d.numLeaves());
	up.pool.run([&](int t) {
		const size_t T = (size_t) up.pool.size(), lo = dref.size() * (size_t) t / T, hi = dref.size() * ((size_t) t + 1) / T;
		for (size_t i = lo; i < hi; i++)
			dref[i] = make_uint2(f.d.ref_id1[i], f.d.ref_id2[i]);
	});
	if ((rc = uploadArray(&c->d_leaf_d_ref, dref.data(), dref.size(), up)) != 0) return rc;
	CQ_CUDA(cudaStreamSynchronize(c->stream));
	lap("tries and leaf records");

	c->h = f.hash_len;
	c->n_genomes = n_genomes;
	c->n_leaves_u = f.u.numLeaves();
	c->n_leaves_d = f.d.numLeaves();
	c->table_mask = f.n_table_buckets - 1;
	c->table_shift = f.table_shift;
	const size_t ncnt = 2 * ((size_t) n_genomes + 1);
	CQ_CUDA(cudaMalloc((void **) &c->d_counts, (ncnt + 4) * sizeof(unsigned long long)));
	CQ_CUDA(cudaMalloc((void **) &c->d_rcount_u, std::max<size_t>(c->n_leaves_u, 1) * 4));
	CQ_CUDA(cudaMalloc((void **) &c->d_rcount_d, std::max<size_t>(c->n_leaves_d, 1) * 4));
	CQ_CUDA(cudaMalloc((void **) &c->d_probe_count, 8 * sizeof(unsigned long long)));
	CQ_CUDA(cudaMemsetAsync(c->d_probe_count, 0, 8 * sizeof(unsigned long long), c->stream));

	if (!f.filter.empty()) {
		if ((rc = uploadArray((uint64_t **) &c->d_filter, f.filter.data(), f.filter.size(), up)) != 0) return rc;
		c->filter_words = f.filter_words;
		c->filter_sel = f.filter_sel_mask;
		c->filter_sieve = f.filter_sieve;
		CQ_CUDA(cudaStreamSynchronize(c->stream));
	}
	// block-private genome counters in shared memory when they fit, global atomics otherwise;
	// the grid itself is sized per launch (it depends on the tile's shared-memory footprint)
	// (a large counter block would eat the L1 the in-flight probes need: beyond kMaxSmemGenomes the
	// warps' match_any-combined adds go to global memory instead)
	c->smem_counters = n_genomes <= kMaxSmemGenomes;
	c->smem_bytes = c->smem_counters ? ncnt * sizeof(uint32_t) : 0;
	c->max_grid = kMaxBlocksPerSM * c->n_sms;
	if (c->smem_counters)
		CQ_CUDA(cudaMalloc((void **) &c->d_partials, (size_t) c->max_grid * ncnt * sizeof(uint32_t)));
	c->has_index = true;
	lap("filter, counters, scratch");
	return cq_reset(c);
}

extern "C" int cq_reset(cq_ctx *c) {
	if (c == NULL || !c->has_index)
		return fail(CQ_ESTATE, "cq_reset: no index resident.");
	CQ_CUDA(cudaSetDevice(c->device));
	const size_t ncnt = 2 * ((size_t) c->n_genomes + 1);
	CQ_CUDA(cudaMemsetAsync(c->d_counts, 0, (ncnt + 4) * sizeof(unsigned long long), c->stream));
	CQ_CUDA(cudaMemsetAsync(c->d_rcount_u, 0, std::max<size_t>(c->n_leaves_u, 1) * 4, c->stream));
	CQ_CUDA(cudaMemsetAsync(c->d_rcount_d, 0, std::max<size_t>(c->n_leaves_d, 1) * 4, c->stream));
	CQ_CUDA(cudaMemsetAsync(c->d_probe_count + 6, 0, 16, c->stream)); // the scan's tile counter, the pair records held
	c->sc_reads_since_reset = 0;
	return CQ_OK; // stream-ordered; every reader of the counters is on the same stream
}

// a stage buffer of the chunk pipeline may still be read by the scan that used it kStages chunks
// ago: a buffer that has to grow waits for that scan first (and grows with headroom, so that
// chunks of slightly different extents do not reallocate at all)
template <typename T>
static int ensureStage(cq_ctx *c, int b, T **ptr, size_t *cap, size_t need) {
	if (need <= *cap && *ptr != NULL)
		return CQ_OK;
	if (*ptr != NULL)
		CQ_CUDA(cudaEventSynchronize(c->ev_free[b]));
	if (*ptr) cudaFree(*ptr);
	*ptr = NULL;
	*cap = 0;
	const size_t n = std::max<size_t>(need + need / 8, 1);
	CQ_CUDA(cudaMalloc((void **) ptr, n * sizeof(T)));
	*cap = n;
	return CQ_OK;
}

template <typename T>
static int ensure(T **ptr, size_t *cap, size_t need) {
	if (need <= *cap && *ptr != NULL)
		return CQ_OK;
	if (*ptr) cudaFree(*ptr);
	*ptr = NULL;
	*cap = 0;
	size_t n = std::max<size_t>(need, 1);
	CQ_CUDA(cudaMalloc((void **) ptr, n * sizeof(T)));
	*cap = n;
	return CQ_OK;
}

// ------------------------------------------------------------------------------- query

// Fold the CUDA events of the steps issued so far into the timing sums (synchronises).
static int foldStepEvents(cq_ctx *c) {
	if (c->steps_used == 0)
		return CQ_OK;
	CQ_CUDA(cudaStreamSynchronize(c->stream));
	for (size_t i = 0; i < c->steps_used; i++) {
		float pack = 0, scan = 0, red = 0;
		CQ_CUDA(cudaEventElapsedTime(&pack, c->steps[i].e[0], c->steps[i].e[1]));
		CQ_CUDA(cudaEventElapsedTime(&scan, c->steps[i].e[1], c->steps[i].e[2]));
		CQ_CUDA(cudaEventElapsedTime(&red, c->steps[i].e[2], c->steps[i].e[3]));
		c->timing.pack_ms = pack;
		c->timing.scan_ms = scan;
		c->timing.reduce_ms = red;
		c->timing.pack_ms_sum += pack;
		c->timing.scan_ms_sum += scan;
		c->timing.reduce_ms_sum += red;
		c->timing.steps++;
	}
	c->steps_used = 0;
	return CQ_OK;
}

static int stageReads(cq_ctx *c, bool packed, const uint8_t *bases, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t n_reads) {
	if (c == NULL || !c->has_index)
		return fail(CQ_ESTATE, "cq_reads_stage: no index resident.");
	if (n_reads > 0 && (bases == NULL || lengths == NULL))
		return fail(CQ_EINVAL, "cq_reads_stage: NULL read buffers.");
	CQ_CUDA(cudaSetDevice(c->device));
	// byte range the reads span in the caller's buffer, and the longest read (sets the tile size)
	uint64_t lo = ~0ull, hi = 0;
	uint32_t max_len = 1;
	if (offsets != NULL) {
		for (uint64_t i = 0; i < n_reads; i++) {
			lo = std::min(lo, offsets[i]);
			hi = std::max(hi, offsets[i] + (packed ? packedBytes(lengths[i]) : lengths[i]));
			max_len = std::max<uint32_t>(max_len, lengths[i]);
		}
	} else if (n_reads > 0) {
		// fixed stride: the span starts at 0 and only the reads within 255 bytes of the end can set its extent
		for (uint64_t i = 0; i < n_reads; i++)
			max_len = std::max<uint32_t>(max_len, lengths[i]);
		lo = 0;
		for (uint64_t i = n_reads; i-- > 0;) {
			hi = std::max(hi, i * stride + (packed ? packedBytes(lengths[i]) : lengths[i]));
			if ((n_reads - 1 - i) * stride >= 255)
				break;
		}
	}
	if (hi < lo) lo = hi = 0;
	const uint64_t copy_lo = lo & ~15ull, total = hi - copy_lo;
	int rc;
	if ((rc = ensure(&c->d_bases, &c->cap_bases, total + 64)) != 0) return rc;
	if ((rc = ensure(&c->d_lengths, &c->cap_reads_len, n_reads)) != 0) return rc;
	if (offsets && (rc = ensure(&c->d_offsets, &c->cap_reads_off, n_reads)) != 0) return rc;
	CQ_CUDA(cudaEventRecord(c->ev[0], c->stream));
	if (n_reads > 0) {
		if (total > 0)
			CQ_CUDA(cudaMemcpyAsync(c->d_bases, bases + copy_lo, total, cudaMemcpyHostToDevice, c->stream));
		CQ_CUDA(cudaMemcpyAsync(c->d_lengths, lengths, n_reads, cudaMemcpyHostToDevice, c->stream));
		if (offsets)
			CQ_CUDA(cudaMemcpyAsync(c->d_offsets, offsets, n_reads * 8, cudaMemcpyHostToDevice, c->stream));
	}
	CQ_CUDA(cudaEventRecord(c->ev[1], c->stream));
	c->staged_shift = copy_lo;
	c->staged_reads = n_reads;
	c->staged_stride = stride;
	c->staged_has_offsets = offsets != NULL;
	c->staged_bytes = total;
	c->staged_max_len = max_len;
	c->staged_packed = packed;
	return CQ_OK;
}

extern "C" int cq_reads_stage(cq_ctx *c, const uint8_t *bases, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t n_reads) {
	return stageReads(c, false, bases, offsets, stride, lengths, n_reads);
}

extern "C" int cq_reads_stage_packed(cq_ctx *c, const uint8_t *packed, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t n_reads) {
	return stageReads(c, true, packed, offsets, stride, lengths, n_reads);
}

// A batch of reads resident on the device as the caller holds them.  `bases` is the address read
// offset 0 would have (the batch may hold only a slice of the caller's base buffer), `first` is
// the caller's index of the batch's first read, `slot` the tile-layout buffer to use.
struct ReadBatch {
	const uint8_t *bases;
	const uint64_t *offsets;
	uint64_t stride;
	const uint8_t *lengths;
	uint64_t n, first;
	uint32_t max_len;
	bool packed;               // 2-bit bytes from the host packer instead of ASCII
	const uint32_t *offsets32; // packed batches: batch-relative 32-bit offsets
	int slot;                  // 0..kStages-1: pipeline stage, kStages: the staged reads
};

static uint32_t wordsPerRead(uint32_t longest) { return (((std::max<uint32_t>(longest, 1) + 15) / 16) + 2) | 1u; }

// pack_tiles_kernel: the batch -> the scan's tile layout (+ validated lengths for ASCII input)
static int launchPack(cq_ctx *c, const ReadBatch &rb) {
	if (rb.n == 0)
		return CQ_OK;
	const int s = rb.slot;
	PackParams q;
	memset(&q, 0, sizeof(q));
	q.words_per_read = wordsPerRead(rb.max_len);
	q.n_reads = rb.n;
	q.n_padded = (rb.n + 31) & ~31ull;
	int rc;
	if ((rc = ensure(&c->d_words[s], &c->cap_words[s], (size_t) q.n_padded * q.words_per_read)) != 0) return rc;
	q.bases = rb.bases;
	q.offsets = rb.offsets;
	q.offsets32 = rb.offsets32;
	q.stride = rb.stride;
	q.read_base = rb.first;
	q.lengths_in = rb.lengths;
	q.words = c->d_words[s];
	if (!rb.packed) {
		if ((rc = ensure(&c->d_len2[s], &c->cap_len2[s], (size_t) rb.n)) != 0) return rc;
		CQ_CUDA(cudaMemcpyAsync(c->d_len2[s], rb.lengths, rb.n, cudaMemcpyDeviceToDevice, c->stream));
		q.lengths_out = c->d_len2[s];
	}
	q.base_words = (std::max<uint32_t>(rb.max_len, 1) + 15) / 16;
	q.reads_per_block = 256 / q.base_words;
	q.inv_words = (65536 + q.base_words - 1) / q.base_words;
	const uint64_t n_passes = (q.n_padded + q.reads_per_block - 1) / q.reads_per_block;
	const unsigned blocks = (unsigned) std::min<uint64_t>(n_passes, (uint64_t) c->n_sms * 16);
	if (rb.packed)
		pack_tiles_kernel<true><<<blocks, 256, 0, c->stream>>>(q);
	else
		pack_tiles_kernel<false><<<blocks, 256, 0, c->stream>>>(q);
	c->timing.kernel_launches++;
	return CQ_OK;
}

// One launch of the scan (+ partial-count reduction) over a batch launchPack has prepared.
static int launchScan(cq_ctx *c, int mode, const ReadBatch &rb) {
	if (rb.n == 0)
		return CQ_OK;
	const size_t ncnt = 2 * ((size_t) c->n_genomes + 1);
	ScanParams sp;
	memset(&sp, 0, sizeof(sp));
	sp.table = c->d_table;
	sp.table_mask = c->table_mask;
	sp.table_shift = c->table_shift;
	sp.nodes_u = c->d_nodes_u;
	sp.nodes_d = c->d_nodes_d;
	sp.leaf_u_ref = c->d_leaf_u_ref;
	sp.leaf_d_ref = c->d_leaf_d_ref;
	sp.h = c->h;
	sp.n_genomes = c->n_genomes;
	sp.filter = c->d_filter;
	sp.filter_words = c->filter_words;
	sp.filter_sel = c->filter_sel;
	sp.words = c->d_words[rb.slot];
	sp.lengths = rb.packed ? rb.lengths : c->d_len2[rb.slot];
	sp.n_reads = rb.n;
	// two tile buffers per warp: 16 bases per word, two words of slack per read (three-word
	// windows), odd stride so that the lanes' reads start in different banks
	const uint32_t longest = std::max<uint32_t>(rb.max_len, 1);
	sp.words_per_read = wordsPerRead(longest);
	const size_t dyn_smem = (size_t) kWarpsPerBlock * kTileBufs * 32 * sp.words_per_read * 4 + c->smem_bytes;
	// a read can reach 2 tables x 2 strands x (longest - h + 1) leaves; what exceeds the shared
	// hit slots spills to global scratch, sized (and grown) for the longest read seen so far
	const uint32_t max_hits = longest >= c->h ? 4 * (longest - c->h + 1) : 0;
	sp.spill_stride = max_hits > (uint32_t) kHitSeg ? ((max_hits - kHitSeg + 3) & ~3u) : 4;
	// CAMMIQ_LIGHT_HITS (experiments): a huge value sends every hit list through the per-lane quadratic path
	sp.light_hits = getenv("CAMMIQ_LIGHT_HITS") ? (uint32_t) atol(getenv("CAMMIQ_LIGHT_HITS")) : (uint32_t) kLightHits;
	sp.dedup_slots = 64;
	while (sp.dedup_slots < 2 * (kHitSeg + sp.spill_stride))
		sp.dedup_slots <<= 1;
	{
		const size_t warps = (size_t) c->max_grid * kWarpsPerBlock;
		const size_t need_spill = warps * 32 * sp.spill_stride, need_dedup = warps * sp.dedup_slots;
		if (need_spill > c->cap_spill || need_dedup > c->cap_dedup) {
			// scratch of launches still in flight: let them finish before it moves
			CQ_CUDA(cudaStreamSynchronize(c->stream));
			int rc;
			if ((rc = ensure(&c->d_spill, &c->cap_spill, need_spill)) != 0) return rc;
			if ((rc = ensure(&c->d_dedup, &c->cap_dedup, need_dedup)) != 0) return rc;
		}
	}
	sp.smem_counters = c->smem_counters ? 1 : 0;
	sp.partials = c->d_partials;
	sp.counts = c->d_counts;
	sp.rcount_u = c->d_rcount_u;
	sp.rcount_d = c->d_rcount_d;
	sp.pair_records = c->d_pairs;
	sp.pair_count = c->d_probe_count + 7; // last word of the statistics block: survives the per-step clearing of the first five
	sp.hit_spill = c->d_spill;
	sp.dedup_sets = c->d_dedup;
	sp.probe_count = c->d_probe_count;
	sp.tile_counter = reinterpret_cast<uint32_t *>(c->d_probe_count + 6); // 0 between launches (the kernel resets it)
	if (c->want_per_read) {
		sp.read_class = c->d_read_class + rb.first;
		sp.read_rid_a = c->d_read_rid_a + rb.first;
		sp.read_rid_b = c->d_read_rid_b + rb.first;
	}
	if (c->want_sets) {
		sp.leaf_cap = c->leaf_cap;
		sp.read_nleaf_u = c->d_nleaf_u + rb.first;
		sp.read_nleaf_d = c->d_nleaf_d + rb.first;
		sp.read_leaf_u = c->d_leaf_u + rb.first * c->leaf_cap;
		sp.read_leaf_d = c->d_leaf_d + rb.first * c->leaf_cap;
	}
	const int filt = c->d_filter == NULL ? 0 : c->filter_sieve ? 2 : 1;
	static const void *const kernels[6] = {
		(const void *) scan_reads_kernel<CQ_MODE_P, 0>, (const void *) scan_reads_kernel<CQ_MODE_P, 1>,
		(const void *) scan_reads_kernel<CQ_MODE_P, 2>, (const void *) scan_reads_kernel<CQ_MODE_SC, 0>,
		(const void *) scan_reads_kernel<CQ_MODE_SC, 1>, (const void *) scan_reads_kernel<CQ_MODE_SC, 2>};
	const int variant = mode * 3 + filt;
	const void *kern = kernels[variant];
	if (dyn_smem != c->last_dyn_smem[variant]) {
		CQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dyn_smem));
		int per_sm = 0;
		CQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kScanThreads, dyn_smem));
		if (per_sm < 1)
			return fail(CQ_ECUDA, "scan kernel does not fit on an SM.");
		c->last_dyn_smem[variant] = dyn_smem;
		c->last_per_sm[variant] = std::min(per_sm, kMaxBlocksPerSM);
		if (getenv("CAMMIQ_MAX_BLOCKS"))
			c->last_per_sm[variant] = std::min(c->last_per_sm[variant], atoi(getenv("CAMMIQ_MAX_BLOCKS")));
		cudaFuncAttributes fa;
		size_t static_smem = 0;
		if (cudaFuncGetAttributes(&fa, kern) == cudaSuccess) {
			c->timing.regs_per_thread = (uint32_t) fa.numRegs;
			static_smem = fa.sharedSizeBytes;
		}
		// The probes in flight live in the SM's L1 (a random gather holds a 128-byte line while it is
		// outstanding; tools/microbench_l1.py: 284 G gathers/s with the whole 256 KB as L1, 82 G/s with
		// 28 KB), and L1 is what the shared-memory carve-out leaves.  Left to itself the driver sizes
		// the carve-out for the occupancy the kernel COULD reach; ask for what this launch uses.
		const size_t per_block = dyn_smem + static_smem + 1024;
		int pct = (int) ((per_block * (size_t) c->last_per_sm[variant] * 100 + 228 * 1024 - 1) / (228 * 1024));
		pct = std::max(1, std::min(100, pct));
		if (getenv("CAMMIQ_CARVEOUT"))
			pct = atoi(getenv("CAMMIQ_CARVEOUT"));
		if (pct >= 0)
			CQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
		c->last_carveout[variant] = pct;
	}
	c->timing.smem_carveout_pct = (uint32_t) std::max(0, c->last_carveout[variant]);
	const uint64_t n_tiles = (rb.n + kScanThreads - 1) / kScanThreads; // one 32-read sub-tile per warp at least
	c->grid = (int) std::min<uint64_t>((uint64_t) c->last_per_sm[variant] * c->n_sms, n_tiles);
	c->timing.grid_blocks = (uint32_t) c->grid;
	c->timing.blocks_per_sm = (uint32_t) c->last_per_sm[variant];
	c->timing.dyn_smem_bytes = (uint32_t) dyn_smem;
	void *args[] = {&sp};
	CQ_CUDA(cudaLaunchKernel(kern, dim3(c->grid), dim3(kScanThreads), args, dyn_smem, c->stream));
	c->timing.kernel_launches++;
	c->timing.scan_launches++;
	if (c->smem_counters) {
		reduce_partials_kernel<<<(unsigned) ((ncnt + 255) / 256), 256, 0, c->stream>>>(
			c->d_partials, (uint32_t) c->grid, (uint32_t) ncnt, c->d_counts);
		c->timing.kernel_launches++;
	}
	return CQ_OK;
}

// Device buffers a query of n reads needs besides the reads themselves.
static int prepareOutputs(cq_ctx *c, int mode, uint64_t n) {
	int rc;
	const size_t ncnt = 2 * ((size_t) c->n_genomes + 1);
	(void) ncnt;
	if (mode == CQ_MODE_SC) {
		// worst case every read adds a pair record on top of those already held; the number held is
		// bounded by the reads submitted since the last reset (no round trip to the device)
		const size_t have = (size_t) c->sc_reads_since_reset;
		size_t need = have + n;
		c->sc_reads_since_reset += n;
		if (need > c->cap_pairs) {
			need += need / 4;
			unsigned long long *np = NULL;
			CQ_CUDA(cudaMalloc((void **) &np, std::max<size_t>(need, 1) * 8));
			if (have > 0)
				CQ_CUDA(cudaMemcpyAsync(np, c->d_pairs, have * 8, cudaMemcpyDeviceToDevice, c->stream));
			CQ_CUDA(cudaStreamSynchronize(c->stream));
			cudaFree(c->d_pairs);
			c->d_pairs = np;
			c->cap_pairs = need;
		}
	}
	if (c->want_per_read) {
		if ((rc = ensure(&c->d_read_class, &c->cap_read_class, (size_t) n)) != 0) return rc;
		if ((rc = ensure(&c->d_read_rid_a, &c->cap_read_rid_a, (size_t) n)) != 0) return rc;
		if ((rc = ensure(&c->d_read_rid_b, &c->cap_read_rid_b, (size_t) n)) != 0) return rc;
	}
	if (c->want_sets) {
		if ((rc = ensure(&c->d_nleaf_u, &c->cap_nleaf_u, (size_t) n)) != 0) return rc;
		if ((rc = ensure(&c->d_nleaf_d, &c->cap_nleaf_d, (size_t) n)) != 0) return rc;
		if ((rc = ensure(&c->d_leaf_u, &c->cap_leaf_u, (size_t) n * c->leaf_cap)) != 0) return rc;
		if ((rc = ensure(&c->d_leaf_d, &c->cap_leaf_d, (size_t) n * c->leaf_cap)) != 0) return rc;
		CQ_CUDA(cudaMemsetAsync(c->d_leaf_u, 0, std::max<size_t>((size_t) n * c->leaf_cap, 1) * 4, c->stream));
		CQ_CUDA(cudaMemsetAsync(c->d_leaf_d, 0, std::max<size_t>((size_t) n * c->leaf_cap, 1) * 4, c->stream));
	}
	return CQ_OK;
}

static int beginStep(cq_ctx *c, cudaEvent_t **sev) {
	int rc;
	if (c->steps_used == c->steps.size()) {
		if (c->steps.size() >= 1024) {
			if ((rc = foldStepEvents(c)) != 0) return rc;
		} else {
			cq_ctx::StepEvents se;
			for (int i = 0; i < 4; i++)
				CQ_CUDA(cudaEventCreate(&se.e[i]));
			c->steps.push_back(se);
		}
	}
	*sev = c->steps[c->steps_used++].e;
	CQ_CUDA(cudaMemsetAsync(c->d_probe_count, 0, 40, c->stream));
	return CQ_OK;
}

extern "C" int cq_query_staged(cq_ctx *c, int mode) {
	if (c == NULL || !c->has_index)
		return fail(CQ_ESTATE, "cq_query_staged: no index resident.");
	if (mode != CQ_MODE_P && mode != CQ_MODE_SC)
		return fail(CQ_EINVAL, "cq_query_staged: bad mode.");
	CQ_CUDA(cudaSetDevice(c->device));
	int rc;
	if ((rc = prepareOutputs(c, mode, c->staged_reads)) != 0) return rc;
	cudaEvent_t *sev;
	if ((rc = beginStep(c, &sev)) != 0) return rc;
	CQ_CUDA(cudaEventRecord(sev[0], c->stream));
	ReadBatch rb = {c->d_bases - c->staged_shift, c->staged_has_offsets ? c->d_offsets : NULL, c->staged_stride, c->d_lengths,
		c->staged_reads, 0, c->staged_max_len, c->staged_packed, NULL, cq_ctx::kStages};
	if ((rc = launchPack(c, rb)) != 0) return rc;
	CQ_CUDA(cudaEventRecord(sev[1], c->stream));
	if ((rc = launchScan(c, mode, rb)) != 0) return rc;
	CQ_CUDA(cudaEventRecord(sev[2], c->stream));
	CQ_CUDA(cudaEventRecord(sev[3], c->stream));
	CQ_CUDA(cudaGetLastError());
	return CQ_OK;
}

extern "C" int cq_sync(cq_ctx *c) {
	if (c == NULL)
		return fail(CQ_EINVAL, "cq_sync: NULL context.");
	CQ_CUDA(cudaSetDevice(c->device));
	CQ_CUDA(cudaStreamSynchronize(c->stream));
	return CQ_OK;
}

// read_cnts_b (query.cpp:994-997): the per-read pair records are folded into (pair, count)
// entries on the device; only the distinct pairs come back
int cqCollectPairs(cq_ctx *c, std::vector<cq_pair_count> &out) {
	out.clear();
	CQ_CUDA(cudaSetDevice(c->device));
	const size_t ncnt = 2 * ((size_t) c->n_genomes + 1);
	unsigned long long nrec = 0;
	(void) ncnt;
	CQ_CUDA(cudaMemcpyAsync(&nrec, c->d_probe_count + 7, 8, cudaMemcpyDeviceToHost, c->stream));
	CQ_CUDA(cudaStreamSynchronize(c->stream));
	if (nrec == 0)
		return CQ_OK;
	uint64_t slots = 1024;
	while (slots < 2 * nrec)
		slots <<= 1;
	int rc;
	if ((rc = ensure(&c->d_pair_table, &c->cap_pair_table, (size_t) slots)) != 0) return rc;
	if ((rc = ensure(&c->d_pair_out, &c->cap_pair_out, (size_t) nrec + 1)) != 0) return rc;
	init_pairs_kernel<<<(unsigned) ((slots + 255) / 256), 256, 0, c->stream>>>(c->d_pair_table, slots);
	aggregate_pairs_kernel<<<(unsigned) ((nrec + 255) / 256), 256, 0, c->stream>>>(c->d_pairs, nrec, c->d_pair_table, slots - 1);
	unsigned long long *d_n = reinterpret_cast<unsigned long long *>(c->d_pair_out + nrec);
	CQ_CUDA(cudaMemsetAsync(d_n, 0, sizeof(PairSlot), c->stream));
	compact_pairs_kernel<<<(unsigned) ((slots + 255) / 256), 256, 0, c->stream>>>(c->d_pair_table, slots, c->d_pair_out, d_n);
	c->timing.kernel_launches += 3;
	unsigned long long n_distinct = 0;
	CQ_CUDA(cudaMemcpyAsync(&n_distinct, d_n, 8, cudaMemcpyDeviceToHost, c->stream));
	CQ_CUDA(cudaStreamSynchronize(c->stream));
	std::vector<PairSlot> agg((size_t) n_distinct);
	if (n_distinct > 0)
		CQ_CUDA(cudaMemcpy(agg.data(), c->d_pair_out, (size_t) n_distinct * sizeof(PairSlot), cudaMemcpyDeviceToHost));
	out.resize(agg.size());
	for (size_t i = 0; i < agg.size(); i++) {
		out[i].a = (uint32_t) (agg[i].key >> 32);
		out[i].b = (uint32_t) agg[i].key;
		out[i].count = agg[i].count;
	}
	return CQ_OK;
}

static bool pairLess(const cq_pair_count &x, const cq_pair_count &y) { return x.a != y.a ? x.a < y.a : x.b < y.b; }

int cqFetchFrom(cq_ctx *c, int mode, const unsigned long long *d_counts, const uint32_t *d_rcount_u, const uint32_t *d_rcount_d,
		cq_result *out, bool with_pairs) {
	CQ_CUDA(cudaSetDevice(c->device));
	const size_t G1 = (size_t) c->n_genomes + 1, ncnt = 2 * G1;
	std::vector<unsigned long long> counts(ncnt + 4);
	CQ_CUDA(cudaMemcpyAsync(counts.data(), d_counts, (ncnt + 4) * 8, cudaMemcpyDeviceToHost, c->stream));
	if (out->rcount_u && mode == CQ_MODE_P && c->n_leaves_u)
		CQ_CUDA(cudaMemcpyAsync(out->rcount_u, d_rcount_u, c->n_leaves_u * 4, cudaMemcpyDeviceToHost, c->stream));
	if (out->rcount_d && mode == CQ_MODE_P && c->n_leaves_d)
		CQ_CUDA(cudaMemcpyAsync(out->rcount_d, d_rcount_d, c->n_leaves_d * 4, cudaMemcpyDeviceToHost, c->stream));
	CQ_CUDA(cudaStreamSynchronize(c->stream));
	if (out->cnt_u) memcpy(out->cnt_u, counts.data(), G1 * 8);
	if (out->cnt_d) memcpy(out->cnt_d, counts.data() + G1, G1 * 8);
	out->nundet = counts[ncnt];
	out->nconf = counts[ncnt + 1];
	out->n_invalid = counts[ncnt + 2];
	out->n_pairs = 0;
	if (mode == CQ_MODE_SC && with_pairs) {
		std::vector<cq_pair_count> agg;
		int rc = cqCollectPairs(c, agg);
		if (rc != 0) return rc;
		std::sort(agg.begin(), agg.end(), pairLess);
		const uint64_t np = agg.size();
		for (uint64_t i = 0; i < np && out->pairs && i < out->pairs_cap; i++)
			out->pairs[i] = agg[i];
		out->n_pairs = np;
		if (out->pairs && np > out->pairs_cap)
			return fail(CQ_EINVAL, "cq_fetch: pairs_cap too small for the pair map.");
	}
	return CQ_OK;
}

extern "C" int cq_fetch(cq_ctx *c, int mode, cq_result *out) {
	if (c == NULL || !c->has_index || out == NULL)
		return fail(CQ_ESTATE, "cq_fetch: no index resident or NULL result.");
	return cqFetchFrom(c, mode, c->d_counts, c->d_rcount_u, c->d_rcount_d, out, true);
}

static int fetchPerRead(cq_ctx *c, uint64_t n, cq_result *out) {
	if ((c->want_per_read || c->want_sets) && n > 0)
		CQ_CUDA(cudaStreamSynchronize(c->stream)); // the copies below run on the default stream
	if (c->want_per_read && n > 0) {
		CQ_CUDA(cudaMemcpy(out->read_class, c->d_read_class, n, cudaMemcpyDeviceToHost));
		CQ_CUDA(cudaMemcpy(out->read_rid_a, c->d_read_rid_a, n * 4, cudaMemcpyDeviceToHost));
		CQ_CUDA(cudaMemcpy(out->read_rid_b, c->d_read_rid_b, n * 4, cudaMemcpyDeviceToHost));
	}
	if (c->want_sets && n > 0) {
		const uint32_t cap = c->leaf_cap;
		CQ_CUDA(cudaMemcpy(out->read_nleaf_u, c->d_nleaf_u, n * 4, cudaMemcpyDeviceToHost));
		CQ_CUDA(cudaMemcpy(out->read_nleaf_d, c->d_nleaf_d, n * 4, cudaMemcpyDeviceToHost));
		CQ_CUDA(cudaMemcpy(out->read_leaf_u, c->d_leaf_u, n * cap * 4, cudaMemcpyDeviceToHost));
		CQ_CUDA(cudaMemcpy(out->read_leaf_d, c->d_leaf_d, n * cap * 4, cudaMemcpyDeviceToHost));
		// the reference's std::set iterates in a fixed order; present the sets sorted
		for (uint64_t r = 0; r < n; r++) {
			uint32_t nu = std::min(out->read_nleaf_u[r], cap), nd = std::min(out->read_nleaf_d[r], cap);
			std::sort(out->read_leaf_u + r * cap, out->read_leaf_u + r * cap + nu);
			std::sort(out->read_leaf_d + r * cap, out->read_leaf_d + r * cap + nd);
		}
	}
	return CQ_OK;
}

// reads per pipeline stage of cq_query (CAMMIQ_CHUNK_READS: experiments; rounded to whole tiles of 64 reads)
static uint64_t chunkReads() {
	const char *e = getenv("CAMMIQ_CHUNK_READS");
	const uint64_t v = e ? (uint64_t) atoll(e) : 0;
	return v >= (1u << 12) ? std::min<uint64_t>((v + 63) & ~63ull, 1u << 24) : (1u << 20);
}
static const uint64_t kChunkReads = chunkReads();

template <typename T>
static int ensureHost(T **ptr, size_t *cap, size_t need) {
	if (need <= *cap && *ptr != NULL)
		return CQ_OK;
	if (*ptr) cudaFreeHost(*ptr);
	*ptr = NULL;
	*cap = 0;
	size_t n = std::max<size_t>(need + need / 4, 64);
	CQ_CUDA(cudaHostAlloc((void **) ptr, n * sizeof(T), cudaHostAllocDefault));
	*cap = n;
	return CQ_OK;
}

// One chunk of reads, as the caller holds them, copied host -> device on the copy stream into
// stage b and scanned on the compute stream.  `packed`: the caller's buffer holds 2-bit reads.
static int submitChunkDirect(cq_ctx *c, int mode, bool packed, const uint8_t *bases, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t first, uint64_t n, uint64_t k, int b) {
	int rc;
	uint64_t lo = ~0ull, hi = 0;
	uint32_t max_len = 1;
	if (offsets) {
		for (uint64_t i = first; i < first + n; i++) {
			lo = std::min(lo, offsets[i]);
			hi = std::max(hi, offsets[i] + (packed ? packedBytes(lengths[i]) : lengths[i]));
			max_len = std::max<uint32_t>(max_len, lengths[i]);
		}
	} else {
		for (uint64_t i = first; i < first + n; i++)
			max_len = std::max<uint32_t>(max_len, lengths[i]);
		// fixed stride: only the reads within 255 bytes of the chunk's end can set its extent
		lo = first * stride;
		hi = lo;
		for (uint64_t i = first + n; i-- > first;) {
			hi = std::max(hi, i * stride + (packed ? packedBytes(lengths[i]) : lengths[i]));
			if ((first + n - 1 - i) * stride >= 255)
				break;
		}
	}
	if (hi < lo) hi = lo;
	const uint64_t copy_lo = lo & ~15ull, copy_bytes = hi - copy_lo;
	if ((rc = ensureStage(c, b, &c->d_cbases[b], &c->cap_cbases[b], copy_bytes + 64)) != 0) return rc;
	if ((rc = ensureStage(c, b, &c->d_clengths[b], &c->cap_clengths[b], n)) != 0) return rc;
	if (offsets && (rc = ensureStage(c, b, &c->d_coffsets[b], &c->cap_coffsets[b], n)) != 0) return rc;
	if (k >= (uint64_t) cq_ctx::kStages)
		CQ_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_free[b], 0));
	else if (k == 0)
		CQ_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev[0], 0)); // after earlier work on the compute stream
	if (copy_bytes > 0)
		CQ_CUDA(cudaMemcpyAsync(c->d_cbases[b], bases + copy_lo, copy_bytes, cudaMemcpyHostToDevice, c->copy_stream));
	CQ_CUDA(cudaMemcpyAsync(c->d_clengths[b], lengths + first, n, cudaMemcpyHostToDevice, c->copy_stream));
	if (offsets)
		CQ_CUDA(cudaMemcpyAsync(c->d_coffsets[b], offsets + first, n * 8, cudaMemcpyHostToDevice, c->copy_stream));
	c->timing.h2d_bytes += copy_bytes + n + (offsets ? n * 8 : 0);
	CQ_CUDA(cudaEventRecord(c->ev_copied[b], c->copy_stream));
	CQ_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));
	ReadBatch rb = {c->d_cbases[b] - copy_lo, offsets ? c->d_coffsets[b] : NULL, stride, c->d_clengths[b], n, first, max_len,
		packed, NULL, b};
	if ((rc = launchPack(c, rb)) != 0) return rc;
	if ((rc = launchScan(c, mode, rb)) != 0) return rc;
	CQ_CUDA(cudaEventRecord(c->ev_free[b], c->stream));
	return CQ_OK;
}

// Chunks of reads flow host -> device on the copy stream while earlier chunks are scanned on
// the compute stream.
static int pipelineDirect(cq_ctx *c, int mode, bool packed, const uint8_t *bases, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t n_reads) {
	int rc;
	for (uint64_t first = 0, k = 0; first < n_reads; first += kChunkReads, k++) {
		const uint64_t n = std::min<uint64_t>(kChunkReads, n_reads - first);
		if ((rc = submitChunkDirect(c, mode, packed, bases, offsets, stride, lengths, first, n, k, (int) (k % cq_ctx::kStages))) != 0)
			return rc;
	}
	return CQ_OK;
}

// ASCII reads are packed to 2 bits per base by the worker pool, chunk by chunk, into pinned
// staging; chunk k is packed while chunk k-1 crosses PCIe and chunk k-2 is scanned.
static int pipelineHostPack(cq_ctx *c, int mode, const uint8_t *bases, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t n_reads) {
	int rc;
	if (c->pool == NULL || c->pool->size() != c->pack_threads) {
		delete c->pool;
		c->pool = new WorkerPool(c->pack_threads);
	}
	const AsciiReads in = {bases, offsets, stride, lengths};
	const bool dense = offsets != NULL;
	double pack_ms = 0;
	// The packer is bound by the host's memory bandwidth, the ASCII path by PCIe, so with a page-locked
	// caller buffer some chunks could cross PCIe as they are while the workers pack the others
	// (CAMMIQ_DIRECT_EVERY = n: every n-th chunk).  Measured with n = 8 at equal packer speed: 14.8 ms
	// against 13.0 ms per 10M reads without -- the one 100 MB copy lengthens the pipeline's tail by
	// more than it takes off the packers -- so the default is off.
	int direct_every = getenv("CAMMIQ_DIRECT_EVERY") ? atoi(getenv("CAMMIQ_DIRECT_EVERY")) : 0;
	if (direct_every > 0) {
		cudaPointerAttributes attr;
		if (cudaPointerGetAttributes(&attr, bases) != cudaSuccess || attr.type != cudaMemoryTypeHost) {
			cudaGetLastError();
			direct_every = 0; // pageable memory: an "asynchronous" copy of it would block this thread
		}
	}
	for (uint64_t first = 0, k = 0; first < n_reads; first += kChunkReads, k++) {
		const int b = (int) (k % cq_ctx::kStages);
		const uint64_t n = std::min<uint64_t>(kChunkReads, n_reads - first);
		if (direct_every > 0 && k % (uint64_t) direct_every == (uint64_t) direct_every - 1) {
			if ((rc = submitChunkDirect(c, mode, false, bases, offsets, stride, lengths, first, n, k, b)) != 0) return rc;
			continue;
		}
		// the pinned staging of this slot is free once its previous copy has left the host
		if (k >= (uint64_t) cq_ctx::kStages)
			CQ_CUDA(cudaEventSynchronize(c->ev_copied[b]));
		auto p0 = std::chrono::high_resolution_clock::now();
		const PackedLayout layout = planBatch(*c->pool, lengths, first, n, dense);
		if ((rc = ensureHost(&c->h_pbases[b], &c->cap_h_pbases[b], (size_t) layout.total_bytes + 64)) != 0) return rc;
		if ((rc = ensureHost(&c->h_plengths[b], &c->cap_h_plengths[b], (size_t) n)) != 0) return rc;
		if (dense && (rc = ensureHost(&c->h_poffsets[b], &c->cap_h_poffsets[b], (size_t) n)) != 0) return rc;
		packBatch(*c->pool, in, first, n, layout, c->h_pbases[b], c->h_poffsets[b], c->h_plengths[b]);
		pack_ms += std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - p0).count();
		if ((rc = ensureStage(c, b, &c->d_cbases[b], &c->cap_cbases[b], (size_t) layout.total_bytes + 64)) != 0) return rc;
		if ((rc = ensureStage(c, b, &c->d_clengths[b], &c->cap_clengths[b], (size_t) n)) != 0) return rc;
		if (dense && (rc = ensureStage(c, b, &c->d_coffsets32[b], &c->cap_coffsets32[b], (size_t) n)) != 0) return rc;
		if (k >= (uint64_t) cq_ctx::kStages)
			CQ_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_free[b], 0));
		else if (k == 0)
			CQ_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev[0], 0));
		if (layout.total_bytes > 0)
			CQ_CUDA(cudaMemcpyAsync(c->d_cbases[b], c->h_pbases[b], layout.total_bytes, cudaMemcpyHostToDevice, c->copy_stream));
		CQ_CUDA(cudaMemcpyAsync(c->d_clengths[b], c->h_plengths[b], n, cudaMemcpyHostToDevice, c->copy_stream));
		if (dense)
			CQ_CUDA(cudaMemcpyAsync(c->d_coffsets32[b], c->h_poffsets[b], n * 4, cudaMemcpyHostToDevice, c->copy_stream));
		c->timing.h2d_bytes += layout.total_bytes + n + (dense ? n * 4 : 0);
		CQ_CUDA(cudaEventRecord(c->ev_copied[b], c->copy_stream));
		CQ_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));
		// fixed stride: the kernel addresses read r at (first + r) * stride
		ReadBatch rb = {dense ? c->d_cbases[b] : c->d_cbases[b] - first * layout.stride, NULL, layout.stride, c->d_clengths[b], n,
			first, std::max<uint32_t>(layout.max_len, 1), true, dense ? c->d_coffsets32[b] : NULL, b};
		if ((rc = launchPack(c, rb)) != 0) return rc;
		if ((rc = launchScan(c, mode, rb)) != 0) return rc;
		CQ_CUDA(cudaEventRecord(c->ev_free[b], c->stream));
	}
	c->timing.host_pack_ms = pack_ms;
	c->timing.host_pack_threads = (uint32_t) c->pack_threads;
	return CQ_OK;
}

int cqSubmitHost(cq_ctx *c, int mode, bool packed, const uint8_t *bases, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t n_reads, cq_result *per_read, const char *who) {
	if (c == NULL || !c->has_index)
		return fail(CQ_ESTATE, std::string(who) + ": no index resident (call cq_index_upload first).");
	if (mode != CQ_MODE_P && mode != CQ_MODE_SC)
		return fail(CQ_EINVAL, std::string(who) + ": bad mode.");
	if (n_reads > 0 && (bases == NULL || lengths == NULL))
		return fail(CQ_EINVAL, std::string(who) + ": NULL read buffers.");
	if (n_reads >= (1ull << 36))
		return fail(CQ_EINVAL, std::string(who) + ": more than 2^36 reads in one call.");
	CQ_CUDA(cudaSetDevice(c->device));
	c->want_per_read = per_read != NULL && per_read->read_class != NULL && per_read->read_rid_a != NULL && per_read->read_rid_b != NULL;
	c->want_sets = per_read != NULL && per_read->leaf_cap > 0 && per_read->read_nleaf_u && per_read->read_nleaf_d &&
		per_read->read_leaf_u && per_read->read_leaf_d;
	c->leaf_cap = c->want_sets ? per_read->leaf_cap : 0;
	int rc = prepareOutputs(c, mode, n_reads);
	cudaEvent_t *sev = NULL;
	if (rc == 0) rc = beginStep(c, &sev);
	if (rc != 0) return rc;
	CQ_CUDA(cudaEventRecord(sev[0], c->stream));
	CQ_CUDA(cudaEventRecord(sev[1], c->stream));
	CQ_CUDA(cudaEventRecord(c->ev[0], c->stream));
	c->timing.h2d_bytes = 0;
	c->timing.host_pack_ms = 0;
	c->timing.host_pack_threads = 0;
	if (!packed && c->pack_threads > 0)
		rc = pipelineHostPack(c, mode, bases, offsets, stride, lengths, n_reads);
	else
		rc = pipelineDirect(c, mode, packed, bases, offsets, stride, lengths, n_reads);
	if (rc != 0) return rc;
	CQ_CUDA(cudaEventRecord(sev[2], c->stream));
	CQ_CUDA(cudaEventRecord(sev[3], c->stream));
	CQ_CUDA(cudaEventRecord(c->ev[1], c->stream));
	CQ_CUDA(cudaGetLastError());
	if (per_read != NULL)
		rc = fetchPerRead(c, n_reads, per_read);
	c->want_per_read = c->want_sets = false;
	c->leaf_cap = 0;
	return rc;
}

static int queryHost(cq_ctx *c, int mode, bool packed, const uint8_t *bases, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t n_reads, cq_result *out, const char *who) {
	if (out == NULL)
		return fail(CQ_EINVAL, std::string(who) + ": NULL result.");
	auto t0 = std::chrono::high_resolution_clock::now();
	int rc = cqSubmitHost(c, mode, packed, bases, offsets, stride, lengths, n_reads, out, who);
	if (rc != 0) return rc;
	auto t1 = std::chrono::high_resolution_clock::now();
	rc = cq_fetch(c, mode, out);
	auto t2 = std::chrono::high_resolution_clock::now();
	c->timing.d2h_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
	c->timing.total_ms = std::chrono::duration<double, std::milli>(t2 - t0).count();
	return rc;
}

extern "C" int cq_query(cq_ctx *c, int mode, const uint8_t *bases, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t n_reads, cq_result *out) {
	return queryHost(c, mode, false, bases, offsets, stride, lengths, n_reads, out, "cq_query");
}

extern "C" int cq_query_packed(cq_ctx *c, int mode, const uint8_t *packed, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t n_reads, cq_result *out) {
	return queryHost(c, mode, true, packed, offsets, stride, lengths, n_reads, out, "cq_query_packed");
}

extern "C" int cq_query_submit(cq_ctx *c, int mode, const uint8_t *bases, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t n_reads) {
	return cqSubmitHost(c, mode, false, bases, offsets, stride, lengths, n_reads, NULL, "cq_query_submit");
}

extern "C" int cq_query_submit_packed(cq_ctx *c, int mode, const uint8_t *packed, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t n_reads) {
	return cqSubmitHost(c, mode, true, packed, offsets, stride, lengths, n_reads, NULL, "cq_query_submit_packed");
}

extern "C" int cq_ctx_set_host_packing(cq_ctx *c, int threads) {
	if (c == NULL)
		return fail(CQ_EINVAL, "cq_ctx_set_host_packing: NULL context.");
	c->pack_threads = threads < 0 ? defaultPackThreads() : std::min(threads, 256);
	return CQ_OK;
}

extern "C" int cq_pack_reads(const uint8_t *bases, const uint64_t *offsets, uint64_t stride, const uint8_t *lengths,
		uint64_t n_reads, int threads, uint8_t *packed, uint64_t packed_stride, uint8_t *packed_lengths, uint64_t *n_invalid) {
	if (n_reads > 0 && (bases == NULL || lengths == NULL || packed == NULL || packed_lengths == NULL))
		return fail(CQ_EINVAL, "cq_pack_reads: NULL buffers.");
	WorkerPool pool(std::max(1, std::min(threads, 256)));
	const AsciiReads in = {bases, offsets, stride, lengths};
	PackedLayout layout = planBatch(pool, lengths, 0, n_reads, false);
	if (packed_stride < layout.stride)
		return fail(CQ_EINVAL, "cq_pack_reads: packed_stride is smaller than ceil(longest read / 4).");
	layout.stride = packed_stride;
	const uint64_t bad = packBatch(pool, in, 0, n_reads, layout, packed, NULL, packed_lengths);
	if (n_invalid) *n_invalid = bad;
	return CQ_OK;
}

extern "C" const char *cq_pack_isa(void) { return packIsaName(); }

extern "C" int cq_host_alloc(size_t bytes, void **out) {
	if (out == NULL)
		return fail(CQ_EINVAL, "cq_host_alloc: NULL argument.");
	*out = NULL;
	cudaError_t e = cudaHostAlloc(out, std::max<size_t>(bytes, 1), cudaHostAllocDefault);
	if (e != cudaSuccess)
		return fail(e == cudaErrorMemoryAllocation ? CQ_ENOMEM : CQ_ENODEV,
			std::string("cq_host_alloc: ") + cudaGetErrorString(e));
	return CQ_OK;
}

extern "C" void cq_host_free(void *p) {
	if (p != NULL)
		cudaFreeHost(p);
}

// ILP input assembly on the device (query.cpp:1154-1181).
extern "C" int cq_ilp_inputs(cq_ctx *c, const cq_index *idx, cq_ilp_args *io) {
	if (c == NULL || !c->has_index || idx == NULL || io == NULL)
		return fail(CQ_ESTATE, "cq_ilp_inputs: NULL argument or no index resident.");
	if (io->read_length == 0)
		return fail(CQ_EINVAL, "cq_ilp_inputs: read_length is 0.");
	const FlatIndex &f = idx->flat;
	if (f.u.numLeaves() != c->n_leaves_u || f.d.numLeaves() != c->n_leaves_d)
		return fail(CQ_EINVAL, "cq_ilp_inputs: this index is not the one resident on the context.");
	CQ_CUDA(cudaSetDevice(c->device));
	const size_t G1 = (size_t) c->n_genomes + 1;
	for (int t = 0; t < 2; t++) {
		const DecodedIndex &x = t == 0 ? f.u : f.d;
		const uint64_t n = x.numLeaves();
		double *h_w1 = t == 0 ? io->wcov_u : io->wcov_d1, *h_w2 = t == 0 ? NULL : io->wcov_d2;
		double *h_gw = t == 0 ? io->genome_wcov_u : io->genome_wcov_d;
		uint64_t *h_gr = t == 0 ? io->genome_rcount_u : io->genome_rcount_d;
		if (h_gw) memset(h_gw, 0, G1 * sizeof(double));
		if (h_gr) memset(h_gr, 0, G1 * sizeof(uint64_t));
		if (n == 0)
			continue;
		// leaf fields the scan does not need live on the host; they travel for this call only
		uint16_t *d_uc1 = NULL, *d_uc2 = NULL;
		uint8_t *d_depth = NULL;
		double *d_w1 = NULL, *d_w2 = NULL, *d_gw = NULL;
		unsigned long long *d_gr = NULL;
		int rc = CQ_OK;
		auto release = [&]() {
			cudaFree(d_uc1); cudaFree(d_uc2); cudaFree(d_depth); cudaFree(d_w1); cudaFree(d_w2); cudaFree(d_gw); cudaFree(d_gr);
		};
#define CQ_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { release(); return fail(CQ_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
		CQ_TRY(cudaMalloc((void **) &d_uc1, n * 2));
		CQ_TRY(cudaMalloc((void **) &d_depth, n));
		CQ_TRY(cudaMalloc((void **) &d_w1, n * 8));
		CQ_TRY(cudaMalloc((void **) &d_gw, G1 * 8));
		CQ_TRY(cudaMalloc((void **) &d_gr, G1 * 8));
		CQ_TRY(cudaMemcpyAsync(d_uc1, x.ucount1.data(), n * 2, cudaMemcpyHostToDevice, c->stream));
		CQ_TRY(cudaMemcpyAsync(d_depth, x.depth.data(), n, cudaMemcpyHostToDevice, c->stream));
		CQ_TRY(cudaMemsetAsync(d_gw, 0, G1 * 8, c->stream));
		CQ_TRY(cudaMemsetAsync(d_gr, 0, G1 * 8, c->stream));
		if (t == 1) {
			CQ_TRY(cudaMalloc((void **) &d_uc2, n * 2));
			CQ_TRY(cudaMalloc((void **) &d_w2, n * 8));
			CQ_TRY(cudaMemcpyAsync(d_uc2, x.ucount2.data(), n * 2, cudaMemcpyHostToDevice, c->stream));
		}
		IlpParams q;
		memset(&q, 0, sizeof(q));
		// the resident leaf records already hold the genome ids: u32 per U leaf, {u32, u32} per D leaf
		q.ref1 = t == 0 ? c->d_leaf_u_ref : reinterpret_cast<const uint32_t *>(c->d_leaf_d_ref);
		q.ref2 = t == 0 ? NULL : reinterpret_cast<const uint32_t *>(c->d_leaf_d_ref) + 1;
		q.ref_stride = t == 0 ? 1 : 2;
		q.ucount1 = d_uc1;
		q.ucount2 = d_uc2;
		q.depth = d_depth;
		q.rcount = t == 0 ? c->d_rcount_u : c->d_rcount_d;
		q.n = n;
		q.n_genomes = c->n_genomes;
		q.rl = io->read_length;
		q.one_minus_e = 1 - io->erate;
		q.wcov1 = d_w1;
		q.wcov2 = d_w2;
		q.genome_wcov = d_gw;
		q.genome_rcount = d_gr;
		ilp_inputs_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, c->stream>>>(q);
		c->timing.kernel_launches++;
		CQ_TRY(cudaGetLastError());
		if (h_w1) CQ_TRY(cudaMemcpyAsync(h_w1, d_w1, n * 8, cudaMemcpyDeviceToHost, c->stream));
		if (h_w2) CQ_TRY(cudaMemcpyAsync(h_w2, d_w2, n * 8, cudaMemcpyDeviceToHost, c->stream));
		if (h_gw) CQ_TRY(cudaMemcpyAsync(h_gw, d_gw, G1 * 8, cudaMemcpyDeviceToHost, c->stream));
		if (h_gr) CQ_TRY(cudaMemcpyAsync(h_gr, d_gr, G1 * 8, cudaMemcpyDeviceToHost, c->stream));
		CQ_TRY(cudaStreamSynchronize(c->stream));
#undef CQ_TRY
		release();
		(void) rc;
	}
	return CQ_OK;
}

extern "C" int cq_get_device_counters(cq_ctx *c, cq_device_counters *out) {
	if (c == NULL || !c->has_index || out == NULL)
		return fail(CQ_ESTATE, "cq_get_device_counters: no index resident.");
	out->d_counts = c->d_counts;
	out->n_counts = 2 * ((uint64_t) c->n_genomes + 1) + 4;
	out->d_rcount_u = c->d_rcount_u;
	out->n_rcount_u = c->n_leaves_u;
	out->d_rcount_d = c->d_rcount_d;
	out->n_rcount_d = c->n_leaves_d;
	return CQ_OK;
}

extern "C" int cq_swap_accumulators(cq_ctx *c, cq_device_counters *previous) {
	if (c == NULL || !c->has_index)
		return fail(CQ_ESTATE, "cq_swap_accumulators: no index resident.");
	CQ_CUDA(cudaSetDevice(c->device));
	const size_t n_counts = 2 * ((size_t) c->n_genomes + 1) + 4;
	if (c->d_counts_alt == NULL) {
		const size_t nu = std::max<size_t>(c->n_leaves_u, 1), nd = std::max<size_t>(c->n_leaves_d, 1);
		CQ_CUDA(cudaMalloc((void **) &c->d_counts_alt, n_counts * sizeof(unsigned long long)));
		CQ_CUDA(cudaMalloc((void **) &c->d_rcount_u_alt, nu * 4));
		CQ_CUDA(cudaMalloc((void **) &c->d_rcount_d_alt, nd * 4));
		CQ_CUDA(cudaMemsetAsync(c->d_counts_alt, 0, n_counts * sizeof(unsigned long long), c->stream));
		CQ_CUDA(cudaMemsetAsync(c->d_rcount_u_alt, 0, nu * 4, c->stream));
		CQ_CUDA(cudaMemsetAsync(c->d_rcount_d_alt, 0, nd * 4, c->stream));
	}
	if (previous) {
		previous->d_counts = c->d_counts;
		previous->n_counts = n_counts;
		previous->d_rcount_u = c->d_rcount_u;
		previous->n_rcount_u = c->n_leaves_u;
		previous->d_rcount_d = c->d_rcount_d;
		previous->n_rcount_d = c->n_leaves_d;
	}
	std::swap(c->d_counts, c->d_counts_alt);
	std::swap(c->d_rcount_u, c->d_rcount_u_alt);
	std::swap(c->d_rcount_d, c->d_rcount_d_alt);
	return CQ_OK;
}

extern "C" int cq_get_stream(cq_ctx *c, void **stream) {
	if (c == NULL || stream == NULL)
		return fail(CQ_EINVAL, "cq_get_stream: NULL argument.");
	*stream = (void *) c->stream;
	return CQ_OK;
}

extern "C" int cq_get_timing(cq_ctx *c, cq_timing *out) {
	if (c == NULL || out == NULL)
		return fail(CQ_EINVAL, "cq_get_timing: NULL argument.");
	CQ_CUDA(cudaSetDevice(c->device));
	CQ_CUDA(cudaStreamSynchronize(c->stream));
	float ms = 0;
	if (cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]) == cudaSuccess) c->timing.h2d_ms = ms;
	cudaGetLastError();
	int rc = foldStepEvents(c);
	if (rc != 0) return rc;
	if (c->d_probe_count) {
		unsigned long long pc[5] = {0, 0, 0, 0, 0};
		CQ_CUDA(cudaMemcpy(pc, c->d_probe_count, 40, cudaMemcpyDeviceToHost));
		c->timing.sieve_loads = pc[4];
		c->timing.probes = pc[0];
		c->timing.bucket_hits = pc[1];
		c->timing.leaf_hits = pc[2];
		c->timing.chained_loads = pc[3];
	}
	*out = c->timing;
	return CQ_OK;
}

extern "C" int cq_timing_reset(cq_ctx *c) {
	if (c == NULL)
		return fail(CQ_EINVAL, "cq_timing_reset: NULL context.");
	int rc = foldStepEvents(c);
	if (rc != 0) return rc;
	c->timing.pack_ms_sum = c->timing.scan_ms_sum = c->timing.reduce_ms_sum = 0;
	c->timing.steps = 0;
	return CQ_OK;
}

extern "C" int cq_bench_random_sectors(cq_ctx *c, uint64_t n_probes, int iters, double *gsectors_per_s) {
	if (c == NULL || !c->has_index || gsectors_per_s == NULL || iters < 1)
		return fail(CQ_ESTATE, "cq_bench_random_sectors: bad argument or no index resident.");
	CQ_CUDA(cudaSetDevice(c->device));
	unsigned long long *sink = c->d_probe_count;
	cudaEvent_t a, b;
	CQ_CUDA(cudaEventCreate(&a));
	CQ_CUDA(cudaEventCreate(&b));
	int grid = c->n_sms * 8;
	random_sector_kernel<<<grid, 256, 0, c->stream>>>(c->d_table, c->table_mask, n_probes, 12345, sink);
	CQ_CUDA(cudaEventRecord(a, c->stream));
	for (int i = 0; i < iters; i++)
		random_sector_kernel<<<grid, 256, 0, c->stream>>>(c->d_table, c->table_mask, n_probes, 777 + 1000003ull * i, sink);
	CQ_CUDA(cudaEventRecord(b, c->stream));
	CQ_CUDA(cudaStreamSynchronize(c->stream));
	float ms = 0;
	CQ_CUDA(cudaEventElapsedTime(&ms, a, b));
	cudaEventDestroy(a);
	cudaEventDestroy(b);
	*gsectors_per_s = (double) n_probes * iters / (ms * 1e-3) * 1e-9;
	c->timing.kernel_launches += iters + 1;
	return CQ_OK;
}

extern "C" int cq_bench_random_gather(cq_ctx *c, uint64_t region_bytes, int access_bytes, uint64_t n_probes,
		int iters, int persist, double *gaccesses_per_s) {
	if (c == NULL || gaccesses_per_s == NULL || iters < 1 || region_bytes < 4096 ||
		(region_bytes & (region_bytes - 1)) != 0 ||
		(access_bytes != 4 && access_bytes != 8 && access_bytes != 16 && access_bytes != 32))
		return fail(CQ_EINVAL, "cq_bench_random_gather: bad argument.");
	CQ_CUDA(cudaSetDevice(c->device));
	uint8_t *region = NULL;
	unsigned long long *sink = NULL;
	CQ_CUDA(cudaMalloc((void **) &region, region_bytes));
	CQ_CUDA(cudaMalloc((void **) &sink, 8));
	CQ_CUDA(cudaMemsetAsync(region, 1, region_bytes, c->stream));
	if (persist) {
		cudaDeviceProp prop;
		CQ_CUDA(cudaGetDeviceProperties(&prop, c->device));
		CQ_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, prop.persistingL2CacheMaxSize));
		cudaStreamAttrValue attr;
		memset(&attr, 0, sizeof(attr));
		attr.accessPolicyWindow.base_ptr = region;
		attr.accessPolicyWindow.num_bytes = std::min<size_t>(region_bytes, prop.accessPolicyMaxWindowSize);
		attr.accessPolicyWindow.hitRatio = std::min(1.0f, (float) prop.persistingL2CacheMaxSize / (float) region_bytes);
		attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
		attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
		CQ_CUDA(cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
	}
	const uint64_t mask = (region_bytes - 1) & ~(uint64_t) (access_bytes - 1);
	// CAMMIQ_GATHER_SMEM (bytes per block) / CAMMIQ_GATHER_BLOCKS (per SM): the same gathers with part
	// of the SM's unified L1 / shared memory taken away -- how much L1 the in-flight loads need
	const size_t dsm = getenv("CAMMIQ_GATHER_SMEM") ? (size_t) atol(getenv("CAMMIQ_GATHER_SMEM")) : 0;
	const int per_sm = getenv("CAMMIQ_GATHER_BLOCKS") ? std::max(1, atoi(getenv("CAMMIQ_GATHER_BLOCKS"))) : 8;
	const int grid = c->n_sms * per_sm;
	if (getenv("CAMMIQ_GATHER_CARVEOUT")) {
		const int pct = atoi(getenv("CAMMIQ_GATHER_CARVEOUT"));
		cudaFuncSetAttribute((const void *) random_gather_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
		cudaFuncSetAttribute((const void *) random_gather_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
		cudaFuncSetAttribute((const void *) random_gather_kernel<16>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
		cudaFuncSetAttribute((const void *) random_gather_kernel<32>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
	}
	if (dsm > 0) {
		cudaFuncSetAttribute((const void *) random_gather_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dsm);
		cudaFuncSetAttribute((const void *) random_gather_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dsm);
		cudaFuncSetAttribute((const void *) random_gather_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dsm);
		cudaFuncSetAttribute((const void *) random_gather_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dsm);
	}
	auto launch = [&](uint64_t seed) {
		switch (access_bytes) {
		case 4: random_gather_kernel<4><<<grid, 256, dsm, c->stream>>>(region, mask, n_probes, seed, sink); break;
		case 8: random_gather_kernel<8><<<grid, 256, dsm, c->stream>>>(region, mask, n_probes, seed, sink); break;
		case 16: random_gather_kernel<16><<<grid, 256, dsm, c->stream>>>(region, mask, n_probes, seed, sink); break;
		default: random_gather_kernel<32><<<grid, 256, dsm, c->stream>>>(region, mask, n_probes, seed, sink); break;
		}
	};
	launch(1);
	launch(2);
	cudaEvent_t a, b;
	CQ_CUDA(cudaEventCreate(&a));
	CQ_CUDA(cudaEventCreate(&b));
	CQ_CUDA(cudaEventRecord(a, c->stream));
	for (int i = 0; i < iters; i++)
		launch(1000 + 7919ull * i);
	CQ_CUDA(cudaEventRecord(b, c->stream));
	CQ_CUDA(cudaStreamSynchronize(c->stream));
	float ms = 0;
	CQ_CUDA(cudaEventElapsedTime(&ms, a, b));
	cudaEventDestroy(a);
	cudaEventDestroy(b);
	if (persist) {
		cudaStreamAttrValue attr;
		memset(&attr, 0, sizeof(attr));
		attr.accessPolicyWindow.num_bytes = 0;
		cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
		cudaCtxResetPersistingL2Cache();
	}
	cudaFree(region);
	cudaFree(sink);
	*gaccesses_per_s = (double) n_probes * iters / (ms * 1e-3) * 1e-9;
	c->timing.kernel_launches += iters + 2;
	return CQ_OK;
}

extern "C" int cq_get_device_info(cq_ctx *c, cq_device_info *out) {
	if (c == NULL || out == NULL)
		return fail(CQ_EINVAL, "cq_get_device_info: NULL argument.");
	cudaDeviceProp prop;
	CQ_CUDA(cudaGetDeviceProperties(&prop, c->device));
	memset(out, 0, sizeof(*out));
	strncpy(out->name, prop.name, sizeof(out->name) - 1);
	out->sm_count = prop.multiProcessorCount;
	out->cc_major = prop.major;
	out->cc_minor = prop.minor;
	out->l2_bytes = (uint64_t) prop.l2CacheSize;
	out->persisting_l2_max_bytes = (uint64_t) prop.persistingL2CacheMaxSize;
	out->access_policy_max_window_bytes = (uint64_t) prop.accessPolicyMaxWindowSize;
	out->global_mem_bytes = (uint64_t) prop.totalGlobalMem;
	out->sm_clock_khz = prop.clockRate;
	out->mem_clock_khz = prop.memoryClockRate;
	out->mem_bus_bits = prop.memoryBusWidth;
	return CQ_OK;
}
