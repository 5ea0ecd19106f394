// cq_multi_*: the hot path over several GPUs of one box behind the C ABI (SURVEY.md section 8b/8e,
// north_star: "the index is replicated on each GPU, reads are sharded across the GPUs, and only
// the count vectors are combined, with a single NCCL reduce over NVLink").
//
//   cq_multi_create    one cq_ctx per device + one NCCL communicator over them (ncclCommInitAll)
//   cq_multi_upload    the same flattened index to every device
//   cq_multi_query     reads cut into contiguous shards balanced by bases; one host thread per
//                      device runs the ordinary cq_query pipeline on its shard (pack -> copy -> scan,
//                      results left in that device's accumulators); then ONE grouped ncclReduce
//                      (sum) of the counter block -- and, in mode P, of the two per-leaf rcount
//                      arrays -- into totals on device 0, copied to the caller's buffers.
//                      query64_sc's pair map is small and sparse: every device folds its records
//                      into (pair, count) entries and the host merges those.
//
// NCCL is reached through dlopen("libnccl.so.2"): the library must stay loadable next to a
// framework that bundles its own NCCL (torch), and a second copy linked in here would clash with it.
// Integer sums: the result does not depend on the number of devices (exact-equality tests).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <dlfcn.h>

#include <cuda_runtime.h>
#include <nccl.h>

#include "capi_internal.hpp"

namespace {

struct NcclApi {
	void *handle;
	ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *);
	ncclResult_t (*CommDestroy)(ncclComm_t);
	ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*GroupStart)();
	ncclResult_t (*GroupEnd)();
	const char *(*GetErrorString)(ncclResult_t);
	int (*GetVersion)(int *);
};

// resolved once per process; NULL handle = not available
const NcclApi *ncclApi(std::string &err) {
	static NcclApi api;
	static bool tried = false;
	static std::string load_err;
	if (!tried) {
		tried = true;
		memset(&api, 0, sizeof(api));
		// an NCCL the process already holds (torch's bundled one) wins: same SONAME, same handle
		void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
		if (h == NULL) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (h == NULL) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
		if (h == NULL) {
			load_err = std::string("libnccl.so.2 cannot be loaded (") + dlerror() + ")";
		} else {
			api.handle = h;
			bool ok = true;
#define CQ_SYM(field, name)                                       \
	do {                                                          \
		*(void **) (&api.field) = dlsym(h, name);                 \
		if (api.field == NULL) { ok = false; load_err = std::string("libnccl.so.2 lacks ") + name; } \
	} while (0)
			CQ_SYM(CommInitAll, "ncclCommInitAll");
			CQ_SYM(CommDestroy, "ncclCommDestroy");
			CQ_SYM(Reduce, "ncclReduce");
			CQ_SYM(GroupStart, "ncclGroupStart");
			CQ_SYM(GroupEnd, "ncclGroupEnd");
			CQ_SYM(GetErrorString, "ncclGetErrorString");
			CQ_SYM(GetVersion, "ncclGetVersion");
#undef CQ_SYM
			if (!ok) api.handle = NULL;
		}
	}
	if (api.handle == NULL) {
		err = load_err;
		return NULL;
	}
	return &api;
}

} // namespace

struct cq_multi {
	std::vector<cq_ctx *> ctx;
	std::vector<int> devices;
	std::vector<ncclComm_t> comms;
	const NcclApi *nccl = NULL;
	int nccl_version = 0;
	// totals on device 0 (the reduce must not fold into device 0's own accumulators: they keep
	// accumulating across calls like every other device's)
	unsigned long long *d_tot_counts = NULL;
	uint32_t *d_tot_rcount_u = NULL, *d_tot_rcount_d = NULL;
	bool has_index = false;
	uint64_t last_shard_reads[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	double last_reduce_ms = 0;
	cudaEvent_t ev[2] = {NULL, NULL};
};

#define CQ_NCCL(m, call)                                                                                  \
	do {                                                                                                   \
		ncclResult_t r_ = (call);                                                                          \
		if (r_ != ncclSuccess)                                                                             \
			return cqFail(CQ_ECUDA, std::string(#call) + ": " + (m)->nccl->GetErrorString(r_));            \
	} while (0)

static void freeTotals(cq_multi *m) {
	if (m->ctx.empty())
		return;
	cudaSetDevice(m->devices[0]);
	cudaFree(m->d_tot_counts);
	cudaFree(m->d_tot_rcount_u);
	cudaFree(m->d_tot_rcount_d);
	m->d_tot_counts = NULL;
	m->d_tot_rcount_u = m->d_tot_rcount_d = NULL;
}

extern "C" void cq_multi_destroy(cq_multi *m) {
	if (m == NULL)
		return;
	freeTotals(m);
	for (size_t i = 0; i < m->comms.size(); i++)
		if (m->comms[i] != NULL && m->nccl != NULL)
			m->nccl->CommDestroy(m->comms[i]);
	if (!m->ctx.empty()) {
		cudaSetDevice(m->devices[0]);
		for (int i = 0; i < 2; i++)
			if (m->ev[i]) cudaEventDestroy(m->ev[i]);
	}
	for (size_t i = 0; i < m->ctx.size(); i++)
		cq_ctx_destroy(m->ctx[i]);
	delete m;
}

extern "C" int cq_multi_create(int n_gpus, const int *devices, cq_multi **out) {
	if (out == NULL || n_gpus < 1 || n_gpus > 8)
		return cqFail(CQ_EINVAL, "cq_multi_create: n_gpus must be 1..8.");
	*out = NULL;
	int have = 0;
	cudaError_t e = cudaGetDeviceCount(&have);
	if (e != cudaSuccess || have == 0)
		return cqFail(CQ_ENODEV, std::string("no usable CUDA device (") +
			(e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") + "); there is no CPU fallback.");
	cq_multi *m = new (std::nothrow) cq_multi();
	if (m == NULL)
		return cqFail(CQ_ENOMEM, "cq_multi_create: out of memory.");
	for (int i = 0; i < n_gpus; i++) {
		const int dev = devices ? devices[i] : i;
		if (dev < 0 || dev >= have || std::find(m->devices.begin(), m->devices.end(), dev) != m->devices.end()) {
			cq_multi_destroy(m);
			return cqFail(CQ_EINVAL, "cq_multi_create: device ordinal out of range or listed twice (" + std::to_string(have) +
				" device(s) visible).");
		}
		m->devices.push_back(dev);
	}
	// the contexts come up in parallel (a cold CUDA context takes a second on this hardware)
	m->ctx.assign((size_t) n_gpus, (cq_ctx *) NULL);
	std::vector<int> rcs((size_t) n_gpus, 0);
	std::vector<std::string> errs((size_t) n_gpus);
	std::vector<std::thread> pool;
	for (int i = 0; i < n_gpus; i++)
		pool.emplace_back([&, i]() {
			rcs[(size_t) i] = cq_ctx_create(m->devices[(size_t) i], NULL, &m->ctx[(size_t) i]);
			if (rcs[(size_t) i] != 0) errs[(size_t) i] = cq_last_error();
		});
	for (auto &t : pool) t.join();
	for (int i = 0; i < n_gpus; i++)
		if (rcs[(size_t) i] != 0) {
			const int rc = rcs[(size_t) i];
			const std::string msg = errs[(size_t) i];
			cq_multi_destroy(m);
			return cqFail(rc, msg);
		}
	// several contexts pack on the same host: share its cores unless the caller decided otherwise
	if (getenv("CAMMIQ_PACK_THREADS") == NULL && n_gpus > 1) {
		const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
		const int per = (int) std::min(16u, hw / (unsigned) n_gpus);
		for (int i = 0; i < n_gpus; i++)
			cq_ctx_set_host_packing(m->ctx[(size_t) i], per >= 2 ? per : 0);
	}
	if (n_gpus > 1) {
		std::string err;
		m->nccl = ncclApi(err);
		if (m->nccl == NULL) {
			cq_multi_destroy(m);
			return cqFail(CQ_ENODEV, "cq_multi_create: " + err + "; the multi-GPU reduce needs NCCL.");
		}
		m->nccl->GetVersion(&m->nccl_version);
		m->comms.assign((size_t) n_gpus, (ncclComm_t) NULL);
		ncclResult_t r = m->nccl->CommInitAll(m->comms.data(), n_gpus, m->devices.data());
		if (r != ncclSuccess) {
			const std::string msg = std::string("ncclCommInitAll: ") + m->nccl->GetErrorString(r);
			m->comms.clear();
			cq_multi_destroy(m);
			return cqFail(CQ_ECUDA, msg);
		}
	}
	cudaSetDevice(m->devices[0]);
	for (int i = 0; i < 2; i++)
		cudaEventCreate(&m->ev[i]);
	*out = m;
	return CQ_OK;
}

extern "C" int cq_multi_n_gpus(const cq_multi *m) { return m ? (int) m->ctx.size() : 0; }

extern "C" int cq_multi_ctx(cq_multi *m, int i, cq_ctx **out) {
	if (m == NULL || out == NULL || i < 0 || i >= (int) m->ctx.size())
		return cqFail(CQ_EINVAL, "cq_multi_ctx: bad argument.");
	*out = m->ctx[(size_t) i];
	return CQ_OK;
}

// every device runs `fn(i)` on a host thread of its own; the first failure is reported
template <class F>
static int onEveryDevice(cq_multi *m, F fn) {
	const size_t n = m->ctx.size();
	std::vector<int> rcs(n, 0);
	std::vector<std::string> errs(n);
	if (n == 1) {
		rcs[0] = fn(0);
		return rcs[0];
	}
	std::vector<std::thread> pool;
	for (size_t i = 0; i < n; i++)
		pool.emplace_back([&, i]() {
			rcs[i] = fn((int) i);
			if (rcs[i] != 0) errs[i] = cq_last_error(); // the error text is per thread: carry it over
		});
	for (auto &t : pool) t.join();
	for (size_t i = 0; i < n; i++)
		if (rcs[i] != 0)
			return cqFail(rcs[i], "device " + std::to_string(m->devices[i]) + ": " + errs[i]);
	return CQ_OK;
}

extern "C" int cq_multi_upload(cq_multi *m, const cq_index *idx, uint32_t n_genomes) {
	if (m == NULL || idx == NULL)
		return cqFail(CQ_EINVAL, "cq_multi_upload: NULL argument.");
	m->has_index = false;
	int rc = onEveryDevice(m, [&](int i) { return cq_index_upload(m->ctx[(size_t) i], idx, n_genomes); });
	if (rc != 0) return rc;
	freeTotals(m);
	if (m->ctx.size() > 1) {
		cq_ctx *c0 = m->ctx[0];
		CQ_CUDA(cudaSetDevice(m->devices[0]));
		const size_t ncnt = 2 * ((size_t) c0->n_genomes + 1) + 4;
		CQ_CUDA(cudaMalloc((void **) &m->d_tot_counts, ncnt * 8));
		CQ_CUDA(cudaMalloc((void **) &m->d_tot_rcount_u, std::max<size_t>(c0->n_leaves_u, 1) * 4));
		CQ_CUDA(cudaMalloc((void **) &m->d_tot_rcount_d, std::max<size_t>(c0->n_leaves_d, 1) * 4));
	}
	m->has_index = true;
	return CQ_OK;
}

extern "C" int cq_multi_reset(cq_multi *m) {
	if (m == NULL || !m->has_index)
		return cqFail(CQ_ESTATE, "cq_multi_reset: no index resident.");
	for (size_t i = 0; i < m->ctx.size(); i++) {
		int rc = cq_reset(m->ctx[i]);
		if (rc != 0) return rc;
	}
	return CQ_OK;
}

// totals of all devices -> host buffers
static int multiFetch(cq_multi *m, int mode, cq_result *out) {
	const size_t n = m->ctx.size();
	cq_ctx *c0 = m->ctx[0];
	if (n == 1)
		return cqFetchFrom(c0, mode, c0->d_counts, c0->d_rcount_u, c0->d_rcount_d, out, true);
	// ONE grouped NCCL sum-reduce into device 0: the counter block and, in mode P, the per-leaf counts
	const size_t ncnt = 2 * ((size_t) c0->n_genomes + 1) + 4;
	CQ_CUDA(cudaSetDevice(m->devices[0]));
	CQ_CUDA(cudaEventRecord(m->ev[0], c0->stream));
	CQ_NCCL(m, m->nccl->GroupStart());
	for (size_t i = 0; i < n; i++) {
		cq_ctx *c = m->ctx[i];
		CQ_NCCL(m, m->nccl->Reduce(c->d_counts, m->d_tot_counts, ncnt, ncclUint64, ncclSum, 0, m->comms[i], c->stream));
		if (mode == CQ_MODE_P && c0->n_leaves_u)
			CQ_NCCL(m, m->nccl->Reduce(c->d_rcount_u, m->d_tot_rcount_u, c0->n_leaves_u, ncclUint32, ncclSum, 0, m->comms[i], c->stream));
		if (mode == CQ_MODE_P && c0->n_leaves_d)
			CQ_NCCL(m, m->nccl->Reduce(c->d_rcount_d, m->d_tot_rcount_d, c0->n_leaves_d, ncclUint32, ncclSum, 0, m->comms[i], c->stream));
	}
	CQ_NCCL(m, m->nccl->GroupEnd());
	CQ_CUDA(cudaSetDevice(m->devices[0]));
	CQ_CUDA(cudaEventRecord(m->ev[1], c0->stream));
	int rc = cqFetchFrom(c0, mode, m->d_tot_counts, m->d_tot_rcount_u, m->d_tot_rcount_d, out, false);
	if (rc != 0) return rc;
	float ms = 0;
	if (cudaEventElapsedTime(&ms, m->ev[0], m->ev[1]) == cudaSuccess) m->last_reduce_ms = ms;
	for (size_t i = 1; i < n; i++) { // the peers' part of the collective is done once their streams drain
		CQ_CUDA(cudaSetDevice(m->devices[i]));
		CQ_CUDA(cudaStreamSynchronize(m->ctx[i]->stream));
	}
	if (mode == CQ_MODE_SC) {
		std::map<std::pair<uint32_t, uint32_t>, uint64_t> merged;
		for (size_t i = 0; i < n; i++) {
			std::vector<cq_pair_count> part;
			if ((rc = cqCollectPairs(m->ctx[i], part)) != 0) return rc;
			for (size_t k = 0; k < part.size(); k++)
				merged[std::make_pair(part[k].a, part[k].b)] += part[k].count;
		}
		uint64_t k = 0;
		for (auto it = merged.begin(); it != merged.end(); ++it, ++k)
			if (out->pairs && k < out->pairs_cap) {
				out->pairs[k].a = it->first.first;
				out->pairs[k].b = it->first.second;
				out->pairs[k].count = it->second;
			}
		out->n_pairs = merged.size();
		if (out->pairs && merged.size() > out->pairs_cap)
			return cqFail(CQ_EINVAL, "cq_multi_query: pairs_cap too small for the pair map.");
	}
	return CQ_OK;
}

static int multiQuery(cq_multi *m, int mode, bool packed, const uint8_t *bases, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t n_reads, cq_result *out, const char *who) {
	if (m == NULL || !m->has_index)
		return cqFail(CQ_ESTATE, std::string(who) + ": no index resident (call cq_multi_upload first).");
	if (out == NULL || (mode != CQ_MODE_P && mode != CQ_MODE_SC))
		return cqFail(CQ_EINVAL, std::string(who) + ": NULL result or bad mode.");
	if (n_reads > 0 && (bases == NULL || lengths == NULL))
		return cqFail(CQ_EINVAL, std::string(who) + ": NULL read buffers.");
	const size_t n = m->ctx.size();
	// contiguous shards with (nearly) equal numbers of bases
	std::vector<uint64_t> cut(n + 1, 0);
	cut[n] = n_reads;
	if (n > 1) {
		uint64_t total = 0;
		for (uint64_t i = 0; i < n_reads; i++)
			total += lengths[i];
		uint64_t acc = 0;
		size_t next = 1;
		for (uint64_t i = 0; i < n_reads && next < n; i++) {
			acc += lengths[i];
			while (next < n && acc * n >= total * next)
				cut[next++] = i + 1;
		}
		for (; next < n; next++)
			cut[next] = n_reads;
	}
	const bool per_read = out->read_class != NULL || out->read_nleaf_u != NULL;
	int rc = onEveryDevice(m, [&](int i) {
		const uint64_t lo = cut[(size_t) i], hi = cut[(size_t) i + 1];
		m->last_shard_reads[i] = hi - lo;
		cq_result part;
		memset(&part, 0, sizeof(part));
		if (per_read) { // this shard's slice of the caller's per-read buffers
			part.read_class = out->read_class ? out->read_class + lo : NULL;
			part.read_rid_a = out->read_rid_a ? out->read_rid_a + lo : NULL;
			part.read_rid_b = out->read_rid_b ? out->read_rid_b + lo : NULL;
			part.leaf_cap = out->leaf_cap;
			part.read_nleaf_u = out->read_nleaf_u ? out->read_nleaf_u + lo : NULL;
			part.read_nleaf_d = out->read_nleaf_d ? out->read_nleaf_d + lo : NULL;
			part.read_leaf_u = out->read_leaf_u ? out->read_leaf_u + lo * out->leaf_cap : NULL;
			part.read_leaf_d = out->read_leaf_d ? out->read_leaf_d + lo * out->leaf_cap : NULL;
		}
		// offsets are absolute positions in `bases`; with a fixed stride the shard starts lo reads in
		const uint8_t *b = offsets ? bases : bases + lo * stride;
		int r = cqSubmitHost(m->ctx[(size_t) i], mode, packed, b, offsets ? offsets + lo : NULL, stride, lengths + lo, hi - lo,
			per_read ? &part : NULL, who);
		if (r == 0) r = cq_sync(m->ctx[(size_t) i]);
		return r;
	});
	if (rc != 0) return rc;
	return multiFetch(m, mode, out);
}

extern "C" int cq_multi_query(cq_multi *m, int mode, const uint8_t *bases, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t n_reads, cq_result *out) {
	return multiQuery(m, mode, false, bases, offsets, stride, lengths, n_reads, out, "cq_multi_query");
}

extern "C" int cq_multi_query_packed(cq_multi *m, int mode, const uint8_t *packed, const uint64_t *offsets, uint64_t stride,
		const uint8_t *lengths, uint64_t n_reads, cq_result *out) {
	return multiQuery(m, mode, true, packed, offsets, stride, lengths, n_reads, out, "cq_multi_query_packed");
}

extern "C" int cq_multi_get_info(const cq_multi *m, cq_multi_info *out) {
	if (m == NULL || out == NULL)
		return cqFail(CQ_EINVAL, "cq_multi_get_info: NULL argument.");
	memset(out, 0, sizeof(*out));
	out->n_gpus = (int) m->ctx.size();
	out->nccl_version = m->nccl_version;
	out->reduce_ms = m->last_reduce_ms;
	for (size_t i = 0; i < m->ctx.size() && i < 8; i++) {
		out->devices[i] = m->devices[i];
		out->shard_reads[i] = m->last_shard_reads[i];
	}
	return CQ_OK;
}
