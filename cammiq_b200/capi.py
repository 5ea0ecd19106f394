"""ctypes binding of libcammiq_gpu.so -- mirrors include/cammiq_gpu.h one to one.

Host-side mirror of the reference's query interface for this path:

    Index(path_u, path_d)            ~ FqReader::loadIdx_p            (query.cpp:109-123)
    Context(device).upload(index, G) ~ index resident for the scan
    Context.query(mode, reads...)    ~ query64_p / query64mt_p / query64_sc (query.cpp:458-1080)
    Context.reset()                  ~ resetCounters(_sc)             (query.cpp:1820-1858)

Errors raise CammiqError carrying the ABI's code and cq_last_error() text; there is no CPU
path behind any of these calls.
"""
import ctypes as C
import os

import numpy as np

MODE_P, MODE_SC = 0, 1
TABLE_U, TABLE_D = 0, 1
CLASS_UNLABELED, CLASS_CONFLICT, CLASS_U, CLASS_D_PAIR, CLASS_UD, CLASS_D_INTER = range(6)
ABI_VERSION = 3

_HERE = os.path.dirname(os.path.abspath(__file__))


def library_path():
    # CAMMIQ_LIB: experiments with alternative builds of the same ABI
    return os.environ.get("CAMMIQ_LIB") or os.path.join(_HERE, "libcammiq_gpu.so")


class CammiqError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("cammiq_gpu error %d: %s" % (code, msg))
        self.code = code


class IndexInfo(C.Structure):
    _fields_ = [("hash_len", C.c_uint32), ("n_leaves_u", C.c_uint64), ("n_leaves_d", C.c_uint64),
                ("n_buckets_u", C.c_uint64), ("n_buckets_d", C.c_uint64), ("n_keys", C.c_uint64),
                ("n_table_buckets", C.c_uint64), ("n_nodes_u", C.c_uint64), ("n_nodes_d", C.c_uint64),
                ("n_cnodes_u", C.c_uint64), ("n_cnodes_d", C.c_uint64),
                ("max_ref_id", C.c_uint32), ("filter_bytes", C.c_uint64), ("device_bytes", C.c_uint64),
                ("decode_ms", C.c_double), ("flatten_ms", C.c_double)]


class LeafView(C.Structure):
    _fields_ = [("n", C.c_uint64), ("ref_id1", C.c_void_p), ("ref_id2", C.c_void_p),
                ("ucount1", C.c_void_p), ("ucount2", C.c_void_p), ("depth", C.c_void_p)]


class PairCount(C.Structure):
    _fields_ = [("a", C.c_uint32), ("b", C.c_uint32), ("count", C.c_uint64)]


class Result(C.Structure):
    _fields_ = [("cnt_u", C.c_void_p), ("cnt_d", C.c_void_p), ("rcount_u", C.c_void_p),
                ("rcount_d", C.c_void_p), ("pairs", C.c_void_p), ("pairs_cap", C.c_uint64),
                ("read_class", C.c_void_p), ("read_rid_a", C.c_void_p), ("read_rid_b", C.c_void_p),
                ("leaf_cap", C.c_uint32), ("read_nleaf_u", C.c_void_p), ("read_nleaf_d", C.c_void_p),
                ("read_leaf_u", C.c_void_p), ("read_leaf_d", C.c_void_p),
                ("nundet", C.c_uint64), ("nconf", C.c_uint64), ("n_invalid", C.c_uint64),
                ("n_pairs", C.c_uint64)]


class DeviceCounters(C.Structure):
    _fields_ = [("d_counts", C.c_void_p), ("n_counts", C.c_uint64), ("d_rcount_u", C.c_void_p),
                ("n_rcount_u", C.c_uint64), ("d_rcount_d", C.c_void_p), ("n_rcount_d", C.c_uint64)]


class DeviceInfo(C.Structure):
    _fields_ = [("name", C.c_char * 128), ("sm_count", C.c_int), ("cc_major", C.c_int), ("cc_minor", C.c_int),
                ("l2_bytes", C.c_uint64), ("persisting_l2_max_bytes", C.c_uint64),
                ("access_policy_max_window_bytes", C.c_uint64), ("global_mem_bytes", C.c_uint64),
                ("sm_clock_khz", C.c_int), ("mem_clock_khz", C.c_int), ("mem_bus_bits", C.c_int)]


class Timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("pack_ms", C.c_double), ("scan_ms", C.c_double),
                ("reduce_ms", C.c_double), ("d2h_ms", C.c_double), ("total_ms", C.c_double),
                ("scan_launches", C.c_uint64), ("kernel_launches", C.c_uint64), ("probes", C.c_uint64),
                ("bucket_hits", C.c_uint64), ("leaf_hits", C.c_uint64), ("chained_loads", C.c_uint64),
                ("pack_ms_sum", C.c_double), ("scan_ms_sum", C.c_double), ("reduce_ms_sum", C.c_double),
                ("steps", C.c_uint64), ("grid_blocks", C.c_uint32), ("blocks_per_sm", C.c_uint32),
                ("dyn_smem_bytes", C.c_uint32), ("regs_per_thread", C.c_uint32),
                ("host_pack_ms", C.c_double), ("host_pack_threads", C.c_uint32), ("smem_carveout_pct", C.c_uint32),
                ("h2d_bytes", C.c_uint64), ("sieve_loads", C.c_uint64)]


class MultiInfo(C.Structure):
    _fields_ = [("n_gpus", C.c_int), ("nccl_version", C.c_int), ("devices", C.c_int * 8),
                ("shard_reads", C.c_uint64 * 8), ("reduce_ms", C.c_double)]


class IlpArgs(C.Structure):
    _fields_ = [("erate", C.c_double), ("read_length", C.c_uint32), ("wcov_u", C.c_void_p), ("wcov_d1", C.c_void_p),
                ("wcov_d2", C.c_void_p), ("genome_wcov_u", C.c_void_p), ("genome_wcov_d", C.c_void_p),
                ("genome_rcount_u", C.c_void_p), ("genome_rcount_d", C.c_void_p)]


# every symbol include/cammiq_gpu.h declares: (restype, argtypes)
SYMBOLS = {
    "cq_last_error": (C.c_char_p, []),
    "cq_abi_version": (C.c_int, []),
    "cq_index_load": (C.c_int, [C.c_char_p, C.c_char_p, C.c_double, C.POINTER(C.c_void_p)]),
    "cq_index_free": (None, [C.c_void_p]),
    "cq_index_get_info": (C.c_int, [C.c_void_p, C.POINTER(IndexInfo)]),
    "cq_index_set_filter_budget": (C.c_int, [C.c_void_p, C.c_uint64]),
    "cq_index_leaves": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(LeafView)]),
    "cq_index_map_sp": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p,
                                   C.POINTER(C.c_uint64)]),
    "cq_index_find_host": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_char_p, C.c_size_t,
                                      C.POINTER(C.c_uint64)]),
    "cq_ctx_create": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "cq_ctx_destroy": (None, [C.c_void_p]),
    "cq_index_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "cq_query": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                           C.c_uint64, C.POINTER(Result)]),
    "cq_ctx_set_host_packing": (C.c_int, [C.c_void_p, C.c_int]),
    "cq_pack_reads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p,
                                C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64)]),
    "cq_pack_isa": (C.c_char_p, []),
    "cq_query_packed": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                  C.c_uint64, C.POINTER(Result)]),
    "cq_reset": (C.c_int, [C.c_void_p]),
    "cq_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "cq_host_free": (None, [C.c_void_p]),
    "cq_reads_stage": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]),
    "cq_reads_stage_packed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]),
    "cq_query_staged": (C.c_int, [C.c_void_p, C.c_int]),
    "cq_query_submit": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]),
    "cq_query_submit_packed": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]),
    "cq_sync": (C.c_int, [C.c_void_p]),
    "cq_fetch": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Result)]),
    "cq_get_device_counters": (C.c_int, [C.c_void_p, C.POINTER(DeviceCounters)]),
    "cq_swap_accumulators": (C.c_int, [C.c_void_p, C.POINTER(DeviceCounters)]),
    "cq_get_stream": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "cq_get_timing": (C.c_int, [C.c_void_p, C.POINTER(Timing)]),
    "cq_timing_reset": (C.c_int, [C.c_void_p]),
    "cq_bench_random_sectors": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_double)]),
    "cq_bench_random_gather": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_int, C.c_int,
                                          C.POINTER(C.c_double)]),
    "cq_get_device_info": (C.c_int, [C.c_void_p, C.POINTER(DeviceInfo)]),
    "cq_multi_create": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "cq_multi_destroy": (None, [C.c_void_p]),
    "cq_multi_n_gpus": (C.c_int, [C.c_void_p]),
    "cq_multi_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "cq_multi_query": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                 C.c_uint64, C.POINTER(Result)]),
    "cq_multi_query_packed": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                        C.c_uint64, C.POINTER(Result)]),
    "cq_multi_reset": (C.c_int, [C.c_void_p]),
    "cq_multi_ctx": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "cq_multi_get_info": (C.c_int, [C.c_void_p, C.POINTER(MultiInfo)]),
    "cq_ilp_inputs": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(IlpArgs)]),
}

_LIB = None


def lib():
    """Load libcammiq_gpu.so or fail loudly -- there is no fallback implementation."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise ImportError(
                "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C cammiq_b200/csrc` (nvcc, sm_100a). There is no CPU fallback." % path)
        L = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)  # AttributeError = header / library mismatch
            f.restype, f.argtypes = res, args
        if L.cq_abi_version() != ABI_VERSION:
            raise ImportError("libcammiq_gpu.so ABI version mismatch")
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise CammiqError(rc, lib().cq_last_error().decode(errors="replace"))


def _view(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype)


class Index:
    """Decoded + flattened index pair (host memory)."""

    def __init__(self, path_u, path_d, load_factor=0.0):
        self._h = C.c_void_p()
        _check(lib().cq_index_load(os.fsencode(path_u), os.fsencode(path_d), load_factor, C.byref(self._h)))
        info = IndexInfo()
        _check(lib().cq_index_get_info(self._h, C.byref(info)))
        self.info = info
        self.hash_len = info.hash_len
        self.n_leaves_u, self.n_leaves_d = info.n_leaves_u, info.n_leaves_d

    def set_filter_budget(self, max_bytes):
        _check(lib().cq_index_set_filter_budget(self._h, max_bytes))
        _check(lib().cq_index_get_info(self._h, C.byref(self.info)))

    def leaves(self, table):
        v = LeafView()
        _check(lib().cq_index_leaves(self._h, table, C.byref(v)))
        return dict(ref_id1=_view(v.ref_id1, v.n, np.uint32), ref_id2=_view(v.ref_id2, v.n, np.uint32),
                    ucount1=_view(v.ucount1, v.n, np.uint16), ucount2=_view(v.ucount2, v.n, np.uint16),
                    depth=_view(v.depth, v.n, np.uint8))

    def map_sp(self, table, n_genomes):
        off = np.zeros(n_genomes + 2, dtype=np.uint64)
        total = C.c_uint64()
        _check(lib().cq_index_map_sp(self._h, table, n_genomes, off.ctypes.data, None, C.byref(total)))
        ids = np.zeros(max(total.value, 1), dtype=np.uint64)
        _check(lib().cq_index_map_sp(self._h, table, n_genomes, off.ctypes.data, ids.ctypes.data, C.byref(total)))
        return off, ids[:total.value]

    def find_host(self, table, bucket, cand=b""):
        leaf = C.c_uint64()
        _check(lib().cq_index_find_host(self._h, table, bucket, cand, len(cand), C.byref(leaf)))
        return leaf.value

    def close(self):
        if self._h:
            lib().cq_index_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _CudaArray:
    """Zero-copy view of a device buffer through __cuda_array_interface__ (for torch.as_tensor)."""

    def __init__(self, ptr, n, typestr):
        self.ptr = int(ptr or 0)
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2, "strides": None}


class Context:
    """One CUDA device with a resident index and the accumulated counters."""

    def __init__(self, device=0, stream=None):
        self._h = C.c_void_p()
        self.device = device
        # stream: None -> the context creates its own; an integer cudaStream_t handle otherwise
        # (0, the legacy default stream as torch reports it, is passed as cudaStreamLegacy = 1)
        handle = None if stream is None else C.c_void_p(stream if stream != 0 else 1)
        _check(lib().cq_ctx_create(device, handle, C.byref(self._h)))
        self.n_genomes = 0
        self.index = None
        self._pinned = []

    def upload(self, index, n_genomes):
        _check(lib().cq_index_upload(self._h, index._h, n_genomes))
        self.index, self.n_genomes = index, n_genomes
        return self

    def reset(self):
        _check(lib().cq_reset(self._h))

    def pinned_result_buffers(self):
        """Reusable page-locked buffers for cnt_u/cnt_d/rcount_u/rcount_d (pass as `buffers=`)."""
        G, idx = self.n_genomes, self.index
        out = {}
        for name, n, dt in (("cnt_u", G + 1, np.uint64), ("cnt_d", G + 1, np.uint64),
                            ("rcount_u", max(idx.n_leaves_u, 1), np.uint32),
                            ("rcount_d", max(idx.n_leaves_d, 1), np.uint32)):
            p = C.c_void_p()
            nbytes = n * np.dtype(dt).itemsize
            _check(lib().cq_host_alloc(nbytes, C.byref(p)))
            self._pinned.append(p)
            out[name] = np.frombuffer((C.c_char * nbytes).from_address(p.value), dtype=dt)
        return out

    def _result(self, mode, n_reads, per_read, leaf_cap, want_rcount, pairs_cap, buffers=None):
        G, idx = self.n_genomes, self.index
        if idx is None:
            raise CammiqError(-7, "no index resident (call upload first)")
        res, keep = Result(), {}
        buffers = buffers or {}
        keep["cnt_u"] = buffers["cnt_u"] if "cnt_u" in buffers else np.zeros(G + 1, dtype=np.uint64)
        keep["cnt_d"] = buffers["cnt_d"] if "cnt_d" in buffers else np.zeros(G + 1, dtype=np.uint64)
        res.cnt_u, res.cnt_d = keep["cnt_u"].ctypes.data, keep["cnt_d"].ctypes.data
        if mode == MODE_P and want_rcount:
            keep["rcount_u"] = buffers["rcount_u"] if "rcount_u" in buffers else np.zeros(max(idx.n_leaves_u, 1), dtype=np.uint32)
            keep["rcount_d"] = buffers["rcount_d"] if "rcount_d" in buffers else np.zeros(max(idx.n_leaves_d, 1), dtype=np.uint32)
            res.rcount_u, res.rcount_d = keep["rcount_u"].ctypes.data, keep["rcount_d"].ctypes.data
        if mode == MODE_SC:
            keep["_pairs"] = (PairCount * pairs_cap)()
            res.pairs, res.pairs_cap = C.addressof(keep["_pairs"]), pairs_cap
        n1 = max(n_reads, 1)
        if per_read:
            keep["read_class"] = np.zeros(n1, dtype=np.uint8)
            keep["read_rid_a"] = np.zeros(n1, dtype=np.uint32)
            keep["read_rid_b"] = np.zeros(n1, dtype=np.uint32)
            res.read_class = keep["read_class"].ctypes.data
            res.read_rid_a, res.read_rid_b = keep["read_rid_a"].ctypes.data, keep["read_rid_b"].ctypes.data
            if leaf_cap > 0:
                keep["read_nleaf_u"] = np.zeros(n1, dtype=np.uint32)
                keep["read_nleaf_d"] = np.zeros(n1, dtype=np.uint32)
                keep["read_leaf_u"] = np.zeros((n1, leaf_cap), dtype=np.uint32)
                keep["read_leaf_d"] = np.zeros((n1, leaf_cap), dtype=np.uint32)
                res.leaf_cap = leaf_cap
                res.read_nleaf_u, res.read_nleaf_d = keep["read_nleaf_u"].ctypes.data, keep["read_nleaf_d"].ctypes.data
                res.read_leaf_u, res.read_leaf_d = keep["read_leaf_u"].ctypes.data, keep["read_leaf_d"].ctypes.data
        return res, keep

    def _finish(self, mode, res, keep, n_reads):
        out = {k: v for k, v in keep.items() if not k.startswith("_")}
        for k in ("read_class", "read_rid_a", "read_rid_b", "read_nleaf_u", "read_nleaf_d",
                  "read_leaf_u", "read_leaf_d"):
            if k in out:
                out[k] = out[k][:n_reads]
        if "rcount_u" in out:
            out["rcount_u"] = out["rcount_u"][:self.index.n_leaves_u]
            out["rcount_d"] = out["rcount_d"][:self.index.n_leaves_d]
        out.update(nundet=res.nundet, nconf=res.nconf, n_invalid=res.n_invalid)
        out["pairs"] = {}
        if mode == MODE_SC:
            p = keep["_pairs"]
            out["pairs"] = {(p[i].a, p[i].b): p[i].count for i in range(res.n_pairs)}
        return out

    def query(self, mode, bases, offsets, lengths, stride=0, per_read=False, leaf_cap=0,
              want_rcount=True, pairs_cap=1 << 16, buffers=None):
        """bases uint8[], offsets uint64[] or None (fixed stride), lengths uint8[] (host arrays)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        lengths = np.ascontiguousarray(lengths, dtype=np.uint8)
        if offsets is not None:
            offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(lengths)
        res, keep = self._result(mode, n, per_read, leaf_cap, want_rcount, pairs_cap, buffers)
        _check(lib().cq_query(self._h, mode, bases.ctypes.data,
                              offsets.ctypes.data if offsets is not None else None, stride,
                              lengths.ctypes.data, n, C.byref(res)))
        return self._finish(mode, res, keep, n)

    def set_host_packing(self, threads):
        """threads > 0: cq_query packs reads to 2 bits per base on the host before the copy;
        0: ASCII crosses PCIe; < 0: library default."""
        _check(lib().cq_ctx_set_host_packing(self._h, threads))
        return self

    def query_packed(self, mode, packed, offsets, lengths, stride=0, per_read=False, leaf_cap=0,
                     want_rcount=True, pairs_cap=1 << 16, buffers=None):
        """Reads already packed by pack_reads(): packed uint8[], offsets uint64[] or None (fixed
        stride in bytes), lengths uint8[] (0 = invalid read)."""
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        lengths = np.ascontiguousarray(lengths, dtype=np.uint8)
        if offsets is not None:
            offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(lengths)
        res, keep = self._result(mode, n, per_read, leaf_cap, want_rcount, pairs_cap, buffers)
        _check(lib().cq_query_packed(self._h, mode, packed.ctypes.data,
                                     offsets.ctypes.data if offsets is not None else None, stride,
                                     lengths.ctypes.data, n, C.byref(res)))
        return self._finish(mode, res, keep, n)

    # device-resident plumbing -------------------------------------------------------------
    def stage(self, bases, offsets, lengths, stride=0):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        lengths = np.ascontiguousarray(lengths, dtype=np.uint8)
        if offsets is not None:
            offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        _check(lib().cq_reads_stage(self._h, bases.ctypes.data,
                                    offsets.ctypes.data if offsets is not None else None, stride,
                                    lengths.ctypes.data, len(lengths)))
        _check(lib().cq_sync(self._h))

    def stage_packed(self, packed, offsets, lengths, stride=0):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        lengths = np.ascontiguousarray(lengths, dtype=np.uint8)
        if offsets is not None:
            offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        _check(lib().cq_reads_stage_packed(self._h, packed.ctypes.data,
                                           offsets.ctypes.data if offsets is not None else None, stride,
                                           lengths.ctypes.data, len(lengths)))
        _check(lib().cq_sync(self._h))

    def query_staged(self, mode):
        _check(lib().cq_query_staged(self._h, mode))

    def submit(self, mode, bases, offsets, lengths, stride=0, packed=False):
        """query() without the copy of the totals to the host (cq_query_submit[_packed]); the caller's
        arrays must stay alive until sync()."""
        fn = lib().cq_query_submit_packed if packed else lib().cq_query_submit
        _check(fn(self._h, mode, bases.ctypes.data, offsets.ctypes.data if offsets is not None else None, stride,
                  lengths.ctypes.data, len(lengths)))

    def sync(self):
        _check(lib().cq_sync(self._h))

    def fetch(self, mode, want_rcount=True, pairs_cap=1 << 16):
        res, keep = self._result(mode, 0, False, 0, want_rcount, pairs_cap)
        _check(lib().cq_fetch(self._h, mode, C.byref(res)))
        return self._finish(mode, res, keep, 0)

    def device_counters(self):
        dc = DeviceCounters()
        _check(lib().cq_get_device_counters(self._h, C.byref(dc)))
        return dc

    def device_counter_arrays(self):
        """(counts as int64 view, rcount_u, rcount_d as int32 views) exposing
        __cuda_array_interface__; integer sums are bit-identical in two's complement."""
        dc = self.device_counters()
        return (_CudaArray(dc.d_counts, dc.n_counts, "<i8"),
                _CudaArray(dc.d_rcount_u, max(dc.n_rcount_u, 1), "<i4"),
                _CudaArray(dc.d_rcount_d, max(dc.n_rcount_d, 1), "<i4"))

    def swap_accumulators(self):
        """Make the context's other accumulator set current (cq_swap_accumulators); returns the arrays of
        the set that was current, as device_counter_arrays() does."""
        dc = DeviceCounters()
        _check(lib().cq_swap_accumulators(self._h, C.byref(dc)))
        return (_CudaArray(dc.d_counts, dc.n_counts, "<i8"),
                _CudaArray(dc.d_rcount_u, max(dc.n_rcount_u, 1), "<i4"),
                _CudaArray(dc.d_rcount_d, max(dc.n_rcount_d, 1), "<i4"))

    def timing(self):
        t = Timing()
        _check(lib().cq_get_timing(self._h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in Timing._fields_}

    def timing_reset(self):
        _check(lib().cq_timing_reset(self._h))

    def bench_random_sectors(self, n_probes, iters=3):
        v = C.c_double()
        _check(lib().cq_bench_random_sectors(self._h, n_probes, iters, C.byref(v)))
        return v.value

    def bench_random_gather(self, region_bytes, access_bytes, n_probes, iters=3, persist=False):
        v = C.c_double()
        _check(lib().cq_bench_random_gather(self._h, region_bytes, access_bytes, n_probes, iters,
                                            1 if persist else 0, C.byref(v)))
        return v.value

    def ilp_inputs(self, erate, read_length):
        """ILP set-up coefficients from the accumulated rcount (cq_ilp_inputs): per-leaf wcov in
        file order, per-genome coverage sums and rcount sums."""
        G, idx = self.n_genomes, self.index
        out = dict(wcov_u=np.zeros(max(idx.n_leaves_u, 1)), wcov_d1=np.zeros(max(idx.n_leaves_d, 1)),
                   wcov_d2=np.zeros(max(idx.n_leaves_d, 1)), genome_wcov_u=np.zeros(G + 1), genome_wcov_d=np.zeros(G + 1),
                   genome_rcount_u=np.zeros(G + 1, dtype=np.uint64), genome_rcount_d=np.zeros(G + 1, dtype=np.uint64))
        a = IlpArgs(erate=erate, read_length=read_length, **{k: v.ctypes.data for k, v in out.items()})
        _check(lib().cq_ilp_inputs(self._h, idx._h, C.byref(a)))
        out["wcov_u"] = out["wcov_u"][:idx.n_leaves_u]
        out["wcov_d1"], out["wcov_d2"] = out["wcov_d1"][:idx.n_leaves_d], out["wcov_d2"][:idx.n_leaves_d]
        return out

    def device_info(self):
        d = DeviceInfo()
        _check(lib().cq_get_device_info(self._h, C.byref(d)))
        out = {k: getattr(d, k) for k, _ in DeviceInfo._fields_}
        out["name"] = out["name"].decode()
        return out

    def close(self):
        if self._h:
            lib().cq_ctx_destroy(self._h)
            self._h = C.c_void_p()
        for p in self._pinned:
            lib().cq_host_free(p)
        self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiContext(Context):
    """cq_multi: the same interface as Context over several GPUs of one box -- reads sharded,
    index replicated, one NCCL reduce of the counters (include/cammiq_gpu.h)."""

    def __init__(self, n_gpus, devices=None):  # noqa: super().__init__ would create a single-device context
        self._m = C.c_void_p()
        arr = (C.c_int * n_gpus)(*devices) if devices is not None else None
        _check(lib().cq_multi_create(n_gpus, arr, C.byref(self._m)))
        self._h = C.c_void_p()
        self.n_genomes, self.index, self._pinned = 0, None, []

    def upload(self, index, n_genomes):
        _check(lib().cq_multi_upload(self._m, index._h, n_genomes))
        self.index, self.n_genomes = index, n_genomes
        return self

    def reset(self):
        _check(lib().cq_multi_reset(self._m))

    def _run(self, fn, mode, bases, offsets, lengths, stride, per_read, leaf_cap, want_rcount, pairs_cap, buffers):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        lengths = np.ascontiguousarray(lengths, dtype=np.uint8)
        if offsets is not None:
            offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(lengths)
        res, keep = self._result(mode, n, per_read, leaf_cap, want_rcount, pairs_cap, buffers)
        _check(fn(self._m, mode, bases.ctypes.data, offsets.ctypes.data if offsets is not None else None, stride,
                  lengths.ctypes.data, n, C.byref(res)))
        return self._finish(mode, res, keep, n)

    def query(self, mode, bases, offsets, lengths, stride=0, per_read=False, leaf_cap=0, want_rcount=True,
              pairs_cap=1 << 16, buffers=None):
        return self._run(lib().cq_multi_query, mode, bases, offsets, lengths, stride, per_read, leaf_cap, want_rcount,
                         pairs_cap, buffers)

    def query_packed(self, mode, packed, offsets, lengths, stride=0, per_read=False, leaf_cap=0, want_rcount=True,
                     pairs_cap=1 << 16, buffers=None):
        return self._run(lib().cq_multi_query_packed, mode, packed, offsets, lengths, stride, per_read, leaf_cap,
                         want_rcount, pairs_cap, buffers)

    def set_host_packing(self, threads):
        for i in range(lib().cq_multi_n_gpus(self._m)):
            h = C.c_void_p()
            _check(lib().cq_multi_ctx(self._m, i, C.byref(h)))
            _check(lib().cq_ctx_set_host_packing(h, threads))
        return self

    def info(self):
        mi = MultiInfo()
        _check(lib().cq_multi_get_info(self._m, C.byref(mi)))
        n = mi.n_gpus
        return dict(n_gpus=n, nccl_version=mi.nccl_version, devices=list(mi.devices)[:n],
                    shard_reads=list(mi.shard_reads)[:n], reduce_ms=mi.reduce_ms)

    def close(self):
        if self._m:
            lib().cq_multi_destroy(self._m)
            self._m = C.c_void_p()
        for p in self._pinned:
            lib().cq_host_free(p)
        self._pinned = []


def pack_isa():
    return lib().cq_pack_isa().decode()


def pack_reads(bases, offsets, lengths, stride=0, threads=1, packed_stride=None):
    """Host-only: ASCII reads -> (packed uint8[n, packed_stride], lengths uint8[n] with 0 for
    invalid reads, n_invalid).  Layout: include/cammiq_gpu.h (cq_pack_reads)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    lengths = np.ascontiguousarray(lengths, dtype=np.uint8)
    if offsets is not None:
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(lengths)
    if packed_stride is None:
        packed_stride = (int(lengths.max()) + 3) // 4 if n else 1
    packed = np.zeros((max(n, 1), max(packed_stride, 1)), dtype=np.uint8)
    out_len = np.zeros(max(n, 1), dtype=np.uint8)
    bad = C.c_uint64()
    _check(lib().cq_pack_reads(bases.ctypes.data, offsets.ctypes.data if offsets is not None else None, stride,
                               lengths.ctypes.data, n, threads, packed.ctypes.data, packed_stride,
                               out_len.ctypes.data, C.byref(bad)))
    return packed[:n], out_len[:n], bad.value

