"""ctypes binding of oracle/liboracle.so (TEST INFRASTRUCTURE: the CPU restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(REPO, "oracle")
_LIB = None

NONE = 2 ** 64 - 1
CLASS_UNLABELED, CLASS_CONFLICT, CLASS_U, CLASS_D_PAIR, CLASS_UD, CLASS_D_INTER = range(6)
MODE_P, MODE_SC = 0, 1


class _Result(C.Structure):
    _fields_ = [
        ("cnt_u", C.c_void_p), ("cnt_d", C.c_void_p),
        ("nundet", C.c_uint64), ("nconf", C.c_uint64), ("n_invalid", C.c_uint64),
        ("rcount_u", C.c_void_p), ("rcount_d", C.c_void_p),
        ("pair_a", C.c_void_p), ("pair_b", C.c_void_p), ("pair_cnt", C.c_void_p),
        ("n_pairs", C.c_uint64), ("pairs_cap", C.c_uint64),
        ("read_class", C.c_void_p), ("read_rid_a", C.c_void_p), ("read_rid_b", C.c_void_p),
        ("leaf_cap", C.c_uint32),
        ("read_nleaf_u", C.c_void_p), ("read_nleaf_d", C.c_void_p),
        ("read_leaf_u", C.c_void_p), ("read_leaf_d", C.c_void_p),
    ]


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        src = os.path.join(ORACLE_DIR, "cammiq_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        L.cqo_index_load.restype = C.c_void_p
        L.cqo_index_load.argtypes = [C.c_char_p]
        L.cqo_index_free.argtypes = [C.c_void_p]
        for name, rt in (("hash_len", C.c_uint32), ("is_doubly_unique", C.c_int),
                         ("num_buckets", C.c_uint64), ("num_leaves", C.c_uint64)):
            f = getattr(L, "cqo_index_" + name)
            f.restype, f.argtypes = rt, [C.c_void_p]
        for name in ("ref1", "ref2", "ucount1", "ucount2", "depth"):
            f = getattr(L, "cqo_leaf_" + name)
            f.restype, f.argtypes = C.c_void_p, [C.c_void_p]
        L.cqo_canonical_ids.argtypes = [C.c_void_p, C.c_void_p]
        L.cqo_map_sp.restype = C.c_uint64
        L.cqo_map_sp.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.cqo_hash.restype = C.c_uint64
        L.cqo_hash.argtypes = [C.c_char_p, C.c_uint32]
        L.cqo_find.restype = C.c_uint64
        L.cqo_find.argtypes = [C.c_void_p, C.c_uint64, C.c_char_p, C.c_size_t]
        L.cqo_query.restype = C.c_int
        L.cqo_query.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_uint64, C.POINTER(_Result)]
        _LIB = L
    return _LIB


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).copy()


class OracleIndex:
    def __init__(self, path):
        self._h = lib().cqo_index_load(path.encode())
        if not self._h:
            raise RuntimeError("oracle: cannot decode index " + path)
        L = lib()
        self.h = L.cqo_index_hash_len(self._h)
        self.is_d = bool(L.cqo_index_is_doubly_unique(self._h))
        self.n_buckets = L.cqo_index_num_buckets(self._h)
        self.n_leaves = L.cqo_index_num_leaves(self._h)
        n = self.n_leaves
        self.ref1 = _arr(L.cqo_leaf_ref1(self._h), n, np.uint32)
        self.ref2 = _arr(L.cqo_leaf_ref2(self._h), n, np.uint32)
        self.ucount1 = _arr(L.cqo_leaf_ucount1(self._h), n, np.uint16)
        self.ucount2 = _arr(L.cqo_leaf_ucount2(self._h), n, np.uint16)
        self.depth = _arr(L.cqo_leaf_depth(self._h), n, np.uint8)

    def canonical_ids(self):
        out = np.zeros(max(self.n_leaves, 1), dtype=np.uint64)
        lib().cqo_canonical_ids(self._h, out.ctypes.data)
        return out[:self.n_leaves]

    def map_sp(self, G):
        off = np.zeros(G + 2, dtype=np.uint64)
        total = lib().cqo_map_sp(self._h, G, off.ctypes.data, None)
        ids = np.zeros(max(total, 1), dtype=np.uint64)
        lib().cqo_map_sp(self._h, G, off.ctypes.data, ids.ctypes.data)
        return off, ids[:total]

    def find(self, bucket, cand):
        return lib().cqo_find(self._h, bucket, cand, len(cand))

    def close(self):
        if self._h:
            lib().cqo_index_free(self._h)
            self._h = None

    def __del__(self):
        self.close()


def pack_reads(reads):
    """list of bytes -> (bases uint8[], offsets uint64[], lengths uint8[]) like the reference's
    reads / rlengths vectors (length truncated to uint8_t, query.cpp:387)."""
    lengths = np.array([len(r) & 0xFF for r in reads], dtype=np.uint8)
    full = np.array([len(r) for r in reads], dtype=np.uint64)
    offsets = np.zeros(len(reads), dtype=np.uint64)
    if len(reads) > 1:
        offsets[1:] = np.cumsum(full)[:-1]
    bases = np.frombuffer(b"".join(reads), dtype=np.uint8).copy() if reads else np.zeros(0, np.uint8)
    return bases, offsets, lengths


def oracle_query(idx_u, idx_d, mode, G, bases, offsets, lengths, per_read=False, leaf_cap=0,
                 pairs_cap=1 << 16):
    n = len(lengths)
    res = _Result()
    cnt_u = np.zeros(G + 1, dtype=np.uint64)
    cnt_d = np.zeros(G + 1, dtype=np.uint64)
    rc_u = np.zeros(max(idx_u.n_leaves, 1), dtype=np.uint32)
    rc_d = np.zeros(max(idx_d.n_leaves, 1), dtype=np.uint32)
    pa = np.zeros(pairs_cap, dtype=np.uint32)
    pb = np.zeros(pairs_cap, dtype=np.uint32)
    pc = np.zeros(pairs_cap, dtype=np.uint64)
    res.cnt_u, res.cnt_d = cnt_u.ctypes.data, cnt_d.ctypes.data
    res.rcount_u, res.rcount_d = rc_u.ctypes.data, rc_d.ctypes.data
    res.pair_a, res.pair_b, res.pair_cnt = pa.ctypes.data, pb.ctypes.data, pc.ctypes.data
    res.pairs_cap = pairs_cap
    keep = [cnt_u, cnt_d, rc_u, rc_d, pa, pb, pc]
    out = {}
    if per_read:
        cls = np.zeros(max(n, 1), dtype=np.uint8)
        ra = np.zeros(max(n, 1), dtype=np.uint32)
        rb = np.zeros(max(n, 1), dtype=np.uint32)
        res.read_class, res.read_rid_a, res.read_rid_b = cls.ctypes.data, ra.ctypes.data, rb.ctypes.data
        out.update(read_class=cls[:n], read_rid_a=ra[:n], read_rid_b=rb[:n])
        if leaf_cap > 0:
            nlu = np.zeros(max(n, 1), dtype=np.uint32)
            nld = np.zeros(max(n, 1), dtype=np.uint32)
            lu = np.zeros(max(n, 1) * leaf_cap, dtype=np.uint32)
            ld = np.zeros(max(n, 1) * leaf_cap, dtype=np.uint32)
            res.leaf_cap = leaf_cap
            res.read_nleaf_u, res.read_nleaf_d = nlu.ctypes.data, nld.ctypes.data
            res.read_leaf_u, res.read_leaf_d = lu.ctypes.data, ld.ctypes.data
            out.update(read_nleaf_u=nlu[:n], read_nleaf_d=nld[:n],
                       read_leaf_u=lu.reshape(-1, leaf_cap)[:n], read_leaf_d=ld.reshape(-1, leaf_cap)[:n])
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    lengths = np.ascontiguousarray(lengths, dtype=np.uint8)
    rc = lib().cqo_query(idx_u._h, idx_d._h, mode, G, bases.ctypes.data, offsets.ctypes.data,
                         lengths.ctypes.data, n, C.byref(res))
    if rc != 0:
        raise RuntimeError("oracle query failed (rc=%d)" % rc)
    out.update(cnt_u=cnt_u, cnt_d=cnt_d, nundet=res.nundet, nconf=res.nconf, n_invalid=res.n_invalid,
               rcount_u=rc_u[:idx_u.n_leaves], rcount_d=rc_d[:idx_d.n_leaves],
               pairs={(int(pa[i]), int(pb[i])): int(pc[i]) for i in range(res.n_pairs)})
    del keep
    return out
