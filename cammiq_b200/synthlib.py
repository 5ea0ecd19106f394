"""ctypes binding of libcammiq_synth.so -- the seeded synthetic workload generator
(cammiq_b200/csrc/synth.cpp).  TOOLING for bench.py and tests; never on the query path."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class Params(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_genomes", C.c_uint32), ("genome_len", C.c_uint32),
                ("cluster_size", C.c_uint32), ("block_len", C.c_uint32),
                ("permille_private", C.c_uint32), ("permille_pair", C.c_uint32),
                ("u_per_block", C.c_uint32), ("d_per_block", C.c_uint32), ("k", C.c_uint32),
                ("lmax", C.c_uint32), ("permille_deep", C.c_uint32), ("threads", C.c_uint32)]


class IndexStats(C.Structure):
    _fields_ = [("n_leaves_u", C.c_uint64), ("n_leaves_d", C.c_uint64), ("n_dropped", C.c_uint64),
                ("gen_ms", C.c_double), ("sort_ms", C.c_double), ("write_ms", C.c_double)]


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libcammiq_synth.so")
        if not os.path.exists(path):
            raise ImportError(path + " is missing: run `make -C cammiq_b200/csrc`.")
        L = C.CDLL(path)
        L.cqs_write_index.restype = C.c_int
        L.cqs_write_index.argtypes = [C.POINTER(Params), C.c_char_p, C.POINTER(IndexStats)]
        L.cqs_make_reads.restype = C.c_int
        L.cqs_make_reads.argtypes = [C.POINTER(Params), C.c_uint64, C.c_uint64, C.c_uint32, C.c_double,
                                     C.c_void_p, C.c_void_p]
        L.cqs_write_fastq.restype = C.c_int
        L.cqs_write_fastq.argtypes = [C.POINTER(Params), C.c_uint64, C.c_uint64, C.c_uint32, C.c_double,
                                      C.c_char_p]
        _LIB = L
    return _LIB


def params(seed=1, n_genomes=10, genome_len=1 << 20, cluster_size=4, block_len=1024,
           permille_private=450, permille_pair=250, u_per_block=36, d_per_block=64, k=26, lmax=50,
           permille_deep=50, threads=0):
    """Defaults reproduce the densities SURVEY.md section 8 measured on strain clusters: about
    17 unique leaves per kbp of private sequence, 10 doubly-unique leaves per kbp of pair-shared
    sequence (20 per block, emitted once per pair), 5% of keys longer than k."""
    if threads <= 0:
        threads = min(os.cpu_count() or 1, 32)
    return Params(seed, n_genomes, genome_len, cluster_size, block_len, permille_private,
                  permille_pair, u_per_block, d_per_block, k, lmax, permille_deep, threads)


def write_index(p, out_dir):
    os.makedirs(out_dir, exist_ok=True)
    st = IndexStats()
    rc = lib().cqs_write_index(C.byref(p), os.fsencode(out_dir), C.byref(st))
    if rc != 0:
        raise RuntimeError("cqs_write_index failed: %d" % rc)
    return {k: getattr(st, k) for k, _ in IndexStats._fields_}


def make_reads(p, first, n, read_len, erate, out=None, want_src=False):
    """Returns uint8[n, read_len] ASCII reads (and the 1-based source genome ids)."""
    if out is None:
        out = np.empty((n, read_len), dtype=np.uint8)
    src = np.empty(n, dtype=np.uint32) if want_src else None
    rc = lib().cqs_make_reads(C.byref(p), first, n, read_len, erate, out.ctypes.data,
                              src.ctypes.data if want_src else None)
    if rc != 0:
        raise RuntimeError("cqs_make_reads failed: %d" % rc)
    return (out, src) if want_src else out


def write_fastq(p, first, n, read_len, erate, path):
    rc = lib().cqs_write_fastq(C.byref(p), first, n, read_len, erate, os.fsencode(path))
    if rc != 0:
        raise RuntimeError("cqs_write_fastq failed: %d" % rc)
