"""Pin the C restatement (oracle/cammiq_oracle.c) against the UNMODIFIED reference compiled
in oracle/_ref, live, on seeded inputs.  Skipped where oracle/_ref is absent (the committed
dumps in tests/golden/ cover that case: test_oracle_golden.py)."""
import os

import numpy as np
import pytest

import oracle_lib as ol
import parity
import synth

pytestmark = pytest.mark.skipif(not synth.have_reference(), reason="oracle/_ref not built")

CASES = [
    # name, genomes, length, cluster, divergence, reads, rl, erate, k, L, Lmax, h
    ("default_h26", 6, 30000, 3, 0.01, 1500, 100, 0.01, 26, 100, 50, 26),
    ("deep_tries_h12", 5, 12000, 5, 0.002, 800, 100, 0.02, 20, 100, 50, 12),
    ("long_reads_150", 4, 20000, 2, 0.005, 600, 150, 0.03, 26, 150, 60, 20),
]


@pytest.fixture(scope="module", params=CASES, ids=[c[0] for c in CASES])
def case(request, tmp_path_factory):
    name, ng, glen, cl, div, nr, rl, er, k, L, Lmax, h = request.param
    d = str(tmp_path_factory.mktemp(name))
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31))
    rng = np.random.default_rng(len(name) * 7919 + ng)
    genomes = synth.make_genomes(rng, ng, glen, cluster_size=cl, divergence=div, private_frac=0.15)
    map_fn = synth.write_fasta_set(os.path.join(d, "fa"), genomes)
    iu, idd = synth.build_reference_index(os.path.join(d, "fa"), map_fn, os.path.join(d, "idx"),
                                          k=k, L=L, Lmax=Lmax, h=h)
    reads, _ = synth.simulate_reads(rng, genomes, nr, rl, erate=er, lower_frac=0.1, len_jitter=rl - h - 5)
    # chimeric reads provoke conflicts and |P|>=2 rows
    for i in range(0, len(reads), 7):
        j = (i * 13 + 5) % len(reads)
        reads[i] = reads[i][:len(reads[i]) // 2] + reads[j][len(reads[j]) // 2:]
    reads[3] = reads[3][:h]          # read of length exactly h
    fq = os.path.join(d, "reads.fq")
    synth.write_fastq(fq, reads)
    return dict(dir=d, map=map_fn, iu=iu, id=idd, fq=fq, reads=reads, G=ng, h=h)


@pytest.mark.parametrize("mode", ["p", "mt", "sc"])
def test_oracle_matches_reference(case, mode):
    dump_fn = os.path.join(case["dir"], "dump_%s.txt" % mode)
    synth.run_ref_dump(case["iu"], case["id"], case["map"], mode, [case["fq"]], dump_fn,
                       threads=4 if mode == "mt" else 1, per_read_n=400)
    dump = synth.parse_ref_dump(dump_fn)
    oi_u, oi_d = ol.OracleIndex(case["iu"]), ol.OracleIndex(case["id"])
    assert (oi_u.h, oi_u.n_leaves, oi_d.n_leaves) == (dump["h"], dump["nu"], dump["nd"])
    assert oi_u.h == case["h"]
    parity.check_leaf_tables(dump, oi_u, oi_d)
    bases, offsets, lengths = ol.pack_reads(case["reads"])
    res = ol.oracle_query(oi_u, oi_d, ol.MODE_SC if mode == "sc" else ol.MODE_P, case["G"],
                          bases, offsets, lengths, per_read=True, leaf_cap=64)
    assert res["n_invalid"] == 0
    parity.check_counters(res, dump["files"][0], oi_u, oi_d, case["G"], mode)
    parity.check_per_read(res, dump, oi_u, oi_d, mode)
    # the scan must have exercised something
    f = dump["files"][0]
    assert sum(f["cu"]) > 0 and f["nundet"] < f["nreads"]
