#!/usr/bin/env python
"""TEST / BENCH INFRASTRUCTURE.  BASELINE.json configs[0] on an index built by the UNMODIFIED
reference builder: 10 synthetic 1 Mbp genomes in strain clusters (shared segments, 0.5 %
divergence), `cammiq --build --both -k 26 -L 100 -Lmax 50 -h 26` (oracle/_ref/cammiq_ref), 100 000
simulated error-free 100-bp reads, and the reference's own answers for them
(oracle/_ref/ref_harness dump: counters, every per-leaf rcount, pair map).

    python oracle/fixtures/make_cfg1.py [out_dir]     # default oracle/_ref/fixtures/cfg1

Needs /root/reference compiled (make -C oracle ref), so it runs in the build container; the
result (about 10 MB, git-ignored like the rest of oracle/_ref) travels to the GPU box.  Seeded."""
import os
import shutil
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "tests"))
import synth  # noqa: E402

PARAMS = dict(ng=10, glen=1_000_000, cl=3, div=0.005, priv=0.1, nr=100_000, rl=100, k=26, L=100, Lmax=50, h=26, seed=101)


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "oracle", "_ref", "fixtures", "cfg1")
    if os.path.exists(os.path.join(out, ".done")):
        print("cfg1 fixture present:", out)
        return
    assert synth.have_reference(), "build oracle/_ref first: make -C oracle ref"
    c = PARAMS
    shutil.rmtree(out, ignore_errors=True)
    work = "/tmp/cammiq_cfg1_fixture"
    shutil.rmtree(work, ignore_errors=True)
    rng = np.random.default_rng(c["seed"])
    t = time.time()
    genomes = synth.make_genomes(rng, c["ng"], c["glen"], cluster_size=c["cl"], divergence=c["div"], private_frac=c["priv"])
    map_fn = synth.write_fasta_set(os.path.join(work, "fa"), genomes)
    print("genomes written %.1f s" % (time.time() - t))
    t = time.time()
    synth.build_reference_index(os.path.join(work, "fa"), map_fn, out, k=c["k"], L=c["L"], Lmax=c["Lmax"], h=c["h"], threads=4)
    print("reference build %.1f s" % (time.time() - t))
    shutil.copy(map_fn, os.path.join(out, "genome_map.out"))
    reads, _ = synth.simulate_reads(rng, genomes, c["nr"], c["rl"], erate=0.0)
    fq = os.path.join(out, "reads.fq")
    synth.write_fastq(fq, reads)
    iu, idd = os.path.join(out, "index_u.bin1"), os.path.join(out, "index_d.bin2")
    t = time.time()
    for mode in ("p", "sc"):
        synth.run_ref_dump(iu, idd, os.path.join(out, "genome_map.out"), mode, [fq], os.path.join(out, "dump_%s.txt" % mode))
    print("reference dumps %.1f s" % (time.time() - t))
    shutil.rmtree(work, ignore_errors=True)
    open(os.path.join(out, ".done"), "w").write(repr(c))
    print("wrote", out, sum(os.path.getsize(os.path.join(out, f)) for f in os.listdir(out)), "bytes")


if __name__ == "__main__":
    main()
