"""Generate the committed golden fixtures with the UNMODIFIED reference (oracle/_ref).

Run in the build container (needs /root/reference compiled by `make -C oracle ref`):
    python tests/golden/make_golden.py
Each case directory holds a real index built by the reference's own builder, the reads, and
the reference's answers (ref_harness dumps: counters, per-leaf rcount, pair map, and for the
first reads the reference's per-read decision and leaf set).  Everything is seeded.
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import synth  # noqa: E402

CASES = {
    # name: genomes, length, cluster, divergence, private, reads, rl, jitter, erate, k, L, Lmax, h
    "cfg1_small": dict(ng=6, glen=30000, cl=3, div=0.01, priv=0.15, nr=1500, rl=100, jit=0, er=0.01,
                       k=26, L=100, Lmax=50, h=26, seed=11),
    "deep_h12": dict(ng=6, glen=12000, cl=2, div=0.002, priv=0.15, nr=900, rl=100, jit=60, er=0.02,
                     k=20, L=100, Lmax=50, h=12, seed=12),
    "long150_h20": dict(ng=4, glen=20000, cl=2, div=0.005, priv=0.1, nr=700, rl=150, jit=100, er=0.03,
                        k=26, L=150, Lmax=60, h=20, seed=13),
    "adversarial_250": dict(ng=6, glen=15000, cl=2, div=0.001, priv=0.05, nr=600, rl=250, jit=5, er=0.05,
                            k=22, L=250, Lmax=80, h=16, seed=14, n_rate=0.01),
}


def main():
    assert synth.have_reference(), "build oracle/_ref first: make -C oracle ref"
    for name, c in CASES.items():
        out = os.path.join(HERE, name)
        shutil.rmtree(out, ignore_errors=True)
        work = os.path.join("/tmp", "golden_" + name)
        shutil.rmtree(work, ignore_errors=True)
        rng = np.random.default_rng(c["seed"])
        genomes = synth.make_genomes(rng, c["ng"], c["glen"], cluster_size=c["cl"],
                                     divergence=c["div"], private_frac=c["priv"])
        map_fn = synth.write_fasta_set(os.path.join(work, "fa"), genomes)
        synth.build_reference_index(os.path.join(work, "fa"), map_fn, out, k=c["k"], L=c["L"],
                                    Lmax=c["Lmax"], h=c["h"])
        shutil.copy(map_fn, os.path.join(out, "genome_map.out"))
        reads, _ = synth.simulate_reads(rng, genomes, c["nr"], c["rl"], erate=c["er"],
                                        n_rate=c.get("n_rate", 0.0), lower_frac=0.1, len_jitter=c["jit"])
        # chimeras provoke conflicts and the |P| >= 2 rows; N is substituted the way the
        # reference reader does (one random base per read, query.cpp:383) but seeded, so the
        # fixture holds POST-substitution reads
        for i in range(0, len(reads), 7):
            j = (i * 13 + 5) % len(reads)
            reads[i] = reads[i][:len(reads[i]) // 2] + reads[j][len(reads[j]) // 2:]
        subs = [b"ACGT"[int(rng.integers(0, 4)):][:1] for _ in reads]
        reads = [r.replace(b"N", s).replace(b"n", s.lower()) for r, s in zip(reads, subs)]
        reads[3] = reads[3][:c["h"]]      # a read of length exactly h
        reads[4] = reads[5]                # duplicate read
        fq = os.path.join(out, "reads.fq")
        synth.write_fastq(fq, reads)
        iu, idd = os.path.join(out, "index_u.bin1"), os.path.join(out, "index_d.bin2")
        for mode in ("p", "sc"):
            synth.run_ref_dump(iu, idd, os.path.join(out, "genome_map.out"), mode, [fq],
                               os.path.join(out, "dump_%s.txt" % mode), per_read_n=400)
        # query64mt_p must agree with query64_p; keep only the proof
        synth.run_ref_dump(iu, idd, os.path.join(out, "genome_map.out"), "mt", [fq],
                           os.path.join(work, "dump_mt.txt"), threads=4)
        a = open(os.path.join(out, "dump_p.txt")).read().split("READ ")[0].replace("MODE p", "MODE x")
        b = open(os.path.join(work, "dump_mt.txt")).read().replace("MODE mt", "MODE x")
        assert a == b, "query64mt_p disagrees with query64_p on " + name
        d = synth.parse_ref_dump(os.path.join(out, "dump_p.txt"))
        f = d["files"][0]
        size = sum(os.path.getsize(os.path.join(out, x)) for x in os.listdir(out))
        print("%-16s h=%d nU=%d nD=%d reads=%d nundet=%d nconf=%d sum_u=%d sum_d=%d bytes=%d" % (
            name, d["h"], d["nu"], d["nd"], f["nreads"], f["nundet"], f["nconf"], sum(f["cu"]),
            sum(f["cd"]), size))


if __name__ == "__main__" and len(sys.argv) == 1:
    main()


def make_cli_expectations():
    """Outputs of the UNMODIFIED reference CLI (oracle/_ref/cammiq_ref) on every case: the
    --read_cnts output file and the counter lines of its stderr, for the drop-in CLI test."""
    import re
    import subprocess
    for name in CASES:
        d = os.path.join(HERE, name)
        base = [synth.CAMMIQ_REF, "--query", "-f", os.path.join(d, "genome_map.out"), "-q",
                os.path.join(d, "reads.fq"), "-i", os.path.join(d, "index_u.bin1"), os.path.join(d, "index_d.bin2")]
        out = os.path.join(d, "ref_cli_read_cnts.out")
        r1 = subprocess.run(base[:2] + ["--read_cnts"] + base[2:] + ["-o", out], capture_output=True, text=True)
        r2 = subprocess.run(base + ["-o", os.path.join("/tmp", "unused.out")], capture_output=True, text=True)
        assert r1.returncode == 0 and r2.returncode == 0, (r1.stderr, r2.stderr)
        keep = re.compile(r"^(Querying|Number of|Completed query|Hash Length)")
        with open(os.path.join(d, "ref_cli_stderr.txt"), "w") as f:
            for tag, r in (("read_cnts", r1), ("standard", r2)):
                for line in r.stderr.replace("\r", "\n").split("\n"):
                    if keep.match(line):
                        f.write(tag + "\t" + line + "\n")
        print(name, open(out).read().strip().split("\n")[-1][:80])


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "cli":
    make_cli_expectations()
