"""Seeded synthetic genomes / reads / FASTA / FASTQ / map files for tests and fixtures.

Test tooling only (never imported by the product path).  The read model follows the
reference's simulator (CAMMiQ-simulate:119-149, 242-273): uniform start, reverse complement
with probability 0.5, i.i.d. substitutions at rate `erate`, optional N rate.
Genomes are strain clusters: an i.i.d. ACGT ancestor per cluster, per-strain SNPs at
`divergence`, plus a private i.i.d. tail per genome.
"""
import os
import subprocess

import numpy as np

ALPHABET = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTacgt", b"TGCATGCA"):
    COMP[_a] = _b

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(REPO, "oracle", "_ref")
CAMMIQ_REF = os.path.join(REF_DIR, "cammiq_ref")
REF_HARNESS = os.path.join(REF_DIR, "ref_harness")


def make_genomes(rng, n_genomes, length, cluster_size=3, divergence=0.005, private_frac=0.1):
    """Return a list of uint8 ASCII arrays (ACGT)."""
    genomes = []
    shared_len = int(length * (1.0 - private_frac))
    ancestor = None
    for g in range(n_genomes):
        if g % cluster_size == 0:
            ancestor = rng.integers(0, 4, size=shared_len, dtype=np.uint8)
        body = ancestor.copy()
        snp = rng.random(shared_len) < divergence
        body[snp] = (body[snp] + rng.integers(1, 4, size=int(snp.sum()), dtype=np.uint8)) & 3
        tail = rng.integers(0, 4, size=length - shared_len, dtype=np.uint8)
        genomes.append(ALPHABET[np.concatenate([body, tail])])
    return genomes


def revcomp(seq):
    return COMP[seq[::-1]]


def simulate_reads(rng, genomes, n_reads, read_len, erate=0.0, n_rate=0.0, rc_prob=0.5,
                   lower_frac=0.0, len_jitter=0):
    """Return (list of bytes reads, list of source genome index)."""
    reads, src = [], []
    for _ in range(n_reads):
        g = int(rng.integers(0, len(genomes)))
        rl = read_len if len_jitter == 0 else int(rng.integers(read_len - len_jitter, read_len + 1))
        start = int(rng.integers(0, len(genomes[g]) - rl + 1))
        r = genomes[g][start:start + rl].copy()
        if rng.random() < rc_prob:
            r = revcomp(r)
        if erate > 0:
            err = rng.random(rl) < erate
            if err.any():
                codes = np.searchsorted(ALPHABET, r[err])
                r[err] = ALPHABET[(codes + rng.integers(1, 4, size=int(err.sum()))) & 3]
        if n_rate > 0:
            r[rng.random(rl) < n_rate] = ord("N")
        if lower_frac > 0 and rng.random() < lower_frac:
            r = np.frombuffer(r.tobytes().lower(), dtype=np.uint8).copy()
        reads.append(r.tobytes())
        src.append(g)
    return reads, src


def write_fasta_set(out_dir, genomes, contigs_per_genome=1, line_len=80):
    """Write g<i>.fna per genome + genome_map.out (fasta, id, taxid, name).  Returns map path."""
    os.makedirs(out_dir, exist_ok=True)
    lines = []
    for i, g in enumerate(genomes, start=1):
        fn = "genome_%04d.fna" % i
        with open(os.path.join(out_dir, fn), "wb") as f:
            bounds = np.linspace(0, len(g), contigs_per_genome + 1).astype(int)
            for c in range(contigs_per_genome):
                f.write(b">g%d_contig%d\n" % (i, c))
                seq = g[bounds[c]:bounds[c + 1]].tobytes()
                for k in range(0, len(seq), line_len):
                    f.write(seq[k:k + line_len] + b"\n")
        lines.append("%s\t%d\t%d\tSynthetic organism %d\n" % (fn, i, 1000 + i, i))
    map_fn = os.path.join(out_dir, "genome_map.out")
    with open(map_fn, "w") as f:
        f.writelines(lines)
    return map_fn


def write_fastq(path, reads):
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            f.write(b"@read%d\n" % i + r + b"\n+\n" + b"I" * len(r) + b"\n")


def read_fastq(path):
    out = []
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    for i in range(1, len(lines), 4):
        if i < len(lines) and (i + 2) < len(lines):
            out.append(lines[i])
    return out


def have_reference():
    return os.access(CAMMIQ_REF, os.X_OK) and os.access(REF_HARNESS, os.X_OK)


def build_reference_index(fasta_dir, map_fn, idx_dir, k=26, L=100, Lmax=50, h=26, threads=2,
                          option="both"):
    """Run the UNMODIFIED reference builder (oracle/_ref/cammiq_ref --build)."""
    os.makedirs(idx_dir, exist_ok=True)
    if not fasta_dir.endswith("/"):
        fasta_dir += "/"
    cmd = [CAMMIQ_REF, "--build", "--" + option, "-k", str(k), "-L", str(L), "-Lmax", str(Lmax),
           "-h", str(h), "-f", map_fn, "-D", fasta_dir,
           "-i", os.path.join(idx_dir, "index_u.bin1"), os.path.join(idx_dir, "index_d.bin2"),
           "-t", str(threads)]
    # the builder drops temp files (gsa.bin, sa0.bin, lcp.bin) in its CWD (gsa.cpp:88,195,811)
    res = subprocess.run(cmd, cwd=idx_dir, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("reference build failed:\n" + res.stderr[-2000:])
    for tmp in ("gsa.bin", "sa0.bin", "lcp.bin"):
        p = os.path.join(idx_dir, tmp)
        if os.path.exists(p):
            os.remove(p)
    return os.path.join(idx_dir, "index_u.bin1"), os.path.join(idx_dir, "index_d.bin2")


def run_ref_dump(idx_u, idx_d, map_fn, mode, fastqs, out_fn, threads=1, per_read_n=0):
    cmd = [REF_HARNESS, "dump", idx_u, idx_d, map_fn, mode, str(threads), str(per_read_n), out_fn]
    cmd += list(fastqs)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("ref_harness failed:\n" + res.stderr[-2000:])
    return res.stderr


def parse_ref_dump(path):
    """Parse an oracle/ref_harness dump into a dict (see oracle/ref_harness.cpp header)."""
    d = {"files": [], "reads": [], "leaf_u": {}, "leaf_d": {}}
    cur = None
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            tag = t[0]
            if tag == "H":
                d["h"] = int(t[1])
            elif tag in ("NU", "ND", "G"):
                d[tag.lower()] = int(t[1])
            elif tag == "MODE":
                d["mode"] = t[1]
            elif tag == "TAXID":
                d["taxid"] = [int(x) for x in t[1:]]
            elif tag in ("LEAFU", "LEAFD"):
                leaves = [tuple(int(v) for v in x.split(":")) for x in t[3:]]
                d["leaf_u" if tag == "LEAFU" else "leaf_d"][int(t[1])] = leaves
            elif tag == "FILE":
                cur = {"name": t[1], "nreads": int(t[2]), "rcu": {}, "rcd": {}, "pairs": {}}
                d["files"].append(cur)
            elif tag == "NUNDET":
                cur["nundet"] = int(t[1])
            elif tag == "NCONF":
                cur["nconf"] = int(t[1])
            elif tag == "CU":
                cur["cu"] = [int(x) for x in t[1:]]
            elif tag == "CD":
                cur["cd"] = [int(x) for x in t[1:]]
            elif tag in ("RCU", "RCD"):
                cur["rcu" if tag == "RCU" else "rcd"][int(t[1])] = [int(x) for x in t[3:]]
            elif tag == "PAIRS":
                for x in t[2:]:
                    a, b, c = (int(v) for v in x.split(":"))
                    cur["pairs"][(a, b)] = c
            elif tag == "READ":
                iu, idd, ilu, ild = t.index("U"), t.index("D"), t.index("LU"), t.index("LD")
                d["reads"].append({
                    "idx": int(t[1]), "nundet": int(t[2]), "nconf": int(t[3]),
                    "u": [int(x) for x in t[iu + 1:idd]],
                    "d": [int(x) for x in t[idd + 1:ilu]],
                    "lu": [int(x) for x in t[ilu + 1:ild]],
                    "ld": [int(x) for x in t[ild + 1:]],
                })
    return d
