"""The drop-in `cammiq --query` CLI (host C++ over the C ABI) against the UNMODIFIED reference
CLI: committed expectations (tests/golden/*/ref_cli_*) and, when oracle/_ref is present, a live
multi-file run.  Output files must be byte-identical, the counter lines of stderr equal."""
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
import synth
from golden_util import GOLD, golden_cases, load_case

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(REPO, "cammiq_b200", "cammiq")
KEEP = re.compile(r"^(Querying|Number of unlabeled|Number of reads with conflict|Completed query|Hash Length)")


@pytest.fixture(scope="module", autouse=True)
def built_cli():
    if not os.access(CLI, os.X_OK):
        subprocess.check_call(["make", "-C", os.path.join(REPO, "cammiq_b200", "csrc"), "../cammiq"])


def run_cli(args, exe=CLI):
    res = subprocess.run([exe] + args, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    return [l for l in res.stderr.replace("\r", "\n").split("\n") if KEEP.match(l)]


@pytest.mark.parametrize("case", golden_cases())
def test_cli_matches_committed_reference_outputs(case, tmp_path):
    d = os.path.join(GOLD, case)
    base = ["--query", "-f", os.path.join(d, "genome_map.out"), "-q", os.path.join(d, "reads.fq"),
            "-i", os.path.join(d, "index_u.bin1"), os.path.join(d, "index_d.bin2")]
    want = {"read_cnts": [], "standard": []}
    for line in open(os.path.join(d, "ref_cli_stderr.txt")):
        tag, text = line.rstrip("\n").split("\t", 1)
        want[tag].append(text)
    out = str(tmp_path / "cnts.out")
    got = run_cli(base[:1] + ["--read_cnts"] + base[1:] + ["-o", out])
    assert got == want["read_cnts"]
    assert open(out, "rb").read() == open(os.path.join(d, "ref_cli_read_cnts.out"), "rb").read()
    got = run_cli(base + ["-o", str(tmp_path / "unused.out")])
    assert got == want["standard"]


@pytest.mark.skipif(not synth.have_reference(), reason="oracle/_ref not present")
def test_cli_multi_file_directory_and_filter_live(tmp_path):
    """Several FASTQ files through -Q (directory), --read_length_filter, counters reset between
    files, output appended per file -- against the reference CLI run on the same inputs."""
    c = load_case("cfg1_small")
    qdir = tmp_path / "q"
    qdir.mkdir()
    reads = c["reads"]
    synth.write_fastq(str(qdir / "sample_a.fq"), reads[:700])
    synth.write_fastq(str(qdir / "sample_b.fastq"), reads[700:1200])
    synth.write_fastq(str(qdir / "sample_c.fq"), reads[1200:])
    outs = {}
    for name, exe in (("ref", synth.CAMMIQ_REF), ("gpu", CLI)):
        out = str(tmp_path / (name + ".out"))
        args = ["--query", "--read_cnts", "--read_length_filter", "40", "-f", c["map"], "-Q", str(qdir) + "/",
                "-i", c["iu"], c["id"], "-o", out]
        err = run_cli(args, exe)
        outs[name] = (sorted(open(out).read().split("\n")), sorted(err))
    assert outs["gpu"] == outs["ref"]


def test_cli_ilp_inputs_dump(tmp_path):
    """--dump_ilp_inputs: per-genome counters and per-leaf rcount / wcov in map_sp order equal
    what the ILP set-up of the reference reads (query.cpp:1100-1181), checked with the oracle."""
    c = load_case("cfg1_small")
    dump = str(tmp_path / "ilp.tsv")
    run_cli(["--query", "-f", c["map"], "-q", c["fq"], "-i", c["iu"], c["id"], "-e", "0.01",
             "--dump_ilp_inputs", dump])
    oi_u, oi_d = ol.OracleIndex(c["iu"]), ol.OracleIndex(c["id"])
    want = ol.oracle_query(oi_u, oi_d, ol.MODE_P, c["G"], c["bases"], c["offsets"], c["lengths"])
    rl = sum(len(r) for r in c["reads"]) // len(c["reads"])
    rows = [l.rstrip("\n").split("\t") for l in open(dump)]
    genomes = [r for r in rows if r[0] == "GENOME"]
    assert [int(r[3]) for r in genomes] == [int(x) for x in want["cnt_u"][1:]]
    assert [int(r[4]) for r in genomes] == [int(x) for x in want["cnt_d"][1:]]
    for tag, oi, rc in (("LEAFU", oi_u, want["rcount_u"]), ("LEAFD", oi_d, want["rcount_d"])):
        off, ids = oi.map_sp(c["G"])
        leaf_rows = [r for r in rows if r[0] == tag]
        assert len(leaf_rows) == len(ids)
        for r, l in zip(leaf_rows, ids):
            l = int(l)
            assert int(r[2]) == l and int(r[8]) == int(rc[l])
            w1 = float(oi.ucount1[l]) * (rl - float(oi.depth[l])) / rl * (1.0 - float(np.float32(0.01))) ** float(oi.depth[l])
            assert abs(float(r[9]) - w1) <= 1e-9 * max(1.0, abs(w1))


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("case", ["cfg1_small", "adversarial_250"])
def test_cli_two_gpus_equal_one(case, tmp_path):
    """--gpus 2: packed reads sharded over two devices, counters combined (NCCL reduce in standard
    mode, host merge of the pair maps with --read_cnts): same files as the single-GPU run."""
    c = load_case(case)
    base = ["--query", "-f", c["map"], "-q", c["fq"], "-i", c["iu"], c["id"]]
    for extra, tag in ((["--read_cnts"], "sc"), (["-e", "0.01"], "p")):
        outs = []
        for g in ("1", "2"):
            out, dump = str(tmp_path / ("%s_%s.out" % (tag, g))), str(tmp_path / ("%s_%s.tsv" % (tag, g)))
            args = base[:1] + extra + base[1:] + ["-o", out, "--gpus", g]
            if tag == "p":
                args += ["--dump_ilp_inputs", dump]
            err = run_cli(args)
            outs.append((err, open(out, "rb").read() if os.path.exists(out) else b"",
                         open(dump, "rb").read() if tag == "p" else b""))
        assert outs[0] == outs[1], (case, tag)
