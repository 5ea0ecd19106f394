"""include/cammiq_gpu.h is a plain C header: examples/minimal.c must compile as strict C99 against
it and link with libcammiq_gpu.so alone.  Without a GPU the program stops at cq_ctx_create with
the no-CPU-fallback error; on a GPU it classifies reads of a golden case like the oracle."""
import os
import subprocess

import pytest

import oracle_lib as ol
from golden_util import load_case

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp_path):
    exe = str(tmp_path / "minimal")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic",
                           "-I" + os.path.join(REPO, "include"), os.path.join(REPO, "examples", "minimal.c"),
                           "-L" + os.path.join(REPO, "cammiq_b200"), "-lcammiq_gpu",
                           "-Wl,-rpath," + os.path.join(REPO, "cammiq_b200"), "-o", exe])
    return exe


def gpu_present():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(gpu_present(), reason="a GPU is visible: see the gpu test below")
def test_c_example_builds_and_fails_loudly_without_gpu(tmp_path):
    c = load_case("cfg1_small")
    res = subprocess.run([build(tmp_path), c["iu"], c["id"], str(c["G"]), "ACGT" * 20], capture_output=True, text=True)
    assert res.returncode == 1
    assert "leaves" in res.stdout and "no CPU fallback" in res.stderr


@pytest.mark.gpu
def test_c_example_classifies_like_the_oracle(tmp_path):
    c = load_case("cfg1_small")
    reads = [r for r in c["reads"][:40]]
    res = subprocess.run([build(tmp_path), c["iu"], c["id"], str(c["G"])] + [r.decode() for r in reads],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    b, o, l = ol.pack_reads(reads)
    want = ol.oracle_query(ol.OracleIndex(c["iu"]), ol.OracleIndex(c["id"]), ol.MODE_P, c["G"], b, o, l, per_read=True)
    got = [line.split() for line in res.stdout.split("\n") if line.startswith("read ")]
    assert len(got) == len(reads)
    for i, f in enumerate(got):
        assert (int(f[3].rstrip(",")), int(f[5]), int(f[6])) == (
            int(want["read_class"][i]), int(want["read_rid_a"][i]), int(want["read_rid_b"][i])), i
    assert "unlabeled %d, conflicting %d" % (want["nundet"], want["nconf"]) in res.stdout
