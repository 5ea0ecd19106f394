"""The UNMODIFIED reference program hosting libcammiq_gpu.so.

oracle/Makefile links the reference's own main.cpp / query.cpp / hashtrie.cpp / binaryio.cpp with
cammiq_b200/csrc/host/reference_binding.cpp, whose FqReader::query64_p / query64mt_p / query64_sc
(query.hpp:113-115) call the C ABI.  Everything around the three members is the reference's code:
its loaders, its FASTQ reader, its counters, outputUniqueCnts.  Checked here:

  * oracle/_ref/ref_harness_hosted: the REFERENCE's state after a GPU scan -- Genome::read_cnts_u/d,
    nundet, nconf, every pleafNode::rcount in map_sp order, read_cnts_b, and the per-read decisions
    of the first reads -- dumped by the same harness that produced tests/golden/*/dump_{p,sc}.txt
    from the reference's CPU scan.  The dump files must be byte-identical.
  * oracle/_ref/cammiq_hosted: the reference CLI; `--query --read_cnts` output file byte-identical
    to the committed output of oracle/_ref/cammiq_ref, same stderr counter lines in both modes.
"""
import os
import re
import subprocess

import pytest

from golden_util import GOLD, golden_cases

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(REPO, "oracle", "_ref", "ref_harness_hosted")
CLI = os.path.join(REPO, "oracle", "_ref", "cammiq_hosted")
KEEP = re.compile(r"^(Querying|Number of unlabeled|Number of reads with conflict|Completed query|Hash Length)")

needs_hosted = pytest.mark.skipif(not (os.access(HARNESS, os.X_OK) and os.access(CLI, os.X_OK)),
                                  reason="oracle/_ref/*_hosted not built (needs /root/reference at build time)")


@needs_hosted
@pytest.mark.parametrize("case", golden_cases())
@pytest.mark.parametrize("mode", ["p", "mt", "sc"])
def test_reference_state_after_gpu_scan_equals_reference_dump(case, mode, tmp_path):
    d = os.path.join(GOLD, case)
    out = str(tmp_path / "dump.txt")
    threads = 4 if mode == "mt" else 1
    res = subprocess.run([HARNESS, "dump", os.path.join(d, "index_u.bin1"), os.path.join(d, "index_d.bin2"),
                          os.path.join(d, "genome_map.out"), mode, str(threads), "400" if mode != "mt" else "0", out,
                          os.path.join(d, "reads.fq")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    got = open(out).read()
    if mode == "mt":
        # query64mt_p leaves the state of query64_p (tests/golden/make_golden.py asserts it for the reference)
        want = open(os.path.join(d, "dump_p.txt")).read().split("READ ")[0].replace("MODE p", "MODE mt")
    else:
        want = open(os.path.join(d, "dump_%s.txt" % mode)).read()
    assert got == want


@needs_hosted
@pytest.mark.parametrize("case", golden_cases())
def test_reference_cli_with_gpu_scan_matches_reference_cli(case, tmp_path):
    d = os.path.join(GOLD, case)
    base = ["--query", "-f", os.path.join(d, "genome_map.out"), "-q", os.path.join(d, "reads.fq"),
            "-i", os.path.join(d, "index_u.bin1"), os.path.join(d, "index_d.bin2")]
    want = {"read_cnts": [], "standard": []}
    for line in open(os.path.join(d, "ref_cli_stderr.txt")):
        tag, text = line.rstrip("\n").split("\t", 1)
        want[tag].append(text)

    def run(args):
        res = subprocess.run([CLI] + args, capture_output=True, text=True)
        assert res.returncode == 0, res.stderr[-2000:]
        return [l for l in res.stderr.replace("\r", "\n").split("\n") if KEEP.match(l)]

    out = str(tmp_path / "cnts.out")
    assert run(base[:1] + ["--read_cnts"] + base[1:] + ["-o", out]) == want["read_cnts"]
    assert open(out, "rb").read() == open(os.path.join(d, "ref_cli_read_cnts.out"), "rb").read()
    assert run(base + ["-t", "4", "-o", str(tmp_path / "unused.out")]) == want["standard"]
