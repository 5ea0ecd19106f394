"""Corrupted index files must be rejected (or decoded) without crashing: the decoder computes
INT-stream offsets from the AUX shape, so every mutation below goes through its bounds checks.
Each case runs in a child process so that a crash shows up as a failed assertion."""
import os
import subprocess
import sys

import numpy as np
import pytest

from golden_util import GOLD

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys
sys.path.insert(0, %r)
import cammiq_b200 as cq
ok = bad = 0
for i in range(int(sys.argv[2])):
    base = "%%s/m%%d" %% (sys.argv[1], i)
    try:
        idx = cq.Index(base + ".bin1", base + ".bin2")
        n = idx.info.n_keys          # decoded: the flattened layout must be usable
        idx.find_host(cq.TABLE_U, 12345, b"ACGTACGT")
        ok += 1
    except cq.CammiqError as e:
        assert e.code in (-3, -2), e
        bad += 1
print("decoded", ok, "rejected", bad)
""" % REPO


def mutate(rng, data, kind):
    b = bytearray(data)
    if kind == "truncate":
        return bytes(b[:int(rng.integers(0, len(b)))])
    if kind == "flip":
        for _ in range(int(rng.integers(1, 6))):
            b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
        return bytes(b)
    if kind == "zero_run":
        a = int(rng.integers(0, len(b)))
        n = int(rng.integers(1, 64))
        b[a:a + n] = bytes(len(b[a:a + n]))
        return bytes(b)
    if kind == "ones_run":
        a = int(rng.integers(0, len(b)))
        n = int(rng.integers(1, 64))
        b[a:a + n] = b"\xff" * len(b[a:a + n])
        return bytes(b)
    raise AssertionError(kind)


@pytest.mark.parametrize("case", ["cfg1_small", "deep_h12"])
def test_mutated_indices_never_crash_the_loader(tmp_path, case):
    d = os.path.join(GOLD, case)
    files = {ext: open(os.path.join(d, name), "rb").read() for ext, name in (
        (".bin1", "index_u.bin1"), (".bin1.aux", "index_u.bin1.aux"),
        (".bin2", "index_d.bin2"), (".bin2.aux", "index_d.bin2.aux"))}
    rng = np.random.default_rng(len(case))
    n = 80
    kinds = ["truncate", "flip", "zero_run", "ones_run"]
    for i in range(n):
        victim = list(files)[int(rng.integers(0, 4))]
        for ext, data in files.items():
            out = mutate(rng, data, kinds[i % 4]) if ext == victim else data
            with open(str(tmp_path / ("m%d%s" % (i, ext))), "wb") as f:
                f.write(out)
    res = subprocess.run([sys.executable, "-c", CHILD, str(tmp_path), str(n)], capture_output=True, text=True)
    assert res.returncode == 0, (res.returncode, res.stdout[-500:], res.stderr[-1500:])
    assert "rejected" in res.stdout
