"""Without a usable sm_100 device the drop-in CLI must stop with an error that says so: there
is no CPU path behind it (the reference binary is the CPU implementation)."""
import os
import subprocess

import pytest

from golden_util import GOLD

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(REPO, "cammiq_b200", "cammiq")


def gpu_present():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(gpu_present(), reason="a GPU is visible: covered by the gpu tests")
def test_cli_refuses_to_run_without_a_gpu(tmp_path):
    if not os.access(CLI, os.X_OK):
        subprocess.check_call(["make", "-C", os.path.join(REPO, "cammiq_b200", "csrc"), "../cammiq"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    d = os.path.join(GOLD, "cfg1_small")
    res = subprocess.run([CLI, "--query", "--read_cnts", "-f", os.path.join(d, "genome_map.out"), "-q",
                          os.path.join(d, "reads.fq"), "-i", os.path.join(d, "index_u.bin1"),
                          os.path.join(d, "index_d.bin2"), "-o", str(tmp_path / "o.out")],
                         capture_output=True, text=True)
    assert res.returncode != 0
    assert "Cannot create the GPU context" in res.stderr and "no CPU fallback" in res.stderr
    assert not os.path.exists(str(tmp_path / "o.out"))
