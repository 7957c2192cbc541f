"""The CTA reductions of csrc/hmx_cell_common.cuh (``warp_sum_multi``: transposing butterfly for several values per
thread; ``block_partials`` / ``block_sum``) on the CPU fiber emulation, for 1 .. 32 values per thread, in three
thread schedules."""
import os
import subprocess

import pytest

from hommx_b200 import native

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "cpu_emu")


@pytest.fixture(scope="module")
def binary(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("bsum") / "block_sum_check")
    cmd = ["g++", "-O1", "-std=c++17", "-DHMX_EMULATE", "-I", EMU, "-I", native.CSRC, os.path.join(EMU, "block_sum_check.cpp"), "-o", out, "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    return out


@pytest.mark.parametrize("order", ["forward", "reverse", "shuffle:7"])
def test_block_sums_match_plain_sums(binary, order):
    r = subprocess.run([binary], capture_output=True, text=True, env={**os.environ, "HMX_EMU_ORDER": order, "HMX_EMU_POISON": "1"})
    assert r.returncode == 0, r.stdout + r.stderr
    assert "failures 0" in r.stdout
