"""GPU micro-tests of the Blackwell packed-fp32 building blocks the BP kernel relies on (compiled with nvcc on the box):
* FMUL2 / FFMA2 / FADD2 halves are bit-identical to scalar round-to-nearest operations (denormals, NaN, zero);
* the packed helpers of bp_kernel.cuh (division refinement, 1-r, check-node chain) equal their scalar forms --
  this is the guard against ptxas contracting mul.rn.f32x2 + add.rn.f32x2 into one FFMA2."""
import os
import re
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_and_run(name, tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / name)
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.run([nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                    os.path.join(ROOT, "tools", "micro", name + ".cu")], check=True, env=env)
    return subprocess.run([exe], check=True, capture_output=True, text=True).stdout


def test_packed_fp32x2_halves_are_ieee(tmp_path):
    out = build_and_run("f32x2_exact", tmp_path)
    m = re.search(r"mul2 (\d+)\s+fma2 (\d+)\s+add2 (\d+).*seen: (\d+)", out)
    assert m and [int(m.group(i)) for i in (1, 2, 3)] == [0, 0, 0] and int(m.group(4)) > 1000, out


def test_packed_helpers_equal_scalar_forms(tmp_path):
    out = build_and_run("pack_vs_scalar", tmp_path)
    m = re.search(r"div (\d+)\s+one_minus (\d+)\s+check_chain (\d+)", out)
    assert m and [int(m.group(i)) for i in (1, 2, 3)] == [0, 0, 0], out
