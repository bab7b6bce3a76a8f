"""The quantiser the kernels use (r-star-tree_b200/csrc/rt_quant.cuh), compiled for the host from that very
header and checked on millions of random and adversarial floats for the properties the exactness argument of
DESIGN.md 3.1 rests on: monotone, a conservative filter for the reference's closed overlap and point-in-box
tests, strictly-inside-on-the-grid implies the reference's is_inside, empty slots never pass, NaN bounds
constrain nothing, and the packed 8-D leaf record derives to the record that would have been stored
(tests/cpp/quant_properties.cpp).  CPU only: no device, no kernel -- the GPU parity tests are the end-to-end
check, this one pins the arithmetic underneath them."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INCLUDE = "/usr/local/cuda/include"
CXX = shutil.which("g++")


@pytest.mark.skipif(CXX is None or not os.path.exists(os.path.join(CUDA_INCLUDE, "cuda_runtime.h")),
                    reason="g++ or the CUDA headers are not present")
def test_quantiser_properties_on_the_host(tmp_path):
    exe = str(tmp_path / "quant_properties")
    subprocess.run([CXX, "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-Wno-unknown-pragmas", "-Werror",
                    "-I", CUDA_INCLUDE, "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(ROOT, "r-star-tree_b200", "csrc"), "-x", "c++",
                    os.path.join(ROOT, "tests", "cpp", "quant_properties.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe, "20000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    lines = [line for line in out.stdout.splitlines() if line.startswith("dim ")]
    assert [line.split(":")[0] for line in lines] == ["dim 2", "dim 3", "dim 8"]
    assert "VIOLATION" not in out.stdout
