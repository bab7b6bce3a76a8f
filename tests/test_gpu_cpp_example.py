"""The C++ example (examples/range_and_knn.cpp) -- host header + C ABI, no Python in the data path --
is compiled with g++ against the in-tree library and run on the GPU; it compares every answer with
RTree::search() on its own host tree and fails on any difference."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_example_builds_and_agrees_with_host_search(rtb, tmp_path):
    exe = str(tmp_path / "range_and_knn")
    csrc = os.path.join(ROOT, "r-star-tree_b200", "csrc")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off",
                    "-I", os.path.join(ROOT, "r-star-tree_b200", "include"), "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "range_and_knn.cpp"), "-L", csrc, "-lrtree_b200",
                    "-Wl,-rpath," + csrc, "-o", exe], check=True)
    out = subprocess.run([exe, "100000", "20000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 mismatching hit lists" in out.stdout and "0 ids differing" in out.stdout
