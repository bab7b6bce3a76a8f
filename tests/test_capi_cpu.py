"""CPU tests of the C-ABI boundary: the CUDA library loads without a GPU, exports every symbol
include/rtree_b200.h declares, its structs have the documented layout, and the product fails
LOUDLY (no fallback) when no device is present.  No compute calls are made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest


def declared_symbols(header_path):
    text = open(header_path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    inline = set(re.findall(r"static\s+inline\s+\w+\s+(rt_[a-z_]+)\s*\(", text))  # header-only helpers
    return sorted(set(re.findall(r"\b(rt_[a-z_]+)\s*\(", text)) - inline)


def test_library_is_built(rtb):
    assert rtb.capi.have_library(), "run __graft_entry__.build()"
    assert rtb.host.have_library()


def test_every_declared_symbol_is_exported(rtb):
    names = declared_symbols(rtb.capi.HEADER_PATH)
    assert {"rt_upload", "rt_query_boxes", "rt_query_rays", "rt_free"} <= set(names)
    assert sorted(names) == sorted(rtb.capi.EXPORTS)
    lib = C.CDLL(rtb.capi.LIB_PATH)
    for name in names:
        assert hasattr(lib, name), name


def test_abi_version_and_struct_sizes(rtb):
    capi = rtb.capi
    assert capi.lib().rt_abi_version() == capi.RT_ABI_VERSION == 2
    # x86-64 layouts of the structs in include/rtree_b200.h
    assert C.sizeof(capi.RtTreeDesc) == 88
    assert C.sizeof(capi.RtHits) == 48
    assert C.sizeof(capi.RtRayResult) == 96
    assert C.sizeof(capi.RtStats) == 80
    assert C.sizeof(capi.RtKnnResult) == 32


def test_header_compiles_as_plain_c(rtb, tmp_path):
    src = tmp_path / "abi.c"
    src.write_text('#include <rtree_b200.h>\n#include <stddef.h>\n'
                   'int sizes[5] = { sizeof(rt_tree_desc), sizeof(rt_hits), sizeof(rt_ray_result), sizeof(rt_stats),\n'
                   '                 sizeof(rt_knn_result) };\n'
                   'int main(void) { return sizes[0] == 88 && sizes[1] == 48 && sizes[2] == 96 && sizes[3] == 80\n'
                   '                        && sizes[4] == 32 ? 0 : 1; }\n')
    exe = tmp_path / "abi"
    inc = os.path.dirname(rtb.capi.HEADER_PATH)
    assert os.system("/usr/bin/gcc -std=c99 -Wall -Werror -I%s -o %s %s" % (inc, exe, src)) == 0
    assert os.system(str(exe)) == 0


def _no_gpu(rtb):
    return rtb.capi.lib().rt_device_count() <= 0


def test_fails_loudly_without_a_device(rtb):
    if not _no_gpu(rtb):
        pytest.skip("a CUDA device is present")
    image = rtb.RTree(4).insert(np.random.default_rng(0).random((100, 3)).astype(np.float32)).flatten_device()
    with pytest.raises(rtb.RtError) as err:
        image.upload(0)
    assert err.value.code == rtb.capi.RT_E_CUDA
    assert "CUDA" in str(err.value) or "device" in str(err.value)


def test_argument_validation(rtb):
    capi = rtb.capi
    lib = capi.lib()
    handle = C.c_void_p()
    assert lib.rt_upload(None, 0, C.byref(handle)) == capi.RT_E_INVALID
    image = rtb.RTree(4).insert(np.random.default_rng(0).random((100, 3)).astype(np.float32)).flatten_device()
    bad = capi.RtTreeDesc.from_buffer_copy(image.desc)
    bad.abi_version = 99
    assert lib.rt_upload(C.byref(bad), 0, C.byref(handle)) == capi.RT_E_INVALID
    bad = capi.RtTreeDesc.from_buffer_copy(image.desc)
    bad.width = 12
    assert lib.rt_upload(C.byref(bad), 0, C.byref(handle)) == capi.RT_E_UNSUPPORTED
    bad = capi.RtTreeDesc.from_buffer_copy(image.desc)
    bad.blocks_bytes = 16
    assert lib.rt_upload(C.byref(bad), 0, C.byref(handle)) == capi.RT_E_INVALID
    assert b"blocks_bytes" in lib.rt_last_error()
    out = capi.RtHits()
    assert lib.rt_query_boxes(None, None, 0, 0, 0, C.byref(out)) == capi.RT_E_INVALID
    lib.rt_free(None)  # documented no-op


def test_missing_library_is_an_error_not_a_fallback(rtb, monkeypatch):
    capi = rtb.capi
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", capi.LIB_PATH + ".absent")
    with pytest.raises(rtb.RtError) as err:
        capi.lib()
    assert "no CPU fallback" in str(err.value)


def test_batch_planner_never_divides_by_zero_and_covers_the_call(rtb):
    """the batch planner of pipelined calls (host arithmetic, no device): every plan covers exactly the n
    queries, tiny or absurd RT_OPT_PIPELINE_BATCH values (the environment override clamps them to 4096
    instead of refusing them) neither crash nor explode into millions of batches, and the documented
    ramp B/2, B, 2 B, ~4 B ..., 2 B, B, B/2 comes out for the headline call"""
    plan = rtb.capi.plan_batches
    for n in (0, 1, 5, 6, 4095, 4096, 8192, 47003, 1 << 20, 16_000_000, 3_999_999_999):
        for b in (0, 1, 2, 7, 4095, 4096, 5000, 1 << 20, 1 << 40):
            sizes = plan(n, b)
            assert sum(sizes) == n and len(sizes) >= 1
            assert len(sizes) <= 1100
            assert all(s > 0 for s in sizes) or n == 0
            if b == 0:
                assert sizes == [n]
    mi = 1 << 20
    assert plan(16_000_000, mi)[:3] == [mi // 2, mi, 2 * mi] and plan(16_000_000, mi)[-3:] == [2 * mi, mi, mi // 2]
    assert len(plan(16_000_000, mi)) == 9
    assert plan(16_000_000, 1) == plan(16_000_000, 4096)          # clamped, not divided by
    with pytest.raises(rtb.RtError):
        plan(1 << 32, mi)


def test_cpp_example_links_and_fails_loudly_without_a_device(rtb, tmp_path):
    """examples/range_and_knn.cpp (host header + C ABI from C++) compiles and links
    against the in-tree library on a CPU box; run without a GPU it must stop at rt_upload with the library's error
    text, not fall back to anything (tests/test_gpu_cpp_example.py runs it for real)"""
    import os
    import subprocess
    if not _no_gpu(rtb):
        pytest.skip("a CUDA device is present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "r-star-tree_b200", "csrc")
    exe = str(tmp_path / "range_and_knn")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror",
                    "-I", os.path.join(root, "r-star-tree_b200", "include"), "-I", os.path.join(root, "include"),
                    os.path.join(root, "examples", "range_and_knn.cpp"), "-L", csrc, "-lrtree_b200",
                    "-Wl,-rpath," + csrc, "-o", exe], check=True)
    out = subprocess.run([exe, "2000", "100"], capture_output=True, text=True, timeout=120)
    assert out.returncode != 0
    assert "rt_upload" in out.stdout + out.stderr
    assert "mismatching" not in out.stdout          # it never got as far as answering anything
