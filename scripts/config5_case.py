"""BASELINE config 5 on one GPU: 3-D tree of 200 M uniform points (MAX_ENTRIES 8), the 64 M-query shard one of
8 GPUs would answer (512 M cubes / 8, ~10 hits each): device-resident and end-to-end timing, visit counts,
parity of a sample against brute force over the raw points."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
rtb = importlib.import_module("r-star-tree_b200")
W, capi = rtb.workloads, rtb.capi
N = int(os.environ.get("PROFILE_N", 200_000_000)); Q = int(os.environ.get("PROFILE_Q", 64_000_000))
with open("/proc/meminfo") as f:
    avail_gb = [int(l.split()[1]) for l in f if l.startswith("MemAvailable")][0] / 1e6
print("host memory available: %.0f GB" % avail_gb, flush=True)
if avail_gb < N * 200 / 1e9:
    raise SystemExit("not enough host memory for a %d-point build" % N)
pts = W.uniform_points(N, seed=42)
t0 = time.time()
host = rtb.RTree(4).insert(pts)
t_build = time.time() - t0
t0 = time.time()
image = host.flatten_device()
print("host build %.0f s, flatten_device %.1f s: %d nodes, %d levels, %.2f GB of blocks"
      % (t_build, time.time() - t0, image.desc.num_nodes, image.desc.num_levels, image.desc.blocks_bytes / 1e9), flush=True)
del host
tree = image.upload(0)
q = W.cube_queries(Q, N, seed=43)
hq = torch.from_numpy(q).pin_memory()
dq = hq.cuda()
flags = capi.RT_QUERIES_ON_DEVICE | capi.RT_RESULT_ON_DEVICE
tree.query_boxes_raw(dq.data_ptr(), Q, capi.RT_POINT_IN_BOX, flags | capi.RT_COLLECT_VISITS)
v = tree.stats()
print("visits/query: internal %.1f leaf %.1f, algorithmic bytes/query %.0f" % (v["visits_internal"] / Q, v["visits_leaf"] / Q,
                                                                             v["algorithmic_bytes"] / Q), flush=True)
for i in range(3):
    out = tree.query_boxes_raw(dq.data_ptr(), Q, capi.RT_POINT_IN_BOX, flags)
    s = tree.stats()
    print("device call %d: hits/query %.2f, reorder %.2f traverse %.2f finish %.2f ms -> %.1f M q/s, %.0f GB/s algorithmic"
          % (i, out.total / Q, s["reorder_ms"], s["traverse_ms"], s["finish_ms"], Q / s["total_ms"] / 1e3,
             v["algorithmic_bytes"] / s["traverse_ms"] / 1e6), flush=True)
for i in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = tree.query_boxes_raw(hq.data_ptr(), Q, capi.RT_POINT_IN_BOX, 0)
    wall = (time.perf_counter() - t0) * 1e3
    print("e2e call %d: %.1f ms (%d batches) -> %.1f M q/s" % (i, wall, tree.stats()["batches"], Q / wall / 1e3), flush=True)
import ctypes as C
off = np.ctypeslib.as_array(C.cast(out.offsets, C.POINTER(C.c_uint64)), shape=(Q + 1,))
ids = np.ctypeslib.as_array(C.cast(out.ids, C.POINTER(C.c_uint32)), shape=(int(out.total),))
ok = True
for qi in np.random.default_rng(1).integers(0, Q, 12):
    inside = np.nonzero(np.all((pts >= q[qi, 0]) & (pts <= q[qi, 1]), axis=1))[0].astype(np.uint32)
    ok = ok and np.array_equal(inside, ids[off[qi]:off[qi + 1]])
print("brute-force parity on 12 sampled queries: %s" % ok, flush=True)
