"""Traversal time against the number of persistent CTAs per SM (occupancy sensitivity), device-resident,
then the pipelined end-to-end time with a CTA slot per SM left free for the small kernels of other batches."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
rtb = importlib.import_module("r-star-tree_b200")
W, capi = rtb.workloads, rtb.capi
N = int(os.environ.get("PROFILE_N", 10_000_000)); Q = int(os.environ.get("PROFILE_Q", 16_000_000))
tree = rtb.RTree(4).insert(W.uniform_points(N, seed=42)).flatten_device().upload(0)
hq = torch.from_numpy(W.cube_queries(Q, N, seed=43)).pin_memory()
dq = hq.cuda()
flags = capi.RT_QUERIES_ON_DEVICE | capi.RT_RESULT_ON_DEVICE
tree.set_option(capi.RT_OPT_QUERY_CHUNK, 4)
for per_sm in (8, 7, 6, 5, 4):
    tree.set_option(capi.RT_OPT_TRAVERSE_GRID, 148 * per_sm)
    ms = []
    for i in range(3):
        out = tree.query_boxes_raw(dq.data_ptr(), Q, capi.RT_POINT_IN_BOX, flags)
        ms.append(tree.stats()["traverse_ms"])
    print("CTAs/SM %d: traverse ms %s" % (per_sm, " ".join("%.2f" % m for m in ms)), flush=True)
for per_sm in (8, 7, 6):
    tree.set_option(capi.RT_OPT_TRAVERSE_GRID, 148 * per_sm)
    for batch in (1000000, 2000000, 4000000):
        tree.set_option(capi.RT_OPT_PIPELINE_BATCH, batch)
        times = []
        for i in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = tree.query_boxes_raw(hq.data_ptr(), Q, capi.RT_POINT_IN_BOX, 0)
            times.append((time.perf_counter() - t0) * 1e3)
        s = tree.stats()
        print("CTAs/SM %d batch %9d: wall ms %s | engine total %.2f ms, batches %d, sum h2d %.2f reorder %.2f traverse %.2f "
              "finish %.2f d2h %.2f" % (per_sm, batch, " ".join("%.1f" % t for t in times), s["total_ms"], s["batches"], s["h2d_ms"],
                                        s["reorder_ms"], s["traverse_ms"], s["finish_ms"], s["d2h_ms"]), flush=True)
