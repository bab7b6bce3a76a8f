"""End-to-end (host buffers in, host CSR out) time of rt_query_boxes on BASELINE config 2 for a
sweep of RT_OPT_PIPELINE_BATCH values; prints one line per setting (wall clock around the call,
which is synchronous for host results, and the engine's own event time)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
rtb = importlib.import_module("r-star-tree_b200")
W, capi = rtb.workloads, rtb.capi
N = int(os.environ.get("PROFILE_N", 10_000_000)); Q = int(os.environ.get("PROFILE_Q", 16_000_000))
BATCHES = [int(x) for x in os.environ.get("SWEEP", "0,1048576,2097152,4194304,8000000").split(",")]
t0 = time.time()
tree = rtb.RTree(4).insert(W.uniform_points(N, seed=42)).flatten_device().upload(0)
print("built + uploaded N=%d in %.1f s" % (N, time.time() - t0), flush=True)
hq = torch.from_numpy(W.cube_queries(Q, N, seed=43)).pin_memory()
for batch in BATCHES:
    tree.set_option(capi.RT_OPT_PIPELINE_BATCH, batch)
    times = []
    for i in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = tree.query_boxes_raw(hq.data_ptr(), Q, capi.RT_POINT_IN_BOX, 0)
        times.append((time.perf_counter() - t0) * 1e3)
    s = tree.stats()
    print("batch %9d: wall ms %s | engine total %.2f ms, batches %d, sum h2d %.2f reorder %.2f traverse %.2f "
          "finish %.2f d2h %.2f, launches %d, total hits %d"
          % (batch, " ".join("%.1f" % t for t in times), s["total_ms"], s["batches"], s["h2d_ms"], s["reorder_ms"],
             s["traverse_ms"], s["finish_ms"], s["d2h_ms"], s["kernel_launches"], out.total), flush=True)
