"""Device-resident rt_query_boxes on BASELINE config 2 for a sweep of the work-distribution knobs
(RT_OPT_WORK_RANGES x RT_OPT_QUERY_CHUNK), then the end-to-end pipeline batch sweep with the best."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
rtb = importlib.import_module("r-star-tree_b200")
W, capi = rtb.workloads, rtb.capi
N = int(os.environ.get("PROFILE_N", 10_000_000)); Q = int(os.environ.get("PROFILE_Q", 16_000_000))
t0 = time.time()
tree = rtb.RTree(4).insert(W.uniform_points(N, seed=42)).flatten_device().upload(0)
print("built + uploaded N=%d in %.1f s" % (N, time.time() - t0), flush=True)
hq = torch.from_numpy(W.cube_queries(Q, N, seed=43)).pin_memory()
dq = hq.cuda()
flags = capi.RT_QUERIES_ON_DEVICE | capi.RT_RESULT_ON_DEVICE
best = (1e9, 0, 0)
for ranges in (1, 0, 148, 296, 1184):
    for chunk in (8, 4, 2):
        tree.set_option(capi.RT_OPT_WORK_RANGES, ranges)
        tree.set_option(capi.RT_OPT_QUERY_CHUNK, chunk)
        ms = []
        for i in range(4):
            out = tree.query_boxes_raw(dq.data_ptr(), Q, capi.RT_POINT_IN_BOX, flags)
            ms.append(tree.stats()["traverse_ms"])
        print("ranges %4d chunk %2d: traverse ms %s (hits %d)" % (ranges, chunk, " ".join("%.2f" % m for m in ms), out.total), flush=True)
        if ranges != 1 and min(ms[1:]) < best[0]:
            best = (min(ms[1:]), ranges, chunk)
print("best", best, flush=True)
tree.set_option(capi.RT_OPT_WORK_RANGES, best[1])
tree.set_option(capi.RT_OPT_QUERY_CHUNK, best[2])
for batch in [int(x) for x in os.environ.get("SWEEP", "0,1000000,2000000,4000000,8000000").split(",")]:
    tree.set_option(capi.RT_OPT_PIPELINE_BATCH, batch)
    times = []
    for i in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = tree.query_boxes_raw(hq.data_ptr(), Q, capi.RT_POINT_IN_BOX, 0)
        times.append((time.perf_counter() - t0) * 1e3)
    s = tree.stats()
    print("batch %9d: wall ms %s | engine total %.2f ms, batches %d, sum h2d %.2f reorder %.2f traverse %.2f "
          "finish %.2f d2h %.2f" % (batch, " ".join("%.1f" % t for t in times), s["total_ms"], s["batches"], s["h2d_ms"],
                                    s["reorder_ms"], s["traverse_ms"], s["finish_ms"], s["d2h_ms"]), flush=True)
