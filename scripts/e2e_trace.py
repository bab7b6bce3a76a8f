"""Timeline of one pipelined rt_query_boxes call on BASELINE config 2 (host queries in, host CSR out):
RTB_TRACE_BATCHES makes the engine print where every batch's phases lie (ms after the call's fork)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
rtb = importlib.import_module("r-star-tree_b200")
W, capi = rtb.workloads, rtb.capi
N = int(os.environ.get("PROFILE_N", 10_000_000)); Q = int(os.environ.get("PROFILE_Q", 16_000_000))
path = os.environ.get("AB_IMAGE", "/tmp/ab_config2_%d.rtb" % N)
if os.path.exists(path):
    image = rtb.DeviceImage.load(path)
else:
    image = rtb.RTree(4).insert(W.uniform_points(N, seed=42)).flatten_device()
    image.save(path)
tree = image.upload(0)
hq = torch.from_numpy(W.cube_queries(Q, N, seed=43)).pin_memory()
for batch in [int(x) for x in os.environ.get("SWEEP", "1048576").split(",")]:
    tree.set_option(capi.RT_OPT_PIPELINE_BATCH, batch)
    for i in range(4):
        if i == 3:
            os.environ["RTB_TRACE_BATCHES"] = "1"
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = tree.query_boxes_raw(hq.data_ptr(), Q, capi.RT_POINT_IN_BOX, 0)
        wall = (time.perf_counter() - t0) * 1e3
        os.environ.pop("RTB_TRACE_BATCHES", None)
        print("batch %d call %d: wall %.2f ms, engine %.2f ms" % (batch, i, wall, tree.stats()["total_ms"]), flush=True)
