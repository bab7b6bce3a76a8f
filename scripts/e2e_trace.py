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
def timed_calls(label, trace_last=True):
    best = 1e9
    for i in range(4):
        if i == 3 and trace_last:
            os.environ["RTB_TRACE_BATCHES"] = "1"
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = tree.query_boxes_raw(hq.data_ptr(), Q, capi.RT_POINT_IN_BOX, 0)
        wall = (time.perf_counter() - t0) * 1e3
        os.environ.pop("RTB_TRACE_BATCHES", None)
        engine = tree.stats()["total_ms"]
        best = min(best, engine) if i else best
        print("%s call %d: wall %.2f ms, engine %.2f ms" % (label, i, wall, engine), flush=True)
    return best

# PLANS: ';'-separated RTB_BATCH_PLAN strings (sizes in units of 1024 queries), each timed like a SWEEP entry
if os.environ.get("PLANS"):
    results = []
    for plan in os.environ["PLANS"].split(";"):
        os.environ["RTB_BATCH_PLAN"] = ",".join(x if x.startswith("*") else str(int(float(x) * 1024)) for x in plan.split(","))
        results.append((timed_calls("plan " + plan, trace_last=os.environ.get("TRACE_PLANS") == "1"), plan))
    os.environ.pop("RTB_BATCH_PLAN", None)
    results.append((timed_calls("default plan", trace_last=False), "default"))
    for ms, plan in sorted(results):
        print("best of 3: %.2f ms  %s" % (ms, plan))
    sys.exit(0)
for batch in [int(x) for x in os.environ.get("SWEEP", "1048576").split(",")]:
    tree.set_option(capi.RT_OPT_PIPELINE_BATCH, batch)
    timed_calls("batch %d" % batch)
