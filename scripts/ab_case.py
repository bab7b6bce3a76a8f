"""A/B harness for the traversal kernel: BASELINE config 2 (or a cut: PROFILE_N / PROFILE_Q), device-resident
queries, PROFILE_CALLS calls of rt_query_boxes; prints the per-call phase times and the visit / step counters
on both images.  RTB_LIB picks the build of the library (csrc/Makefile `variants`); the tree image is cached in
AB_IMAGE (default /tmp/ab_config2.rtb) so that several builds can be compared on one box without re-inserting."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
rtb = importlib.import_module("r-star-tree_b200")
W, capi = rtb.workloads, rtb.capi
N = int(os.environ.get("PROFILE_N", 10_000_000)); Q = int(os.environ.get("PROFILE_Q", 16_000_000))
CALLS = int(os.environ.get("PROFILE_CALLS", 4))
path = os.environ.get("AB_IMAGE", "/tmp/ab_config2_%d.rtb" % N)
t0 = time.time()
if os.path.exists(path):
    image = rtb.DeviceImage.load(path)
else:
    image = rtb.RTree(4).insert(W.uniform_points(N, seed=42)).flatten_device()
    image.save(path)
tree = image.upload(0)
print("[%s] tree N=%d ready in %.1f s" % (os.path.basename(capi.LIB_PATH), N, time.time() - t0), flush=True)
dq = torch.from_numpy(W.cube_queries(Q, N, seed=43)).cuda()
flags = capi.RT_QUERIES_ON_DEVICE | capi.RT_RESULT_ON_DEVICE
if os.environ.get("AB_VISITS", "1") == "1":
    for quant in (1, 0):
        tree.set_option(capi.RT_OPT_QUANTISED, quant)
        out = tree.query_boxes_raw(dq.data_ptr(), Q, capi.RT_POINT_IN_BOX, flags | capi.RT_COLLECT_VISITS)
        s = tree.stats()
        print("visits/query on the %s image: internal %.3f leaf %.3f, steps %.2f, float lookups %.3f, hits %.3f"
              % ("quantised" if quant else "float", s["visits_internal"] / Q, s["visits_leaf"] / Q, s["traverse_steps"] / Q,
                 s["leaf_lookups"] / Q, out.total / Q), flush=True)
    tree.set_option(capi.RT_OPT_QUANTISED, 1)
ms = []
for i in range(CALLS):
    out = tree.query_boxes_raw(dq.data_ptr(), Q, capi.RT_POINT_IN_BOX, flags)
    s = tree.stats()
    ms.append(s["traverse_ms"])
    print("call %d: total hits %d, reorder %.3f ms, traverse %.3f ms, finish %.3f ms, total %.3f ms, launches %d, deferred %d"
          % (i, out.total, s["reorder_ms"], s["traverse_ms"], s["finish_ms"], s["total_ms"], s["kernel_launches"],
             s["deferred_queries"]), flush=True)
print("[%s] traverse ms: min %.3f median %.3f" % (os.path.basename(capi.LIB_PATH), min(ms[1:] or ms), float(np.median(ms[1:] or ms))), flush=True)
