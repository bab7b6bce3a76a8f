"""BASELINE config 4 at a cut size: 8-D clustered Gaussian mixture, MAX_ENTRIES 16 QuadraticSplit, box
queries calibrated to ~10 hits; device-resident timing + visit counts + parity against the reference's search()."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
rtb = importlib.import_module("r-star-tree_b200")
W, capi = rtb.workloads, rtb.capi
N = int(os.environ.get("PROFILE_N", 2_000_000)); Q = int(os.environ.get("PROFILE_Q", 4_000_000))
pts = W.mixture_points(N)
t0 = time.time(); h = W.calibrate_half_width(pts); print("half width %.5f (%.1f s)" % (h, time.time() - t0), flush=True)
q = W.mixture_queries(Q, pts, h)
t0 = time.time()
host = rtb.RTree(6).insert(pts)
print("host build %.1f s" % (time.time() - t0), flush=True)
tree = host.flatten_device().upload(0)
dq = torch.from_numpy(q).cuda()
flags = capi.RT_QUERIES_ON_DEVICE | capi.RT_RESULT_ON_DEVICE
tree.query_boxes_raw(dq.data_ptr(), Q, capi.RT_POINT_IN_BOX, flags | capi.RT_COLLECT_VISITS)
s = tree.stats()
print("visits/query: internal %.1f leaf %.1f, algorithmic bytes/query %.0f, steps/query %.1f"
      % (s["visits_internal"] / Q, s["visits_leaf"] / Q, s["algorithmic_bytes"] / Q, s["traverse_steps"] / Q), flush=True)
alg = s["algorithmic_bytes"]
for i in range(3):
    out = tree.query_boxes_raw(dq.data_ptr(), Q, capi.RT_POINT_IN_BOX, flags)
    s = tree.stats()
    print("call %d: hits/query %.2f, reorder %.2f ms, traverse %.2f ms, finish %.2f ms -> %.1f M q/s, %.0f GB/s algorithmic"
          % (i, out.total / Q, s["reorder_ms"], s["traverse_ms"], s["finish_ms"], Q / s["total_ms"] / 1e3,
             alg / s["traverse_ms"] / 1e6), flush=True)
if os.environ.get("PARITY", "1") == "1":
    import oracles as O
    n_par = 20000
    got = tree.query_boxes(q[:n_par])
    t0 = time.time()
    ref = O.RefTree(6).insert(pts)
    want = ref.search_boxes(q[:n_par], threads=O.ref_max_threads())
    print("parity on %d queries vs reference search(): %s (reference build + search %.1f s)"
          % (n_par, bool(np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])), time.time() - t0), flush=True)
