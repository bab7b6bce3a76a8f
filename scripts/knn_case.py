"""kNN timing on the BASELINE config 2 tree: device-resident rt_query_knn for a few k, next to the
reference's search() driven by the same functors (oracle/_ref) on a sample of the queries."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
rtb = importlib.import_module("r-star-tree_b200")
W, capi = rtb.workloads, rtb.capi
N = int(os.environ.get("PROFILE_N", 10_000_000)); Q = int(os.environ.get("PROFILE_Q", 16_000_000))
pts = W.uniform_points(N, seed=42)
t0 = time.time()
tree = rtb.RTree(4).insert(pts).flatten_device().upload(0)
print("built + uploaded N=%d in %.1f s" % (N, time.time() - t0), flush=True)
q = W.uniform_points(Q, seed=49)
dq = torch.from_numpy(q).cuda()
flags = capi.RT_QUERIES_ON_DEVICE | capi.RT_RESULT_ON_DEVICE
for k in (1, 8, 32):
    tree.query_knn_raw(dq.data_ptr(), Q, k, flags | capi.RT_COLLECT_VISITS)
    v = tree.stats()
    for i in range(3):
        tree.query_knn_raw(dq.data_ptr(), Q, k, flags)
        s = tree.stats()
    print("k=%2d: reorder %.2f ms, traverse %.2f ms -> %.1f M queries/s; visits/query internal %.1f leaf %.1f, "
          "%.0f GB/s algorithmic" % (k, s["reorder_ms"], s["traverse_ms"], Q / s["total_ms"] / 1e3,
                                     v["visits_internal"] / Q, v["visits_leaf"] / Q,
                                     v["algorithmic_bytes"] / s["traverse_ms"] / 1e6), flush=True)
if os.environ.get("CPU", "1") == "1":
    import oracles as O
    t0 = time.time()
    ref = O.RefTree(4).insert(pts)
    print("reference tree built in %.1f s" % (time.time() - t0), flush=True)
    n = 200_000
    sec = ref.search_knn(q[:n], 8, threads=O.ref_max_threads(), timed_only=True)
    print("reference search() kNN k=8: %d queries on %d threads in %.2f s -> %.3f M queries/s"
          % (n, O.ref_max_threads(), sec, n / sec / 1e6), flush=True)
    got = tree.query_knn(q[:20000], 8)
    want = ref.search_knn(q[:20000], 8)
    print("parity on 20000 queries: %s" % bool(np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])), flush=True)
