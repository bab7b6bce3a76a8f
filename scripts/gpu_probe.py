"""First-contact GPU script: small parity checks against the oracles, then a mid-size timing."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracles as O
rtb = importlib.import_module("r-star-tree_b200")
W = rtb.workloads
print("cpu cores", os.cpu_count(), "ref threads", O.ref_max_threads(), flush=True)

def check(name, a, b):
    ok = all((np.asarray(x) == np.asarray(y)).all() and len(x) == len(y) for x, y in zip(a, b))
    print(("OK   " if ok else "FAIL ") + name, flush=True)
    return ok

# 1. Flatten replay (1-D int)
t = rtb.RTree(0).insert(np.arange(3000, dtype=np.int32))
dt = t.flatten_device().upload(0)
rng = np.random.default_rng(1)
a = rng.integers(1, 3000, size=(200, 2)).astype(np.int32); a.sort(axis=1)
off, ids = dt.query_boxes(a.reshape(-1, 2, 1))
ok = all((ids[off[i]:off[i+1]] == np.arange(a[i,0], a[i,1]+1)).all() for i in range(200))
print("OK   flatten replay" if ok else "FAIL flatten replay", dt.stats(), flush=True)

# 2. 3-D points vs reference search, all flag combos
N = 200000
pts = W.uniform_points(N)
ref = O.RefTree(4).insert(pts)
mine = rtb.RTree(4).insert(pts)
q = W.cube_queries(20000, N)
ro, ri = ref.search_boxes(q)
dt = mine.flatten_device().upload(0)
for flags, name in [(0, "default"), (rtb.RT_NO_REORDER, "no_reorder"), (rtb.RT_COLLECT_VISITS, "two_pass+visits")]:
    o, i = dt.query_boxes(q, flags=flags)
    check("3d points " + name, (o, i), (ro, ri))
    print("     stats", dt.stats(), flush=True)
# big queries (deferred + big sort)
qb = W.cube_queries(300, N, hits_per_query=5000.0)
ro, ri = ref.search_boxes(qb)
o, i = dt.query_boxes(qb)
check("3d big lists", (o, i), (ro, ri)); print("     stats", dt.stats(), flush=True)
o, i = dt.query_boxes(qb, flags=rtb.RT_COLLECT_VISITS)
check("3d big lists two-pass", (o, i), (ro, ri))
# whole tree
qa = np.array([[[-1,-1,-1],[2,2,2]]], np.float32)
o, i = dt.query_boxes(qa)
print("OK   whole tree" if (i == np.arange(N)).all() else "FAIL whole tree", flush=True)

# 3. rays vs reference on teapot
tris = W.teapot_triangles(1.0); boxes = W.triangle_aabbs(tris)
print("teapot triangles", len(boxes))
reft = O.RefTree(5).insert(boxes); minet = rtb.RTree(5).insert(boxes)
dtt = minet.flatten_device().upload(0)
rays = W.random_rays(20000, boxes, tmax=1e30)
ro, ri = reft.search_rays(rays, O.RAY_ALL); o, i = dtt.query_rays(rays, rtb.RT_RAY_ALL_HITS)
check("rays all-hits", (o, i), (ro, ri))
rt_, rid = reft.search_rays(rays, O.RAY_CLOSEST); t_, id_ = dtt.query_rays(rays, rtb.RT_RAY_CLOSEST)
check("rays closest", (t_.view(np.uint32), id_), (rt_.view(np.uint32), rid))
ra = reft.search_rays(rays, O.RAY_ANY); a_ = dtt.query_rays(rays, rtb.RT_RAY_ANY)
check("rays any", (a_,), (ra,)); print("     hit fraction", a_.mean(), flush=True)
qb = W.cube_queries(5000, 10, hits_per_query=1.0) * 4 - 1
ro, ri = reft.search_boxes(qb, O.MODE_INTERSECTS); o, i = dtt.query_boxes(qb, rtb.RT_INTERSECTS)
check("box keys intersects", (o, i), (ro, ri))
ro, ri = reft.search_boxes(qb, O.MODE_CONTAINS); o, i = dtt.query_boxes(qb, rtb.RT_CONTAINS)
check("box keys contains", (o, i), (ro, ri)); print("     hits", len(ri))

# 4. timing at mid size
N = int(os.environ.get("PROBE_N", 2000000)); Q = int(os.environ.get("PROBE_Q", 4000000))
pts = W.uniform_points(N)
t0 = time.time(); mine = rtb.RTree(4).insert(pts); print("build %d pts: %.1fs" % (N, time.time()-t0), flush=True)
dt = mine.flatten_device().upload(0)
q = W.cube_queries(Q, N)
for flags, name in [(rtb.RT_COLLECT_VISITS, "two_pass+visits"), (0, "default"), (0, "default"), (rtb.RT_NO_REORDER, "no_reorder"),
                    (rtb.RT_COUNT_ONLY, "count_only"), (rtb.RT_COUNT_ONLY | rtb.RT_NO_REORDER, "count_only no_reorder")]:
    o, i = dt.query_boxes(q, flags=flags)
    s = dt.stats()
    print("%-22s hits/q %.2f  total %.1f ms  trav %.1f ms  reorder %.2f finish %.2f h2d %.1f d2h %.1f  deferred %d  -> %.1f Mq/s (traverse only %.1f Mq/s)" % (
        name, len(i)/Q if len(i) else o[-1]/Q, s["total_ms"], s["traverse_ms"], s["reorder_ms"], s["finish_ms"], s["h2d_ms"], s["d2h_ms"], s["deferred_queries"],
        Q/s["total_ms"]/1e3, Q/max(s["traverse_ms"],1e-9)/1e3), flush=True)
    if s["algorithmic_bytes"]:
        print("     visits/q int %.2f leaf %.2f  alg bytes/q %.0f" % (s["visits_internal"]/Q, s["visits_leaf"]/Q, s["algorithmic_bytes"]/Q), flush=True)
sec, tot = O.RefTree(4).insert(pts[:1]).time_search_boxes(q[:10])  # warm
reft = O.RefTree(4).insert(pts)
sec, tot = reft.time_search_boxes(q[:200000], threads=1); print("ref 1 thread: %.1f kq/s" % (200000/sec/1e3))
sec, tot = reft.time_search_boxes(q[:1000000], threads=0); print("ref %d threads: %.1f kq/s" % (O.ref_max_threads(), 1000000/sec/1e3))
