"""BASELINE config 3 timing: teapot triangle AABBs, random-direction rays, closest-hit / any-hit / all-hits,
device-resident and end to end (host rays in, host result out)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
rtb = importlib.import_module("r-star-tree_b200")
W, capi = rtb.workloads, rtb.capi
R = int(os.environ.get("PROFILE_RAYS", 16_000_000)); RES = float(os.environ.get("PROFILE_RES", 1.0))
CALLS = int(os.environ.get("PROFILE_CALLS", 3))
boxes = W.triangle_aabbs(W.teapot_triangles(RES))
t0 = time.time()
tree = rtb.RTree(5).insert(boxes).flatten_device().upload(0)
print("teapot res %.1f: %d boxes, built + uploaded in %.2f s" % (RES, len(boxes), time.time() - t0), flush=True)
t0 = time.time()
hr = torch.from_numpy(W.random_rays(R, boxes, seed=44)).pin_memory()
print("%d rays generated in %.1f s" % (R, time.time() - t0), flush=True)
dr = hr.cuda()
dev = capi.RT_QUERIES_ON_DEVICE | capi.RT_RESULT_ON_DEVICE
for name, mode in (("closest", capi.RT_RAY_CLOSEST), ("any", capi.RT_RAY_ANY), ("all", capi.RT_RAY_ALL_HITS)):
    for where, ptr, flags in (("device", dr.data_ptr(), dev), ("host", hr.data_ptr(), 0)):
        for i in range(CALLS):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = tree.query_rays_raw(ptr, R, mode, flags)
            wall = (time.perf_counter() - t0) * 1e3
            s = tree.stats()
        print("%-7s %-6s: wall %.2f ms, engine total %.2f (h2d %.2f traverse %.2f finish %.2f d2h %.2f), n_hit %d, launches %d -> %.0f M rays/s"
              % (name, where, wall, s["total_ms"], s["h2d_ms"], s["traverse_ms"], s["finish_ms"], s["d2h_ms"], out.n_hit,
                 s["kernel_launches"], R / wall / 1e3), flush=True)
