#!/usr/bin/env python
"""bench.py -- box-queries/s and rays/s of the batched R-tree query path on N B200s (BASELINE.json).

A step is one pass of the hot path over one batch of synthetic queries.  The headline line is
config 2 of BASELINE.json -- a 3-D tree of 10 M uniform points (MAX_ENTRIES 8, built on the CPU
by the host library's insert / RStarSplit / reinsert) searched by 16 M small cubes (~10 hits
each), results as per-query sorted id lists in CSR form.

  value    whole-job queries/s with the queries already resident in HBM and the CSR result
           left in HBM (rt_query_boxes with RT_QUERIES_ON_DEVICE | RT_RESULT_ON_DEVICE);
           CUDA events on the launching stream, barrier + synchronize on both sides, max
           over ranks.  Inputs (384 MB of queries + 531 MB of tree) exceed the 126 MB L2.
  e2e      the same metric through the C ABI with HOST buffers: pinned queries in, pinned
           offsets + ids out, both copies inside the timed region (the engine cuts such a call
           into batches and overlaps their copies with the traversal of their neighbours);
           `e2e.counts32` is the same with RT_RESULT_COUNTS32 (32-bit hits per query instead of
           64-bit offsets: 64 MB less D2H per 16 M queries).
  roofline the traversal kernel is bound by instruction issue, not by HBM: `frac` = warp
           instructions per second (instructions per launch from the committed ncu capture of this
           very workload, profiles/traffic.json, over the kernel time measured live here) over the
           issue peak (SMs x 4 schedulers x SM clock sampled during the run); next to it the
           HBM view the contract asks for: algorithmic bytes (SURVEY.md 8d formula on the node
           visits the engine counts with RT_COLLECT_VISITS), `algorithmic_over_hbm`, the DRAM
           `traffic` of the capture and `dram_frac`.
  cpu_baseline  the reference's own RTree::search() (oracle/_ref, the unmodified reference)
           on this box's host cores, OpenMP over a bounded prefix of the same queries.
  parity   every rank's first 200 k queries against the reference's search(), bit-exact; the
           run exits non-zero, with `value` null, when any rank differs.

Further sections, each with its own parity bit, CPU arm (N = 1) and algorithmic bytes:
  rays     config 3 on the 9.6 k-triangle teapot and on a 333 k-triangle refinement, closest-hit
           and any-hit, 64 M rays split over the GPUs
  config4  8-D Gaussian mixture, MAX_ENTRIES 16 QuadraticSplit (a cut: RTB_C4_POINTS)
  config5  strong scaling: ONE batch of 512 M queries over a replicated tree of RTB_C5_POINTS
           points (default 50 M; 200 M is the full configuration), split evenly over the GPUs
  knn      k = 8 nearest neighbours on the config-2 tree

Multi-GPU (torchrun, one rank per GPU): rank 0 builds the trees, the block buffers are
broadcast GPU-to-GPU over NCCL, every rank answers its own batch (config 2 / 4: weak scaling;
config 5 / rays: one batch split), no data-path collective; value = all ranks' queries /
max-over-ranks time.

`--impl reference` times the reference's CPU search on the same workloads (rank 0 only).
"""
import argparse
import ctypes as C
import datetime
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "box_queries_per_sec"
UNIT = "queries/s"
KIND = 4      # RTree<aabb_t<point_t<float,3>>, point_t<float,3>, uint32_t, DefaultConfig>
KIND_BOX = 5  # RTree<aabb_t<point_t<float,3>>, aabb_t<...>, uint32_t, DefaultConfig>  (triangle AABBs)
KIND_8D = 6   # 8-D points, MIN 8 / MAX 16 / REINSERT 6, QuadraticSplit
C5_CHUNK = 16_000_000  # config 5 queries are generated (and seeded) in chunks of this many


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def env_int(name, default):
    return int(float(os.environ.get(name, default)))


def host_threads():
    """cores this process may use; NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=env_int("RTB_BENCH_POINTS", 10_000_000))
    ap.add_argument("--queries", type=int, default=env_int("RTB_BENCH_QUERIES", 16_000_000))
    ap.add_argument("--cpu-sample", type=int, default=2_000_000,
                    help="queries the OpenMP CPU baseline runs (a prefix of the batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true",
                    help="skip the CPU arms and the parity checks against the reference (profiling runs)")
    ap.add_argument("--rays", type=int, default=env_int("RTB_BENCH_RAYS", 64_000_000),
                    help="rays of the config-3 section (split evenly over the GPUs); 0 = skip")
    ap.add_argument("--c4-points", type=int, default=env_int("RTB_C4_POINTS", 2_000_000),
                    help="points of the config-4 section (8-D mixture); 0 = skip")
    ap.add_argument("--c4-queries", type=int, default=env_int("RTB_C4_QUERIES", 4_000_000))
    ap.add_argument("--c5-points", type=int, default=env_int("RTB_C5_POINTS", 50_000_000),
                    help="points of the config-5 tree (200 M = the full configuration); 0 = skip")
    ap.add_argument("--c5-queries", type=int, default=env_int("RTB_C5_QUERIES", 512_000_000),
                    help="queries of the ONE config-5 batch that is split over the GPUs")
    ap.add_argument("--knn-queries", type=int, default=env_int("RTB_KNN_QUERIES", 4_000_000),
                    help="kNN queries per GPU on the config-2 tree (k = 8); 0 = skip")
    return ap.parse_args()


def workload_config(args, world):
    return {
        "workload": "BASELINE config 2: 3-D uniform points N=%d, %d cube queries/GPU (~10 hits/query), "
                    "MAX_ENTRIES=8 RStarSplit, point-in-box, sorted CSR ids" % (args.points, args.queries),
        "points": args.points,
        "queries_per_gpu": args.queries,
        "hits_per_query_target": 10,
        "tree_replicated": True,
        "query_split": "%d x %d (one batch per GPU)" % (world, args.queries),
        "l2": "inputs larger than L2 (queries %.0f MB + tree blocks > 126 MB); no flush needed"
              % (args.queries * 24 / 1e6),
    }


# ---------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = [l for t, l in self.lines if t_begin - 0.05 <= t <= t_end + 0.15] or [l for _, l in self.lines]
        sm, sm_max, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])), sm_max.append(float(parts[1])), power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(sm_max), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def ncu_capture():
    """what the committed `ncu --set full` capture of the traversal kernel on THIS workload says per
    launch (profiles/traffic.json, written by scripts/summarise_ncu.py --traffic): DRAM bytes, warp
    instructions, L1 bytes.  {} when there is none."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def host_view(ptr, count, ctype, dtype):
    if not ptr or count == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(int(count),))


def csr_prefix(out, n, p):
    """copies of offsets[0..p] and ids[0..offsets[p]) of a host rt_hits result over n queries"""
    off = host_view(out.offsets, n + 1, C.c_uint64, np.uint64)[:p + 1].copy()
    ids = host_view(out.ids, int(out.total), C.c_uint32, np.uint32)[:int(off[p])].copy()
    return off, ids


def c5_queries(chunk, count, n_points):
    """`count` (<= C5_CHUNK) queries of chunk `chunk` of the config-5 batch"""
    W = importlib.import_module("r-star-tree_b200").workloads
    return W.cube_queries(count, n_points, seed=[46, int(chunk)])


# ---------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference: the unmodified reference's RTree::search() on the host cores."""
    if rank != 0:
        return
    import oracles as O
    rtb = importlib.import_module("r-star-tree_b200")
    if not O.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_oracle.so was not built"}))
        return
    W = rtb.workloads
    threads = host_threads()
    sample = min(args.queries, max(100_000, env_int("RTB_REF_STEP_QUERIES", 1_000_000)))
    log("[reference] building the reference tree: %d inserts ..." % args.points)
    t0 = time.time()
    tree = O.RefTree(KIND).insert(W.uniform_points(args.points, seed=42))
    log("[reference] built in %.1f s; %d threads; %d queries per step" % (time.time() - t0, threads, sample))
    q = W.cube_queries(min(args.queries, 4 * sample), args.points, seed=43)
    n_chunks = max(1, len(q) // sample)
    secs, hits = [], 0
    for i in range(args.warmup + args.steps):
        chunk = q[(i % n_chunks) * sample:(i % n_chunks + 1) * sample]
        s, h = tree.time_search_boxes(chunk, threads=threads)
        if i >= args.warmup:
            secs.append(s)
            hits += h
    total = sum(secs)
    value = sample * args.steps / total
    del tree

    # the ray half of the metric: reference search() + the slab functor of oracle/slab.h, both meshes
    rays_section = None
    if args.rays > 0:
        rays_section = {"unit": "rays/s", "cores": threads}
        n_r = min(args.rays, env_int("RTB_REF_RAYS", 2_000_000))
        for mesh, res in (("teapot_9k", 1.0), ("teapot_333k", 5.9)):
            boxes = W.triangle_aabbs(W.teapot_triangles(res))
            rt = O.RefTree(KIND_BOX).insert(boxes)
            rays = W.random_rays(n_r, boxes, seed=44, tmax=1e30)
            entry = {"triangles": len(boxes), "sample_rays": n_r}
            for name, mode in (("closest_hit", O.RAY_CLOSEST), ("any_hit", O.RAY_ANY)):
                s, _ = rt.search_rays(rays, mode, threads=threads, timed_only=True)
                entry[name] = n_r / s
            rays_section[mesh] = entry

    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": "each step = RTree::search() over %d queries of the workload (the GPU arm's "
                                   "step is the whole batch of %d; throughput is per query), OpenMP "
                                   "schedule(dynamic,1024) on %d threads; %.2f hits/query"
                                   % (sample, args.queries, threads, hits / float(sample * args.steps))},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "rays": rays_section,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------
class Ctx:
    """what every section needs"""

    def __init__(self, args, torch, dist, rtb, rank, world, local_rank, device, stream):
        self.args, self.torch, self.dist, self.rtb = args, torch, dist, rtb
        self.rank, self.world, self.local_rank, self.device, self.stream = rank, world, local_rank, device, stream
        self.capi, self.W = rtb.capi, rtb.workloads
        self.want_cpu = not args.no_cpu_baseline      # parity on every rank, every N
        self.cpu_arm = self.want_cpu and world == 1   # the timed CPU baselines: N = 1 only (contract)
        self.dev_flags = rtb.capi.RT_QUERIES_ON_DEVICE | rtb.capi.RT_RESULT_ON_DEVICE
        self.parity_ok = True

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def sum_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    def timed(self, fn, steps, warmup):
        """`steps` calls of fn on the section's stream between two events, barriers on both sides;
        returns (this rank's ms, result of the last call)"""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            r = None
            for _ in range(warmup):
                r = fn()
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            for _ in range(steps):
                r = fn()
            e1.record(self.stream)
            self.barrier()
        return e0.elapsed_time(e1), r

    def gather(self, obj):
        """every rank's object, in rank order, on every rank"""
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def replicate_tree(self, image):
        """rank 0 uploads `image` (a DeviceImage; None elsewhere); the block buffer goes to the other
        GPUs with ONE broadcast over NCCL -- the only collective of the path.  Returns (tree, ms)."""
        capi, torch, dist = self.capi, self.torch, self.dist
        gpu_tree = image.upload(self.local_rank) if self.rank == 0 else None
        if self.world == 1:
            return gpu_tree, 0.0
        meta = [bytes(image.desc), image.level_node_begin.tobytes()] if self.rank == 0 else [None, None]
        dist.broadcast_object_list(meta, 0)
        desc = capi.RtTreeDesc.from_buffer_copy(meta[0])
        levels = np.frombuffer(meta[1], np.uint32).copy()
        if self.rank == 0:
            blocks = torch.as_tensor(gpu_tree.device_blocks(), device=self.device)
        else:
            blocks = torch.empty(desc.blocks_bytes, dtype=torch.uint8, device=self.device)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.broadcast(blocks, 0)
        e1.record()
        torch.cuda.synchronize()
        if self.rank != 0:
            desc.blocks = blocks.data_ptr()
            desc.blocks_memory = capi.RT_MEM_DEVICE
            desc.level_node_begin = levels.ctypes.data_as(C.POINTER(C.c_uint32))
            gpu_tree = capi.DeviceTree(desc, self.local_rank)
        del blocks
        return gpu_tree, e0.elapsed_time(e1)

    def box_parity(self, ref_tree, my_queries_prefix, my_result_prefix, regenerate, what):
        """every rank's CSR for the first queries of ITS batch against the reference's search() of the
        same queries.  The reference tree lives on rank 0, which regenerates each rank's queries from
        their seed (`regenerate(rank, count)`) and compares.  Returns the `parity` object (rank 0)."""
        parts = self.gather((len(my_queries_prefix), my_result_prefix[0], my_result_prefix[1]))
        if self.rank != 0:
            return None
        if ref_tree is None:
            return {"checked": False, "why": "oracle/_ref was not built"}
        threads = host_threads()
        per_rank = self.rtb.sharding.check_rank_prefixes(
            parts, lambda q: ref_tree.search_boxes(q, threads=threads),
            lambda r, count: my_queries_prefix if r == 0 else regenerate(r, count))
        for entry in per_rank:
            if not entry["bit_exact"]:
                log("[bench] PARITY FAILURE (%s): rank %d differs from the reference search()" % (what, entry["rank"]))
                self.parity_ok = False
        return {"against": "reference RTree::search() (oracle/_ref), sorted ids per query",
                "per_rank": per_rank, "bit_exact_vs_reference_search": all(p["bit_exact"] for p in per_rank)}


def capture_summary(name):
    """the counters of a committed `ncu --set full` capture (profiles/<name>_summary.json, first launch)
    that say what bounds a kernel: issue slots, L1 wavefronts, hit rates, DRAM bytes.  None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", name + "_summary.json")) as f:
            first = json.load(f)["launches"][0]
    except Exception:
        return None
    keys = ("kernel", "duration_s", "warp_instructions", "issue_active_pct", "lsu_wavefronts_pct", "l1_hit_pct",
            "l2_hit_pct", "l2_throughput_pct", "dram_bytes", "dram_throughput_pct", "eligible_warps_per_cycle",
            "achieved_occupancy_pct", "registers_per_thread")
    out = {k: first.get(k) for k in keys}
    out["source"] = "profiles/%s_summary.json" % name
    return out


def alg_roofline(alg_bytes_per_launch, kernel_ms, note, bound, capture=None):
    """the HBM view the contract asks for (algorithmic bytes over the kernel time against the measured
    copy bandwidth) next to what the committed ncu capture of that kernel says really bounds it"""
    peak, peak_src = measured_peak()
    achieved = alg_bytes_per_launch / (kernel_ms * 1e-3) / 1e9
    return {"bound": bound,
            "achieved": achieved, "peak": peak, "unit": "GB/s (algorithmic bytes; served by L1/L2, see `capture`)",
            "frac": achieved / peak, "traffic": None,
            "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": alg_bytes_per_launch,
            "peak_source": peak_src, "note": note,
            "capture": capture_summary(capture) if capture else None}


# ---------------------------------------------------------------------------------------
def section_rays(cx, mesh_name, res):
    """BASELINE config 3: AABBs of the triangles of the (procedural) teapot mesh, random-direction
    rays, closest-hit and any-hit.  The rays are split evenly over the GPUs (the tree is small and is
    simply uploaded by every rank); value = all rays / max-over-ranks time."""
    args, rtb, torch, capi, W = cx.args, cx.rtb, cx.torch, cx.capi, cx.W
    import oracles as O
    boxes = W.triangle_aabbs(W.teapot_triangles(res))
    tree = rtb.RTree(KIND_BOX).insert(boxes).flatten_device().upload(cx.local_rank)
    tree.set_stream(cx.stream.cuda_stream)
    nr = args.rays // cx.world
    rays = W.random_rays(nr, boxes, seed=44 + 1000 * cx.rank, tmax=1e30)
    h_rays = torch.from_numpy(rays).pin_memory()
    d_rays = h_rays.to(cx.device)
    steps, warmup = max(1, min(args.steps, 5)), 3
    out = {"config": "BASELINE config 3: %d triangle AABBs of the teapot mesh (MAX_ENTRIES 8, box keys), %d "
                     "random-direction rays split over %d GPU(s)" % (len(boxes), nr * cx.world, cx.world),
           "triangles": len(boxes), "unit": "rays/s", "steps": steps, "warmup": warmup, "rays_per_gpu": nr,
           "scaling": "strong"}
    n_par = min(nr, env_int("RTB_PARITY_RAYS", 100_000))
    ref = O.RefTree(KIND_BOX).insert(boxes) if (cx.want_cpu and O.have_ref()) else None
    threads = max(1, host_threads() // cx.world)
    B_int = 8 * (2 * 3 * 4 + 4)
    for name, mode, omode, d2h_per_ray in (("closest_hit", capi.RT_RAY_CLOSEST, O.RAY_CLOSEST, 8),
                                           ("any_hit", capi.RT_RAY_ANY, O.RAY_ANY, 1)):
        res_m = {}
        # node visits of this mode on this batch (outside the timed region): the algorithmic bytes
        with torch.cuda.stream(cx.stream):
            tree.query_rays_raw(d_rays.data_ptr(), nr, mode, cx.dev_flags | capi.RT_COLLECT_VISITS)
        vs = tree.stats()
        for leg, ptr, flags in (("value", d_rays.data_ptr(), cx.dev_flags), ("e2e", h_rays.data_ptr(), 0)):
            trav = []

            def step():
                r = tree.query_rays_raw(ptr, nr, mode, flags)
                trav.append(tree.stats()["traverse_ms"])
                return r
            ms, r = cx.timed(step, steps, warmup)
            ms_all, trav_all = cx.max_over_ranks([ms, float(np.mean(trav[-steps:]))])
            res_m[leg] = float(nr) * cx.world * steps / (ms_all * 1e-3)
            res_m[leg + "_ms_per_step"] = ms_all / steps
            if leg == "value":
                res_m["kernel_ms"] = trav_all
        hits_all, vint, vleaf = cx.sum_over_ranks([r.n_hit, vs["visits_internal"], vs["visits_leaf"]])
        res_m["hit_fraction"] = hits_all / float(nr * cx.world)
        res_m["h2d_bytes_per_step"] = nr * 28 * cx.world
        res_m["d2h_bytes_per_step"] = nr * d2h_per_ray * cx.world
        alg = (vint + vleaf) * B_int / cx.world + nr * (28 + d2h_per_ray)
        res_m["visits_per_ray"] = {"internal": vint / (nr * cx.world), "leaf": vleaf / (nr * cx.world)}
        res_m["roofline"] = alg_roofline(alg, res_m["kernel_ms"],
                                         "SURVEY.md 8d: (V_int + V_leaf) x 224 B (box keys) + 28 B per ray in + %d out, "
                                         "visits counted by the engine for this mode (RT_COLLECT_VISITS)" % d2h_per_ray,
                                         "issue (instruction issue slots; the capture is closest-hit on the 342 k mesh)",
                                         "r2_v3_ray_traverse")
        # parity: this rank's first rays against the reference's search() + the slab functor
        ok = None
        if ref is not None:
            got = tree.query_rays(rays[:n_par], mode)
            if mode == capi.RT_RAY_CLOSEST:
                wt, wid = ref.search_rays(rays[:n_par], omode, threads=threads)
                ok = bool(np.array_equal(got[0].view(np.uint32), wt.view(np.uint32)) and np.array_equal(got[1], wid))
            else:
                ok = bool(np.array_equal(got, ref.search_rays(rays[:n_par], omode, threads=threads)))
        oks = cx.gather(ok)
        res_m["parity"] = {"against": "reference RTree::search() + oracle/slab.h functor; t and id bit-exact",
                           "checked_rays_per_rank": n_par if ref is not None else 0,
                           "per_rank_bit_exact": oks,
                           "bit_exact_vs_reference_search": (all(oks) if ref is not None else None)}
        if ref is not None and not all(oks):
            cx.parity_ok = False
            log("[bench] PARITY FAILURE (rays %s %s)" % (mesh_name, name))
        if cx.cpu_arm and ref is not None:
            n_cpu = min(nr, env_int("RTB_REF_RAYS", 2_000_000))
            all_threads = host_threads()
            s, _ = ref.search_rays(rays[:n_cpu], omode, threads=all_threads, timed_only=True)
            res_m["cpu_baseline"] = {"value": n_cpu / s, "unit": "rays/s", "cores": all_threads, "kind": "reference",
                                     "sample": "reference RTree::search() driven by the stateful slab functor "
                                               "(oracle/slab.h) on the first %d rays, OpenMP over %d threads (%.1f s)"
                                               % (n_cpu, all_threads, s)}
        out[name] = res_m
    tree.close()
    return out


def section_config4(cx, c4):
    """BASELINE config 4 (a cut): 8-D Gaussian mixture, MAX_ENTRIES 16 QuadraticSplit, ~10 hits/query.
    Weak scaling like config 2: the tree is replicated, every rank answers its own batch."""
    args, torch, capi, W = cx.args, cx.torch, cx.capi, cx.W
    n, nq = args.c4_points, args.c4_queries
    tree, bcast_ms = cx.replicate_tree(c4.get("image"))
    tree.set_stream(cx.stream.cuda_stream)
    pts = c4["points"] if cx.rank == 0 else None
    if cx.world > 1:
        # the queries are centred on data points: every rank needs the points (64 MB at 2 M)
        box = [pts]
        cx.dist.broadcast_object_list(box, 0)
        pts = box[0]
    h = c4_half_width(n)

    def make_queries(rank, count):
        # (the generator draws all centres, then all jitters: a prefix is only a prefix of the FULL batch)
        return W.mixture_queries(nq, pts, h, seed=46 + 1000 * rank)[:count]
    q = make_queries(cx.rank, nq)
    h_q = torch.from_numpy(q).pin_memory()
    d_q = h_q.to(cx.device)
    steps, warmup = max(1, min(args.steps, 5)), 3
    visits = {}
    with torch.cuda.stream(cx.stream):
        for image, quant in (("quantised", 1), ("float", 0)):
            tree.set_option(capi.RT_OPT_QUANTISED, quant)
            tree.query_boxes_raw(d_q.data_ptr(), nq, capi.RT_POINT_IN_BOX, cx.dev_flags | capi.RT_COLLECT_VISITS)
            s = tree.stats()
            visits[image] = {"internal": s["visits_internal"] / nq, "leaf": s["visits_leaf"] / nq,
                             "steps": s["traverse_steps"] / nq, "float_lookups": s["leaf_lookups"] / nq}
            alg = s["algorithmic_bytes"]
        tree.set_option(capi.RT_OPT_QUANTISED, 1)
    trav = []

    def dev_step():
        r = tree.query_boxes_raw(d_q.data_ptr(), nq, capi.RT_POINT_IN_BOX, cx.dev_flags)
        trav.append(tree.stats()["traverse_ms"])
        return r
    dev_ms, r = cx.timed(dev_step, steps, warmup)
    hits = r.total
    e2e_ms, out = cx.timed(lambda: tree.query_boxes_raw(h_q.data_ptr(), nq, capi.RT_POINT_IN_BOX, 0), steps, warmup)
    assert out.total == hits
    dev_all, e2e_all, trav_all = cx.max_over_ranks([dev_ms, e2e_ms, float(np.mean(trav[-steps:]))])
    hits_all, alg_all = cx.sum_over_ranks([hits, alg])
    n_par = min(nq, env_int("RTB_PARITY_C4", 20_000))
    parity = None
    if cx.want_cpu:
        parity = cx.box_parity(c4.get("ref"), q[:n_par], csr_prefix(out, nq, n_par), make_queries, "config 4")
    section = None
    if cx.rank == 0:
        total_q = float(nq) * cx.world
        section = {
            "config": "BASELINE config 4 (cut): 8-D mixture of 256 Gaussians sigma 0.02, N=%d, MAX_ENTRIES 16 "
                      "QuadraticSplit (MIN 8, REINSERT 6), %d box queries/GPU centred on data points, half width "
                      "%.5f" % (n, nq, h),
            "points": n, "queries_per_gpu": nq, "unit": UNIT, "scaling": "weak", "steps": steps, "warmup": warmup,
            "value": total_q * steps / (dev_all * 1e-3), "ms_per_step": dev_all / steps,
            "e2e": {"value": total_q * steps / (e2e_all * 1e-3), "ms_per_step": e2e_all / steps,
                    "h2d_bytes_per_step": int(q.nbytes) * cx.world,
                    "d2h_bytes_per_step": int((nq + 1) * 8 * cx.world + hits_all * 4)},
            "hits_per_query": hits_all / total_q,
            "visits_per_query": visits,
            "roofline": alg_roofline(alg_all / cx.world, trav_all,
                                     "SURVEY.md 8d on the float-image visits: V_int x 1088 B + V_leaf x 576 B + 64 B "
                                     "query + 8 + 4 x hits",
                                     "l1-lsu (L1 wavefronts / L2 -> L1 fill of 576-byte quantised nodes)",
                                     "r2_v6_box_traverse_8d"),
            "kernel": "rtb::box_traverse_kernel<float,8,16,false,false,PASS_STASH,QUANT>",
            "tree_broadcast_ms": round(bcast_ms, 3), "host_build_s": c4.get("build_s"),
            "parity": parity,
        }
        if cx.cpu_arm and c4.get("ref") is not None:
            n_cpu = min(nq, env_int("RTB_REF_C4", 100_000))
            threads = host_threads()
            s, _ = c4["ref"].time_search_boxes(q[:n_cpu], threads=threads)
            section["cpu_baseline"] = {"value": n_cpu / s, "unit": UNIT, "cores": threads, "kind": "reference",
                                       "sample": "reference RTree::search() on the first %d queries, OpenMP over %d "
                                                 "threads (%.1f s)" % (n_cpu, threads, s)}
    tree.close()
    return section


def c4_half_width(n):
    # calibrated once by bisection at N = 2 M (0.0164 -> ~10 hits/query, workloads.calibrate_half_width) and
    # scaled with the density: hits ~ N h^8
    return np.float32(float(os.environ.get("RTB_C4_HALF_WIDTH", 0.0164 * (2_000_000 / float(n)) ** (1.0 / 8.0))))


def section_config5(cx, c5):
    """BASELINE config 5: ONE batch of c5_queries cubes over a replicated 3-D tree, split evenly over
    the GPUs (strong scaling): rank r answers the contiguous range shard_range(n, world, r), in calls
    of at most 64 M queries; the tree goes to the other GPUs with one NCCL broadcast."""
    args, torch, capi, W, rtb = cx.args, cx.torch, cx.capi, cx.W, cx.rtb
    n_points, n_total = args.c5_points, args.c5_queries
    tree, bcast_ms = cx.replicate_tree(c5.get("image"))
    tree.set_stream(cx.stream.cuda_stream)
    begin, end = rtb.sharding.shard_range(n_total, cx.world, cx.rank)
    call_max = env_int("RTB_C5_CALL", 64_000_000)
    # this rank's queries: whole generation chunks [begin, end) cut into calls
    calls = []  # (pinned host tensor, device tensor, count)
    t0 = time.time()
    pos = begin
    first_prefix = None
    while pos < end:
        count = min(call_max, end - pos)
        parts = []
        p = pos
        while p < pos + count:
            chunk, inside = divmod(p, C5_CHUNK)
            take = min(C5_CHUNK - inside, pos + count - p)
            qc = c5_queries(chunk, inside + take, n_points)[inside:]
            parts.append(qc)
            p += take
        q = parts[0] if len(parts) == 1 else np.concatenate(parts)
        if first_prefix is None:
            first_prefix = q[:min(len(q), env_int("RTB_PARITY_C5", 100_000))].copy()
        h = torch.from_numpy(q).pin_memory()
        calls.append((h, h.to(cx.device), count))
        pos += count
    t_gen = time.time() - t0
    my_n = end - begin
    steps, warmup = max(1, min(args.steps, env_int("RTB_C5_STEPS", 2))), 1
    with torch.cuda.stream(cx.stream):
        h0, d0, c0 = calls[0]
        tree.query_boxes_raw(d0.data_ptr(), c0, capi.RT_POINT_IN_BOX, cx.dev_flags | capi.RT_COLLECT_VISITS)
        vs = tree.stats()
    trav = []

    def dev_step():
        total = 0
        for _, d, count in calls:
            total += tree.query_boxes_raw(d.data_ptr(), count, capi.RT_POINT_IN_BOX, cx.dev_flags).total
            trav.append(tree.stats()["traverse_ms"])
        return total
    dev_ms, hits = cx.timed(dev_step, steps, warmup)
    first_out = {}

    def host_step():
        total = 0
        for i, (hq, _, count) in enumerate(calls):
            out = tree.query_boxes_raw(hq.data_ptr(), count, capi.RT_POINT_IN_BOX, 0)
            if i == 0 and "csr" not in first_out and len(first_prefix):
                first_out["csr"] = csr_prefix(out, count, len(first_prefix))  # the buffers are reused by the next call
            total += out.total
        return total
    e2e_ms, hits2 = cx.timed(host_step, steps, warmup)
    assert hits == hits2, "device and host legs disagree on the hit total"
    dev_all, e2e_all, trav_all = cx.max_over_ranks([dev_ms, e2e_ms, float(np.sum(trav[-steps * len(calls):])) / steps])
    dev_each = cx.gather(round(dev_ms / steps, 3))
    hits_all, = cx.sum_over_ranks([hits])
    parity = None
    if cx.want_cpu:
        def regenerate(rank, count):
            b, _ = rtb.sharding.shard_range(n_total, cx.world, rank)
            chunk, inside = divmod(b, C5_CHUNK)
            assert inside + count <= C5_CHUNK
            return c5_queries(chunk, inside + count, n_points)[inside:]
        parity = cx.box_parity(c5.get("ref"), first_prefix, first_out.get("csr", (np.zeros(1, np.uint64), np.zeros(0, np.uint32))),
                               regenerate, "config 5")
    section = None
    if cx.rank == 0:
        alg_per_query = vs["algorithmic_bytes"] / float(calls[0][2])
        section = {
            "config": "BASELINE config 5%s: 3-D uniform points N=%d (MAX_ENTRIES 8), ONE batch of %d cube queries "
                      "(~10 hits/query) split evenly over %d GPU(s), tree replicated by one NCCL broadcast"
                      % ("" if n_points >= 200_000_000 else " (scaled: 200 M points is the full configuration)",
                         n_points, n_total, cx.world),
            "points": n_points, "queries_total": n_total, "queries_per_gpu": my_n, "calls_per_gpu": len(calls),
            "unit": UNIT, "scaling": "strong", "steps": steps, "warmup": warmup,
            "value": float(n_total) * steps / (dev_all * 1e-3), "ms_per_step": dev_all / steps,
            "ms_per_step_per_rank": dev_each,
            "e2e": {"value": float(n_total) * steps / (e2e_all * 1e-3), "ms_per_step": e2e_all / steps,
                    "h2d_bytes_per_step": n_total * 24, "d2h_bytes_per_step": int((n_total + 8) * 8 + hits_all * 4)},
            "hits_per_query": hits_all / float(n_total),
            "visits_per_query": {"internal": vs["visits_internal"] / float(calls[0][2]),
                                 "leaf": vs["visits_leaf"] / float(calls[0][2]), "image": "quantised"},
            "roofline": alg_roofline(alg_per_query * my_n, trav_all,
                                     "per GPU: algorithmic bytes/query counted on the first call x this GPU's queries, "
                                     "over the summed traversal-kernel time of a step",
                                     "issue (the config-2 kernel on a deeper tree; capture = config 2)", "r2_v6_box_traverse"),
            "tree": {"nodes": int(tree.num_nodes), "levels": int(tree.leaf_level) + 1, "blocks_bytes": int(tree.blocks_bytes),
                     "host_build_s": c5.get("build_s"), "flatten_device_s": c5.get("flatten_s"),
                     "broadcast_ms": round(bcast_ms, 3)},
            "query_generation_s": round(t_gen, 1),
            "parity": parity,
        }
        if cx.cpu_arm and c5.get("ref") is not None:
            n_cpu = int(min(calls[0][2], env_int("RTB_REF_C5", 1_000_000)))
            threads = host_threads()
            s, _ = c5["ref"].time_search_boxes(calls[0][0].numpy()[:n_cpu], threads=threads)
            section["cpu_baseline"] = {"value": n_cpu / s, "unit": UNIT, "cores": threads, "kind": "reference",
                                       "sample": "reference RTree::search() on the first %d queries of the batch, OpenMP "
                                                 "over %d threads (%.1f s)" % (n_cpu, threads, s)}
    tree.close()
    return section


def section_knn(cx, gpu_tree, ref_tree, points_n):
    """k = 8 nearest neighbours on the config-2 tree, every rank its own batch."""
    args, torch, capi, W = cx.args, cx.torch, cx.capi, cx.W
    nq, k = args.knn_queries, 8
    pts = W.uniform_points(nq, seed=48 + 1000 * cx.rank)
    h_p = torch.from_numpy(pts).pin_memory()
    d_p = h_p.to(cx.device)
    steps, warmup = max(1, min(args.steps, 5)), 3
    with torch.cuda.stream(cx.stream):
        gpu_tree.query_knn_raw(d_p.data_ptr(), nq, k, cx.dev_flags | capi.RT_COLLECT_VISITS)
    vs = gpu_tree.stats()
    trav = []

    def dev_step():
        r = gpu_tree.query_knn_raw(d_p.data_ptr(), nq, k, cx.dev_flags)
        trav.append(gpu_tree.stats()["traverse_ms"])
        return r
    dev_ms, _ = cx.timed(dev_step, steps, warmup)
    e2e_ms, _ = cx.timed(lambda: gpu_tree.query_knn_raw(h_p.data_ptr(), nq, k, 0), steps, warmup)
    dev_all, e2e_all, trav_all = cx.max_over_ranks([dev_ms, e2e_ms, float(np.mean(trav[-steps:]))])
    n_par = min(nq, env_int("RTB_PARITY_KNN", 20_000))
    got = gpu_tree.query_knn(pts[:n_par], k)
    parts = cx.gather((got[0], got[1]))
    section = None
    if cx.rank == 0:
        total_q = float(nq) * cx.world
        section = {"config": "k = %d nearest neighbours on the config-2 tree (N=%d), %d uniform query points/GPU"
                             % (k, points_n, nq),
                   "unit": UNIT, "scaling": "weak", "steps": steps, "warmup": warmup, "k": k,
                   "value": total_q * steps / (dev_all * 1e-3), "ms_per_step": dev_all / steps,
                   "e2e": {"value": total_q * steps / (e2e_all * 1e-3), "ms_per_step": e2e_all / steps,
                           "h2d_bytes_per_step": int(pts.nbytes) * cx.world,
                           "d2h_bytes_per_step": int(nq * k * 8) * cx.world},
                   "visits_per_query": {"internal": vs["visits_internal"] / nq, "leaf": vs["visits_leaf"] / nq},
                   "roofline": alg_roofline(vs["algorithmic_bytes"], trav_all,
                                            "V_int x 224 B + V_leaf x 128 B + 12 B per point in + 8 k out; visits counted "
                                            "by the engine (pruning makes them traversal-order dependent)",
                                            "issue (instruction issue slots)", "r2_v4_knn_traverse")}
        if cx.want_cpu and ref_tree is not None:
            threads = host_threads()
            oks = []
            for r, (ids, d2) in enumerate(parts):
                p = pts[:n_par] if r == 0 else W.uniform_points(nq, seed=48 + 1000 * r)[:n_par]
                wi, wd = ref_tree.search_knn(p, k, threads=threads)
                oks.append(bool(np.array_equal(ids, wi) and np.array_equal(d2.view(np.uint32), wd.view(np.uint32))))
            section["parity"] = {"against": "reference RTree::search() + oracle/knn.h functors; ids and float32 "
                                            "distances bit-exact", "checked_queries_per_rank": n_par,
                                 "per_rank_bit_exact": oks, "bit_exact_vs_reference_search": all(oks)}
            if not all(oks):
                cx.parity_ok = False
                log("[bench] PARITY FAILURE (kNN)")
            if cx.cpu_arm:
                n_cpu = min(nq, env_int("RTB_REF_KNN", 200_000))
                s = ref_tree.search_knn(pts[:n_cpu], k, threads=threads, timed_only=True)
                section["cpu_baseline"] = {"value": n_cpu / s, "unit": UNIT, "cores": threads, "kind": "reference",
                                           "sample": "reference RTree::search() with the kNN functors on the first %d "
                                                     "points, OpenMP over %d threads (%.1f s)" % (n_cpu, threads, s)}
    return section


# ---------------------------------------------------------------------------------------
def main():
    t_start = time.time()
    args = parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    rtb = importlib.import_module("r-star-tree_b200")
    capi, W = rtb.capi, rtb.workloads
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the query path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    # stdout carries the ONE JSON line only: libraries that print there (NCCL writes its version
    # banner with printf when NCCL_DEBUG is set) are sent to stderr until the line is printed
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # rank 0 builds the trees on the host (minutes for config 5) while the others wait
        dist.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(minutes=60))
    stream = torch.cuda.Stream(device)
    cx = Ctx(args, torch, dist, rtb, rank, world, local_rank, device, stream)
    import oracles as O
    have_ref = O.have_ref() and cx.want_cpu

    # ---- host builds, all on rank 0, side by side (ctypes releases the GIL) ---------------------
    # product trees by the host library; reference trees (the parity checkers / CPU arms) by the
    # unmodified reference.  The headline tree first, so its sections start while the rest builds.
    c2, c4, c5 = {}, {}, {}
    threads = []

    def spawn(fn):
        th = threading.Thread(target=fn)
        th.start()
        threads.append(th)
        return th

    if rank == 0:
        points = W.uniform_points(args.points, seed=42)

        def build_c2():
            t0 = time.time()
            host_tree = rtb.RTree(KIND).insert(points)
            c2["build_s"] = time.time() - t0
            t0 = time.time()
            c2["image"] = host_tree.flatten_device(ids_from_mapped=True)
            c2["flatten_s"] = time.time() - t0

        def build_c2_ref():
            t0 = time.time()
            c2["ref"] = O.RefTree(KIND).insert(points)
            c2["ref_build_s"] = time.time() - t0

        def build_c4():
            c4["points"] = W.mixture_points(args.c4_points)
            t0 = time.time()
            c4["image"] = rtb.RTree(KIND_8D).insert(c4["points"]).flatten_device(ids_from_mapped=True)
            c4["build_s"] = round(time.time() - t0, 1)
            if have_ref:
                c4["ref"] = O.RefTree(KIND_8D).insert(c4["points"])

        def build_c5():
            p5 = W.uniform_points(args.c5_points, seed=42)
            c5["points_ready"] = p5
            t0 = time.time()
            host_tree = rtb.RTree(KIND).insert(p5)
            c5["build_s"] = round(time.time() - t0, 1)
            t0 = time.time()
            c5["image"] = host_tree.flatten_device(ids_from_mapped=True)
            c5["flatten_s"] = round(time.time() - t0, 2)

        def build_c5_ref():
            while "points_ready" not in c5:
                time.sleep(0.05)
            c5["ref"] = O.RefTree(KIND).insert(c5["points_ready"])

        th_c2 = spawn(build_c2)
        th_c2_ref = spawn(build_c2_ref) if have_ref else None
        th_c4 = spawn(build_c4) if args.c4_points > 0 else None
        th_c5 = spawn(build_c5) if args.c5_points > 0 else None
        th_c5_ref = spawn(build_c5_ref) if (args.c5_points > 0 and have_ref) else None
        th_c2.join()
        log("[bench] host build %.1f s (%d inserts), flatten_device %.2f s, %d nodes, %.0f MB of blocks"
            % (c2["build_s"], args.points, c2["flatten_s"], c2["image"].desc.num_nodes, c2["image"].desc.blocks_bytes / 1e6))
    gpu_tree, t_bcast_ms = cx.replicate_tree(c2.get("image"))

    # ---- this rank's query batch -----------------------------------------------------------
    nq = args.queries

    def make_queries(r, count):
        return W.cube_queries(count, args.points, seed=43 + 1000 * r)
    queries = make_queries(rank, nq)
    h_queries = torch.from_numpy(queries).pin_memory()
    d_queries = h_queries.to(device)
    gpu_tree.set_stream(stream.cuda_stream)
    dev_flags = cx.dev_flags

    # node visits of one launch over this batch, on both images (outside the timed region): the float
    # image is the reference's exact filter (what SURVEY.md 8d's algorithmic bytes are defined on),
    # the quantised image is what the timed kernel traverses
    visit_stats = {}
    with torch.cuda.stream(stream):
        for image, quant in (("quantised", 1), ("float", 0)):
            gpu_tree.set_option(capi.RT_OPT_QUANTISED, quant)
            gpu_tree.query_boxes_raw(d_queries.data_ptr(), nq, capi.RT_POINT_IN_BOX, dev_flags | capi.RT_COLLECT_VISITS)
            visit_stats[image] = gpu_tree.stats()
        gpu_tree.set_option(capi.RT_OPT_QUANTISED, 1)
    alg_bytes = visit_stats["float"]["algorithmic_bytes"]

    # ---- leg 1: device-resident ------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    traverse_ms, launch_counts, totals = [], [], []

    def device_step():
        out = gpu_tree.query_boxes_raw(d_queries.data_ptr(), nq, capi.RT_POINT_IN_BOX, dev_flags)
        s = gpu_tree.stats()
        traverse_ms.append(s["traverse_ms"])
        launch_counts.append(s["kernel_launches"])
        totals.append(out.total)
        return out

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            device_step()
        cx.barrier()
        sampler.start()
        t_begin = time.time()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            device_step()
        ev1.record(stream)
        cx.barrier()
        t_end = time.time()
    dev_ms = ev0.elapsed_time(ev1)
    traverse_ms, launches, total_hits = traverse_ms[-args.steps:], sum(launch_counts[-args.steps:]), totals[-1]
    last_stats = gpu_tree.stats()

    # ---- leg 2: end to end through the C ABI with host buffers ------------------------------
    def host_step(flags=0):
        return gpu_tree.query_boxes_raw(h_queries.data_ptr(), nq, capi.RT_POINT_IN_BOX, flags)

    e2e_ms, out = cx.timed(host_step, args.steps, args.warmup)
    e2e_stats = gpu_tree.stats()
    clocks = sampler.stop(t_begin, t_end)
    checksum = int(host_view(out.offsets, nq + 1, C.c_uint64, np.uint64)[-1])
    h2d_bytes = int(queries.nbytes)
    d2h_bytes = int((nq + 1) * 8 + out.total * 4 + 64)
    assert checksum == out.total == total_hits, "device and host legs disagree on the hit total"
    n_par = min(nq, env_int("RTB_PARITY_QUERIES", 200_000))
    my_prefix = csr_prefix(out, nq, n_par)
    # the same call delivering 32-bit counts instead of 64-bit offsets
    c32_ms, out32 = cx.timed(lambda: host_step(capi.RT_RESULT_COUNTS32), max(1, min(args.steps, 5)), 2)
    c32_steps = max(1, min(args.steps, 5))
    counts_sum = int(host_view(out32.counts, nq, C.c_uint32, np.uint32).sum(dtype=np.uint64))
    assert counts_sum == total_hits, "counts32 result disagrees on the hit total"

    # ---- max over ranks ---------------------------------------------------------------------
    dev_ms, e2e_ms, trav_ms, c32_ms = cx.max_over_ranks([dev_ms, e2e_ms, float(np.mean(traverse_ms)), c32_ms])
    launches_all, alg_all, hits_all = cx.sum_over_ranks([launches, alg_bytes, total_hits])

    # ---- parity on every rank + CPU baseline on rank 0 (N = 1 only) ---------------------------
    cpu_baseline, parity = None, None
    if cx.want_cpu:
        if rank == 0 and th_c2_ref is not None:
            th_c2_ref.join()
        parity = cx.box_parity(c2.get("ref"), queries[:n_par], my_prefix, make_queries, "config 2")
        if cx.cpu_arm and rank == 0 and c2.get("ref") is not None:
            ref = c2["ref"]
            threads_all = host_threads()
            n1 = min(nq, 300_000)
            s1, _ = ref.time_search_boxes(queries[:n1], threads=1)
            nall = min(nq, args.cpu_sample)
            sall, hall = ref.time_search_boxes(queries[:nall], threads=threads_all)
            cpu_baseline = {
                "value": nall / sall, "unit": UNIT, "cores": threads_all, "kind": "reference",
                "sample": "reference RTree::search() (oracle/_ref) on the first %d queries, OpenMP "
                          "schedule(dynamic,1024) over %d threads (%.1f s); single thread on the first "
                          "%d: %.0f queries/s; reference tree build %.1f s"
                          % (nall, threads_all, sall, n1, n1 / s1, c2.get("ref_build_s", 0.0)),
                "single_thread_value": n1 / s1,
            }

    # ---- the other sections --------------------------------------------------------------------
    knn_section = None
    if args.knn_queries > 0:
        knn_section = section_knn(cx, gpu_tree, c2.get("ref"), args.points)
    main_tree_bytes = int(gpu_tree.blocks_bytes)
    setup = {"host_build_s": round(c2["build_s"], 2) if "build_s" in c2 else None,
             "flatten_device_s": round(c2["flatten_s"], 3) if "flatten_s" in c2 else None,
             "tree_broadcast_ms": round(t_bcast_ms, 3), "tree_bytes": main_tree_bytes}
    gpu_tree.close()
    c2.clear()
    rays_section = None
    if args.rays > 0:
        rays_section = {"teapot_9k": section_rays(cx, "teapot_9k", 1.0),
                        "teapot_333k": section_rays(cx, "teapot_333k", 5.9)}
    c4_section = None
    if args.c4_points > 0:
        if rank == 0:
            th_c4.join()
        c4_section = section_config4(cx, c4)
        c4.clear()
    c5_section = None
    if args.c5_points > 0:
        if rank == 0:
            th_c5.join()
            if th_c5_ref is not None:
                th_c5_ref.join()
            c5.pop("points_ready", None)
        c5_section = section_config5(cx, c5)
        c5.clear()

    if rank == 0:
        peak, peak_src = measured_peak()
        total_q = float(nq) * world
        cap = ncu_capture()
        achieved_alg = (alg_all / world) / (trav_ms * 1e-3) / 1e9  # per-GPU launch: GB/s
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        issue_peak = sms * 4 * sm_mhz * 1e6
        instr = cap.get("warp_instructions_per_launch") if cap.get("queries_per_launch") == nq else None
        if instr is None and cap.get("warp_instructions_per_query"):
            instr = cap["warp_instructions_per_query"] * nq
        issue_rate = instr / (trav_ms * 1e-3) if instr else None
        traffic = cap.get("dram_bytes_per_launch")
        vq = {k: {"internal": v["visits_internal"] / nq, "leaf": v["visits_leaf"] / nq,
                  "loop_steps": v["traverse_steps"] / nq, "float_lookups": v["leaf_lookups"] / nq}
              for k, v in visit_stats.items()}
        roofline = {
            "bound": "issue",
            "achieved": issue_rate, "peak": issue_peak, "unit": "warp-instructions/s",
            "frac": (issue_rate / issue_peak) if issue_rate else None,
            "traffic": traffic,
            "kernel": "rtb::box_traverse_kernel<float,3,8,false,false,PASS_STASH,QUANT>",
            "kernel_ms": trav_ms,
            "instr_per_query": (instr / nq) if instr else None,
            "issue_active_pct_in_capture": cap.get("issue_active_pct"),
            "peak_note": "%d SMs x 4 schedulers x %.0f MHz (SM clock sampled during the timed region)" % (sms, sm_mhz),
            "dram_frac": (traffic / (trav_ms * 1e-3) / 1e9 / peak) if traffic else None,
            "hbm_peak_gbs": peak, "peak_source": peak_src,
            "l1_bytes": cap.get("l1_bytes_per_launch"),
            "algorithmic_bytes_per_launch": alg_all / world,
            "algorithmic_bytes_per_query": alg_all / world / nq,
            "algorithmic_gbs": achieved_alg,
            "algorithmic_over_hbm": achieved_alg / peak,
            "visits_per_query": vq,
            "capture": cap.get("source"),
            "note": "the kernel is bound by instruction issue: its node blocks are served by L1/L2 (DRAM traffic "
                    "is about 1 % of the algorithmic bytes), so the HBM fraction the contract asks for is "
                    "`dram_frac`, and `algorithmic_over_hbm` (> 1) only says how much faster than a DRAM-streaming "
                    "implementation the cache hierarchy lets it run.  Algorithmic bytes use the float-image visits "
                    "(the reference's exact filter); the timed kernel traverses the quantised image, whose visit "
                    "counts are listed beside them.",
        }
        line = {
            "metric": METRIC, "value": (total_q * args.steps / (dev_ms * 1e-3)) if cx.parity_ok else None,
            "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": (total_q * args.steps / (e2e_ms * 1e-3)) if cx.parity_ok else None, "unit": UNIT,
                    "h2d_bytes_per_step": h2d_bytes * world, "d2h_bytes_per_step": d2h_bytes * world,
                    "ms_per_step": e2e_ms / args.steps,
                    "batches": int(e2e_stats["batches"]),
                    "note": "pipelined: the call is cut into batches whose H2D / traversal / D2H overlap; "
                            "breakdown_ms are per-phase SUMS over the batches (they overlap one another)",
                    "breakdown_ms": {k: round(e2e_stats[k], 3) for k in
                                     ("h2d_ms", "reorder_ms", "traverse_ms", "finish_ms", "d2h_ms")},
                    "counts32": {"value": total_q * c32_steps / (c32_ms * 1e-3), "ms_per_step": c32_ms / c32_steps,
                                 "d2h_bytes_per_step": int(nq * 4 + out.total * 4 + 64) * world,
                                 "note": "RT_RESULT_COUNTS32: 32-bit hits per query instead of 64-bit offsets"}},
            "gpu_launches": int(launches_all),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "clocks": clocks,
            "hits_per_query": hits_all / total_q,
            "device_breakdown_ms": {k: round(last_stats[k], 3) for k in
                                    ("reorder_ms", "traverse_ms", "finish_ms")},
            "setup": setup,
            "parity": parity,
            "rays": rays_section,
            "config4": c4_section,
            "config5": c5_section,
            "knn": knn_section,
            "host": {"threads": host_threads(), "wall_s": round(time.time() - t_start, 1)},
        }
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    ok = cx.parity_ok
    if world > 1:
        flag = cx.max_over_ranks([0.0 if ok else 1.0])[0]
        ok = flag == 0.0
        dist.destroy_process_group()
    if not ok:
        raise SystemExit(3)


if __name__ == "__main__":
    main()
