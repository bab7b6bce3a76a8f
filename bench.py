#!/usr/bin/env python
"""bench.py -- box-queries/s of the batched R-tree query path on N B200s (BASELINE.json).

A step is one pass of the hot path over one batch of synthetic queries:
config 2 of BASELINE.json -- a 3-D tree of 10 M uniform points (MAX_ENTRIES 8, built on the
CPU by the host library's insert / RStarSplit / reinsert) searched by 16 M small cubes
(~10 hits each), results as per-query sorted id lists in CSR form.

  value    whole-job queries/s with the queries already resident in HBM and the CSR result
           left in HBM (rt_query_boxes with RT_QUERIES_ON_DEVICE | RT_RESULT_ON_DEVICE);
           CUDA events on the launching stream, barrier + synchronize on both sides, max
           over ranks.  Inputs (384 MB of queries + 429 MB of tree) exceed the 126 MB L2.
  e2e      the same metric through the C ABI with HOST buffers: pinned queries in, pinned
           offsets + ids out, both copies inside the timed region (the engine cuts such a call
           into batches and overlaps their copies with the traversal of their neighbours).
  roofline algorithmic bytes per launch (SURVEY.md 8d formula on the node-visit totals the
           engine counts with RT_COLLECT_VISITS) / average duration of the traversal kernel
           (CUDA events inside the timed steps), against MEASURED_PEAKS.json's HBM number.
  cpu_baseline  the reference's own RTree::search() (oracle/_ref, the unmodified reference)
           on this box's host cores, OpenMP over a bounded prefix of the same queries.

Multi-GPU (torchrun, one rank per GPU): rank 0 builds the tree, the block buffer is
broadcast GPU-to-GPU over NCCL, every rank answers its own 16 M-query batch (weak scaling,
no data-path collective), value = all ranks' queries / max-over-ranks time.

`--impl reference` times the reference's CPU search on the same workload (rank 0 only).
"""
import argparse
import ctypes as C
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
KIND = 4  # RTree<aabb_t<point_t<float,3>>, point_t<float,3>, uint32_t, DefaultConfig>


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=int(os.environ.get("RTB_BENCH_POINTS", 10_000_000)))
    ap.add_argument("--queries", type=int, default=int(os.environ.get("RTB_BENCH_QUERIES", 16_000_000)))
    ap.add_argument("--cpu-sample", type=int, default=2_000_000,
                    help="queries the OpenMP CPU baseline runs (a prefix of the batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rays", type=int, default=int(os.environ.get("RTB_BENCH_RAYS", 64_000_000)),
                    help="rays of the secondary config-3 section (split evenly over the GPUs); 0 = skip")
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


def dram_traffic_per_launch():
    """dram bytes per launch of the traversal kernel from the committed ncu --set full capture
    (profiles/traffic.json, written when a capture is summarised), else null."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


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
    threads = O.ref_max_threads()
    sample = min(args.queries, max(100_000, int(os.environ.get("RTB_REF_STEP_QUERIES", 1_000_000))))
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
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": "each step = RTree::search() over %d queries of the workload, "
                                   "OpenMP schedule(dynamic,1024) on %d threads; %.2f hits/query"
                                   % (sample, threads, hits / float(sample * args.steps))},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------
def run_rays(args, rtb, torch, dist, device, local_rank, rank, world, stream, barrier):
    """BASELINE config 3: AABBs of the triangles of the (procedural) teapot mesh, random-direction
    rays, closest-hit and any-hit.  The 64 M rays are split evenly over the GPUs (the tree is a few
    hundred kB and is simply uploaded by every rank); value = all rays / max-over-ranks time."""
    capi, W = rtb.capi, rtb.workloads
    boxes = W.triangle_aabbs(W.teapot_triangles(1.0))
    tree = rtb.RTree(5).insert(boxes).flatten_device().upload(local_rank)
    tree.set_stream(stream.cuda_stream)
    nr = args.rays // world
    h_rays = torch.from_numpy(W.random_rays(nr, boxes, seed=44 + 1000 * rank, tmax=1e30)).pin_memory()
    d_rays = h_rays.to(device)
    steps, warmup = max(1, min(args.steps, 5)), max(3, min(args.warmup, 3))
    dev_flags = capi.RT_QUERIES_ON_DEVICE | capi.RT_RESULT_ON_DEVICE
    out = {"config": "BASELINE config 3: %d triangle AABBs of the teapot mesh (MAX_ENTRIES 8, box keys), %d "
                     "random-direction rays split over %d GPU(s)" % (len(boxes), nr * world, world),
           "unit": "rays/s", "steps": steps, "warmup": warmup, "rays_per_gpu": nr}
    for name, mode, d2h_per_ray in (("closest_hit", capi.RT_RAY_CLOSEST, 8), ("any_hit", capi.RT_RAY_ANY, 1)):
        res = {}
        for leg, ptr, flags in (("value", d_rays.data_ptr(), dev_flags), ("e2e", h_rays.data_ptr(), 0)):
            with torch.cuda.stream(stream):
                for _ in range(warmup):
                    tree.query_rays_raw(ptr, nr, mode, flags)
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(steps):
                    r = tree.query_rays_raw(ptr, nr, mode, flags)
                e1.record(stream)
                barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            res[leg] = float(nr) * world * steps / (float(ms.item()) * 1e-3)
            res[leg + "_ms_per_step"] = float(ms.item()) / steps
        res["hit_fraction"] = r.n_hit / float(nr)
        res["h2d_bytes_per_step"] = nr * 28 * world
        res["d2h_bytes_per_step"] = nr * d2h_per_ray * world
        out[name] = res
    tree.close()
    return out


# ---------------------------------------------------------------------------------------
def main():
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
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the tree: built once on rank 0 by the host library, broadcast to the other GPUs ----
    ref_holder = {}
    want_cpu = (world == 1 and not args.no_cpu_baseline)
    points = W.uniform_points(args.points, seed=42) if rank == 0 else None

    def build_reference_tree():
        import oracles as O
        if O.have_ref():
            t0 = time.time()
            ref_holder["tree"] = O.RefTree(KIND).insert(points)
            ref_holder["build_s"] = time.time() - t0

    ref_thread = None
    if want_cpu and rank == 0:
        ref_thread = threading.Thread(target=build_reference_tree)
        ref_thread.start()  # ctypes releases the GIL: runs beside the product build

    t_build = t_flatten = 0.0
    if rank == 0:
        t0 = time.time()
        host_tree = rtb.RTree(KIND).insert(points)
        t_build = time.time() - t0
        t0 = time.time()
        image = host_tree.flatten_device(ids_from_mapped=True)
        t_flatten = time.time() - t0
        log("[bench] host build %.1f s (%d inserts), flatten_device %.2f s, %d nodes, %.0f MB of blocks"
            % (t_build, args.points, t_flatten, image.desc.num_nodes, image.desc.blocks_bytes / 1e6))
        gpu_tree = image.upload(local_rank)
        meta = [bytes(image.desc), image.level_node_begin.tobytes()]
    else:
        meta = [None, None]
    t_bcast_ms = 0.0
    if world > 1:
        dist.broadcast_object_list(meta, 0)
        desc = capi.RtTreeDesc.from_buffer_copy(meta[0])
        levels = np.frombuffer(meta[1], np.uint32).copy()
        if rank == 0:
            blocks = torch.as_tensor(gpu_tree.device_blocks(), device=device)
        else:
            blocks = torch.empty(desc.blocks_bytes, dtype=torch.uint8, device=device)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.broadcast(blocks, 0)  # the one collective of the path: tree replication over NVLink
        e1.record()
        torch.cuda.synchronize()
        t_bcast_ms = e0.elapsed_time(e1)
        if rank != 0:
            desc.blocks = blocks.data_ptr()
            desc.blocks_memory = capi.RT_MEM_DEVICE
            desc.level_node_begin = levels.ctypes.data_as(C.POINTER(C.c_uint32))
            gpu_tree = capi.DeviceTree(desc, local_rank)
            del blocks

    # ---- this rank's query batch -----------------------------------------------------------
    nq = args.queries
    queries = W.cube_queries(nq, args.points, seed=43 + 1000 * rank)
    h_queries = torch.from_numpy(queries).pin_memory()
    d_queries = h_queries.to(device)
    stream = torch.cuda.Stream(device)
    gpu_tree.set_stream(stream.cuda_stream)
    dev_flags = capi.RT_QUERIES_ON_DEVICE | capi.RT_RESULT_ON_DEVICE

    # algorithmic bytes of one launch over this batch (outside the timed region)
    with torch.cuda.stream(stream):
        gpu_tree.query_boxes_raw(d_queries.data_ptr(), nq, capi.RT_POINT_IN_BOX, dev_flags | capi.RT_COLLECT_VISITS)
    visit_stats = gpu_tree.stats()
    alg_bytes = visit_stats["algorithmic_bytes"]

    # ---- leg 1: device-resident ------------------------------------------------------------
    def device_step():
        return gpu_tree.query_boxes_raw(d_queries.data_ptr(), nq, capi.RT_POINT_IN_BOX, dev_flags)

    sampler = ClockSampler(local_rank)
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            device_step()
        barrier()
        sampler.start()
        t_begin = time.time()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        traverse_ms, launches, total_hits = [], 0, 0
        for _ in range(args.steps):
            out = device_step()
            s = gpu_tree.stats()
            traverse_ms.append(s["traverse_ms"])
            launches += s["kernel_launches"]
            total_hits = out.total
        ev1.record(stream)
        barrier()
        t_end = time.time()
    dev_ms = ev0.elapsed_time(ev1)
    last_stats = gpu_tree.stats()

    # ---- leg 2: end to end through the C ABI with host buffers ------------------------------
    def host_step():
        return gpu_tree.query_boxes_raw(h_queries.data_ptr(), nq, capi.RT_POINT_IN_BOX, 0)

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            host_step()
        barrier()
        ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev2.record(stream)
        for _ in range(args.steps):
            out = host_step()
            checksum = int(np.ctypeslib.as_array(C.cast(out.offsets, C.POINTER(C.c_uint64)), shape=(nq + 1,))[-1])
        ev3.record(stream)
        barrier()
    e2e_ms = ev2.elapsed_time(ev3)
    e2e_stats = gpu_tree.stats()
    clocks = sampler.stop(t_begin, t_end)
    h2d_bytes = int(queries.nbytes)
    d2h_bytes = int((nq + 1) * 8 + out.total * 4 + 64)
    assert checksum == out.total == total_hits, "device and host legs disagree on the hit total"

    # ---- secondary section: BASELINE config 3, rays/s (closest-hit and any-hit) ----------------
    rays_section = run_rays(args, rtb, torch, dist, device, local_rank, rank, world, stream, barrier) if args.rays > 0 else None

    # ---- max over ranks ---------------------------------------------------------------------
    times = torch.tensor([dev_ms, e2e_ms, float(np.mean(traverse_ms))], dtype=torch.float64, device=device)
    sums = torch.tensor([float(launches), float(alg_bytes), float(total_hits)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms, trav_ms = [float(x) for x in times.tolist()]
    launches_all, alg_all, hits_all = [float(x) for x in sums.tolist()]

    # ---- parity + CPU baseline on rank 0 (N = 1 only) ---------------------------------------
    cpu_baseline, parity = None, None
    if want_cpu and rank == 0:
        ref_thread.join()
        import oracles as O
        ref = ref_holder.get("tree")
        if ref is not None:
            threads = O.ref_max_threads()
            n_par = min(nq, 200_000)
            want = ref.search_boxes(queries[:n_par], threads=threads)
            off = np.ctypeslib.as_array(C.cast(out.offsets, C.POINTER(C.c_uint64)), shape=(nq + 1,))
            ids = np.ctypeslib.as_array(C.cast(out.ids, C.POINTER(C.c_uint32)), shape=(max(int(out.total), 1),))
            ok = bool(np.array_equal(off[:n_par + 1], want[0]) and np.array_equal(ids[:int(off[n_par])], want[1]))
            parity = {"checked_queries": n_par, "bit_exact_vs_reference_search": ok}
            n1 = min(nq, 300_000)
            s1, _ = ref.time_search_boxes(queries[:n1], threads=1)
            nall = min(nq, args.cpu_sample)
            sall, hall = ref.time_search_boxes(queries[:nall], threads=threads)
            cpu_baseline = {
                "value": nall / sall, "unit": UNIT, "cores": threads, "kind": "reference",
                "sample": "reference RTree::search() (oracle/_ref) on the first %d queries, OpenMP "
                          "schedule(dynamic,1024) over %d threads (%.1f s); single thread on the first "
                          "%d: %.0f queries/s; reference tree build %.1f s"
                          % (nall, threads, sall, n1, n1 / s1, ref_holder.get("build_s", 0.0)),
                "single_thread_value": n1 / s1,
            }
            if not ok:
                log("[bench] PARITY FAILURE against the reference search on the first %d queries" % n_par)

    if rank == 0:
        peak, peak_src = measured_peak()
        total_q = float(nq) * world
        achieved = (alg_all / world) / (trav_ms * 1e-3) / 1e9  # per-GPU launch: GB/s
        line = {
            "metric": METRIC, "value": total_q * args.steps / (dev_ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": total_q * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d_bytes * world, "d2h_bytes_per_step": d2h_bytes * world,
                    "ms_per_step": e2e_ms / args.steps,
                    "batches": int(e2e_stats["batches"]),
                    "note": "pipelined: the call is cut into batches whose H2D / traversal / D2H overlap; "
                            "breakdown_ms are per-phase SUMS over the batches (they overlap one another)",
                    "breakdown_ms": {k: round(e2e_stats[k], 3) for k in
                                     ("h2d_ms", "reorder_ms", "traverse_ms", "finish_ms", "d2h_ms")}},
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": dram_traffic_per_launch(),
                         "kernel": "rtb::box_traverse_kernel<float,3,8,false,false,PASS_STASH,QUANT>",
                         "kernel_ms": trav_ms, "algorithmic_bytes_per_launch": alg_all / world,
                         "algorithmic_bytes_per_query": alg_all / world / nq,
                         "visits_per_query": {"internal": visit_stats["visits_internal"] / nq,
                                              "leaf": visit_stats["visits_leaf"] / nq},
                         "peak_source": peak_src,
                         "note": "algorithmic bytes count every node block a query must read; most of "
                                 "them are served by L1/L2 (the internal levels fit the 126 MB L2), so "
                                 "the fraction can exceed 1 -- `traffic` is what actually reached HBM"},
            "cpu_baseline": cpu_baseline,
            "clocks": clocks,
            "hits_per_query": hits_all / total_q,
            "device_breakdown_ms": {k: round(last_stats[k], 3) for k in
                                    ("reorder_ms", "traverse_ms", "finish_ms")},
            "setup": {"host_build_s": round(t_build, 2), "flatten_device_s": round(t_flatten, 3),
                      "tree_broadcast_ms": round(t_bcast_ms, 3), "tree_bytes": int(gpu_tree.blocks_bytes)},
            "parity": parity,
            "rays": rays_section,
        }
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    gpu_tree.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
