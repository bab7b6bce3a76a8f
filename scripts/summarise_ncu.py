#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, on the CPU box) into profiles/.

  python scripts/summarise_ncu.py gpurun_out/box_traverse_r1b.ncu-rep profiles/r1_v3_box_traverse \
      [--traffic [--queries N]]   # also (re)write profiles/traffic.json: what bench.py's roofline takes from
                                  # the capture, per launch of the first kernel (DRAM bytes, warp instructions,
                                  # L1 bytes, issue utilisation); N = queries that launch answered

Writes <out>_summary.json (the counters DESIGN.md / bench.py quote, one object per profiled launch)
and <out>_details.csv (ncu's own details page)."""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_instruction",
    "smsp__thread_inst_executed_per_inst_executed.pct": "warp_execution_efficiency_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__warps_eligible.avg.per_cycle_active": "eligible_warps_per_cycle",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "lsu_wavefronts_pct",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum": "global_load_requests",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum": "global_load_sectors",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__sass_average_branch_targets_threads_uniform.pct": "branch_uniformity_pct",
    "smsp__sass_branch_targets_threads_divergent.sum": "divergent_branch_targets",
    "smsp__sass_branch_targets.sum": "branch_targets",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__occupancy_limit_registers": "occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_shared_mem",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9,
         "s": 1.0, "second": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}


def ncu(*args):
    return subprocess.run(["ncu", *args], check=True, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    head, units = rows[0], rows[1]
    launches = []
    for row in rows[2:]:
        d = {"kernel": row[head.index("Kernel Name")]}
        for k, name in KEYS.items():
            if k not in head:
                continue
            i = head.index(k)
            try:
                v = float(row[i].replace(",", ""))
            except ValueError:
                continue
            u = units[i]
            if name in ("duration", "dram_read", "dram_write") and u in SCALE:
                v *= SCALE[u]
                name_u = name + ("_s" if name == "duration" else "_bytes")
            else:
                name_u = name
            d[name_u] = v
        if "dram_read_bytes" in d and "dram_write_bytes" in d:
            d["dram_bytes"] = d["dram_read_bytes"] + d["dram_write_bytes"]
        launches.append(d)
    with open(out + "_summary.json", "w") as f:
        json.dump({"report": os.path.basename(rep), "launches": launches}, f, indent=1)
    with open(out + "_details.csv", "w") as f:
        f.write(ncu("-i", rep, "--page", "details", "--csv"))
    if "--traffic" in sys.argv and launches:
        path = os.path.join(os.path.dirname(out) or ".", "traffic.json")
        first = launches[0]
        queries = int(sys.argv[sys.argv.index("--queries") + 1]) if "--queries" in sys.argv else None
        instr = first.get("warp_instructions")
        sectors = first.get("global_load_sectors")
        with open(path, "w") as f:
            json.dump({"kernel": first["kernel"], "dram_bytes_per_launch": first.get("dram_bytes"),
                       "warp_instructions_per_launch": instr,
                       "queries_per_launch": queries,
                       "warp_instructions_per_query": (instr / queries) if (instr and queries) else None,
                       "l1_bytes_per_launch": (sectors * 32.0) if sectors else None,
                       "issue_active_pct": first.get("issue_active_pct"),
                       "l1_hit_pct": first.get("l1_hit_pct"), "l2_hit_pct": first.get("l2_hit_pct"),
                       "kernel_ms_under_ncu": first.get("duration_s", 0.0) * 1e3,
                       "source": "profiles/" + os.path.basename(out) + "_summary.json"}, f, indent=1)
    for d in launches:
        print(json.dumps(d))


if __name__ == "__main__":
    main()
