set -x
cd $GRAFT_REPO_ROOT
# bench.py after config 5 gained its CPU arm: a small full-path run (every section, CPU arms and parity on)
timeout 280 python bench.py --points 1000000 --queries 2000000 --rays 2000000 --c4-points 200000 --c4-queries 400000 --c5-points 2000000 --c5-queries 8000000 --knn-queries 400000 --steps 3 --warmup 3 > gpurun_out/r2z2_bench_small.json 2> gpurun_out/r2z2_bench_small.err; echo "bench exit $?"
tail -3 gpurun_out/r2z2_bench_small.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2z2_bench_small.json'))
print("value", d["value"], "parity", d["parity"]["bit_exact_vs_reference_search"])
for k in ("config4","config5","knn"):
    print(k, d[k]["value"], d[k].get("cpu_baseline"), d[k]["parity"]["bit_exact_vs_reference_search"])
P
