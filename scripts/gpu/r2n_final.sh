set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench exit $?"
grep "\[bench\]" gpurun_out/r2n_bench.err; head -c 300 gpurun_out/r2n_bench.json; echo
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1; tail -1 gpurun_out/r2n_smoke.log
BENCH_ARGS="--steps 2 --warmup 3 --no-cpu-baseline --rays 0 --c4-points 0 --c5-points 0 --knn-queries 0"
timeout 600 python bench.py $BENCH_ARGS > gpurun_out/r2n_bench_short.json 2> gpurun_out/r2n_bench_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2n_launches_config2.csv python bench.py $BENCH_ARGS > gpurun_out/r2n_ncu_launches.log 2>&1
