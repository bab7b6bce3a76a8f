set -x
cd $GRAFT_REPO_ROOT
RTB_C4_POINTS=50000000 RTB_C4_QUERIES=32000000 timeout 1500 python bench.py --points 1000000 --queries 1000000 --rays 0 --c5-points 0 --knn-queries 0 --steps 2 --warmup 3 > gpurun_out/r2o_bench_config4_full.json 2> gpurun_out/r2o_bench_config4_full.err; echo "config4 full exit $?"
tail -2 gpurun_out/r2o_bench_config4_full.err
