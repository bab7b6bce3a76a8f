set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc; free -g | head -2
timeout 1000 python -m pytest tests -m gpu -q > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest_gpu.log
tail -5 gpurun_out/r2a_pytest_gpu.log
L=r-star-tree_b200/csrc
timeout 300 python scripts/ab_case.py > gpurun_out/r2a_ab_default.log 2>&1
for v in nosync stage stagenosync; do
  AB_VISITS=0 RTB_LIB=$PWD/$L/librtree_b200_$v.so timeout 300 python scripts/ab_case.py > gpurun_out/r2a_ab_$v.log 2>&1
done
AB_VISITS=0 timeout 300 python scripts/ab_case.py > gpurun_out/r2a_ab_default2.log 2>&1
tail -3 gpurun_out/r2a_ab_*.log
PARITY=1 timeout 400 python scripts/config4_case.py > gpurun_out/r2a_c4_progressive.log 2>&1
PARITY=0 RTB_LIB=$PWD/$L/librtree_b200_wholerows.so timeout 400 python scripts/config4_case.py > gpurun_out/r2a_c4_wholerows.log 2>&1
tail -4 gpurun_out/r2a_c4_*.log
timeout 300 python scripts/knn_case.py > gpurun_out/r2a_knn.log 2>&1; tail -8 gpurun_out/r2a_knn.log
RTB_C5_POINTS=2000000 RTB_C5_QUERIES=32000000 RTB_C5_CALL=8000000 timeout 600 python bench.py --points 2000000 --queries 4000000 --rays 8000000 --steps 3 --warmup 3 --c4-points 300000 --c4-queries 1000000 --knn-queries 500000 > gpurun_out/r2a_bench_small.json 2> gpurun_out/r2a_bench_small.err; echo "bench exit $?"
tail -5 gpurun_out/r2a_bench_small.err; head -c 600 gpurun_out/r2a_bench_small.json
