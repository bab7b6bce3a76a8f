set -x
cd $GRAFT_REPO_ROOT
L=$PWD/r-star-tree_b200/csrc
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest_gpu.log
tail -4 gpurun_out/r2c_pytest_gpu.log
# the driver's bench command, default sizes, N = 1
timeout 1500 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench exit $?"
tail -3 gpurun_out/r2c_bench.err; head -c 400 gpurun_out/r2c_bench.json; echo
# A/B of the loop barriers on the final kernel
AB_VISITS=1 timeout 300 python scripts/ab_case.py > gpurun_out/r2c_ab_default.log 2>&1
AB_VISITS=0 RTB_LIB=$L/librtree_b200_sync.so timeout 300 python scripts/ab_case.py > gpurun_out/r2c_ab_sync.log 2>&1
grep -h "traverse ms" gpurun_out/r2c_ab_*.log
# host side of the copies (one GPU here; the multi-GPU run repeats it)
nvcc -O2 -std=c++17 -o /tmp/pcie_probe scripts/pcie_probe.cu && timeout 300 /tmp/pcie_probe > gpurun_out/r2c_pcie_probe_1gpu.log 2>&1; tail -30 gpurun_out/r2c_pcie_probe_1gpu.log
# ncu: launch list of a short bench, then full captures of the traversal kernels
BENCH_ARGS="--steps 2 --warmup 3 --no-cpu-baseline --rays 0 --c4-points 0 --c5-points 0 --knn-queries 0"
timeout 600 python bench.py $BENCH_ARGS > gpurun_out/r2c_bench_short.json 2> gpurun_out/r2c_bench_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches_config2.csv python bench.py $BENCH_ARGS > gpurun_out/r2c_ncu_launches.log 2>&1
export AB_VISITS=0 PROFILE_CALLS=2
timeout 300 python scripts/ab_case.py > gpurun_out/r2c_profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:box_traverse -s 1 -c 1 -o gpurun_out/r2c_box_traverse python scripts/ab_case.py > gpurun_out/r2c_ncu_box.log 2>&1
PARITY=0 timeout 400 python scripts/config4_case.py > gpurun_out/r2c_c4_plain.log 2>&1 &&
PARITY=0 ncu --set full --clock-control none --import-source on -k regex:box_traverse -s 2 -c 1 -o gpurun_out/r2c_box_traverse_8d python scripts/config4_case.py > gpurun_out/r2c_ncu_c4.log 2>&1
export PROFILE_N=10000000 PROFILE_Q=4000000 CPU=0
timeout 300 python scripts/knn_case.py > gpurun_out/r2c_knn_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:knn_traverse -s 5 -c 1 -o gpurun_out/r2c_knn_traverse python scripts/knn_case.py > gpurun_out/r2c_ncu_knn.log 2>&1
export PROFILE_RAYS=8000000 PROFILE_RES=5.9 PROFILE_CALLS=2
timeout 300 python scripts/ray_case.py > gpurun_out/r2c_ray_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ray_traverse -s 1 -c 1 -o gpurun_out/r2c_ray_traverse python scripts/ray_case.py > gpurun_out/r2c_ncu_ray.log 2>&1
ls -la gpurun_out/ | tail -30
