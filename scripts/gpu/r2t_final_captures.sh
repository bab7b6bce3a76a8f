set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2t_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2t_pytest_gpu.log
tail -3 gpurun_out/r2t_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2t_smoke.log 2>&1; tail -1 gpurun_out/r2t_smoke.log
# full captures of the two box kernels as shipped (each after its plain run has exited 0)
export AB_VISITS=0 PROFILE_CALLS=2
timeout 300 python scripts/ab_case.py > gpurun_out/r2t_profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:box_traverse -s 1 -c 1 -o gpurun_out/r2t_box_traverse python scripts/ab_case.py > gpurun_out/r2t_ncu_box.log 2>&1
PARITY=0 timeout 400 python scripts/config4_case.py > gpurun_out/r2t_c4_plain.log 2>&1 &&
PARITY=0 ncu --set full --clock-control none --import-source on -k regex:box_traverse -s 2 -c 1 -o gpurun_out/r2t_box_traverse_8d python scripts/config4_case.py > gpurun_out/r2t_ncu_c4.log 2>&1
# launch list of a short bench (config 2 only)
BENCH_ARGS="--steps 2 --warmup 3 --no-cpu-baseline --rays 0 --c4-points 0 --c5-points 0 --knn-queries 0"
timeout 600 python bench.py $BENCH_ARGS > gpurun_out/r2t_bench_short.json 2> gpurun_out/r2t_bench_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2t_launches_config2.csv python bench.py $BENCH_ARGS > gpurun_out/r2t_ncu_launches.log 2>&1
ls -la gpurun_out | grep r2t
