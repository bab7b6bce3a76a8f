set -x
cd $GRAFT_REPO_ROOT
L=$PWD/r-star-tree_b200/csrc
free -g | head -2; nproc
timeout 1500 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench exit $?"
grep "\[bench\]" gpurun_out/r2f_bench.err; head -c 300 gpurun_out/r2f_bench.json; echo
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err; echo "reference exit $?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; tail -1 gpurun_out/r2f_smoke.log
export AB_VISITS=0 PROFILE_CALLS=2
timeout 300 python scripts/ab_case.py > gpurun_out/r2f_profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:box_traverse -s 1 -c 1 -o gpurun_out/r2f_box_traverse python scripts/ab_case.py > gpurun_out/r2f_ncu_box.log 2>&1
export PROFILE_N=10000000 PROFILE_Q=4000000 CPU=0
timeout 300 python scripts/knn_case.py > gpurun_out/r2f_knn_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:knn_traverse -s 5 -c 1 -o gpurun_out/r2f_knn_traverse python scripts/knn_case.py > gpurun_out/r2f_ncu_knn.log 2>&1
# BASELINE config 4 at FULL size (50 M 8-D points, 32 M queries), the other sections off
RTB_C4_POINTS=50000000 RTB_C4_QUERIES=32000000 timeout 1500 python bench.py --points 1000000 --queries 1000000 --rays 0 --c5-points 0 --knn-queries 0 --steps 2 --warmup 3 > gpurun_out/r2f_bench_config4_full.json 2> gpurun_out/r2f_bench_config4_full.err; echo "config4 full exit $?"
tail -2 gpurun_out/r2f_bench_config4_full.err
