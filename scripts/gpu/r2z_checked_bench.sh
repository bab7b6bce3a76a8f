set -x
cd $GRAFT_REPO_ROOT
# the assertions of the checked build at full size: config 2 (10 M points, 16 M queries, device-resident and the
# pipelined host call), rays on both meshes, the config-4 cut (8-D, 16-wide) and kNN -- numbers from this run are NOT
# bench values (assertions on), only "no assertion tripped" is
RTB_LIB=$GRAFT_REPO_ROOT/r-star-tree_b200/csrc/librtree_b200_checked.so timeout 300 python bench.py --no-cpu-baseline --steps 2 --warmup 3 --rays 16000000 --c5-points 0 > gpurun_out/r2z_checked_bench.json 2> gpurun_out/r2z_checked_bench.err; echo "bench exit $?"
grep -c "rtb check" gpurun_out/r2z_checked_bench.err
tail -5 gpurun_out/r2z_checked_bench.err
head -c 400 gpurun_out/r2z_checked_bench.json
