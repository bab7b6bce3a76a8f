set -x
cd $GRAFT_REPO_ROOT
L=$PWD/r-star-tree_b200/csrc
timeout 1000 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest_gpu.log
tail -4 gpurun_out/r2b_pytest_gpu.log
# A/B: default, no explicit loop barriers, top levels staged in shared memory (with / without barriers)
AB_VISITS=0 timeout 300 python scripts/ab_case.py > gpurun_out/r2b_ab_default.log 2>&1
for v in nosync stage stagenosync; do
  AB_VISITS=0 RTB_LIB=$L/librtree_b200_$v.so timeout 300 python scripts/ab_case.py > gpurun_out/r2b_ab_$v.log 2>&1
done
grep -h "traverse ms" gpurun_out/r2b_ab_*.log
# 8-D: whole records (default), row by row, deferred link row
PARITY=0 timeout 400 python scripts/config4_case.py > gpurun_out/r2b_c4_default.log 2>&1
for v in progressive deferlink; do
  PARITY=0 RTB_LIB=$L/librtree_b200_$v.so timeout 400 python scripts/config4_case.py > gpurun_out/r2b_c4_$v.log 2>&1
done
grep -h "call 2" gpurun_out/r2b_c4_*.log
# ncu: the traversal kernel of the second call, default build and the build without loop barriers
export AB_VISITS=0 PROFILE_CALLS=2
timeout 300 python scripts/ab_case.py > gpurun_out/r2b_profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:box_traverse -s 1 -c 1 -o gpurun_out/r2b_box_traverse python scripts/ab_case.py > gpurun_out/r2b_ncu_default.log 2>&1
RTB_LIB=$L/librtree_b200_nosync.so timeout 300 python scripts/ab_case.py > gpurun_out/r2b_profile_plain_nosync.log 2>&1 &&
RTB_LIB=$L/librtree_b200_nosync.so ncu --set full --clock-control none --import-source on -k regex:box_traverse -s 1 -c 1 -o gpurun_out/r2b_box_traverse_nosync python scripts/ab_case.py > gpurun_out/r2b_ncu_nosync.log 2>&1
ls -la gpurun_out/
