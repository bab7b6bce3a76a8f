set -x
cd $GRAFT_REPO_ROOT
L=$GRAFT_REPO_ROOT/r-star-tree_b200/csrc
timeout 900 python -m pytest tests/test_gpu_boxes.py -m gpu -q > gpurun_out/r2s_pytest_boxes.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2s_pytest_boxes.log
tail -3 gpurun_out/r2s_pytest_boxes.log
# 8-D kernel: short step for node groups that are all leaves (this build) against the committed build
PARITY=1 timeout 400 python scripts/config4_case.py > gpurun_out/r2s_c4_new.log 2>&1; tail -4 gpurun_out/r2s_c4_new.log
PARITY=0 RTB_LIB=$L/librtree_b200_prev.so timeout 400 python scripts/config4_case.py > gpurun_out/r2s_c4_prev.log 2>&1; tail -3 gpurun_out/r2s_c4_prev.log
PARITY=0 timeout 400 python scripts/config4_case.py > gpurun_out/r2s_c4_new2.log 2>&1; tail -3 gpurun_out/r2s_c4_new2.log
