set -x
cd $GRAFT_REPO_ROOT
# the GPU suite against the checked build (make -C r-star-tree_b200/csrc checked): every traversal kernel asserts its
# stack bound (32 entries per level), the validity of every popped link and its stash / list cursors, and traps otherwise
ls -la r-star-tree_b200/csrc/librtree_b200_checked.so
RTB_LIB=$GRAFT_REPO_ROOT/r-star-tree_b200/csrc/librtree_b200_checked.so timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r2y_pytest_gpu_checked.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2y_pytest_gpu_checked.log
grep -c "rtb check" gpurun_out/r2y_pytest_gpu_checked.log
tail -5 gpurun_out/r2y_pytest_gpu_checked.log
