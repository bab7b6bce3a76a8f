set -x
cd $GRAFT_REPO_ROOT
# the shipped build once more through the whole GPU suite and smoke() (the last GPU minutes of the round)
timeout 420 python -m pytest tests -m gpu -q > gpurun_out/r2w_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2w_pytest_gpu.log
tail -4 gpurun_out/r2w_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2w_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r2w_smoke.log
tail -3 gpurun_out/r2w_smoke.log
