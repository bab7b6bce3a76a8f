set -x
cd $GRAFT_REPO_ROOT
SWEEP=1048576,2097152,524288 timeout 600 python scripts/e2e_trace.py > gpurun_out/r2k_e2e_trace.log 2>&1; tail -40 gpurun_out/r2k_e2e_trace.log
