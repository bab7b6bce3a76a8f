set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2i_pytest_gpu.log
tail -3 gpurun_out/r2i_pytest_gpu.log
PARITY=1 timeout 400 python scripts/config4_case.py > gpurun_out/r2i_c4_packed_leaves.log 2>&1; tail -5 gpurun_out/r2i_c4_packed_leaves.log
AB_VISITS=0 timeout 300 python scripts/ab_case.py > gpurun_out/r2i_ab_default.log 2>&1; tail -1 gpurun_out/r2i_ab_default.log
PARITY=0 timeout 400 python scripts/config4_case.py > gpurun_out/r2i_c4_plain.log 2>&1 &&
PARITY=0 ncu --set full --clock-control none --import-source on -k regex:box_traverse -s 2 -c 1 -o gpurun_out/r2i_box_traverse_8d python scripts/config4_case.py > gpurun_out/r2i_ncu_c4.log 2>&1
