set -x
cd $GRAFT_REPO_ROOT
L=$GRAFT_REPO_ROOT/r-star-tree_b200/csrc
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2p_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2p_pytest_gpu.log
tail -3 gpurun_out/r2p_pytest_gpu.log
# headline kernel: stack pointer as an address + split leaf bookkeeping (new) against the previous build
AB_VISITS=0 timeout 300 python scripts/ab_case.py > gpurun_out/r2p_ab_new.log 2>&1; tail -1 gpurun_out/r2p_ab_new.log
AB_VISITS=0 RTB_LIB=$L/librtree_b200_prev.so timeout 300 python scripts/ab_case.py > gpurun_out/r2p_ab_prev.log 2>&1; tail -1 gpurun_out/r2p_ab_prev.log
AB_VISITS=0 timeout 300 python scripts/ab_case.py > gpurun_out/r2p_ab_new2.log 2>&1; tail -1 gpurun_out/r2p_ab_new2.log
# 8-D kernel: new, the 8-D kernel left as it was apart from the leaf bookkeeping (relwide), previous build
PARITY=0 timeout 400 python scripts/config4_case.py > gpurun_out/r2p_c4_new.log 2>&1; tail -2 gpurun_out/r2p_c4_new.log
PARITY=0 RTB_LIB=$L/librtree_b200_relwide.so timeout 400 python scripts/config4_case.py > gpurun_out/r2p_c4_relwide.log 2>&1; tail -2 gpurun_out/r2p_c4_relwide.log
PARITY=0 RTB_LIB=$L/librtree_b200_prev.so timeout 400 python scripts/config4_case.py > gpurun_out/r2p_c4_prev.log 2>&1; tail -2 gpurun_out/r2p_c4_prev.log
# batch plans of the pipelined host call (sizes in K queries)
PLANS="512,1024,2048,*3,2048,1024,512;512,1024,2048,*3,1024,512;512,2048,*3,2048,512;512,2048,*3,1024,512;256,1024,*4,1024,256;512,1536,*4,1024,512;1024,*4,1024;512,*5,512;1024,2048,*3,2048,1024;512,1024,*4,1536,768,384;256,512,1024,2048,*3,2048,1024,512,256;512,1024,*5,1024,512;768,2304,*3,1536,768;512,*4,512;1024,*3,1024;256,768,2048,*3,2048,768,256;384,1152,*4,1152,384" \
  timeout 600 python scripts/e2e_trace.py > gpurun_out/r2p_plans.log 2>&1; grep "best of" gpurun_out/r2p_plans.log
