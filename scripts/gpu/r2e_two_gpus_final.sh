set -x
cd $GRAFT_REPO_ROOT
free -g | head -2; nproc
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2e_pytest_gpu.log
tail -3 gpurun_out/r2e_pytest_gpu.log
PROFILE_N=10000000 PROFILE_Q=16000000 CPU=0 timeout 300 python scripts/knn_case.py > gpurun_out/r2e_knn.log 2>&1; tail -4 gpurun_out/r2e_knn.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 > gpurun_out/r2e_bench_2gpu.json 2> gpurun_out/r2e_bench_2gpu.err; echo "bench exit $?"
grep "\[bench\]" gpurun_out/r2e_bench_2gpu.err; head -c 300 gpurun_out/r2e_bench_2gpu.json; echo
