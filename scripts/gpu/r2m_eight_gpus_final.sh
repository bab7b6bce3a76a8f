set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l; free -g | head -2; nproc
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 > gpurun_out/r2m_bench_8gpu.json 2> gpurun_out/r2m_bench_8gpu.err; echo "bench exit $?"
grep "\[bench\]" gpurun_out/r2m_bench_8gpu.err; head -c 300 gpurun_out/r2m_bench_8gpu.json; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 4 --rays 0 --c4-points 0 --c5-points 0 --knn-queries 0 > gpurun_out/r2m_bench_4gpu_short.json 2> gpurun_out/r2m_bench_4gpu_short.err; echo "bench exit $?"
