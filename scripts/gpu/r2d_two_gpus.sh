set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 600 python -m pytest tests -m gpu -q -k "strict or counts32 or any_hit" > gpurun_out/r2d_pytest_subset.log 2>&1; tail -3 gpurun_out/r2d_pytest_subset.log
nvcc -O2 -std=c++17 -o /tmp/pcie_probe scripts/pcie_probe.cu && timeout 400 /tmp/pcie_probe > gpurun_out/r2d_pcie_probe_2gpu.log 2>&1; grep -E "2 GPU" gpurun_out/r2d_pcie_probe_2gpu.log
export RTB_C5_POINTS=10000000
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2d_bench_2gpu.json 2> gpurun_out/r2d_bench_2gpu.err; echo "bench exit $?"
tail -5 gpurun_out/r2d_bench_2gpu.err; head -c 300 gpurun_out/r2d_bench_2gpu.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2d_bench_2gpu_reference.json 2> gpurun_out/r2d_bench_2gpu_reference.err; echo "reference arm exit $?"
tail -3 gpurun_out/r2d_bench_2gpu_reference.err; head -c 600 gpurun_out/r2d_bench_2gpu_reference.json; echo
