set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l; free -g | head -2; nproc
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 > gpurun_out/r2g_bench_8gpu.json 2> gpurun_out/r2g_bench_8gpu.err; echo "bench exit $?"
grep "\[bench\]" gpurun_out/r2g_bench_8gpu.err; head -c 300 gpurun_out/r2g_bench_8gpu.json; echo
nvcc -O2 -std=c++17 -o /tmp/pcie_probe scripts/pcie_probe.cu && timeout 400 /tmp/pcie_probe > gpurun_out/r2g_pcie_probe_8gpu.log 2>&1; grep -E " [48] GPU\(s\) D2H  " gpurun_out/r2g_pcie_probe_8gpu.log
