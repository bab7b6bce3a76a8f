set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2l_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2l_pytest_gpu.log
tail -3 gpurun_out/r2l_pytest_gpu.log
SWEEP=1048576 timeout 600 python scripts/e2e_trace.py > gpurun_out/r2l_e2e_trace.log 2>&1; tail -14 gpurun_out/r2l_e2e_trace.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --rays 0 --c4-points 0 --c5-points 0 --knn-queries 0 > gpurun_out/r2l_bench_2gpu_short.json 2> gpurun_out/r2l_bench_2gpu_short.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r2l_bench_2gpu_short.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'c32', d['e2e']['counts32']['value'], d['e2e']['counts32']['ms_per_step'], d['parity']['bit_exact_vs_reference_search'])"
