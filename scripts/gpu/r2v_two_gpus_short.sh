set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --rays 0 --c4-points 0 --c5-points 0 --knn-queries 0 > gpurun_out/r2v_bench_2gpu_short.json 2> gpurun_out/r2v_bench_2gpu_short.err; echo "bench exit $?"
head -c 300 gpurun_out/r2v_bench_2gpu_short.json; echo
