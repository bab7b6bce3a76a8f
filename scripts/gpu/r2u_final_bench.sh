set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench exit $?"
grep "\[bench\]" gpurun_out/r2u_bench.err; head -c 300 gpurun_out/r2u_bench.json; echo
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2u_bench_reference.json 2> gpurun_out/r2u_bench_reference.err; echo "reference exit $?"
