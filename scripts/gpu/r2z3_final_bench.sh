set -x
cd $GRAFT_REPO_ROOT
timeout 175 python bench.py > gpurun_out/r2z3_bench.json 2> gpurun_out/r2z3_bench.err; echo "bench exit $?"
grep "\[bench\]" gpurun_out/r2z3_bench.err | tail -5; head -c 300 gpurun_out/r2z3_bench.json; echo
