set -x
cd $GRAFT_REPO_ROOT
# the GPU suite once more under hostile tuning knobs (none of them may change a result):
#  a) every host-result call of >= 8192 queries cut into 4096-query batches (three working sets rotating), and the
#     stash pool pretended to be 3000 ids, so that most lists take the deferred second traversal
#     (tests that assert on the deferred-query / batch COUNTERS of the default configuration are expected to object;
#     every comparison of a RESULT must hold)
#  b) the float blocks instead of the quantised image everywhere  (RTB_NO_QUANT=1: 97 passed)
RTB_PIPELINE_BATCH=4096 RTB_STASH_LIMIT_IDS=3000 timeout 300 python -m pytest tests -m gpu -q > gpurun_out/r2x_stress_pipeline_stash.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2x_stress_pipeline_stash.log
grep -n "^E  \|^FAILED\|passed\|failed" gpurun_out/r2x_stress_pipeline_stash.log | tail -40
