set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2j_pytest_gpu.log
tail -3 gpurun_out/r2j_pytest_gpu.log
PROFILE_RAYS=16000000 PROFILE_RES=1.0 PROFILE_CALLS=3 timeout 300 python scripts/ray_case.py > gpurun_out/r2j_rays_9k.log 2>&1; grep device gpurun_out/r2j_rays_9k.log
PROFILE_RAYS=16000000 PROFILE_RES=5.9 PROFILE_CALLS=3 timeout 300 python scripts/ray_case.py > gpurun_out/r2j_rays_342k.log 2>&1; grep device gpurun_out/r2j_rays_342k.log
PROFILE_N=10000000 PROFILE_Q=16000000 CPU=0 timeout 300 python scripts/knn_case.py > gpurun_out/r2j_knn.log 2>&1; tail -3 gpurun_out/r2j_knn.log
