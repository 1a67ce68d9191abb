set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_slabs.py tests/test_gpu_more_schemes.py tests/test_gpu_capi_c.py -m gpu -x -q > gpurun_out/r2aa_tests.log 2>&1; echo "tests rc=$?"
tail -30 gpurun_out/r2aa_tests.log
