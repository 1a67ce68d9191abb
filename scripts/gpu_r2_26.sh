set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_slabs_nccl.py -m gpu -x -q > gpurun_out/r2ac_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2ac_tests.log
