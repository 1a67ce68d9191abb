set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_slabs_nccl.py -m gpu -x -q -k "flow or open_box" > gpurun_out/r2af_tests.log 2>&1; echo "tests rc=$?"
tail -25 gpurun_out/r2af_tests.log
