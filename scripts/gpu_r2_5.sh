set -x
cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2g_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -25 gpurun_out/r2g_gputests.log
