set -x
cd $GRAFT_REPO_ROOT
CMD="python bench.py --flags 0 --no-cpu-baseline --no-e2e --no-strict --steps 3 --warmup 1 --device-gen"
$CMD > gpurun_out/r2ae_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_binary_ -s 2 -c 2 -o gpurun_out/r2ae_final64M $CMD > gpurun_out/r2ae_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2ae_ncu.log
