set -x
cd $GRAFT_REPO_ROOT
CMD="python bench.py --workload bell_hill_3d_8M --flags 1 --no-cpu-baseline --no-e2e --no-strict --steps 2 --warmup 1 --device-gen"
$CMD > gpurun_out/r2d_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_tile_ -s 2 -c 2 -o gpurun_out/r2d_tiles $CMD > gpurun_out/r2d_ncu.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/r2d_ncu.log
