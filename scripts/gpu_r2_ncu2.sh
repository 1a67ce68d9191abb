set -x
cd $GRAFT_REPO_ROOT
CMD="python bench.py --workload bell_hill_3d_8M --flags 33 --no-cpu-baseline --no-e2e --no-strict --steps 2 --warmup 1 --device-gen"
$CMD > gpurun_out/r2e_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_binary_ -s 4 -c 2 -o gpurun_out/r2e_records $CMD > gpurun_out/r2e_ncu.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/r2e_ncu.log
