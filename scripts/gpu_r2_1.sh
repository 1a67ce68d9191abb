set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
SPHMW_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_pair_list.py -m gpu -x -q > gpurun_out/r2_packed_tests.log 2>&1; echo "packed tests rc=$?"
tail -3 gpurun_out/r2_packed_tests.log
for f in 1 33 0 32; do
  timeout 300 python bench.py --workload bell_hill_3d_8M --flags $f --no-cpu-baseline --no-e2e --no-strict --steps 10 --device-gen > gpurun_out/r2_8M_f$f.json 2> gpurun_out/r2_8M_f$f.err; echo "8M flags $f rc=$?"
done
for f in 1 33; do
  timeout 400 python bench.py --flags $f --no-cpu-baseline --no-e2e --no-strict --steps 10 --device-gen > gpurun_out/r2_64M_f$f.json 2> gpurun_out/r2_64M_f$f.err; echo "64M flags $f rc=$?"
done
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/r2_*_f*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print(p, round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['per_kernel_ms_per_step'].items()})
    except Exception as e:
        print(p,'ERR',e)
PY
