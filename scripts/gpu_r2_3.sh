set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pair_list.py tests/test_gpu_slabs.py tests/test_gpu_edge_cases.py -m gpu -x -q > gpurun_out/r2c_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2c_gputests.log
for f in 1 0; do
  timeout 300 python bench.py --workload bell_hill_3d_8M --flags $f --no-cpu-baseline --no-e2e --no-strict --steps 10 --device-gen > gpurun_out/r2c_8M_f$f.json 2> gpurun_out/r2c_8M_f$f.err; echo "8M flags $f rc=$?"
done
for f in 1; do
  timeout 400 python bench.py --flags $f --no-cpu-baseline --no-e2e --no-strict --steps 10 --device-gen > gpurun_out/r2c_64M_f$f.json 2> gpurun_out/r2c_64M_f$f.err; echo "64M flags $f rc=$?"
done
timeout 300 python bench.py --workload witch_2d_4M --flags 1 --no-cpu-baseline --no-e2e --no-strict --steps 20 > gpurun_out/r2c_2d_f1.json 2> gpurun_out/r2c_2d_f1.err; echo "2D rc=$?"
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/r2c_*_f*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print(p, round(d['ms_per_step'],3), d['config']['pair_list'], {k:round(v,3) for k,v in d['roofline']['per_kernel_ms_per_step'].items() if 'wcsph' in k or 'tile' in k})
    except Exception as e:
        print(p,'ERR',e)
PY
