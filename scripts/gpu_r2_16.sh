set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2u_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2u_64M.json 2> gpurun_out/r2u_64M.err; echo "64M rc=$?"
timeout 400 python bench.py --workload bell_hill_3d_8M --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/r2u_8M.json 2> gpurun_out/r2u_8M.err; echo "8M rc=$?"
timeout 300 python bench.py --workload witch_2d_4M --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/r2u_2d.json 2> gpurun_out/r2u_2d.err; echo "2d rc=$?"
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/r2u_*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print(p, round(d['ms_per_step'],3), d['config'].get('fast_arithmetic_ms_per_step'), 'e2e', d.get('e2e',{}).get('value'), d['value'], {k:round(v,3) for k,v in d['roofline']['per_kernel_ms_per_step'].items()})
    except Exception as e:
        print(p,'ERR',e)
PY
