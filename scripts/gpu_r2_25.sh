set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ab_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2ab_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ab_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2ab_smoke.log
timeout 600 python bench.py > gpurun_out/r2ab_default.json 2> gpurun_out/r2ab_default.err; echo "default bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2ab_reference.json 2> gpurun_out/r2ab_reference.err; echo "reference arm rc=$?"
timeout 400 python bench.py --workload bell_hill_3d_8M --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/r2ab_8M.json 2> gpurun_out/r2ab_8M.err; echo "8M rc=$?"
timeout 400 python bench.py --workload bell_hill_3d_1M --no-cpu-baseline --steps 40 --warmup 3 > gpurun_out/r2ab_1M.json 2> gpurun_out/r2ab_1M.err; echo "1M rc=$?"
timeout 300 python bench.py --workload witch_2d_4M --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/r2ab_2d.json 2> gpurun_out/r2ab_2d.err; echo "2d rc=$?"
timeout 300 python bench.py --workload static_2d_250k --no-cpu-baseline --steps 100 --warmup 5 > gpurun_out/r2ab_c2.json 2> gpurun_out/r2ab_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/r2ab_*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print(p, round(d['ms_per_step'],3), d['config'].get('fast_arithmetic_ms_per_step'), 'e2e', (d.get('e2e') or {}).get('value'), d['value'], (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e:
        print(p,'ERR',e)
PY
