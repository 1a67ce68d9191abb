set -x
cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -m gpu -q --deselect tests/test_gpu_full_size.py::test_bell_hill_3d_steps_vs_oracle_at_bench_sizes > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2q_tests.log
python scripts/e2e_probe.py 2>&1 | grep -E "step\(8\)|capture|e2e|kernel sum|host time"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --device-gen > gpurun_out/r2q_64M.json 2> gpurun_out/r2q_64M.err; echo "64M rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload bell_hill_3d_8M > gpurun_out/r2q_8M.json 2> gpurun_out/r2q_8M.err; echo "8M rc=$?"
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/r2q_*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print(p, round(d['ms_per_step'],4), '%.4g'%d['value'], 'e2e', d['e2e'] and '%.4g'%d['e2e']['value'], 'ratio %.3f' % (d['e2e']['value']/d['value']))
    except Exception as e:
        print(p,'ERR',e)
PY
