set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_pair_list.py tests/test_gpu_more_schemes.py -m gpu -q > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2p_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --device-gen > gpurun_out/r2p_64M.json 2> gpurun_out/r2p_64M.err; echo "64M rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload bell_hill_3d_8M > gpurun_out/r2p_8M.json 2> gpurun_out/r2p_8M.err; echo "8M rc=$?"
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --workload witch_2d_4M > gpurun_out/r2p_2d.json 2> gpurun_out/r2p_2d.err; echo "2D rc=$?"
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/r2p_*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print(p, round(d['ms_per_step'],4), '%.4g'%d['value'], 'e2e', d['e2e'] and '%.4g'%d['e2e']['value'], 'ratio %.3f' % (d['e2e']['value']/d['value']))
    except Exception as e:
        print(p,'ERR',e)
PY
