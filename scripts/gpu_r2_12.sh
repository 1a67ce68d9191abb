set -x
cd $GRAFT_REPO_ROOT
for f in 0 128 1 129; do
  timeout 300 python bench.py --workload witch_2d_4M --flags $f --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-strict > gpurun_out/r2o_2d_f$f.json 2> gpurun_out/r2o_2d_f$f.err; echo "2D flags $f rc=$?"
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --device-gen > gpurun_out/r2o_64M.json 2> gpurun_out/r2o_64M.err; echo "64M rc=$?"
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/r2o_*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print(p, round(d['ms_per_step'],4), '%.4g'%d['value'], 'e2e', d['e2e'] and '%.4g'%d['e2e']['value'], {k:round(v,3) for k,v in d['roofline']['per_kernel_ms_per_step'].items() if 'wcsph' in k or 'gather' in k})
    except Exception as e:
        print(p,'ERR',e)
PY
