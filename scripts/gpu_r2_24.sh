cd $GRAFT_REPO_ROOT
for q in 32 34 36 40; do
  SPHMW_PAIR_QUEUE_ROWS=$q timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-strict --steps 10 --warmup 3 --device-gen > gpurun_out/q3_$q.json 2> gpurun_out/q3_$q.err; echo "3d rows $q rc=$?"
done
for q in 20 24 28 36; do
  SPHMW_PAIR_QUEUE_ROWS=$q timeout 300 python bench.py --workload witch_2d_4M --no-cpu-baseline --no-e2e --no-strict --steps 20 --warmup 3 > gpurun_out/q2_$q.json 2> gpurun_out/q2_$q.err; echo "2d rows $q rc=$?"
done
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/q[23]_*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        k=d['roofline']['per_kernel_ms_per_step']
        print(p, round(d['ms_per_step'],3), 'density', round(k['wcsph.density_fused'],3), 'force', round(k['wcsph.momentum_fused'],3), d['config']['pair_list'])
    except Exception as e: print(p,'ERR',e)
PY
timeout 900 python -m pytest tests/test_gpu_pair_list.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -3
