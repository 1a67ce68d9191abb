cd $GRAFT_REPO_ROOT
for st in 28 32 36 40 48; do
  SPHMW_PAIR_LIST_STRIDE=$st timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-strict --steps 10 --warmup 3 --device-gen > gpurun_out/st_$st.json 2> gpurun_out/st_$st.err; echo "stride $st rc=$?"
done
python - <<'PY'
import json
for st in (28,32,36,40,48):
    try:
        d=json.loads(open(f'gpurun_out/st_{st}.json').read().strip().splitlines()[-1])
        k=d['roofline']['per_kernel_ms_per_step']
        print(st, round(d['ms_per_step'],3), 'density', round(k['wcsph.density_fused'],3), 'force', round(k['wcsph.momentum_fused'],3), d['config']['pair_list'])
    except Exception as e: print(st,'ERR',e)
PY
