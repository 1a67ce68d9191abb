# A/B of library variants built by scripts/build_variant.sh: VARIANTS="a b c" bash scripts/gpu_ab.sh [bench args]
cd $GRAFT_REPO_ROOT
cp sph_mountain_waves_b200/libsphmw.so /tmp/libsphmw_default.so
for v in $VARIANTS; do
  cp sph_mountain_waves_b200/build/variants/libsphmw_$v.so sph_mountain_waves_b200/libsphmw.so
  timeout 400 python bench.py --no-cpu-baseline --no-e2e --no-strict --steps 10 --warmup 3 --device-gen "$@" > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err; echo "$v rc=$?"
done
cp /tmp/libsphmw_default.so sph_mountain_waves_b200/libsphmw.so
python - <<'PY'
import json,glob,os
for v in os.environ["VARIANTS"].split():
    p=f'gpurun_out/ab_{v}.json'
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        k=d['roofline']['per_kernel_ms_per_step']
        print(v, round(d['ms_per_step'],3), 'density', round(k['wcsph.density_fused'],3), 'force', round(k['wcsph.momentum_fused'],3), 'gather', round(k['cell_gather'],3))
    except Exception as e:
        print(p,'ERR',e)
PY
