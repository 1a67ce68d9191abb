set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_more_schemes.py tests/test_gpu_pair_list.py tests/test_gpu_capi_c.py -m gpu -q > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2m_tests.log
./profiles/microbench/fp64_peak > gpurun_out/r2m_fp64_peak.json; cat gpurun_out/r2m_fp64_peak.json
for w in static_2d_250k witch_2d_4M bell_hill_3d_1M bell_hill_3d_8M; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_$w.json 2> gpurun_out/r2m_$w.err; echo "$w rc=$?"
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2m_bell_hill_3d_64M.json 2> gpurun_out/r2m_64M.err; echo "64M rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2m_reference_arm.json 2> gpurun_out/r2m_ref.err; echo "ref rc=$?"
CMD="python bench.py --no-cpu-baseline --no-e2e --no-strict --steps 2 --warmup 1 --device-gen"
$CMD > gpurun_out/r2m_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2m_launches_64M.csv $CMD > gpurun_out/r2m_ncu.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/r2m_*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        if 'roofline' not in d: print(p, d); continue
        print(p, d['config']['particles'], round(d['ms_per_step'],3), '%.4g'%d['value'], 'e2e %.4g'%(d['e2e']['value'] if d['e2e'] else 0), 'frac %.4f'%d['roofline']['frac'], d['config'].get('fast_arithmetic_ms_per_step'), d.get('cpu_baseline'))
    except Exception as e:
        print(p,'ERR',e)
PY
