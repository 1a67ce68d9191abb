set -x
cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -m gpu -q --deselect tests/test_gpu_full_size.py::test_bell_hill_3d_steps_vs_oracle_at_bench_sizes > gpurun_out/r2k_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -12 gpurun_out/r2k_gputests.log
timeout 600 python bench.py --steps 10 --device-gen --no-cpu-baseline > gpurun_out/r2k_64M_f0.json 2> gpurun_out/r2k_64M_f0.err; echo "64M strict rc=$?"; tail -3 gpurun_out/r2k_64M_f0.err
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/r2k_*_f*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print(p, round(d['ms_per_step'],3), d['value'], d['e2e'], d['config'].get('fast_arithmetic_ms_per_step'), {k:round(v,3) for k,v in d['roofline']['per_kernel_ms_per_step'].items()})
    except Exception as e:
        print(p,'ERR',e)
PY
