set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_full_size.py::test_bell_hill_3d_steps_vs_oracle_at_bench_sizes > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/r2s_tests.log
timeout 400 python bench.py --workload bell_hill_3d_8M --no-cpu-baseline --no-e2e --steps 20 --warmup 3 > gpurun_out/r2s_8M.json 2> gpurun_out/r2s_8M.err; echo "8M rc=$?"
timeout 400 python bench.py --no-cpu-baseline --no-e2e --steps 10 --warmup 3 --device-gen > gpurun_out/r2s_64M.json 2> gpurun_out/r2s_64M.err; echo "64M rc=$?"
SPHMW_CELL_ORDER=xchunk timeout 400 python bench.py --no-cpu-baseline --no-e2e --steps 10 --warmup 3 --device-gen > gpurun_out/r2s_64M_xchunk.json 2> gpurun_out/r2s_64M_xchunk.err; echo "64M xchunk rc=$?"
timeout 300 python bench.py --workload witch_2d_4M --no-cpu-baseline --no-e2e --steps 20 --warmup 3 > gpurun_out/r2s_2d.json 2> gpurun_out/r2s_2d.err; echo "2d rc=$?"
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/r2s_*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print(p, round(d['ms_per_step'],3), d['config'].get('fast_arithmetic_ms_per_step'), {k:round(v,3) for k,v in d['roofline']['per_kernel_ms_per_step'].items()})
    except Exception as e:
        print(p,'ERR',e)
PY
tail -3 gpurun_out/r2s_*.err
