set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_slabs.py tests/test_gpu_slabs_nccl.py tests/test_gpu_capi_c.py -m gpu -x -q > gpurun_out/r2v_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2v_tests.log
N=${NGPU:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --device-gen > gpurun_out/r2v_64M_n$N.json 2> gpurun_out/r2v_64M_n$N.err; echo "bench rc=$?"
SPHMW_NO_FUSED_ADVANCE=1 SPHMW_NO_COLUMN_RANGES=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --device-gen > gpurun_out/r2v_64M_n${N}_nofold.json 2> gpurun_out/r2v_64M_n${N}_nofold.err; echo "bench nofold rc=$?"
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/r2v_*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print(p, round(d['ms_per_step'],3), d['config'].get('fast_arithmetic_ms_per_step'), 'e2e', d.get('e2e',{}).get('value'), d['value'], {k:round(v,3) for k,v in d['roofline']['per_kernel_ms_per_step'].items()})
    except Exception as e:
        print(p,'ERR',e)
PY
tail -3 gpurun_out/r2v_64M_n$N.err
