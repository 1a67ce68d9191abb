set -x
cd $GRAFT_REPO_ROOT
N=${NGPU:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-strict --device-gen > gpurun_out/r2x_64M_n$N.json 2> gpurun_out/r2x_64M_n$N.err; echo "bench rc=$?"
timeout 600 python -m pytest tests/test_gpu_slabs_nccl.py tests/test_gpu_capi_c.py -m gpu -x -q > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2x_tests.log
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/r2x_*.json')):
    d=json.loads(open(p).read().strip().splitlines()[-1])
    print(p, round(d['ms_per_step'],3), d['value'], (d.get('e2e') or {}).get('value'))
PY
tail -3 gpurun_out/r2x_64M_n$N.err
