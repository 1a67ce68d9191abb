set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_slabs.py tests/test_gpu_slabs_nccl.py -m gpu -x -q > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2w_tests.log
run() { # N name extra...
  N=$1; name=$2; shift; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N$N bench.py --gpus $N --warmup 3 --no-cpu-baseline --device-gen "$@" > gpurun_out/r2w_$name.json 2> gpurun_out/r2w_$name.err; echo "$name rc=$?"
}
run 8 64M_n8 --steps 20
run 4 64M_n4 --steps 20
run 8 256M_n8 --steps 10 --workload bell_hill_3d_256M --no-e2e
run 8 8M_n8 --steps 40 --workload bell_hill_3d_8M --no-e2e
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/r2w_*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print(p, round(d['ms_per_step'],3), d['config'].get('fast_arithmetic_ms_per_step'), 'e2e', d.get('e2e',{}).get('value'), d['value'], {k:round(v,3) for k,v in d['roofline']['per_kernel_ms_per_step'].items()})
    except Exception as e:
        print(p,'ERR',e)
PY
