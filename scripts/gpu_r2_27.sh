set -x
cd $GRAFT_REPO_ROOT
run() { # N name extra...
  N=$1; name=$2; shift; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 296$N$N bench.py --gpus $N --warmup 3 --no-cpu-baseline --device-gen "$@" > gpurun_out/r2ad_$name.json 2> gpurun_out/r2ad_$name.err; echo "$name rc=$?"
}
run 8 64M_n8 --steps 20
run 4 64M_n4 --steps 20
run 2 64M_n2 --steps 10 --no-e2e
run 8 256M_n8 --steps 10 --workload bell_hill_3d_256M --no-e2e --no-strict
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/r2ad_*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print(p, round(d['ms_per_step'],3), d['config'].get('fast_arithmetic_ms_per_step'), 'e2e', (d.get('e2e') or {}).get('value'), d['value'])
    except Exception as e:
        print(p,'ERR',e)
PY
