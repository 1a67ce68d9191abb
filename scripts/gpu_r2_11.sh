set -x
cd $GRAFT_REPO_ROOT
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29632"
timeout 600 $TR8 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --device-gen > gpurun_out/r2n_64M_n8.json 2> gpurun_out/r2n_64M_n8.err; echo "n8 64M rc=$?"
timeout 600 $TR8 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --device-gen --workload bell_hill_3d_8M > gpurun_out/r2n_8M_n8.json 2> gpurun_out/r2n_8M_n8.err; echo "n8 8M rc=$?"
timeout 900 $TR8 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --device-gen --workload bell_hill_3d_256M > gpurun_out/r2n_256M_n8.json 2> gpurun_out/r2n_256M_n8.err; echo "n8 256M rc=$?"
timeout 600 $TR4 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline --device-gen > gpurun_out/r2n_64M_n4.json 2> gpurun_out/r2n_64M_n4.err; echo "n4 64M rc=$?"
tail -3 gpurun_out/r2n_256M_n8.err
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/r2n_*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        pk=d['roofline']['per_kernel_ms_per_step']
        print(p, d['config']['particles'], round(d['ms_per_step'],3), '%.4g'%d['value'], 'kernel sum', round(sum(pk.values()),3), 'e2e', d['e2e'] and '%.4g'%d['e2e']['value'], d['config'].get('comm'))
    except Exception as e:
        print(p,'ERR',e)
PY
