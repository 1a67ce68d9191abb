set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_slabs_nccl.py tests/test_gpu_capi_c.py -m gpu -q > gpurun_out/r2i_multi_tests.log 2>&1; echo "multi-gpu tests rc=$?"; tail -8 gpurun_out/r2i_multi_tests.log
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29622"
timeout 600 $TR8 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --device-gen > gpurun_out/r2i_64M_n8_lib.json 2> gpurun_out/r2i_64M_n8_lib.err; echo "n8 lib rc=$?"
timeout 600 $TR8 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --device-gen --python-transport > gpurun_out/r2i_64M_n8_py.json 2> gpurun_out/r2i_64M_n8_py.err; echo "n8 py rc=$?"
timeout 600 $TR8 bench.py --gpus 8 --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --device-gen > gpurun_out/r2i_64M_n8_lib100.json 2> gpurun_out/r2i_64M_n8_lib100.err; echo "n8 lib 100 rc=$?"
timeout 600 $TR4 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline --device-gen > gpurun_out/r2i_64M_n4_lib.json 2> gpurun_out/r2i_64M_n4_lib.err; echo "n4 lib rc=$?"
tail -3 gpurun_out/r2i_64M_n8_lib.err
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/r2i_*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        pk=d['roofline']['per_kernel_ms_per_step']
        print(p, round(d['ms_per_step'],3), 'kernel sum', round(sum(pk.values()),3), 'e2e', d['e2e'] and d['e2e']['value'], d['config'].get('comm'), {k:round(v,3) for k,v in pk.items()})
    except Exception as e:
        print(p,'ERR',e)
PY
