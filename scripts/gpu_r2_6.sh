set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_slabs_nccl.py -m gpu -x -q > gpurun_out/r2h_nccl_tests.log 2>&1; echo "nccl tests rc=$?"; tail -30 gpurun_out/r2h_nccl_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --device-gen > gpurun_out/r2h_64M_n2_lib.json 2> gpurun_out/r2h_64M_n2_lib.err; echo "n2 lib rc=$?"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --device-gen --python-transport > gpurun_out/r2h_64M_n2_py.json 2> gpurun_out/r2h_64M_n2_py.err; echo "n2 py rc=$?"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --device-gen --workload bell_hill_3d_8M > gpurun_out/r2h_8M_n2_lib.json 2> gpurun_out/r2h_8M_n2_lib.err; echo "8M n2 lib rc=$?"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --device-gen --workload bell_hill_3d_8M --python-transport > gpurun_out/r2h_8M_n2_py.json 2> gpurun_out/r2h_8M_n2_py.err; echo "8M n2 py rc=$?"
tail -5 gpurun_out/r2h_64M_n2_lib.err
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/r2h_*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        pk=d['roofline']['per_kernel_ms_per_step']
        print(p, round(d['ms_per_step'],3), 'kernel sum', round(sum(pk.values()),3), 'e2e', d['e2e'] and d['e2e']['value'], d['config'].get('comm'))
    except Exception as e:
        print(p,'ERR',e)
PY
