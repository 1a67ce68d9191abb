set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ag_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2ag_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-strict --device-gen > gpurun_out/r2ag_64M.json 2> gpurun_out/r2ag_64M.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2ag_64M.json').read().strip().splitlines()[-1]); print('ms', d['ms_per_step'], d['value'])"
