cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest16.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest16.log
timeout 600 python tools/shard_probe.py 8 5 4 2>&1 | grep -E "^rank|group 0" | cut -c1-300
timeout 300 python tools/heavy_probe.py 63739 10000 2>&1 | tail -1
timeout 600 python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config5 N=1 slice', d['value'], d['ms_per_step'], d['e2e']['value'])"
