cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest9.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest9.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench9.json 2> gpurun_out/bench9.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench9.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:v['ms_per_pass'] for k,v in d['kernels'].items()})
P
