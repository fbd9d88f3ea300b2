set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest2.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/pytest2.log
timeout 300 python tools/heavy_probe.py 63739,12860 10000 > gpurun_out/heavy2.log 2>&1; cat gpurun_out/heavy2.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:v['ms_per_pass'] for k,v in d['kernels'].items()})
P
timeout 300 python tools/tail_one.py 63739 256 > gpurun_out/tail_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:contact_kernel -s 1 -c 1 -f -o gpurun_out/r02_tail python tools/tail_one.py 63739 256 > gpurun_out/ncu_tail.log 2>&1; echo "ncu rc=$?"; tail -5 gpurun_out/ncu_tail.log
