cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest15.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest15.log
timeout 300 python tools/first_chunk.py 96 2>&1 | tail -1 | cut -c1-50
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench15.json 2> gpurun_out/bench15.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench15.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('traffic'), {k:v['ms_per_pass'] for k,v in d['kernels'].items()}, {k:(v['frac'],v.get('traffic')) for k,v in d['roofline_by_kernel'].items()})
for k,v in d['extra']['configs'].items(): print(k, '%.3g'%v['value'], '%.1f ms'%v['ms_per_pass'], 'e2e %.3g'%v['e2e']['value'], v['roofline_whole_pass']['frac'], {a:round(b['ms_per_pass'],1) for a,b in v['kernels'].items()})
P
