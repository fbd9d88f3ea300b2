cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest6.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest6.log
timeout 300 python tools/heavy_probe.py 63739,12860 10000 2>&1 | tail -2
echo "=== new lib timeline"; FIZZ_B200_LIB=$PWD/tmp_variants/libfizz_dbg.so timeout 300 python tools/timeline_probe.py 96 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench6.json 2> gpurun_out/bench6.err; echo "bench rc=$?"; tail -3 gpurun_out/bench6.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench6.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], {k:v['ms_per_pass'] for k,v in d['kernels'].items()})
print(d['cpu_baseline'])
for k,v in d['extra']['configs'].items(): print(k, v['value'], v['ms_per_pass'], v['e2e']['value'], v['roofline_whole_pass'], {a:b['ms_per_pass'] for a,b in v['kernels'].items()})
P
