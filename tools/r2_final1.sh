cd $GRAFT_REPO_ROOT
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/b1.err; echo "bench rc=$?"
timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/bref.err; echo "ref rc=$?"; tail -2 gpurun_out/bref.err
timeout 1500 python bench.py --steps 2 --warmup 3 --no-extra --sweep-worlds 131072,262144,524288,1048576 > gpurun_out/r02_bench_worlds_sweep.json 2> gpurun_out/bsw.err; echo "sweep rc=$?"; tail -2 gpurun_out/bsw.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_1gpu.json').read().strip().splitlines()[-1])
print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('traffic'), d['clocks'])
r=json.loads(open('gpurun_out/r02_bench_reference_arm.json').read().strip().splitlines()[-1])
print('ref', r['value'], r['cpu_baseline']['cores'], r.get('port',{}).get('value'))
s=json.loads(open('gpurun_out/r02_bench_worlds_sweep.json').read().strip().splitlines()[-1])
for x in s['extra']['worlds_sweep']: print('sweep', x['worlds'], '%.3g'%x['value'], '%.1f ms'%x['ms_per_pass'], x['roofline_whole_pass_frac'], x['kernels_ms'])
P
