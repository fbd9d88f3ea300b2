set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest1.log 2>&1; echo "pytest rc=$?" 
tail -15 gpurun_out/pytest1.log
timeout 300 python tools/heavy_probe.py 63739,12860 10000 > gpurun_out/heavy1.log 2>&1; cat gpurun_out/heavy1.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench1.err
python - <<'P'
import json
for f in ['gpurun_out/bench1.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('roofline',{}).get('frac'), d['kernels'])
        print(d.get('cpu_baseline'))
    except Exception as e: print(f, 'ERR', e)
P
for c in 32 48 64; do timeout 300 python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --chunk $c > gpurun_out/bench_c$c.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/bench_c$c.json').read().strip().splitlines()[-1]);print('chunk',$c,d['value'],d['ms_per_step'],{k:v['ms_per_pass'] for k,v in d['kernels'].items()})"; done
