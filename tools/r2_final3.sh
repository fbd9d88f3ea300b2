cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2c.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2c.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02c_bench_1gpu.json 2> gpurun_out/r02c_bench_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c_bench_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'])
print(d['work'])
print({k:(v['value'],v['ms_per_pass']) for k,v in d['extra']['configs'].items()})
print(d['roofline_by_kernel'])
PY
