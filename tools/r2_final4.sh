cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_final.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_final_bench_1gpu.json 2> gpurun_out/r02_final_bench_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_final_bench_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'], d['source_id'], d['spec_rounds_tuning'])
print(d['work']); print(d['roofline']['traffic'], d['roofline']['frac'], d['roofline_by_kernel']['ballistic']['frac'], d['roofline_whole_pass'])
print({k:(v['value'],v['ms_per_pass']) for k,v in d['extra']['configs'].items()})
PY
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
