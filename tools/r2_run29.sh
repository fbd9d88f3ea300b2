cd $GRAFT_REPO_ROOT
T='tests/test_gpu_parity.py::test_batch_vs_oracle'
FIZZ_SPEC=0 timeout 300 python -m pytest "$T" -m gpu -q -x -k "notrace and hybrid50" 2>&1 | tail -2
timeout 300 python -m pytest "$T" -m gpu -q -x -k "notrace and hybrid50" 2>&1 | tail -2
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest "$T" -m gpu -q -x -k "notrace and hybrid50 and RAMPWORLD" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; grep -m 30 -B2 -A12 "Invalid\|ERROR SUMMARY\|at .*contact_kernel\|at .*spec" gpurun_out/sanitizer.log | head -80
