cd $GRAFT_REPO_ROOT
T='tests/test_gpu_parity.py::test_runner_mode_mixed_groups'
for i in 1 2 3; do FIZZ_B200_LIB=$PWD/tmp_variants/libfizz_assert.so timeout 200 python -m pytest "$T" tests/test_gpu_spec.py -m gpu -q 2>&1 | grep -m6 "FIZZ_ASSERT\|passed\|failed" | cut -c1-250; done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_r2d.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2d.log
timeout 1500 python -m pytest tests -m gpu -q -x -p no:randomly > gpurun_out/pytest_r2e.log 2>&1; echo "pytest (2nd run) rc=$?"; tail -2 gpurun_out/pytest_r2e.log
