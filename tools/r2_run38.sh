cd $GRAFT_REPO_ROOT
T='tests/test_gpu_parity.py::test_runner_mode_mixed_groups'
for i in 1 2 3; do FIZZ_B200_LIB=$PWD/tmp_variants/libfizz_assert.so timeout 200 python -m pytest "$T" -m gpu -q -s 2>&1 | grep -m6 "FIZZ_ASSERT\|passed\|failed\|out of range" | cut -c1-250; done
