cd $GRAFT_REPO_ROOT
T='tests/test_gpu_parity.py::test_batch_vs_oracle'
for e in "FIZZ_SPEC_NOPRED=1" "FIZZ_SPEC_S=2" "FIZZ_SPEC_CTAS=1" "FIZZ_SPEC_COST=1000000"; do
echo "== $e"
env $e timeout 300 python -m pytest "$T" -m gpu -q -k "notrace and hybrid50" 2>&1 | tail -4 | cut -c1-200
done
