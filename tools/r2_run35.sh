cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "batch_vs_oracle or trapped or critical or runner or snapshot or hybrid_equals" > gpurun_out/pytest35.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest35.log
for cfg in "32 200 48 2" "32 400 48 2" "32 200 32 2" "32 200 48 3"; do
  set -- $cfg
  echo "== S=$1 D=$2 COST=$3 CTAS=$4"
  FIZZ_SPEC_S=$1 FIZZ_SPEC_D=$2 FIZZ_SPEC_COST=$3 FIZZ_SPEC_CTAS=$4 FIZZ_B200_LIB=$PWD/tmp_variants/libfizz_dbg.so timeout 120 python tools/timeline_probe.py 96 hybrid 2>&1 | tail -64 | cut -c1-330 | grep -v "^   [0-9]*: [0-9]\|ordinary-launch\|cumulative: ball\|spec rounds per chunk" | grep -v "^   chunk  *[0-9]*: [0-9]* / "
done
