cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/pytest28.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest28.log
for cfg in "1 32 700 48 2" "1 64 700 48 2" "1 32 700 32 2" "1 32 350 48 2" "1 32 700 48 1"; do
  set -- $cfg
  echo "== FIZZ_SPEC=$1 S=$2 D=$3 COST=$4 CTAS=$5"
  FIZZ_SPEC=$1 FIZZ_SPEC_S=$2 FIZZ_SPEC_D=$3 FIZZ_SPEC_COST=$4 FIZZ_SPEC_CTAS=$5 FIZZ_B200_LIB=$PWD/tmp_variants/libfizz_dbg.so timeout 120 python tools/timeline_probe.py 96 hybrid 2>&1 | tail -64 | cut -c1-400 | grep -v "^   [0-9]*:\|worst worlds\|ordinary-launch"
done
