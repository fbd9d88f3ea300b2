cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/pytest19.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest19.log
for cfg in "0 32 700 48" "1 32 700 48" "1 64 700 48" "1 32 700 100" "1 32 1400 48" "1 16 700 48" "1 32 700 24"; do
  set -- $cfg
  echo "== FIZZ_SPEC=$1 S=$2 D=$3 COST=$4"
  FIZZ_SPEC=$1 FIZZ_SPEC_S=$2 FIZZ_SPEC_D=$3 FIZZ_SPEC_COST=$4 FIZZ_B200_LIB=$PWD/tmp_variants/libfizz_dbg.so timeout 300 python tools/timeline_probe.py 96 hybrid 2>&1 | tail -4 | cut -c1-900
done
