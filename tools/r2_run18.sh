cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/pytest18.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest18.log
for cfg in "1 32 1" "1 32 2" "1 32 4" "1 16 2" "1 16 4" "1 8 4"; do
  set -- $cfg
  echo "== FIZZ_SPEC=$1 K=$2 L=$3"
  FIZZ_SPEC=$1 FIZZ_SPEC_K=$2 FIZZ_SPEC_L=$3 FIZZ_B200_LIB=$PWD/tmp_variants/libfizz_dbg.so timeout 300 python tools/timeline_probe.py 96 hybrid 2>&1 | tail -16 | cut -c1-700
done
