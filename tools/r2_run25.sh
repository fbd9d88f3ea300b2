cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q > gpurun_out/pytest25.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest25.log
for cfg in "1 32 700 48" "1 32 350 48" "1 32 700 24" "1 64 350 32"; do
  set -- $cfg
  echo "== FIZZ_SPEC=$1 S=$2 D=$3 COST=$4"
  FIZZ_SPEC=$1 FIZZ_SPEC_S=$2 FIZZ_SPEC_D=$3 FIZZ_SPEC_COST=$4 FIZZ_B200_LIB=$PWD/tmp_variants/libfizz_dbg.so timeout 300 python tools/timeline_probe.py 96 hybrid 2>&1 | tail -51 | cut -c1-400 | grep -v "^   [0-9]*:\|worst worlds"
done
