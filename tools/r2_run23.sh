cd $GRAFT_REPO_ROOT
for cfg in "1 32 700 48"; do
  set -- $cfg
  echo "== FIZZ_SPEC=$1 S=$2 D=$3 COST=$4"
  FIZZ_SPEC=$1 FIZZ_SPEC_S=$2 FIZZ_SPEC_D=$3 FIZZ_SPEC_COST=$4 FIZZ_B200_LIB=$PWD/tmp_variants/libfizz_dbg.so timeout 300 python tools/timeline_probe.py 96 hybrid 2>&1 | grep -A17 "worst worlds" | head -18
done
