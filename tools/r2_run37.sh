cd $GRAFT_REPO_ROOT
for cfg in "12 0 16" "8 0 16" "6 0 16" "8 400 16" "8 0 8" "8 0 64"; do
  set -- $cfg
  echo "== ABORT_X4=$1 BUDGET=$2 S=$3"
  FIZZ_SPEC_ABORT_X4=$1 FIZZ_SPEC_BUDGET=$2 FIZZ_SPEC_S=$3 FIZZ_B200_LIB=$PWD/tmp_variants/libfizz_dbg.so timeout 120 python tools/timeline_probe.py 96 hybrid 2>&1 | grep "^pass\|spec (cumulative)" | tail -2 | cut -c1-250
done
