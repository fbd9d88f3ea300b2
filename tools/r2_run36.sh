cd $GRAFT_REPO_ROOT
for cfg in "32 200 48 2" "32 400 48 2" "32 100 48 2"; do
  set -- $cfg
  echo "== S=$1 D=$2 COST=$3 CTAS=$4"
  FIZZ_SPEC_S=$1 FIZZ_SPEC_D=$2 FIZZ_SPEC_COST=$3 FIZZ_SPEC_CTAS=$4 FIZZ_B200_LIB=$PWD/tmp_variants/libfizz_dbg.so timeout 120 python tools/timeline_probe.py 96 hybrid 2>&1 | tail -64 | cut -c1-330 | grep -v "^   [0-9]*: [0-9]\|ordinary-launch\|cumulative: ball\|spec rounds per chunk" | grep -v "^   chunk  *[0-9]*: [0-9]* / " | grep -v "^   [0-9]*: (window" | tail -22
  FIZZ_SPEC_S=$1 FIZZ_SPEC_D=$2 FIZZ_SPEC_COST=$3 FIZZ_SPEC_CTAS=$4 FIZZ_B200_LIB=$PWD/tmp_variants/libfizz_dbg.so timeout 120 python tools/timeline_probe.py 96 hybrid 2>&1 | grep -A6 "worst worlds" | tail -7 | cut -c1-200
done
