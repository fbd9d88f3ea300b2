cd $GRAFT_REPO_ROOT
FIZZ_SPEC_S=32 FIZZ_SPEC_D=700 FIZZ_B200_LIB=$PWD/tmp_variants/libfizz_dbg.so timeout 120 python tools/timeline_probe.py 96 hybrid 2>&1 | grep -A17 "worst worlds" | tail -18 | cut -c1-200
