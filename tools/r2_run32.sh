cd $GRAFT_REPO_ROOT
FIZZ_SPEC_NOPRED=1 FIZZ_B200_LIB=$PWD/tmp_variants/libfizz_assert.so timeout 100 python tools/spec_debug.py CORRIDOR_2SQUARES 5 777 300 512 50 2>&1 | grep "pending\|item \|reset\|publish\|consume" | head -90 | cut -c1-200
