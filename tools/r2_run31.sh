cd $GRAFT_REPO_ROOT
echo "== asserts build, NOPRED"
FIZZ_SPEC_NOPRED=1 FIZZ_B200_LIB=$PWD/tmp_variants/libfizz_assert.so timeout 200 python tools/spec_debug.py CORRIDOR_2SQUARES 5 777 300 512 50 2>&1 | tail -12
echo "== release, NOPRED, n=64"
FIZZ_SPEC_NOPRED=1 timeout 200 python tools/spec_debug.py CORRIDOR_2SQUARES 5 64 300 512 50 2>&1 | tail -12
echo "== release, NOPRED, n=777 CTAS=1"
FIZZ_SPEC_NOPRED=1 FIZZ_SPEC_CTAS=1 timeout 200 python tools/spec_debug.py CORRIDOR_2SQUARES 5 777 300 512 50 2>&1 | tail -12
