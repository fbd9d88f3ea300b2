cd $GRAFT_REPO_ROOT
echo "=== r01 lib timeline"; FIZZ_B200_LIB=$PWD/tmp_variants/libfizz_r01_dbg.so timeout 300 python tools/timeline_probe.py 96 2>&1 | tail -4
echo "=== new lib timeline"; FIZZ_B200_LIB=$PWD/tmp_variants/libfizz_dbg.so timeout 300 python tools/timeline_probe.py 96 2>&1 | tail -4
echo "=== new lib timeline chunk 32"; FIZZ_B200_LIB=$PWD/tmp_variants/libfizz_dbg.so timeout 300 python tools/timeline_probe.py 32 2>&1 | tail -2
echo "=== new lib hybrid_sync"; FIZZ_B200_LIB=$PWD/tmp_variants/libfizz_dbg.so timeout 300 python tools/timeline_probe.py 96 hybrid_sync 2>&1 | tail -2
