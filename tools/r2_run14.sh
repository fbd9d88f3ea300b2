cd $GRAFT_REPO_ROOT
for n in base u1 u1c3 c3 c5; do echo "== $n"; FIZZ_B200_LIB=$PWD/tmp_variants/libfizz_$n.so timeout 300 python tools/first_chunk.py 96 2>&1 | tail -1 | cut -c1-60; FIZZ_B200_LIB=$PWD/tmp_variants/libfizz_$n.so timeout 300 python tools/heavy_probe.py 63739 3000 2>&1 | tail -1; done
