cd $GRAFT_REPO_ROOT
FIZZ_B200_LIB=$PWD/tmp_variants/libfizz_dbg2.so timeout 300 python tools/heavy_probe.py 235581 300 2>&1 | head -6
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest12.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest12.log
timeout 300 python tools/heavy_probe.py 63739,235581,309402 10000 2>&1 | tail -3
timeout 600 python tools/shard_probe.py 8 3 5,2 2>&1 | grep -A1 rank | cut -c1-330
