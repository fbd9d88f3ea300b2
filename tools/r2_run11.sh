cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest11.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest11.log
timeout 300 python tools/heavy_probe.py 63739,235581,309402 10000 2>&1 | tail -3
timeout 600 python tools/shard_probe.py 8 3 5,2 2>&1 | grep -A1 rank | cut -c1-330
