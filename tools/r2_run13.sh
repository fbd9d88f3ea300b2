cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest13.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest13.log
timeout 300 python tools/heavy_probe.py 63739 10000 2>&1 | tail -1
