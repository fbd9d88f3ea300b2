cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest7.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest7.log
timeout 600 python tools/fp32_probe.py 2>&1 | tail -12
timeout 300 python tools/heavy_probe.py 63739 10000 2>&1 | tail -1
