cd $GRAFT_REPO_ROOT
echo "=== new lib timeline"; FIZZ_B200_LIB=$PWD/tmp_variants/libfizz_dbg.so timeout 300 python tools/timeline_probe.py 96 2>&1 | tail -4
timeout 300 python tools/first_chunk.py 96 > gpurun_out/fc_plain.log 2>&1 && cat gpurun_out/fc_plain.log && timeout 900 ncu --set full --clock-control none --import-source on -k regex:contact_kernel -s 1 -c 1 -f -o gpurun_out/r02_chunk1 python tools/first_chunk.py 96 > gpurun_out/ncu_fc.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_fc.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "runner or trapped or hybrid_equals" 2>&1 | tail -3
