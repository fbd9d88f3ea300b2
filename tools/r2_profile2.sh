# final profiles of the benchmarked build: bench line, launch list of the bench command, ncu --set full of its kernels
cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_final_bench_1gpu.json 2> gpurun_out/r02_final_bench_1gpu.err; echo "bench rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --no-extra --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_final_launches_bench.csv $CMD > gpurun_out/prof_l.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ballistic_kernel -s 4 -c 1 -f -o gpurun_out/r02_final_ballistic $CMD > gpurun_out/prof_b.log 2>&1; echo "ballistic rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:contact_kernel -s 0 -c 1 -f -o gpurun_out/r02_final_contact_chunk0 $CMD > gpurun_out/prof_c0.log 2>&1; echo "contact chunk0 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:contact_kernel -s 34 -c 6 -f -o gpurun_out/r02_final_contact_late $CMD > gpurun_out/prof_c1.log 2>&1; echo "contact late rc=$?"
ls -la gpurun_out/r02_final*.ncu-rep
