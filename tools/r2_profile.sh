# profiles of the benchmarked build: launch list of the bench command + ncu --set full of its kernels (B200_PROFILING.md recipe)
cd $GRAFT_REPO_ROOT
CMD="python bench.py --steps 1 --warmup 1 --no-extra --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv $CMD > gpurun_out/prof_l.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:ballistic_kernel -s 4 -c 1 -f -o gpurun_out/r02_ballistic $CMD > gpurun_out/prof_b.log 2>&1; echo "ballistic rc=$?"
ncu --set full --clock-control none --import-source on -k regex:contact_kernel -s 0 -c 2 -f -o gpurun_out/r02_contact $CMD > gpurun_out/prof_c.log 2>&1; echo "contact rc=$?"
ls -la gpurun_out/*.ncu-rep
