# profiling aid: one full ncu capture of the step kernels + the launch list of a few steady-state steps
TAG=${1:-prof}
CMD="python bench.py --steps 12 --warmup 5 --burn-in 300 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:mg_step_ -s 1240 -c 4 -o gpurun_out/${TAG}_step $CMD > gpurun_out/prof_ncu.log 2>&1
tail -3 gpurun_out/prof_ncu.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 1530 -c 50 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/prof_ncu2.log 2>&1
tail -2 gpurun_out/prof_ncu2.log
