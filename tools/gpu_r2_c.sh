#!/bin/bash
# round 2: ncu launch list + full capture of the per-player step kernel at the bench shape
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-learner"
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_bench_physics.csv $CMD > gpurun_out/r2_ncu_list.log 2>&1
$CMD > gpurun_out/r2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_pp -s 50 -c 2 -o gpurun_out/r2_prof_step_pp $CMD > gpurun_out/r2_ncu_full.log 2>&1
tail -2 gpurun_out/r2_ncu_full.log; ls -la gpurun_out/r2_prof_step_pp.ncu-rep
