#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_env_parity.py -m gpu -x -q > gpurun_out/r2_pytest_env.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_env.log
tail -4 gpurun_out/r2_pytest_env.log
for pdl in 1 0; do echo "SS_STEP_PDL=$pdl"; SS_STEP_PDL=$pdl SS_E=65536,262144,1048576 SS_K=1 SS_ONLY=physics timeout 200 python tools/explore_step.py; done > gpurun_out/r2_step_pdl.txt 2>&1
cat gpurun_out/r2_step_pdl.txt
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-learner"
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_bench_physics.csv $CMD > gpurun_out/r2_ncu_list.log 2>&1
$CMD > gpurun_out/r2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_pp -s 50 -c 2 -o gpurun_out/r2_prof_step_pp $CMD > gpurun_out/r2_ncu_full.log 2>&1
tail -2 gpurun_out/r2_plain.log | head -c 400; tail -2 gpurun_out/r2_ncu_full.log
