#!/bin/bash
# tensor-core actor kernel: diagnostics, tests, then (each only after the plain run exited 0) the ncu
# launch list and one full capture.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 180 python tools/tc_diag.py > gpurun_out/tc_diag.log 2>&1; echo "diag rc=$?"; tail -8 gpurun_out/tc_diag.log
timeout 300 python -m pytest tests/test_gpu_learner_parity.py -x -q -k "tensor_core" 2>&1 | tail -3
timeout 120 python tools/prof_tc.py > gpurun_out/prof_tc_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:actor_fwd_tc --csv \
    --log-file gpurun_out/tc_launches.csv python tools/prof_tc.py > gpurun_out/ncu_tc_list.log 2>&1
tail -7 gpurun_out/tc_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:actor_fwd_tc -s 3 -c 1 \
    -o gpurun_out/prof_tc -f python tools/prof_tc.py > gpurun_out/ncu_tc_full.log 2>&1
echo "ncu full rc=$?"
