#!/bin/bash
# round 2, session k: the overlap test at its small size, fixed cost of the tensor-core launches, phase trace of the
# gradient kernel, launch list of one rollout tick + one update
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_learner_parity.py -m gpu -x -q -k "rollout or fused" > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
timeout 300 python tools/tc_fixed.py 2>&1 | tee gpurun_out/r2_tc_fixed.txt
timeout 120 python tools/tc_grad_trace.py 4 K 2>&1 | tee gpurun_out/r2_grad_trace_K.txt
timeout 120 python tools/tc_grad_trace.py 1 K 2>&1 | tee -a gpurun_out/r2_grad_trace_K.txt
timeout 300 python tools/prof_all.py > /dev/null 2>&1 && timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_tick_and_update.csv python tools/prof_all.py > gpurun_out/ncu.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2_launches_tick_and_update.csv")) if len(r) > 10 and r[0].isdigit()]
for r in rows:
    print("%-70s %8.1f us" % (r[4][:70], float(r[-1]) / 1e3))
PY
