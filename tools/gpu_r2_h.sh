#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -12 gpurun_out/r2_pytest_gpu.log
for pp in 1 0; do echo "SS_STEP_PP=$pp"; SS_STEP_PP=$pp timeout 200 python tools/rollout_parts.py 3456 2>&1 | tail -6; done | tee gpurun_out/r2_rollout_parts.txt
