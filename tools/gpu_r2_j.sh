#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_learner_parity.py -m gpu -x -q -k "rollout or fused or checkpoint or trainer" > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -8 gpurun_out/r2_pytest_gpu.log
for f in 1 0; do echo "SS_ROLLOUT_OVERLAP=$f"; SS_ROLLOUT_OVERLAP=$f timeout 120 python tools/rollout_parts.py 3584 2>&1 | tail -3; done | tee gpurun_out/r2_rollout_overlap.txt
