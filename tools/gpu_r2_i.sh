#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_learner_parity.py tests/test_gpu_env_parity.py -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -12 gpurun_out/r2_pytest_gpu.log
for f in 1 0; do echo "SS_ROLLOUT_FUSED=$f"; SS_ROLLOUT_FUSED=$f timeout 200 python tools/rollout_parts.py 3584 2>&1 | tail -6; done | tee gpurun_out/r2_rollout_parts.txt
