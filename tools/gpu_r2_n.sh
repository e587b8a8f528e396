#!/bin/bash
# round 2, session n: dependent-launch chain in the rollout loop: tests, A/B timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_learner_parity.py -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
for f in 0 1 0 1; do echo "SS_ROLLOUT_PDL=$f"; SS_ROLLOUT_PDL=$f timeout 120 python tools/rollout_parts.py 3584 2>&1 | tail -1; done | tee gpurun_out/r2_rollout_pdl.txt
