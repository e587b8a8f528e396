#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_env_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for m in 0 8 10 12 16; do echo "MINB=$m"; SS_OBS_MINB=$m SS_E=262144,1048576 SS_K=1 timeout 120 python tools/explore_step.py | grep " 1 1 looking"; done > gpurun_out/minb.log 2>&1
tail -2 gpurun_out/pytest_gpu.log; cat gpurun_out/minb.log
