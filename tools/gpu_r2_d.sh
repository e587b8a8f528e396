#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_env_parity.py -m gpu -x -q > gpurun_out/r2_pytest_env.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_env.log
tail -3 gpurun_out/r2_pytest_env.log
SS_E=65536,262144,1048576 SS_K=32,128,256 SS_ONLY=physics timeout 300 python tools/explore_step.py > gpurun_out/r2_step_pp_v3.txt 2>&1
cat gpurun_out/r2_step_pp_v3.txt
