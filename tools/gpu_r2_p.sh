#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_learner_parity.py -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -8 gpurun_out/r2_pytest_gpu.log
