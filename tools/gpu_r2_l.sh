#!/bin/bash
# round 2, session l: the update's launches as one programmatic-dependent-launch chain: tests, then A/B timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_learner_parity.py -m gpu -x -q -k "update or trainer or surface or tensor_core or replay" > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
for f in 0 1 0 1; do SS_UPDATE_PDL=$f timeout 300 python tools/update_time.py 2>&1 | grep SS_UPDATE; done | tee gpurun_out/r2_update_pdl.txt
