#!/bin/bash
# round 2, session m: gradient kernel with the reordered weight-gradient group and the G1 wait elided: tests, timing, trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_learner_parity.py -m gpu -x -q -k "update or trainer or surface or tensor_core or replay or additive" > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
timeout 300 python tools/update_time.py 2>&1 | grep SS_UPDATE | tee gpurun_out/r2_update_time.txt
timeout 300 python tools/update_time.py 2>&1 | grep SS_UPDATE | tee -a gpurun_out/r2_update_time.txt
timeout 120 python tools/tc_grad_trace.py 4 2>&1 | head -75 | tee gpurun_out/r2_grad_trace.txt
SS_TRACE_RATE=0 timeout 120 python tools/tc_grad_trace.py 4 2>&1 | head -75 | tee gpurun_out/r2_grad_trace_nodrop.txt
