#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_learner_parity.py tests/test_gpu_frames_learner.py -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
for f in 0 1 0 1; do SS_UPDATE_SAMPLE_EARLY=$f timeout 300 python tools/update_time.py 65536 524288 2>&1 | grep SS_UPDATE | sed "s/^/SAMPLE_EARLY=$f /"; done | tee gpurun_out/r2_update_sample_early.txt
