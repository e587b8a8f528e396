#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_learner_parity.py -m gpu -x -q -k "one_launch or update" > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -3 gpurun_out/r2_pytest_gpu.log
timeout 200 python tools/pair_time.py 2>&1 | tail -4 | tee gpurun_out/r2_pair_time.txt
for f in 0 1; do SS_UPDATE_PAIR=$f timeout 300 python tools/update_time.py 65536 131072 2>&1 | grep SS_UPDATE | sed "s/^/PAIR=$f /"; done | tee gpurun_out/r2_update_pair.txt
