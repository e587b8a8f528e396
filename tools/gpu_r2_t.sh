#!/bin/bash
mkdir -p gpurun_out
for b in 0 256 448 896 0 448; do echo "SS_STEP1_BLK=$b"; SS_STEP1_BLK=$b SS_ONLY=physics SS_E=65536,262144,1048576 SS_K=1 timeout 600 python tools/explore_step.py 2>&1 | tail -3; done | tee gpurun_out/r2_step1_blk.txt
