#!/bin/bash
mkdir -p gpurun_out
for b in 64 896 448 64 896; do echo "SS_STEP_BLK=$b"; SS_STEP_BLK=$b SS_ONLY=physics SS_E=65536,131072,262144,1048576 SS_K=32,256 timeout 600 python tools/explore_step.py 2>&1 | tail -8; done | tee gpurun_out/r2_step_blk_sizes.txt
