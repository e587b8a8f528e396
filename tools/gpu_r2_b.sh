#!/bin/bash
# round 2: per-player step kernel -- parity, then A/B timing against the one-thread-per-env kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_env_parity.py tests/test_gpu_reference_callers.py -m gpu -x -q > gpurun_out/r2_pytest_env.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_env.log
tail -4 gpurun_out/r2_pytest_env.log
for pp in 1 0; do
  echo "SS_STEP_PP=$pp"
  SS_STEP_PP=$pp SS_E=65536,262144 SS_K=1,32,128 SS_ONLY=physics timeout 300 python tools/explore_step.py
done > gpurun_out/r2_step_ab.txt 2>&1
cat gpurun_out/r2_step_ab.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-learner --no-cpu-baseline > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; echo "bench rc=$?"; head -c 600 gpurun_out/r2_bench_b.json
