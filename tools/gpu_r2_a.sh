#!/bin/bash
# round 2, first GPU session: op costs, the whole GPU test suite, the bench line (both arms)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi.txt 2>&1
timeout 120 tools/bin/op_probe > gpurun_out/r2_op_probe.txt 2>&1; echo "op_probe rc=$?" >> gpurun_out/r2_op_probe.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "bench rc=$?" >> gpurun_out/r2_bench_a.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_bench_ref_a.json 2> gpurun_out/r2_bench_ref_a.err
tail -5 gpurun_out/r2_pytest_gpu.log; cat gpurun_out/r2_op_probe.txt; tail -3 gpurun_out/r2_bench_a.err; head -c 1500 gpurun_out/r2_bench_a.json
