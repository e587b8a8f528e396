#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_env_parity.py -m gpu -x -q -k "not bench_shape_full" 2>&1 | tail -2
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -4
