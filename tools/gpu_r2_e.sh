#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -15 gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_e.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_e.json'))
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["actual_bound"])
print(json.dumps(d["learner"]["planning_actor_speed_sweep"],indent=1))
print(d["cpu_baseline"]["python_reference"])
PY
