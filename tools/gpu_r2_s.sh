#!/bin/bash
# round 2: CTA sizes of the physics kernels chosen from the env count: tests, sweep, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_env_parity.py tests/test_gpu_reference_callers.py -m gpu -x -q > gpurun_out/r2_pytest_env.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_env.log
tail -5 gpurun_out/r2_pytest_env.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-learner > gpurun_out/r2_bench_physics.json 2> gpurun_out/r2_bench_physics.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_physics.json').read().strip().splitlines()[-1])
print("value %.4g e2e %.4g launch_us %.1f frac %.3f"%(d["value"], d["e2e"]["value"], d["roofline"]["launch_us"], d["roofline"]["frac"]), d["roofline"].get("actual_bound",{}).get("frac"), d["roofline"]["one_tick_per_launch"])
PY
