#!/bin/bash
# both bench arms as the driver runs them
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_line.json 2> gpurun_out/r2_bench_line.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_line.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_line.json').read().strip().splitlines()[-1])
print("value %.4g e2e %.4g full %.4g"%(d["value"], d["e2e"]["value"], d["e2e"]["full_outputs"]["value"]), d["gpu_launches"], d["ms_per_step"])
R=d["roofline"]; print({k:R.get(k) for k in ("frac","dram_frac","launch_us","traffic")}, R["actual_bound"]["frac"], R["actual_bound"]["fp64"]["frac"], R["one_tick_per_launch"]["frac"])
L=d["learner"]; print("rollout %.4g train %.4g (%.4f ms, tf %.3f) cfg4 %.4f ms cfg5 %.4g cfg5train %.4g"%(L["rollout"]["env_steps_per_sec"], L["train"]["samples_per_sec"], L["train"]["ms_per_update"], L["train"]["tensor_frac"], L["selfplay_training"]["ms_per_iteration"], L["planning_actor_speed_sweep"]["env_steps_per_sec"], L["planning_actor_speed_sweep"]["train"]["samples_per_sec"]))
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["python_reference"].get("env_steps_per_sec"), d["clocks"])
r=json.loads(open('gpurun_out/r2_bench_reference_arm.json').read().strip().splitlines()[-1]); print("ref", r["value"], r["cpu_baseline"]["cores"], r["config"]==d["config"])
PY
