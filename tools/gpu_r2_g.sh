#!/bin/bash
# round 2: the whole GPU suite, both bench arms, then profiles of the shipped kernels (step kernel at the bench's launch shape,
# tensor-core critic gradient at 65,536 rows)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -6 gpurun_out/r2_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_line.json 2> gpurun_out/r2_bench_line.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_line.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-learner"
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_bench_physics.csv $CMD > gpurun_out/r2_ncu_list.log 2>&1
$CMD > gpurun_out/r2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_pp -s 25 -c 1 -o gpurun_out/r2_prof_step_pp -f $CMD > gpurun_out/r2_ncu_full.log 2>&1
python tools/prof_grad_tc.py > gpurun_out/r2_prof_grad_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mlp_grad_tc -s 4 -c 2 -o gpurun_out/r2_prof_grad_tc -f python tools/prof_grad_tc.py > gpurun_out/r2_ncu_grad_full.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_line.json'))
print("value %.4g e2e %.4g full %.4g"%(d["value"], d["e2e"]["value"], d["e2e"]["full_outputs"]["value"]))
print("roofline", {k:d["roofline"][k] for k in ("frac","dram_frac","launch_us")}, d["roofline"]["one_tick_per_launch"]["frac"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["python_reference"].get("env_steps_per_sec"))
L=d["learner"]; print("rollout %.4g train %.4g (%.4f ms) cfg4 %.4f ms cfg5 %.4g cfg5train %.4g"%(L["rollout"]["env_steps_per_sec"], L["train"]["samples_per_sec"], L["train"]["ms_per_update"], L["selfplay_training"]["ms_per_iteration"], L["planning_actor_speed_sweep"]["env_steps_per_sec"], L["planning_actor_speed_sweep"]["train"]["samples_per_sec"]))
r=json.load(open('gpurun_out/r2_bench_reference_arm.json')); print("ref", r["value"], r["cpu_baseline"]["cores"])
PY
