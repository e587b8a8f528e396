#!/bin/bash
# round 2, final session: the whole GPU suite, smoke, both bench arms, launch lists and full captures of the shipped kernels
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -6 gpurun_out/r2_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_line.json 2> gpurun_out/r2_bench_line.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_line.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-learner"
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_bench_physics.csv $CMD > gpurun_out/r2_ncu_list.log 2>&1
$CMD > gpurun_out/r2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_pp -s 6 -c 1 -o gpurun_out/r2_prof_step_pp -f $CMD > gpurun_out/r2_ncu_full.log 2>&1
timeout 300 python tools/prof_all.py > /dev/null 2>&1 && timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_tick_and_update.csv python tools/prof_all.py > gpurun_out/ncu.log 2>&1
python tools/prof_grad_tc.py > gpurun_out/r2_prof_grad_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mlp_grad_tc -s 4 -c 2 -o gpurun_out/r2_prof_grad_tc -f python tools/prof_grad_tc.py > gpurun_out/r2_ncu_grad_full.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
python - <<'PY'
import json, csv
d=json.load(open('gpurun_out/r2_bench_line.json'))
print("value %.4g e2e %.4g full %.4g"%(d["value"], d["e2e"]["value"], d["e2e"]["full_outputs"]["value"]))
print("roofline", {k:d["roofline"].get(k) for k in ("frac","dram_frac","launch_us","actual_bound")}, d["roofline"]["one_tick_per_launch"]["frac"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["python_reference"].get("env_steps_per_sec"))
L=d["learner"]; print("rollout %.4g train %.4g (%.4f ms) cfg4 %.4f ms cfg5 %.4g cfg5train %.4g"%(L["rollout"]["env_steps_per_sec"], L["train"]["samples_per_sec"], L["train"]["ms_per_update"], L["selfplay_training"]["ms_per_iteration"], L["planning_actor_speed_sweep"]["env_steps_per_sec"], L["planning_actor_speed_sweep"]["train"]["samples_per_sec"]))
print("train", json.dumps(L["train"])[:600])
r=json.load(open('gpurun_out/r2_bench_reference_arm.json')); print("ref", r["value"], r["cpu_baseline"]["cores"])
rows = [r for r in csv.reader(open("gpurun_out/r2_launches_tick_and_update.csv")) if len(r) > 10 and r[0].isdigit()]
for r in rows:
    print("%-70s %8.1f us" % (r[4][:70], float(r[-1]) / 1e3))
PY
