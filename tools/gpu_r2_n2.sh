#!/bin/bash
# round 2, N GPUs of one box: peer-exchange correctness (pytest), the bench under torchrun (both arms)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_pytest_multi_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_multi_n$N.log
tail -5 gpurun_out/r2_pytest_multi_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n$N.err
[ -n "$SKIP_REF" ] || timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_n$N.json 2> gpurun_out/r2_bench_ref_n$N.err; echo "ref rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"], "peer_check", d.get("peer_check"), d.get("peer_check_detail"))
print(json.dumps(d["learner"].get("scaling_in_run"), indent=1))
print("train", d["learner"]["train"]["ms_per_update"], d["learner"]["train"]["samples_per_sec"], "cfg4", d["learner"]["selfplay_training"]["ms_per_iteration"])
import os
if not os.environ.get("SKIP_REF"):
    r=json.loads(open('gpurun_out/r2_bench_ref_n$N.json').read().strip().splitlines()[-1])
    print("reference arm:", r["value"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["sample"][:120])
PY
