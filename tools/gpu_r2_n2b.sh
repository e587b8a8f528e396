#!/bin/bash
# round 2, N GPUs: the one-kernel peer exchange against the two-kernel one on the same box (learner legs only)
N=${1:-2}
mkdir -p gpurun_out
for f in 1 0 1 0; do
SS_PEER_STAGED=$f timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$f bench.py --gpus $N --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_n${N}_fused$f.json 2> gpurun_out/r2_bench_n${N}_fused$f.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n${N}_fused$f.json').read().strip().splitlines()[-1])
L=d["learner"]
print("SS_PEER_STAGED=$f N=$N update ms", L["train"]["ms_per_update"], "local", L["scaling_in_run"]["update_ms_without_exchange"], "cfg4 ms", L["selfplay_training"]["ms_per_iteration"], "peer_check", d.get("peer_check"))
PY
done | tee gpurun_out/r2_peer_staged_ab_n$N.txt
