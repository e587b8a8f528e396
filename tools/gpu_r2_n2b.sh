#!/bin/bash
# round 2, N GPUs: A/B of one switch of the peer-exchange path on the same box, alternating (learner legs only).
#   bash tools/gpu_r2_n2b.sh 2 SS_PEER_FUSED     one-kernel exchange (1) against push + Adam (0)
#   bash tools/gpu_r2_n2b.sh 2 SS_PEER_STAGED    early actor forward between push and Adam (1) against the paired forward after Adam (0)
N=${1:-2}
VAR=${2:-SS_PEER_FUSED}
mkdir -p gpurun_out
for f in 1 0 1 0; do
env $VAR=$f timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$f bench.py --gpus $N --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_n${N}_ab$f.json 2> gpurun_out/r2_bench_n${N}_ab$f.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n${N}_ab$f.json').read().strip().splitlines()[-1])
L=d["learner"]
print("$VAR=$f N=$N update ms", L["train"]["ms_per_update"], "local", L["scaling_in_run"]["update_ms_without_exchange"], "cfg4 ms", L["selfplay_training"]["ms_per_iteration"], "peer_check", d.get("peer_check"))
PY
done | tee gpurun_out/r2_peer_ab_${VAR}_n$N.txt
