"""One workload for ncu: the tensor-core actor forward on 524,288 rows (BASELINE.json config 3's
2 x 262,144 observations per tick), a few launches.  Optional argv[1] = rows."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from skillshot_learning_b200 import ActorCritic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
noise = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
ac = ActorCritic(device="cuda:0", seed=1)
obs = torch.rand((n, 12), device="cuda")
out = torch.empty((n, 2), device="cuda")
for _ in range(6):
    ac.actor_forward(obs, out=out, precision="bf16", param_noise_sd=noise, noise_group=max(128, n // 148 // 128 * 128))
torch.cuda.synchronize()
print("done", float(out.abs().mean()))
