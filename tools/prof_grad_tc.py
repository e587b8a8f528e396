"""One workload for ncu: the tensor-core critic gradient on 65,536 rows (the bench's update batch)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from skillshot_learning_b200 import ActorCritic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
ac = ActorCritic(device="cuda:0", seed=1, update_precision="bf16")
s = torch.rand((n, 12), device="cuda"); a = torch.rand((n, 2), device="cuda") * 2 - 1; r = -torch.rand(n, device="cuda")
for _ in range(4):
    ac.critic_grad(s, a, r)
    ac.actor_grad(s)
torch.cuda.synchronize()
print("done", float(ac.grads.abs().mean()))
