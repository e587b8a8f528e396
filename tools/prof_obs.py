"""Short run of the observation-writing step kernel for ncu (E=1M, one tick per launch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from skillshot_learning_b200 import SkillshotEnvs
E = int(os.environ.get("SS_E", "1048576"))
envs = SkillshotEnvs(E, device="cuda:0", random_positions=True, seed=1, reward_mode="looking", tick_limit=2000, auto_reset=True)
acts = torch.rand((8, E, 2, 2), device="cuda:0") * 2.4 - 1.2
for j in range(8):
    envs.step(acts[j], want_obs=True)
torch.cuda.synchronize()
print("ok")
