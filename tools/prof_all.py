"""Profiling driver: one rollout tick (262,144 envs) and one 65,536-row DDPG update inside a cudaProfilerStart/Stop window
(run under `ncu --profile-from-start off`); every kernel of the learner path appears once."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from skillshot_learning_b200 import SelfPlayTrainer
E = 262144
tr = SelfPlayTrainer(E, device="cuda:0", seed=0, replay_capacity=2 * E * 4, batch_size=65536, gamma=0.99, tau=0.005, precision="bf16",
                     noise_group=4096, tick_limit=200)
tr.rollout(4)
for _ in range(2):
    tr.update()
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.rollout(1)
tr.update()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(tr.networks.stats[0]))
