"""Development: host time of one SelfPlayTrainer.update() (tiny batch: the device is never the limit)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from skillshot_learning_b200 import SelfPlayTrainer

for batch in (1024, 65536):
    tr = SelfPlayTrainer(8192, device="cuda:0", seed=0, replay_capacity=8192 * 2 * 16, batch_size=batch, gamma=0.99, tau=0.005,
                         precision="bf16", noise_group=1024)
    tr.rollout(16)
    for _ in range(5):
        tr.update()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        tr.update()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("batch %6d: host %.1f us per update() (enqueue only), %.1f us with the final sync" % (batch, (t1 - t0) / 200 * 1e6, (t2 - t0) / 200 * 1e6))
