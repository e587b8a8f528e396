"""Profiling driver: the tensor-core frame-stacked actor forward, 131,072 rows x 20 frames."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from skillshot_learning_b200 import FrameStackActor
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
fa = FrameStackActor(n, frames=20, device="cuda:0", seed=1, precision="bf16")
g = torch.Generator(device="cuda").manual_seed(1)
fa.push(torch.rand((n, 12), device="cuda", generator=g))
out = torch.empty((n, 2), device="cuda")
for _ in range(4):
    fa.forward(out=out)
for _ in range(2):
    fa.forward(param_noise_sd=0.5, noise_group=1024, out=out)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
