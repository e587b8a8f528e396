"""Development: launch times of the frame-stacked actor (20 frames), float32 kernels vs tensor cores."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from skillshot_learning_b200 import FrameStackActor

def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

for n in (131072, 524288):
    fl = 2 * (240 * 256 + 256 * 128 + 128 * 2)
    for prec in ("bf16", "f32"):
        fa = FrameStackActor(n, frames=20, device="cuda:0", seed=1, precision=prec)
        g = torch.Generator(device="cuda").manual_seed(1)
        fa.push(torch.rand((n, 12), device="cuda", generator=g))
        out = torch.empty((n, 2), device="cuda")
        t = timeit(lambda: fa.forward(out=out), iters=20 if prec == "bf16" else 3)
        tn = timeit(lambda: fa.forward(param_noise_sd=0.5, noise_group=1024, out=out), iters=20 if prec == "bf16" else 3)
        s = torch.rand((n, 12), device="cuda")
        tp = timeit(lambda: fa.push(s))
        print("n=%d %s: forward %.1f us (%.1f TFLOP/s)  with noise groups of 1024: %.1f us   push %.1f us" % (
            n, prec, t, n * fl / t * 1e6 / 1e12, tn, tp), flush=True)
