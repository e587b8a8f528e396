"""Development: fixed (size-independent) cost of the tensor-core launches -- forward and gradient
kernels timed at 1, 2, 4, 8 tiles per CTA; the intercept of the line is prologue + weight staging + tail."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from skillshot_learning_b200 import ActorCritic

ac = ActorCritic(device="cuda:0", seed=1, update_precision="bf16")

def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

empty = torch.empty(1, device="cuda")
print("torch fill_ (tiny kernel) back-to-back: %.2f us" % timeit(lambda: empty.fill_(1.0)))
for k in (1, 2, 4, 8, 16):
    n = 148 * 128 * k
    obs = torch.rand((n, 12), device="cuda"); act = torch.rand((n, 2), device="cuda") * 2 - 1; y = torch.rand(n, device="cuda")
    out = torch.empty((n, 2), device="cuda")
    t_f = timeit(lambda: ac.actor_forward(obs, out=out, precision="bf16"))
    t_c = timeit(lambda: ac.critic_forward(obs, act, precision="bf16"))
    t_g = timeit(lambda: ac.critic_grad(obs, act, y, slices_only=True))
    t_a = timeit(lambda: ac.actor_grad(obs, slices_only=True))
    print("tiles/CTA %2d  n=%7d  actor fwd %.1f us  critic fwd %.1f us  critic grad (slices) %.1f us  actor grad (slices) %.1f us" % (k, n, t_f, t_c, t_g, t_a), flush=True)
