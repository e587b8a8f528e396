"""Diagnostics of the tensor-core gradient kernels on the GPU box (run under `timeout`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import learner_oracle as lo
from tests.test_gpu_learner_parity import _batch
from skillshot_learning_b200 import ActorCritic

rng = np.random.default_rng(5)
theta, phi = lo.init_actor(rng), lo.init_critic(rng)
theta[3072:3328] = rng.normal(0, 0.05, 256); theta[36096:36224] = rng.normal(0, 0.05, 128); theta[36480:] = 0.02
phi[3072:3328] = rng.normal(0, 0.05, 256); phi[36352:36480] = rng.normal(0, 0.05, 128); phi[36608] = 0.1
ac = ActorCritic(device="cuda:0", seed=11, update_precision="bf16")
ac.set_weights(theta, phi)
names_c = [("W1", 0, 3072), ("b1", 3072, 3328), ("W2", 3328, 36096), ("W2act", 36096, 36352), ("b2", 36352, 36480), ("W3", 36480, 36608), ("b3", 36608, 36609)]
names_a = [("W1", 0, 3072), ("b1", 3072, 3328), ("W2", 3328, 36096), ("b2", 36096, 36224), ("W3", 36224, 36480), ("b3", 36480, 36482)]

def report(tag, g, want, names):
    for nm, a, b in names:
        sc = np.abs(want[a:b]).max() + 1e-30
        print("  %-6s %-6s max|d|/scale %.3g   scale %.3g   got-scale %.3g" % (tag, nm, np.abs(g[a:b] - want[a:b]).max() / sc, sc, np.abs(g[a:b]).max()), flush=True)

for n in (128, 300, 5000):
    s, a, r = _batch(n, n)
    q, up = ac.critic_forward(s, a, precision="bf16", want_dq_da=True)
    torch.cuda.synchronize()
    wq = lo.critic_forward(phi, s, a)
    print("n=%d critic fwd tc: max|q-q32| %.3g" % (n, np.abs(q.cpu().numpy() - wq).max()), flush=True)
    keep = (np.random.default_rng(n).uniform(size=(n, 256)) >= 0.2).astype(np.uint8)
    g = ac.critic_grad(s, a, r, keep=keep).cpu().numpy()
    want, wsse = lo.critic_grad(phi.astype(np.float64), s, a, r, keep.astype(np.float64), 0.2, dtype=torch.float64)
    print("n=%d critic grad tc: sse %.6g want %.6g" % (n, float(ac.stats[0]), wsse), flush=True)
    report("critic", g, want, names_c)
    g = ac.actor_grad(s).cpu().numpy()
    want, wq = lo.actor_grad(theta.astype(np.float64), phi.astype(np.float64), s, dtype=torch.float64)
    print("n=%d actor grad tc: qsum %.6g want %.6g" % (n, float(ac.stats[1]), wq), flush=True)
    report("actor", g, want, names_a)

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

ac32 = ActorCritic(device="cuda:0", seed=11)
ac32.set_weights(theta, phi)
for n in (65536, 524288):
    s = torch.rand((n, 12), device="cuda"); a = torch.rand((n, 2), device="cuda") * 2 - 1; r = -torch.rand(n, device="cuda")
    for nm, net in (("bf16", ac), ("f32", ac32)):
        if nm == "f32" and n > 65536:
            continue
        tc = timeit(lambda: net.critic_grad(s, a, r), 10)
        ta = timeit(lambda: net.actor_grad(s), 10)
        net.gamma = 0.99
        tt = timeit(lambda: net.td_targets(r, s), 10)
        print("n=%d %s: critic_grad %.1f us (%.3g rows/s)  actor_grad %.1f us (%.3g rows/s)  targets %.1f us" % (
            n, nm, tc, n / tc * 1e6, ta, n / ta * 1e6, tt), flush=True)
