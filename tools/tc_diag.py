"""Diagnostics of the tensor-core actor kernel on the GPU box: error against the restated
bf16 arithmetic for structured and random inputs, and launch times.  Run under `timeout`."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import learner_oracle as lo
from tests.test_gpu_learner_parity import actor_forward_bf16_model, _batch
from skillshot_learning_b200 import ActorCritic

ac = ActorCritic(device="cuda:0", seed=1)
rng = np.random.default_rng(0)
theta = lo.init_actor(rng)
theta[3072:3328] = rng.normal(0, 0.05, 256); theta[36096:36224] = rng.normal(0, 0.05, 128); theta[36480:] = 0.01
ac.set_weights(theta, None)
for n in (128, 1, 300, 20000):
    s, _, _ = _batch(n, n)
    torch.cuda.synchronize()
    got = ac.actor_forward(s, precision="bf16")
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    model = actor_forward_bf16_model(theta, s)
    exact = lo.actor_forward(theta, s)
    print("n=%d  max|tc-model|=%.3g  max|tc-f32|=%.3g  max|model-f32|=%.3g  nan=%d" % (
        n, np.abs(got - model).max(), np.abs(got - exact).max(), np.abs(model - exact).max(), int(np.isnan(got).sum())), flush=True)
    if np.abs(got - model).max() > 1e-2:
        bad = np.argwhere(np.abs(got - model) > 1e-2)
        print("   first bad rows:", bad[:10].tolist(), "got", got[bad[0][0]], "want", model[bad[0][0]], flush=True)

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

for n in (131072, 524288, 2097152):
    obs = torch.rand((n, 12), device="cuda")
    out = torch.empty((n, 2), device="cuda")
    t_tc = timeit(lambda: ac.actor_forward(obs, out=out, precision="bf16"))
    t_tcn = timeit(lambda: ac.actor_forward(obs, out=out, precision="bf16", param_noise_sd=0.5, noise_group=-(-(n // 128) // 148) * 128))
    t_f32 = timeit(lambda: ac.actor_forward(obs, out=out), iters=5)
    print("n=%d  tc %.1f us (%.3g samples/s, %.1f TFLOP/s)  tc+noise %.1f us  f32 %.1f us (%.3g samples/s)" % (
        n, t_tc, n / t_tc * 1e6, n * 72192 / t_tc * 1e6 / 1e12, t_tcn, t_f32, n / t_f32 * 1e6), flush=True)
