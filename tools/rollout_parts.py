"""Development: the pieces of one rollout tick at 262,144 envs, timed alone."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from skillshot_learning_b200 import SelfPlayTrainer, _lib
from skillshot_learning_b200._lib import lib, check
E = 262144
NG = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
tr = SelfPlayTrainer(E, device="cuda:0", seed=0, replay_capacity=2 * E * 4, batch_size=65536, gamma=0.99, tau=0.005, precision="bf16",
                     noise_group=NG, tick_limit=200)
net, envs = tr.networks, tr.envs
def timeit(fn, iters=30):
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
obs, act = tr.obs.view(-1, 12), tr.actions.view(-1, 2)
n = obs.shape[0]
groups = -(-n // NG)
stride = 36484
scratch = torch.empty(groups * stride, device="cuda")
st = torch.cuda.current_stream().cuda_stream
print("noise groups:", groups)
print("forward, no noise           %.1f us" % timeit(lambda: net.actor_forward(obs, out=act, precision="bf16")))
print("forward, in-kernel noise    %.1f us" % timeit(lambda: net.actor_forward(obs, out=act, precision="bf16", param_noise_sd=0.5, noise_group=NG)))
print("noise vectors kernel        %.1f us" % timeit(lambda: check(lib.ss_param_noise_groups(net.actor.data_ptr(), scratch.data_ptr(), 36482, groups, stride, 0.5, 1, 2, st), "n")))
print("env step + observations     %.1f us" % timeit(lambda: envs.step(tr.actions, obs_out=tr.obs)))
print("rollout tick (16 per call)  %.1f us" % (timeit(lambda: tr.rollout(16), 10) / 16))
# Measured once (then removed): drawing tick t + 1's perturbed vectors on a second stream (ss_param_noise_groups -> a forward
# that reads ready-made vectors: 47.4 us instead of 56.0) gave 81.7 us per tick when the draw ran beside the env step and
# 87.2 us when it ran beside the forward kernel, against 84.6 us with the in-kernel noise: the 10.7 us draw is not hidden,
# the kernels do not share SMs usefully, and the cross-stream events cost what the shorter forward saves.
