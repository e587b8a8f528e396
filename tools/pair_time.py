"""Development: a = actor(s), critic(s, a) as two launches vs one pair launch (ss_actor_critic_forward_tc), graph-replayed."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from skillshot_learning_b200 import ActorCritic
from skillshot_learning_b200._lib import lib, check
ac = ActorCritic(device="cuda:0", seed=1, update_precision="bf16")
for n in (65536, 524288):
    s = torch.rand((n, 12), device="cuda"); a = torch.empty((n, 2), device="cuda"); q = torch.empty(n, device="cuda")
    up = torch.empty((n, 2), device="cuda"); flags = torch.full((n, 2), 0x7fc0dead, dtype=torch.int32, device="cuda")
    def two(st):
        check(lib.ss_actor_forward_tc(ac.actor.data_ptr(), s.data_ptr(), a.data_ptr(), n, 0.0, 0, 0.0, 0, 0, st), "a")
        check(lib.ss_critic_forward_tc(ac.critic.data_ptr(), s.data_ptr(), a.data_ptr(), n, q.data_ptr(), up.data_ptr(), None, None, 0.0, None, st), "c")
    def pair(st):
        check(lib.ss_actor_critic_forward_tc(ac.actor.data_ptr(), ac.critic.data_ptr(), s.data_ptr(), a.data_ptr(), n, q.data_ptr(), up.data_ptr(),
                                             None, None, 0.0, None, flags.data_ptr(), st), "p")
    for name, fn in (("two launches", two), ("pair launch", pair)):
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            fn(side.cuda_stream); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for _ in range(20):
                    fn(torch.cuda.current_stream().cuda_stream)
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                g.replay()
            e1.record(); torch.cuda.synchronize()
        print("n %7d  %-13s %.1f us" % (n, name, e0.elapsed_time(e1) / 200 * 1e3), flush=True)
