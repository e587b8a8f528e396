"""Run under torchrun with >= 2 GPUs (tests/test_gpu_multi.py launches it; `gpurun --gpus 2`):
the fused NVLink peer exchange (ss_peer_reduce_push + ss_peer_adam_tf) against the NCCL all-reduce
path and against a single-rank update of the whole batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import learner_oracle as lo  # noqa: E402
from skillshot_learning_b200 import ActorCritic  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(0)                       # identical data on every rank
    theta, phi = lo.init_actor(rng), lo.init_critic(rng)
    for precision, tol in (("f32", 2e-6), ("bf16", 2e-6)):
        nets = {c: ActorCritic(device=dev, seed=5, gamma=0.9, tau=0.1, process_group=True, update_precision=precision,
                               collective=c) for c in ("nccl", "peer")}
        solo = ActorCritic(device=dev, seed=5, gamma=0.9, tau=0.1, update_precision=precision)
        for ac in list(nets.values()) + [solo]:
            ac.set_weights(theta, phi)
        n_local = 2048
        n = n_local * world
        for it in range(6):
            s = rng.uniform(0, 1, (n, 12)).astype(np.float32)
            a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
            r = (-rng.uniform(0, 1, n)).astype(np.float32)
            keep = (rng.uniform(size=(n, 256)) >= 0.2).astype(np.uint8)
            sl = slice(rank * n_local, (rank + 1) * n_local)
            for ac in nets.values():
                ac.critic_step(s[sl], a[sl], r[sl], keep=keep[sl])
                ac.actor_step(s[sl])
            solo.critic_step(s, a, r, keep=keep)
            solo.actor_step(s)
        nets["peer"].peer.check_status()
        p_peer, p_nccl, p_solo = nets["peer"].params, nets["nccl"].params, solo.params
        moved = float((p_solo - torch.cat([torch.from_numpy(theta), torch.zeros(2), torch.from_numpy(phi)]).to(dev)).abs().max())
        d_nccl = float((p_peer - p_nccl).abs().max())
        d_solo = float((p_peer - p_solo).abs().max())
        # every rank must hold bit-identical parameters after the peer exchange (same summation order everywhere)
        gathered = [torch.empty_like(p_peer) for _ in range(world)]
        dist.all_gather(gathered, p_peer)
        identical = all(torch.equal(gathered[0], g) for g in gathered)
        tgt_ok = torch.allclose(nets["peer"].target, nets["nccl"].target, atol=tol)
        if rank == 0:
            print("%s: moved %.3g  |peer-nccl| %.3g  |peer-solo| %.3g  ranks identical %s  targets %s" % (
                precision, moved, d_nccl, d_solo, identical, tgt_ok), flush=True)
        assert identical, "ranks diverged"
        # Adam turns a last-bit difference of a near-zero gradient into a visible step, so the bound is on the
        # worst parameter relative to the distance moved; the summation orders differ (all-reduce tree vs rank order)
        assert d_nccl <= 0.02 * moved + tol and d_solo <= 0.02 * moved + tol, (d_nccl, d_solo, moved)
        assert tgt_ok

        # timing of the exchange step itself: critic step with a small batch (gradient kernel cost is the same in both)
        s, a, r = (torch.rand((512, 12), device=dev), torch.rand((512, 2), device=dev), torch.rand(512, device=dev))
        res = {}
        for c, ac in nets.items():
            for _ in range(5):
                ac.critic_step(s, a, r)
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                ac.critic_step(s, a, r)
            e1.record(); torch.cuda.synchronize()
            res[c] = e0.elapsed_time(e1) / 50 * 1e3
        if rank == 0:
            print("%s: 512-row critic step  nccl %.1f us   peer %.1f us" % (precision, res["nccl"], res["peer"]), flush=True)
        nets["peer"].peer.close()

    # the whole sharded update from one library call (ss_ddpg_update with the peer exchange inside) against the same
    # update made of separate calls: two trainers per rank with the same seeds must stay bit-identical
    from skillshot_learning_b200 import SelfPlayTrainer
    for precision in ("f32", "bf16"):
        a, b = [SelfPlayTrainer(1024, device=dev, seed=40 + rank, batch_size=3000, noise_group=128, tick_limit=30, gamma=0.95,
                                tau=0.05, precision=precision, process_group=True, collective="peer") for _ in range(2)]
        for tr in (a, b):
            tr.rollout(5)
        for it in range(4):
            sa, qa = a.update()
            sb, qb = b.update_stepwise()
            torch.cuda.synchronize()
            for name in ("params", "target", "adam_m", "adam_v", "grads", "stats"):
                assert torch.equal(getattr(a.networks, name), getattr(b.networks, name)), (precision, it, name)
        for tr in (a, b):
            tr.networks.peer.check_status()
        gathered = [torch.empty_like(a.networks.params) for _ in range(world)]
        dist.all_gather(gathered, a.networks.params)
        assert all(torch.equal(gathered[0], g) for g in gathered), "ranks diverged"
        if rank == 0:
            print("%s: single-call sharded update == stepwise, ranks identical" % precision, flush=True)
        for tr in (a, b):
            tr.networks.peer.close()
    if rank == 0:
        print("PEER OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
