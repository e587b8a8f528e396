"""World-size-2 test (gloo, CPU) of the sharded update protocol the GPU path uses
(skillshot_learning_b200/learner.py: shard_info + allreduce_sum around the gradient kernels):
each rank takes the gradient of its half of the batch with the GLOBAL mean divisor, one
all-reduce sums the flat gradient, every rank applies the same Adam step.  The kernels are
stood in for by the oracle here (no GPU in this container); the protocol is the product's."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import learner_oracle as lo


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from skillshot_learning_b200.learner import allreduce_sum, shard_info
        rng = np.random.default_rng(0)                       # same data on every rank, each takes its slice
        theta, phi = lo.init_actor(rng), lo.init_critic(rng)
        n = 48
        s = rng.uniform(0, 1, (n, 12)).astype(np.float32)
        a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        y = rng.normal(size=n).astype(np.float32)
        keep = (rng.uniform(size=(n, 256)) > 0.2).astype(np.float32)
        n_local = n // world
        w, n_global, off = shard_info(n_local, True)
        assert (w, n_global, off) == (world, n, rank * n_local)
        sl = slice(off, off + n_local)
        g, _ = lo.critic_grad(phi, s[sl], a[sl], y[sl], keep[sl], n_global=n_global)
        g = allreduce_sum(torch.from_numpy(g), True).numpy()
        full, _ = lo.critic_grad(phi, s, a, y, keep)
        np.testing.assert_allclose(g, full, rtol=2e-5, atol=1e-8)
        ga, _ = lo.actor_grad(theta, phi, s[sl])             # the actor gradient is a batch SUM: no divisor
        ga = allreduce_sum(torch.from_numpy(ga), True).numpy()
        fa, _ = lo.actor_grad(theta, phi, s)
        np.testing.assert_allclose(ga, fa, rtol=2e-5, atol=1e-8)
        new_phi = lo.AdamTF(lo.CRITIC_PARAMS).step(phi, g)   # identical on every rank: no weight broadcast needed
        np.save(os.path.join(out_dir, "phi_%d.npy" % rank), new_phi)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_update_equals_single_rank(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    p0, p1 = (np.load(tmp_path / ("phi_%d.npy" % r)) for r in (0, 1))
    assert np.array_equal(p0, p1)


def test_shard_info_without_a_group():
    from skillshot_learning_b200.learner import shard_info
    assert shard_info(17, None) == (1, 17, 0)
