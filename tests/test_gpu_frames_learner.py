"""The learner of the frame-stacked "planning" networks (readme.md:18-20, BASELINE.json configs[4]; no reference code, parity
unpinned): actor and critic whose first Dense layer reads the player's last `frames` observations, updated exactly as
SkillshotLearner.py:386-443 updates the reference's 12-input networks.  The float32 kernels (ss_*_frames) against the
torch-CPU restatement of the Keras semantics (oracle/learner_oracle.py, gradients from autograd), tolerances as for the
12-input kernels in tests/test_gpu_learner_parity.py; bookkeeping (replay rows, history order, restarts) compared with ==."""
import numpy as np
import pytest
import torch

from oracle import learner_oracle as lo
from tests import philox_ref

pytestmark = pytest.mark.gpu


def _nets(frames, seed=0, **kw):
    from skillshot_learning_b200 import ActorCritic
    rng = np.random.default_rng(seed)
    theta, phi = lo.init_actor(rng, frames), lo.init_critic(rng, frames)
    ac = ActorCritic(device="cuda:0", seed=7, frames=frames, **kw)
    ac.set_weights(theta, phi)
    return ac, theta, phi, rng


@pytest.mark.parametrize("frames", [1, 4, 20])
def test_frames_forward_and_gradients_match_the_oracle(frames):
    ac, theta, phi, rng = _nets(frames)
    assert ac.a_n == len(theta) and ac.c_n == len(phi)
    n, w = 203, 12 * frames                                # not a multiple of the 32-row tile
    s = rng.uniform(0, 1, (n, w)).astype(np.float32)
    a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    y = (-rng.uniform(0, 1, n)).astype(np.float32)
    keep = (rng.uniform(size=(n, 256)) >= 0.2).astype(np.uint8)
    np.testing.assert_allclose(ac.actor_forward(s).cpu().numpy(), lo.frames_actor_forward(theta, s, frames), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(ac.critic_forward(s, a).cpu().numpy(), lo.critic_forward(phi, s, a), rtol=2e-5, atol=2e-6)
    g = ac.critic_grad(s, a, y, keep=keep).cpu().numpy()
    want, sse = lo.critic_grad(phi, s, a, y, keep.astype(np.float32))
    assert np.abs(g - want).max() <= 2e-5 * np.abs(want).max() + 1e-7
    np.testing.assert_allclose(float(ac.stats[0]), sse, rtol=1e-4)
    g = ac.actor_grad(s).cpu().numpy()
    want, q = lo.actor_grad(theta, phi, s)
    assert np.abs(g - want).max() <= 2e-5 * np.abs(want).max() + 1e-7
    np.testing.assert_allclose(float(ac.stats[1]), q, rtol=1e-4, atol=1e-4)
    # TD target through the target networks
    ac.gamma = 0.9
    done = (rng.uniform(size=n) < 0.3).astype(np.uint8)
    yy = ac.td_targets(y, s, torch.from_numpy(done)).cpu().numpy()
    np.testing.assert_allclose(yy, lo.ddpg_targets(theta, phi, y, s, done, 0.9), rtol=2e-5, atol=2e-6)


def test_frames_update_steps_track_the_oracle():
    """Six critic + actor steps (Philox dropout, Adam with eps outside the correction, soft target update) on 20-frame
    networks against the oracle stepping the same minibatches with the same masks."""
    frames = 20
    ac, theta, phi, rng = _nets(frames, gamma=0.0, tau=1.0, dropout=0.2)
    opt_a, opt_c = lo.AdamTF(len(theta)), lo.AdamTF(len(phi))
    n = 96
    for it in range(6):
        s = rng.uniform(0, 1, (n, 12 * frames)).astype(np.float32)
        a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        y = (-rng.uniform(0, 1, n)).astype(np.float32)
        keep = philox_ref.dropout_keep(n, ac.seed, ac.counter, 0.2)
        ac.critic_step(s, a, y)
        gq, _ = lo.critic_grad(phi, s, a, y, keep.astype(np.float32))
        phi = opt_c.step(phi, gq)
        ac.actor_step(s)
        ga, _ = lo.actor_grad(theta, phi, s)
        theta = opt_a.step(theta, ga)
    np.testing.assert_allclose(ac.critic.cpu().numpy(), phi, rtol=0, atol=2e-5)
    np.testing.assert_allclose(ac.actor.cpu().numpy(), theta, rtol=0, atol=2e-5)
    assert torch.equal(ac.target, ac.params)               # tau = 1


def test_frames_replay_ring_rows_and_wrap():
    from skillshot_learning_b200 import ReplayRing
    ring = ReplayRing(10, device="cuda:0", seed=3, frames=3)
    rng = np.random.default_rng(1)
    rows = [rng.uniform(size=(6, 36)).astype(np.float32) for _ in range(2)]
    nxt = [rng.uniform(size=(6, 36)).astype(np.float32) for _ in range(2)]
    for k in range(2):
        ring.push(rows[k], rng.uniform(size=(6, 2)).astype(np.float32), np.arange(6, dtype=np.float32) + 10 * k, nxt[k],
                  torch.tensor([0, 2, 0], dtype=torch.uint8), done_div=2)
    assert ring.size == 10 and ring.pos == 2
    got = ring.obs.cpu().numpy()
    np.testing.assert_array_equal(got[6:10], rows[1][:4])   # the second push wrapped
    np.testing.assert_array_equal(got[0:2], rows[1][4:])
    np.testing.assert_array_equal(got[2:6], rows[0][2:])
    assert ring.done.cpu().numpy().tolist() == [0, 0, 1, 1, 0, 0, 0, 0, 1, 1]   # any non-zero flag is stored as 1
    b = ring.sample(5, indices=[9, 0, 3, 3, 7])
    np.testing.assert_array_equal(b["obs"].cpu().numpy(), got[[9, 0, 3, 3, 7]])
    np.testing.assert_array_equal(b["next_obs"].cpu().numpy(), ring.next_obs.cpu().numpy()[[9, 0, 3, 3, 7]])


@pytest.mark.parametrize("precision", ["f32", "bf16"])
def test_frames_selfplay_trainer_rolls_out_and_learns(precision):
    """SelfPlayTrainer(frames=4): the acting network is the learner's actor vector; the stored stacked observations are
    consistent (a transition's next row is its row shifted by one frame unless the game restarted, in which case every
    frame is the restarted game's first observation); the update changes both networks and, on the minibatch it drew,
    the critic gradient equals the oracle's."""
    from skillshot_learning_b200 import SelfPlayTrainer
    F, n = 4, 256
    tr = SelfPlayTrainer(n, device="cuda:0", seed=11, frames=F, batch_size=512, replay_capacity=2 * n * 8, noise_group=128,
                         tick_limit=9, gamma=0.9, tau=0.05, precision=precision, param_noise_sd=0.5)
    assert tr.stack.params.data_ptr() == tr.networks.actor.data_ptr()
    dones = []
    for t in range(7):
        out = tr.rollout_tick()
        dones.append(out["done"].cpu().numpy().copy())
    assert tr.replay.size == 7 * 2 * n
    obs, nxt = tr.replay.obs[:tr.replay.size].cpu().numpy(), tr.replay.next_obs[:tr.replay.size].cpu().numpy()
    for t in range(7):
        sl = slice(t * 2 * n, (t + 1) * 2 * n)
        restarted = np.repeat(dones[t] != 0, 2)
        np.testing.assert_array_equal(nxt[sl][~restarted][:, :-12], obs[sl][~restarted][:, 12:])
        if restarted.any():
            r = nxt[sl][restarted].reshape(-1, F, 12)
            assert (r == r[:, -1:, :]).all()
        if t + 1 < 7:
            np.testing.assert_array_equal(nxt[sl], obs[(t + 1) * 2 * n:(t + 2) * 2 * n])
    # the newest frame of the stored rows is the observation the envs produced
    np.testing.assert_array_equal(nxt[6 * 2 * n:7 * 2 * n, -12:], tr.obs.reshape(-1, 12).cpu().numpy())
    before = tr.networks.params.clone()
    net = tr.networks
    theta, phi = net.target_actor.cpu().numpy().copy(), net.target_critic.cpu().numpy().copy()
    phi_online = net.critic.cpu().numpy().copy()
    counter = net.counter
    tr.update()
    b = tr._batch
    bs, ba, br, bn, bd = (b[k].cpu().numpy() for k in ("obs", "act", "reward", "next_obs", "done"))
    y = lo.ddpg_targets(theta, phi, br, bn, bd, 0.9)
    keep = philox_ref.dropout_keep(512, net.seed, counter, 0.2)
    gq, _ = lo.critic_grad(phi_online, bs, ba, y, keep.astype(np.float32))
    got = net.grads[net._c_off:net._c_off + net.c_n].cpu().numpy()
    assert np.abs(got - gq).max() <= 5e-5 * np.abs(gq).max() + 1e-7
    assert not torch.equal(net.params, before) and torch.isfinite(net.params).all()
    # the acting path sees the updated actor without a copy
    assert torch.equal(tr.stack.params, net.actor)
    tr.envs.check_status()
    # checkpoint round trip of the frames mode
    sd = tr.state_dict()
    tr2 = SelfPlayTrainer(n, device="cuda:0", seed=11, frames=F, batch_size=512, replay_capacity=2 * n * 8, noise_group=128,
                          tick_limit=9, gamma=0.9, tau=0.05, precision=precision, param_noise_sd=0.5)
    tr2.load_state_dict(sd)
    for _ in range(2):
        tr.rollout_tick(); tr2.rollout_tick()
    assert torch.equal(tr.actions, tr2.actions) and torch.equal(tr.rows, tr2.rows)
