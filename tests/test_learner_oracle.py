"""CPU checks of the learner oracle (oracle/learner_oracle.py) itself.

The reference's learner numerics live in TensorFlow/Keras, absent here and with no
golden vectors in the reference: parity is UNPINNED (see the oracle's header).  These
tests pin the restatement to the published definitions it cites instead: layer shapes
and parameter counts printed by the reference's model.summary(), central-difference
gradients in float64, a hand-computed tf.keras Adam step, and the gamma = 0 / tau = 1
reduction to the reference's update."""
import numpy as np
import pytest
import torch

from oracle import learner_oracle as lo


def _nets(seed=0):
    rng = np.random.default_rng(seed)
    return lo.init_actor(rng), lo.init_critic(rng)


def test_parameter_counts_match_the_reference_summary():
    # SURVEY.md 2 (model.summary() of SkillshotLearner.py:70-121): actor 36,482; critic 36,609
    assert lo.ACTOR_PARAMS == 36482 and lo.CRITIC_PARAMS == 36609
    theta, phi = _nets()
    assert theta.shape == (36482,) and phi.shape == (36609,)
    w = lo.split(theta, lo.ACTOR_SHAPES)
    assert abs(float(w[0].std()) - 0.05) < 0.003 and not w[1].any() and not w[3].any() and not w[5].any()
    c = lo.split(phi, lo.CRITIC_SHAPES)
    lim = np.sqrt(6.0 / (258 + 128))
    assert abs(c[2]).max() <= lim and abs(c[2]).max() > 0.95 * lim          # glorot-uniform
    assert abs(float(c[4].std()) - 0.05) < 0.02


def test_forward_shapes_ranges_and_dropout_scaling():
    theta, phi = _nets(1)
    rng = np.random.default_rng(2)
    s = rng.uniform(0, 1, (7, 12)).astype(np.float32)
    a = lo.actor_forward(theta, s)
    assert a.shape == (7, 2) and np.all(np.abs(a) < 1)
    q0 = lo.critic_forward(phi, s, a)
    q1 = lo.critic_forward(phi, s, a, keep=np.ones((7, 256), np.float32), rate=0.0)
    np.testing.assert_allclose(q0, q1, rtol=1e-6)
    # all units kept at rate 0.2 scales the hidden layer by 1/0.8: not the same output
    q2 = lo.critic_forward(phi, s, a, keep=np.ones((7, 256), np.float32), rate=0.2)
    assert not np.allclose(q0, q2)


def _numeric_grad(f, x, idx, h=1e-6):
    g = np.zeros(len(idx))
    for n, i in enumerate(idx):
        xp, xm = x.copy(), x.copy()
        xp[i] += h
        xm[i] -= h
        g[n] = (f(xp) - f(xm)) / (2 * h)
    return g


def test_critic_gradient_against_central_differences():
    theta, phi = _nets(3)
    rng = np.random.default_rng(4)
    s = rng.uniform(0, 1, (5, 12))
    a = rng.uniform(-1, 1, (5, 2))
    y = rng.normal(size=5)
    keep = (rng.uniform(size=(5, 256)) > 0.2).astype(np.float64)
    g, sse = lo.critic_grad(phi.astype(np.float64), s, a, y, keep, 0.2, dtype=torch.float64)
    idx = rng.choice(lo.CRITIC_PARAMS, 40, replace=False)

    def loss(p):
        q = lo.critic_forward(p, s, a, keep, 0.2, dtype=torch.float64)
        return float(np.mean((q - y) ** 2))
    np.testing.assert_allclose(g[idx], _numeric_grad(loss, phi.astype(np.float64), idx), rtol=1e-4, atol=1e-8)
    assert abs(sse / 5 - loss(phi.astype(np.float64))) < 1e-12


def test_actor_gradient_is_minus_sum_q():
    theta, phi = _nets(5)
    rng = np.random.default_rng(6)
    s = rng.uniform(0, 1, (6, 12))
    g, qsum = lo.actor_grad(theta.astype(np.float64), phi.astype(np.float64), s, dtype=torch.float64)
    idx = rng.choice(lo.ACTOR_PARAMS, 40, replace=False)

    def neg_q(t):
        a = lo.actor_forward(t, s, dtype=torch.float64)
        return -float(lo.critic_forward(phi.astype(np.float64), s, a, dtype=torch.float64).sum())
    np.testing.assert_allclose(g[idx], _numeric_grad(neg_q, theta.astype(np.float64), idx), rtol=1e-4, atol=1e-9)
    assert abs(qsum + neg_q(theta.astype(np.float64))) < 1e-10


def test_adam_is_the_tf_keras_formula():
    # one parameter, two steps by hand: lr_t = lr sqrt(1-b2^t)/(1-b1^t); p -= lr_t m / (sqrt(v) + 1e-7)
    opt = lo.AdamTF(1, dtype=np.float64)
    p = np.array([1.0])
    p = opt.step(p, np.array([0.5]))
    m, v = 0.1 * 0.5, 0.001 * 0.25
    want = 1.0 - 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9) * m / (np.sqrt(v) + 1e-7)
    assert abs(p[0] - want) < 1e-15
    p = opt.step(p, np.array([-0.25]))
    m, v = 0.9 * m + 0.1 * -0.25, 0.999 * v + 0.001 * 0.0625
    want = want - 1e-3 * np.sqrt(1 - 0.999 ** 2) / (1 - 0.9 ** 2) * m / (np.sqrt(v) + 1e-7)
    assert abs(p[0] - want) < 1e-15
    # epsilon sits outside the bias correction: differs from torch.optim.Adam's placement
    t = torch.nn.Parameter(torch.tensor([1.0], dtype=torch.float64))
    o = torch.optim.Adam([t], lr=1e-3, eps=1e-7)
    for gr in (0.5, -0.25):
        t.grad = torch.tensor([gr], dtype=torch.float64)
        o.step()
    d = abs(float(t.detach()) - p[0])
    assert 1e-12 < d < 1e-6


def test_gamma_zero_tau_one_is_the_reference_update():
    theta, phi = _nets(7)
    rng = np.random.default_rng(8)
    r = rng.normal(size=9).astype(np.float32)
    s2 = rng.uniform(0, 1, (9, 12)).astype(np.float32)
    done = (rng.uniform(size=9) < 0.3)
    np.testing.assert_array_equal(lo.ddpg_targets(theta, phi, r, s2, done, 0.0), r)     # critic target = reward (:434)
    y = lo.ddpg_targets(theta, phi, r, s2, done, 0.9)
    assert np.allclose(y[done], r[done]) and not np.allclose(y[~done], r[~done])
    np.testing.assert_array_equal(lo.soft_update(theta * 0, theta, 1.0), theta)          # tau = 1: no target lag


def test_sharded_gradient_sums_to_the_full_batch_gradient():
    theta, phi = _nets(9)
    rng = np.random.default_rng(10)
    s = rng.uniform(0, 1, (32, 12)).astype(np.float32)
    a = rng.uniform(-1, 1, (32, 2)).astype(np.float32)
    y = rng.normal(size=32).astype(np.float32)
    keep = (rng.uniform(size=(32, 256)) > 0.2).astype(np.float32)
    full, _ = lo.critic_grad(phi, s, a, y, keep)
    parts = [lo.critic_grad(phi, s[i:i + 16], a[i:i + 16], y[i:i + 16], keep[i:i + 16], n_global=32)[0] for i in (0, 16)]
    np.testing.assert_allclose(parts[0] + parts[1], full, rtol=2e-5, atol=1e-8)
    fa, _ = lo.actor_grad(theta, phi, s)
    pa = [lo.actor_grad(theta, phi, s[i:i + 16])[0] for i in (0, 16)]
    np.testing.assert_allclose(pa[0] + pa[1], fa, rtol=2e-5, atol=1e-8)


@pytest.mark.parametrize("frames", [1, 3])
def test_two_independent_restatements_agree(frames):
    """oracle/learner_oracle.py (torch autograd) against oracle/learner_oracle_np.py (numpy, hand-derived backward pass,
    written separately): gradients of the critic's fit loss and of the actor step, and three Adam steps, in float64."""
    import torch
    from oracle import learner_oracle_np as ln
    rng = np.random.default_rng(5)
    theta, phi = lo.init_actor(rng, frames).astype(np.float64), lo.init_critic(rng, frames).astype(np.float64)
    n = 37
    s = rng.uniform(0, 1, (n, 12 * frames)); a = rng.uniform(-1, 1, (n, 2)); y = -rng.uniform(0, 1, n)
    keep = (rng.uniform(size=(n, 256)) >= 0.2).astype(np.float64)
    for k in (None, keep):
        g1, sse1 = lo.critic_grad(phi, s, a, y, k, 0.2, n_global=50, dtype=torch.float64)
        g2, sse2 = ln.critic_grad(phi, s, a, y, k, 0.2, n_global=50)
        np.testing.assert_allclose(g2, g1, rtol=1e-9, atol=1e-13)
        assert abs(sse1 - sse2) <= 1e-10 * abs(sse1)
    g1, q1 = lo.actor_grad(theta, phi, s, dtype=torch.float64)
    g2, q2 = ln.actor_grad(theta, phi, s)
    np.testing.assert_allclose(g2, g1, rtol=1e-9, atol=1e-13)
    assert abs(q1 - q2) <= 1e-10 * abs(q1)
    opt = lo.AdamTF(len(phi), dtype=np.float64)
    p1, p2, m, v = phi.copy(), phi.copy(), np.zeros_like(phi), np.zeros_like(phi)
    for t in range(1, 4):
        g = rng.normal(size=len(phi)) * 1e-3
        p1 = opt.step(p1, g)
        p2, m, v = ln.adam_step(p2, g, m, v, t)
    np.testing.assert_allclose(p2, p1, rtol=1e-12, atol=1e-15)
