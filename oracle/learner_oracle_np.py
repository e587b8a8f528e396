"""A SECOND, independent restatement of the learner's arithmetic: plain numpy float64 with a
hand-derived backward pass -- no autograd, no torch -- written from the Keras semantics at the reference's call sites, not
from oracle/learner_oracle.py and not from the kernels.  tests/test_learner_oracle.py requires the two restatements to
agree to float64 rounding; the CUDA kernels are then checked against the first (tests/test_gpu_learner_parity.py).

TEST INFRASTRUCTURE ONLY.  Parity stays unpinned (no TensorFlow here, no stored Keras outputs in the reference); two
restatements that agree only show that neither has a slip in its calculus.

  actor   SkillshotLearner.py:70-96    a = tanh(relu(relu(s W1 + b1) W2 + b2) W3 + b3)
  critic  SkillshotLearner.py:98-121   q = relu([relu(s W1 + b1) * keep / (1 - rate), a] W2 + b2) W3 + b3
  fit     SkillshotLearner.py:118,434  loss = mean_batch (q - y)^2
  actor step  SkillshotLearner.py:395-410   g = d/d theta of -sum_batch q(s, actor(s))     (critic Dropout off)
"""
from __future__ import annotations

import numpy as np

H1, H2, DA = 256, 128, 2


def _split(flat, ds, critic):
    flat = np.asarray(flat, np.float64)
    shapes = [(ds, H1), (H1,), (H1 + (DA if critic else 0), H2), (H2,), (H2, 1 if critic else DA), (1 if critic else DA,)]
    out, o = [], 0
    for sh in shapes:
        n = int(np.prod(sh))
        out.append(flat[o:o + n].reshape(sh))
        o += n
    assert o == len(flat)
    return out


def _flat(parts):
    return np.concatenate([p.ravel() for p in parts])


def critic_grad(phi, s, a, y, keep=None, rate=0.2, n_global=None):
    """(d loss / d phi as a flat vector, sum of squared errors) of one fit batch."""
    s, a, y = (np.asarray(x, np.float64) for x in (s, a, y))
    W1, b1, W2, b2, W3, b3 = _split(phi, s.shape[1], True)
    z1 = s @ W1 + b1
    h1 = np.maximum(z1, 0.0)
    scale = 1.0 if keep is None else np.asarray(keep, np.float64) / (1.0 - rate)
    x2 = np.concatenate([h1 * scale, a], axis=1)
    z2 = x2 @ W2 + b2
    h2 = np.maximum(z2, 0.0)
    q = (h2 @ W3 + b3)[:, 0]
    err = q - y
    dq = (2.0 * err / float(n_global or len(y)))[:, None]          # d mean((q - y)^2) / d q
    gW3, gb3 = h2.T @ dq, dq.sum(0)
    dz2 = (dq @ W3.T) * (z2 > 0)
    gW2, gb2 = x2.T @ dz2, dz2.sum(0)
    dh1 = (dz2 @ W2.T)[:, :H1] * scale
    dz1 = dh1 * (z1 > 0)
    gW1, gb1 = s.T @ dz1, dz1.sum(0)
    return _flat([gW1, gb1, gW2, gb2, gW3, gb3]), float((err ** 2).sum())


def actor_grad(theta, phi, s):
    """(gradient of -sum_batch q(s, actor(s)) with respect to theta, sum q)."""
    s = np.asarray(s, np.float64)
    A1, c1, A2, c2, A3, c3 = _split(theta, s.shape[1], False)
    W1, b1, W2, b2, W3, b3 = _split(phi, s.shape[1], True)
    u1 = s @ A1 + c1; g1 = np.maximum(u1, 0.0)
    u2 = g1 @ A2 + c2; g2 = np.maximum(u2, 0.0)
    act = np.tanh(g2 @ A3 + c3)
    h1 = np.maximum(s @ W1 + b1, 0.0)                               # Dropout off: the model is called directly
    z2 = np.concatenate([h1, act], axis=1) @ W2 + b2
    q = (np.maximum(z2, 0.0) @ W3 + b3)[:, 0]
    dq_da = ((z2 > 0) * W3[:, 0]) @ W2[H1:, :].T                    # only the two action rows of the second kernel matter
    du3 = -dq_da * (1.0 - act ** 2)                                 # output_gradients = -dq/da, through tanh
    gA3, gc3 = g2.T @ du3, du3.sum(0)
    du2 = (du3 @ A3.T) * (u2 > 0)
    gA2, gc2 = g1.T @ du2, du2.sum(0)
    du1 = (du2 @ A2.T) * (u1 > 0)
    gA1, gc1 = s.T @ du1, du1.sum(0)
    return _flat([gA1, gc1, gA2, gc2, gA3, gc3]), float(q.sum())


def adam_step(params, grad, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7):
    """tf.keras Adam, step t >= 1: epsilon is added to sqrt(v) OUTSIDE the bias correction."""
    m = b1 * m + (1.0 - b1) * grad
    v = b2 * v + (1.0 - b2) * grad * grad
    lr_t = lr * np.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    return params - lr_t * m / (np.sqrt(v) + eps), m, v
