"""CPU oracle of the learner's numeric path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's CPU legs, never by skillshot_learning_b200/.

**Parity unpinned.**  The reference evaluates these steps inside TensorFlow /
Keras (version unpinned, not vendored, not installable here: no network), and
its repository holds no test, golden vector or saved model for them.  This file
therefore RESTATES the published Keras semantics at the reference's own call
sites (paths relative to the reference repo root):

  * model_define_actor   SkillshotLearner.py:70-96    12 -> 256 relu -> 128 relu -> 2 tanh,
                         kernels RandomNormal(0, 0.05), biases 0
  * model_define_critic  SkillshotLearner.py:98-121   s -> 256 relu -> Dropout(0.2) -> concat(a)
                         -> 128 relu -> 1 linear; hidden kernels glorot-uniform (the Keras
                         default), last kernel "RandomNormal" = N(0, 0.05); loss "mse"
  * model_act_param_noise SkillshotLearner.py:245-281  w += w * N(0, 0.5) on all six arrays
  * critic fit           SkillshotLearner.py:434      Keras fit: 1 epoch, shuffle, batch 16,
                         short last batch kept, Dropout active, MSE = mean over the batch
  * model_actor_fit_step SkillshotLearner.py:386-417  a = actor(s); q = critic([s, a])
                         (Dropout off); g = d a / d theta with output_gradients = -dq/da,
                         summed over the batch; Adam.apply_gradients
  * tf.keras Adam        defaults lr 1e-3, beta 0.9 / 0.999, epsilon 1e-7:
                         lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);
                         theta -= lr_t * m / (sqrt(v) + eps)       (eps outside the correction)

Dense = x @ W[in, out] + b.  Parameters are kept in Keras get_weights() order as
one flat float32 vector per network, the layout the CUDA library uses
(include/skillshot_b200.h).  Gradients come from torch autograd, i.e. they are
derived independently of the hand-written backward pass of the kernels.

The DDPG pieces the reference's readme points to but never implements (target
networks, tau, gamma, replay) are defined here so that gamma = 0, tau = 1 and a
one-episode buffer reduce exactly to the reference's update.
"""
from __future__ import annotations

import numpy as np
import torch

DIM_S, DIM_A, H1, H2 = 12, 2, 256, 128

ACTOR_SHAPES = [(DIM_S, H1), (H1,), (H1, H2), (H2,), (H2, DIM_A), (DIM_A,)]
CRITIC_SHAPES = [(DIM_S, H1), (H1,), (H1 + DIM_A, H2), (H2,), (H2, 1), (1,)]
ACTOR_PARAMS = sum(int(np.prod(s)) for s in ACTOR_SHAPES)      # 36,482
CRITIC_PARAMS = sum(int(np.prod(s)) for s in CRITIC_SHAPES)    # 36,609


def actor_shapes(frames=1):
    """model_define_actor with a first Dense layer of 12 * frames inputs (readme.md:18-20; frames = 1: the reference)."""
    return [(DIM_S * frames, H1), (H1,), (H1, H2), (H2,), (H2, DIM_A), (DIM_A,)]


def critic_shapes(frames=1):
    return [(DIM_S * frames, H1), (H1,), (H1 + DIM_A, H2), (H2,), (H2, 1), (1,)]


def split(flat, shapes):
    out, o = [], 0
    for s in shapes:
        n = int(np.prod(s))
        out.append(flat[o:o + n].reshape(s))
        o += n
    return out


def init_actor(rng: np.random.Generator, frames=1) -> np.ndarray:
    """SkillshotLearner.py:74-89: every kernel N(0, 0.05), biases zero."""
    parts = []
    for s in actor_shapes(frames):
        parts.append(rng.normal(0.0, 0.05, size=s) if len(s) == 2 else np.zeros(s))
    return np.concatenate([p.ravel() for p in parts]).astype(np.float32)


def init_critic(rng: np.random.Generator, frames=1) -> np.ndarray:
    """SkillshotLearner.py:104-114: glorot-uniform hidden kernels, N(0, 0.05) output kernel."""
    parts = []
    for idx, s in enumerate(critic_shapes(frames)):
        if len(s) == 1:
            parts.append(np.zeros(s))
        elif idx == 4:
            parts.append(rng.normal(0.0, 0.05, size=s))
        else:
            lim = np.sqrt(6.0 / (s[0] + s[1]))
            parts.append(rng.uniform(-lim, lim, size=s))
    return np.concatenate([p.ravel() for p in parts]).astype(np.float32)


def _t(x, dtype):
    return torch.as_tensor(np.asarray(x), dtype=dtype)


def actor_forward_t(theta: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    w1, b1, w2, b2, w3, b3 = split(theta, actor_shapes(s.shape[1] // DIM_S))       # input width 12 * frames
    h = torch.relu(s @ w1 + b1)
    h = torch.relu(h @ w2 + b2)
    return torch.tanh(h @ w3 + b3)


def critic_forward_t(phi: torch.Tensor, s: torch.Tensor, a: torch.Tensor, keep=None, rate: float = 0.2):
    """keep: None = Dropout off (model(x) call / predict); else a {0,1} mask [n,256]
    applied with the inverted scaling 1 / (1 - rate) as Keras does during fit."""
    w1, b1, w2, b2, w3, b3 = split(phi, critic_shapes(s.shape[1] // DIM_S))        # input width 12 * frames
    h = torch.relu(s @ w1 + b1)
    if keep is not None:
        h = h * keep * (1.0 / (1.0 - rate))
    h = torch.relu(torch.cat([h, a], dim=1) @ w2 + b2)
    return (h @ w3 + b3)[:, 0]


def actor_forward(theta, s, dtype=torch.float32) -> np.ndarray:
    with torch.no_grad():
        return actor_forward_t(_t(theta, dtype), _t(s, dtype)).numpy()


def critic_forward(phi, s, a, keep=None, rate=0.2, dtype=torch.float32) -> np.ndarray:
    with torch.no_grad():
        k = None if keep is None else _t(keep, dtype)
        return critic_forward_t(_t(phi, dtype), _t(s, dtype), _t(a, dtype), k, rate).numpy()


def frames_actor_forward(theta, x, frames, dtype=torch.float32) -> np.ndarray:
    """The frame-stacked actor (readme.md:18-20; no reference code): model_define_actor with a first Dense layer of
    12 * frames inputs; x [n, 12 * frames] ordered oldest frame first.  frames = 1 is actor_forward."""
    shapes = [(DIM_S * frames, H1), (H1,), (H1, H2), (H2,), (H2, DIM_A), (DIM_A,)]
    with torch.no_grad():
        w1, b1, w2, b2, w3, b3 = split(_t(theta, dtype), shapes)
        h = torch.relu(_t(x, dtype) @ w1 + b1)
        h = torch.relu(h @ w2 + b2)
        return torch.tanh(h @ w3 + b3).numpy()


def noisy_actor_params(theta: np.ndarray, eps: np.ndarray, sd: float) -> np.ndarray:
    """SkillshotLearner.py:260-265 with the normal draws injected: w += w * (sd * eps)."""
    theta = np.asarray(theta, np.float32)
    return (theta + theta * (np.float32(sd) * np.asarray(eps, np.float32))).astype(np.float32)


def critic_grad(phi, s, a, y, keep=None, rate=0.2, n_global=None, dtype=torch.float32):
    """Gradient of the Keras "mse" loss mean((q - y)^2) of one batch (SkillshotLearner.py:118, 434).
    n_global: the divisor of the mean when the batch is a shard of a larger one.
    Returns (grad flat, sum of squared errors)."""
    p = _t(phi, dtype).clone().requires_grad_(True)
    k = None if keep is None else _t(keep, dtype)
    q = critic_forward_t(p, _t(s, dtype), _t(a, dtype), k, rate)
    sse = ((q - _t(y, dtype)) ** 2).sum()
    (sse / float(n_global or len(y))).backward()
    return p.grad.numpy().copy(), float(sse.detach())


def actor_grad(theta, phi, s, dtype=torch.float32):
    """model_actor_fit_step (SkillshotLearner.py:395-410): gradient of -sum_batch Q(s, actor(s))
    with respect to the actor parameters, critic in inference mode.  Returns (grad flat, sum q)."""
    t = _t(theta, dtype).clone().requires_grad_(True)
    st = _t(s, dtype)
    q = critic_forward_t(_t(phi, dtype), st, actor_forward_t(t, st))
    (-q.sum()).backward()
    return t.grad.numpy().copy(), float(q.sum().detach())


class AdamTF:
    """tf.keras.optimizers.Adam (defaults) on one flat parameter vector."""

    def __init__(self, n, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7, dtype=np.float32):
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.m = np.zeros(n, dtype)
        self.v = np.zeros(n, dtype)
        self.t = 0
        self.dtype = dtype

    def step(self, params: np.ndarray, grad: np.ndarray) -> np.ndarray:
        self.t += 1
        d = self.dtype
        g = grad.astype(d)
        lr_t = d(self.lr * np.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t))
        self.m = (d(self.b1) * self.m + d(1.0 - self.b1) * g).astype(d)
        self.v = (d(self.b2) * self.v + d(1.0 - self.b2) * g * g).astype(d)
        return (params.astype(d) - lr_t * self.m / (np.sqrt(self.v) + d(self.eps))).astype(d)


def soft_update(target, online, tau):
    return (np.float32(tau) * online + np.float32(1.0 - tau) * target).astype(np.float32)


def ddpg_targets(theta_t, phi_t, r, s2, done, gamma, dtype=torch.float32):
    """y = r + gamma * (1 - done) * Q'(s2, mu'(s2)); gamma = 0 gives the reference's y = r."""
    r = np.asarray(r, np.float32)
    if gamma == 0.0:
        return r.copy()
    a2 = actor_forward(theta_t, s2, dtype)
    q2 = critic_forward(phi_t, s2, a2, None, dtype=dtype)
    return (r + np.float32(gamma) * (1.0 - np.asarray(done, np.float32)) * q2).astype(np.float32)


class LearnerOracle:
    """models_fit (SkillshotLearner.py:419-443) with every random choice injected."""

    def __init__(self, theta, phi, batch_size=16, dropout=0.2):
        self.theta = np.asarray(theta, np.float32).copy()
        self.phi = np.asarray(phi, np.float32).copy()
        self.opt_actor = AdamTF(ACTOR_PARAMS)     # self.optimiser, SkillshotLearner.py:68
        self.opt_critic = AdamTF(CRITIC_PARAMS)   # compile(optimizer="adam"), SkillshotLearner.py:118
        self.batch_size, self.dropout = batch_size, dropout

    def critic_fit(self, s, a, y, order, keep):
        """One Keras epoch over the rows in `order` (fit's own shuffle), batches of batch_size,
        keep = dropout masks [n,256] indexed like s.  Returns the per-batch losses."""
        losses = []
        for b in range(0, len(order), self.batch_size):
            idx = order[b:b + self.batch_size]
            g, sse = critic_grad(self.phi, s[idx], a[idx], y[idx], keep[idx], self.dropout)
            self.phi = self.opt_critic.step(self.phi, g)
            losses.append(sse / len(idx))
        return losses

    def actor_fit(self, s):
        """SkillshotLearner.py:440-443: consecutive batches of the (already shuffled) states."""
        qs = []
        for b in range(0, len(s), self.batch_size):
            g, q = actor_grad(self.theta, self.phi, s[b:b + self.batch_size])
            self.theta = self.opt_actor.step(self.theta, g)
            qs.append(q)
        return qs
