"""The learner path on the GPU: actor / critic networks, parameter-noise
exploration, replay ring, critic fit and actor policy-gradient step.

Three layers, all ending in launches of the CUDA library (C ABI:
include/skillshot_b200.h); there is no CPU path.

* :class:`ActorCritic` -- the two networks of SkillshotLearner.model_define_actor /
  model_define_critic (SkillshotLearner.py:70-121) as flat float32 parameter
  vectors with their tf.keras Adam state and (DDPG) target copies, and the batched
  device operations on them.
* :class:`ReplayRing` -- device-resident transitions.
* :class:`SkillshotLearner` -- the reference's class surface (same method names,
  argument meaning and attribute-style hyper-parameters) over one SkillshotGame.
* :class:`SelfPlayTrainer` -- the same loop for many envs at once: rollout of N
  games with the shared actor, replay, sharded update with one gradient all-reduce.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import check, lib
from .game import FEATURE_KEYS, REWARD_MODES, SkillshotEnvs, SkillshotGame

A_N, C_N = _lib.ACTOR_PARAMS, _lib.CRITIC_PARAMS
ACTOR_SHAPES = [(12, 256), (256,), (256, 128), (128,), (128, 2), (2,)]       # Keras get_weights() order
CRITIC_SHAPES = [(12, 256), (256,), (258, 128), (128,), (128, 1), (1,)]


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _f32(x, device, shape=None):
    if not torch.is_tensor(x):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    x = x.to(device=device, dtype=torch.float32).contiguous()
    return x if shape is None else x.reshape(shape)


def split_params(flat, shapes):
    """Views of a flat parameter vector in Keras get_weights() order."""
    out, o = [], 0
    for s in shapes:
        n = int(np.prod(s))
        out.append(flat[o:o + n].reshape(s))
        o += n
    return out


def shard_info(n_local: int, group=None):
    """(world, n_global, row_offset) of a batch sharded evenly over the ranks of `group`
    (None: not distributed; True: the default group).  Rank r owns global rows
    [r * n_local, (r + 1) * n_local): the critic's mean divides by n_global and Philox dropout is
    keyed by the global row, so a sharded update equals the single-GPU update of the whole batch."""
    if group is None:
        return 1, n_local, 0
    import torch.distributed as dist
    g = None if group is True else group
    world, rank = dist.get_world_size(g), dist.get_rank(g)
    return world, n_local * world, rank * n_local


def allreduce_sum(t: torch.Tensor, group=None):
    """The one exchange step of a sharded update: sum the flat gradient over ranks (NCCL on GPUs)."""
    if group is None:
        return t
    import torch.distributed as dist
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=None if group is True else group)
    return t


class PeerExchange:
    """The gradient exchange of a sharded update over NVLink peer memory (csrc/ss_peer.cu): every rank
    owns an inbox {flags | 2 x world x capacity floats} shared with the other processes of the node
    through CUDA IPC; a rank's gradient reduction kernel stores its result into every inbox, and each
    rank's Adam kernel sums its inbox in rank order.  Replaces the NCCL all-reduce call between the
    gradient kernel and Adam (one node, up to 8 ranks)."""

    def __init__(self, group, device, capacity: int = max(A_N, C_N)):      # (ActorCritic passes its own sizes)
        import torch.distributed as dist
        g = None if group is True else group
        self.group, self.device, self.capacity = g, torch.device(device), int(capacity)
        self.world, self.rank = dist.get_world_size(g), dist.get_rank(g)
        if self.world > _lib.PEER_MAX_WORLD:
            raise ValueError("peer exchange supports up to %d ranks of one node" % _lib.PEER_MAX_WORLD)
        self._own = ctypes.c_void_p()
        self._imported = []
        with torch.cuda.device(self.device):
            check(lib.ss_peer_alloc(self.world, self.capacity, ctypes.byref(self._own)), "ss_peer_alloc")
            handle = (ctypes.c_ubyte * _lib.PEER_HANDLE_BYTES)()
            check(lib.ss_peer_export(self._own, handle), "ss_peer_export")
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=g)
            self.bases = (ctypes.c_void_p * self.world)()
            for r, h in enumerate(handles):
                if r == self.rank:
                    self.bases[r] = self._own.value
                else:
                    buf = (ctypes.c_ubyte * _lib.PEER_HANDLE_BYTES).from_buffer_copy(h)
                    ptr = ctypes.c_void_p()
                    check(lib.ss_peer_import(buf, ctypes.byref(ptr)), "ss_peer_import")
                    self.bases[r] = ptr.value
                    self._imported.append(ptr)
        self.done_counter = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.epoch = 0
        dist.barrier(g)

    def reduce_push(self, workspace: torch.Tensor, parts: int, n_params: int, aux: Optional[torch.Tensor]):
        """Slices in `workspace` -> this rank's slot of every inbox; starts a new epoch."""
        self.epoch += 1
        with torch.cuda.device(self.device):
            check(lib.ss_peer_reduce_push(workspace.data_ptr(), int(parts), int(n_params), _ptr(aux), self.bases, self.world,
                                          self.rank, self.capacity, self.epoch, self.done_counter.data_ptr(),
                                          _stream(self.device)), "ss_peer_reduce_push")

    def check_status(self):
        if int(self.status.item()) & _lib.STATUS_PEER_TIMEOUT:
            self.status.zero_()
            raise RuntimeError("peer exchange: a rank's gradient never arrived")

    def close(self):
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for ptr in self._imported:
                lib.ss_peer_close(ptr)
            self._imported = []
            if self._own:
                import torch.distributed as dist
                if dist.is_initialized():
                    dist.barrier(self.group)          # nobody may still be storing into this inbox
                lib.ss_peer_free(self._own)
                self._own = ctypes.c_void_p()


class ActorCritic:
    """Actor and critic parameters, optimiser state and target copies on one GPU.

    Hyper-parameters default to the reference's: tf.keras Adam lr 1e-3, betas
    0.9 / 0.999, eps 1e-7 for both networks (SkillshotLearner.py:68, 118),
    Dropout(0.2) in the critic (SkillshotLearner.py:105).  gamma = 0 and tau = 1
    are the reference's update (critic regresses on the immediate reward, no
    target network); other values give DDPG proper.
    """

    def __init__(self, device="cuda", seed: int = 0, lr_actor: float = 1e-3, lr_critic: float = 1e-3,
                 gamma: float = 0.0, tau: float = 1.0, dropout: float = 0.2, process_group=None,
                 update_precision: str = "f32", collective: str = "nccl", frames: int = 1):
        if not torch.cuda.is_available():
            raise RuntimeError("skillshot_learning_b200 needs a CUDA device (no CPU fallback)")
        # frames > 1: the frame-stacked "planning" networks of readme.md:18-20 (no reference code): both first layers read
        # 12 * frames inputs, oldest frame first.  frames = 1 is the reference.  Their update runs on the exact float32
        # kernels (ss_*_frames); the tensor-core gradient kernels are built for the reference's 12 inputs.
        self.frames = int(frames)
        self.a_n, self.c_n = int(lib.ss_actor_frames_params(self.frames)), int(lib.ss_critic_frames_params(self.frames))
        if self.a_n < 0 or self.c_n < 0:
            raise ValueError("frames must be 1..20")
        if self.frames > 1 and update_precision != "f32":
            raise ValueError("the frame-stacked networks are updated by the float32 kernels: update_precision='f32'")
        self.actor_shapes = [(12 * self.frames, 256)] + ACTOR_SHAPES[1:]
        self.critic_shapes = [(12 * self.frames, 256)] + CRITIC_SHAPES[1:]
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.seed = int(seed)
        self.lr_actor, self.lr_critic = float(lr_actor), float(lr_critic)
        self.beta1, self.beta2, self.eps = 0.9, 0.999, 1e-7
        self.gamma, self.tau, self.dropout = float(gamma), float(tau), float(dropout)
        self.group = process_group
        if update_precision not in ("f32", "bf16"):
            raise ValueError("update_precision must be 'f32' (exact path) or 'bf16' (tensor cores)")
        self.update_precision = update_precision     # gradient / TD-target kernels: float32 CUDA cores or tcgen05
        if collective not in ("nccl", "peer"):
            raise ValueError("collective must be 'nccl' (torch.distributed all-reduce) or 'peer' (fused NVLink exchange)")
        # the exchange step of a sharded update: an all-reduce call, or the fused peer-memory kernels
        self.peer = (PeerExchange(process_group, self.device, capacity=max(self.a_n, self.c_n))
                     if (process_group is not None and collective == "peer") else None)
        dev = self.device
        A_N, C_N = self.a_n, self.c_n
        # one allocation: [actor | pad | critic] so both vectors are 16-byte aligned
        self._a_off, self._c_off = 0, (A_N + 3) // 4 * 4
        total = self._c_off + C_N
        self.params = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grads = torch.zeros(total, dtype=torch.float32, device=dev)
        self.adam_m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.adam_v = torch.zeros(total, dtype=torch.float32, device=dev)
        self.target = torch.zeros(total, dtype=torch.float32, device=dev)
        self.stats = torch.zeros(2, dtype=torch.float32, device=dev)        # [sum sq err, sum q] of the last steps
        # gradient slices: one per CTA of the gradient kernels (at most 160 of them run: 148 SMs, one CTA each for the wide nets)
        self._ws_base = (int(lib.ss_learner_workspace_bytes()) if self.frames == 1
                         else 160 * (max(A_N, C_N) + 1) * 4)
        self.workspace = torch.empty(self._ws_base, dtype=torch.uint8, device=dev)
        self.step_actor = 0
        self.step_critic = 0
        self.counter = 0            # Philox counter: advances with every noisy call
        self._update_args = None    # cached ss_ddpg_update argument block (pointers change rarely)
        self._pair_mail = None      # ss_actor_critic_forward_tc's mailbox, allocated with the argument block
        self.init_weights(seed)

    # -- parameter views -----------------------------------------------------
    @property
    def actor(self):
        return self.params[self._a_off:self._a_off + self.a_n]

    @property
    def critic(self):
        return self.params[self._c_off:self._c_off + self.c_n]

    @property
    def target_actor(self):
        return self.target[self._a_off:self._a_off + self.a_n]

    @property
    def target_critic(self):
        return self.target[self._c_off:self._c_off + self.c_n]

    def _slice(self, buf, which):
        return buf[self._a_off:self._a_off + self.a_n] if which == "actor" else buf[self._c_off:self._c_off + self.c_n]

    def init_weights(self, seed: int):
        """Keras initialisers of the reference's layers: actor kernels RandomNormal(0, 0.05)
        (SkillshotLearner.py:74), critic hidden kernels glorot-uniform (the Dense default),
        critic output kernel "RandomNormal" = N(0, 0.05) (SkillshotLearner.py:113), biases 0."""
        g = torch.Generator(device="cpu")
        g.manual_seed(int(seed))
        a = [torch.randn(s, generator=g) * 0.05 if len(s) == 2 else torch.zeros(s) for s in self.actor_shapes]
        c = []
        for i, s in enumerate(self.critic_shapes):
            if len(s) == 1:
                c.append(torch.zeros(s))
            elif i == 4:
                c.append(torch.randn(s, generator=g) * 0.05)
            else:
                lim = float(np.sqrt(6.0 / (s[0] + s[1])))
                c.append((torch.rand(s, generator=g) * 2.0 - 1.0) * lim)
        self.set_weights(torch.cat([t.reshape(-1) for t in a]), torch.cat([t.reshape(-1) for t in c]))

    def set_weights(self, actor=None, critic=None, reset_optimizer: bool = True):
        """Install flat parameter vectors (Keras get_weights() order); targets are set equal."""
        if actor is not None:
            self.actor.copy_(_f32(actor, self.device, (self.a_n,)))
        if critic is not None:
            self.critic.copy_(_f32(critic, self.device, (self.c_n,)))
        self.target.copy_(self.params)
        if reset_optimizer:
            self.adam_m.zero_()
            self.adam_v.zero_()
            self.step_actor = self.step_critic = 0

    def get_weights(self, which="actor"):
        """List of numpy arrays like keras Model.get_weights()."""
        flat = self._slice(self.params, which).detach().cpu().numpy()
        return [w.copy() for w in split_params(flat, self.actor_shapes if which == "actor" else self.critic_shapes)]

    # -- forward ---------------------------------------------------------------
    def actor_forward(self, obs, param_noise_sd: float = 0.0, noise_group: int = 1, action_noise_sd: float = 0.0,
                      out: Optional[torch.Tensor] = None, target: bool = False, counter: Optional[int] = None,
                      precision: str = "f32"):
        """actions [n,2] = actor(obs [n,12]) (model_act*, SkillshotLearner.py:215-281).  frames > 1: obs [n, 12 * frames],
        oldest frame first, float32 kernels, no noise (FrameStackActor is the acting path of the stacked networks)."""
        obs = _f32(obs, self.device).reshape(-1, 12 * self.frames)
        n = obs.shape[0]
        if out is None:
            out = torch.empty((n, 2), dtype=torch.float32, device=self.device)
        theta = self.target_actor if target else self.actor
        if self.frames > 1:
            if param_noise_sd > 0 or action_noise_sd > 0 or precision != "f32":
                raise ValueError("frame-stacked networks: plain float32 forward only (see FrameStackActor)")
            with torch.cuda.device(self.device):      # dense ordered rows = a history ring read from slot 0
                check(lib.ss_actor_forward_frames(theta.data_ptr(), 0, 0, obs.data_ptr(), self.frames, self.frames - 1,
                                                  out.data_ptr(), n, _stream(self.device)), "ss_actor_forward_frames")
            return out
        if counter is None:
            counter = self.counter
            if param_noise_sd > 0 or action_noise_sd > 0:
                self.counter += 1
        if precision not in ("f32", "bf16"):
            raise ValueError("precision must be 'f32' (exact path) or 'bf16' (tensor cores)")
        fn = lib.ss_actor_forward if precision == "f32" else lib.ss_actor_forward_tc
        with torch.cuda.device(self.device):
            check(fn(theta.data_ptr(), obs.data_ptr(), out.data_ptr(), n, float(param_noise_sd), int(noise_group),
                     float(action_noise_sd), self.seed, int(counter), _stream(self.device)), "ss_actor_forward")
        return out

    def noisy_actor_params(self, sd: float, group: int = 0, counter: Optional[int] = None):
        """The perturbed parameter vector actor_forward uses for one noise group."""
        out = torch.empty(self.a_n, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.ss_param_noise(self.actor.data_ptr(), out.data_ptr(), self.a_n, float(sd), self.seed, int(group),
                                     int(self.counter if counter is None else counter), _stream(self.device)),
                  "ss_param_noise")
        return out

    def critic_forward(self, obs, act, target: bool = False, precision: str = "f32", want_dq_da: bool = False):
        """q [n] = critic([obs, act]) with Dropout off; with want_dq_da (bf16 path) also -dQ/da [n,2]."""
        obs, act = _f32(obs, self.device).reshape(-1, 12 * self.frames), _f32(act, self.device).reshape(-1, 2)
        n = obs.shape[0]
        q = torch.empty(n, dtype=torch.float32, device=self.device)
        phi = self.target_critic if target else self.critic
        with torch.cuda.device(self.device):
            if self.frames > 1:
                if precision != "f32":
                    raise ValueError("frame-stacked networks: float32 kernels only")
                check(lib.ss_critic_forward_frames(phi.data_ptr(), self.frames, obs.data_ptr(), act.data_ptr(), q.data_ptr(), n,
                                                   _stream(self.device)), "ss_critic_forward_frames")
                return q
            if precision == "bf16":
                up = torch.empty((n, 2), dtype=torch.float32, device=self.device) if want_dq_da else None
                check(lib.ss_critic_forward_tc(phi.data_ptr(), obs.data_ptr(), act.data_ptr(), n, q.data_ptr(), _ptr(up),
                                               None, None, 0.0, None, _stream(self.device)), "ss_critic_forward_tc")
                return (q, up) if want_dq_da else q
            check(lib.ss_critic_forward(phi.data_ptr(), obs.data_ptr(), act.data_ptr(), q.data_ptr(), n,
                                        _stream(self.device)), "ss_critic_forward")
        return q

    def td_targets(self, reward, next_obs, done=None):
        """y = r + gamma (1 - done) Q'(s', mu'(s')); gamma = 0 returns the reward itself
        (the reference's critic target, SkillshotLearner.py:434)."""
        reward = _f32(reward, self.device).reshape(-1)
        if self.gamma == 0.0:
            return reward
        next_obs = _f32(next_obs, self.device).reshape(-1, 12 * self.frames)
        y = torch.empty_like(reward)
        if done is not None:
            done = done.to(device=self.device, dtype=torch.uint8).contiguous()
        n = reward.shape[0]
        with torch.cuda.device(self.device):
            if self.update_precision == "bf16":
                ws = self._workspace_for(n)
                check(lib.ss_ddpg_targets_tc(self.target_actor.data_ptr(), self.target_critic.data_ptr(),
                                             reward.data_ptr(), next_obs.data_ptr(), _ptr(done), self.gamma,
                                             y.data_ptr(), n, ws.data_ptr(), ws.numel(), _stream(self.device)),
                      "ss_ddpg_targets_tc")
            else:
                check(lib.ss_ddpg_targets_frames(self.target_actor.data_ptr(), self.target_critic.data_ptr(), self.frames,
                                                 reward.data_ptr(), next_obs.data_ptr(), _ptr(done), self.gamma, y.data_ptr(), n,
                                                 _stream(self.device)), "ss_ddpg_targets")
        return y

    def _workspace_for(self, n: int) -> torch.Tensor:
        """Scratch for the gradient kernels: per-CTA gradient slices (+ 32 bytes per row for the
        tensor-core entry points: actions, -dQ/da and Q of the actor step)."""
        need = self._ws_base + (32 * n if self.update_precision == "bf16" else 0)
        if self.workspace.numel() < need:
            self.workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self.workspace

    # -- update ----------------------------------------------------------------
    def _allreduce(self, t):
        allreduce_sum(t, self.group)

    def critic_grad(self, obs, act, y, keep=None, n_global: int = 0, row_offset: int = 0, slices_only: bool = False):
        """grads[critic] <- d/dphi mean (q - y)^2 of this (shard of a) batch; returns the gradient view.
        slices_only: leave the per-CTA slices in the workspace and return their number (peer exchange)."""
        obs, act = _f32(obs, self.device).reshape(-1, 12 * self.frames), _f32(act, self.device).reshape(-1, 2)
        y = _f32(y, self.device).reshape(-1)
        n = obs.shape[0]
        if keep is not None:
            keep = torch.as_tensor(keep).to(device=self.device, dtype=torch.uint8).contiguous()
        g = self._slice(self.grads, "critic")
        ws = self._workspace_for(n)
        tail = (_ptr(keep), self.dropout, self.seed, self.counter, n, int(n_global), int(row_offset),
                None if slices_only else g.data_ptr(), self.stats[0:1].data_ptr(), ws.data_ptr(), ws.numel(), _stream(self.device))
        with torch.cuda.device(self.device):
            if self.update_precision == "bf16":
                rc = lib.ss_critic_grad_tc(self.critic.data_ptr(), obs.data_ptr(), act.data_ptr(), y.data_ptr(), *tail)
            else:
                rc = lib.ss_critic_grad_frames(self.critic.data_ptr(), self.frames, obs.data_ptr(), act.data_ptr(), y.data_ptr(), *tail)
        self.counter += 1
        if slices_only and rc > 0:
            return rc
        check(rc, "ss_critic_grad")
        return g

    def actor_grad(self, obs, slices_only: bool = False):
        """grads[actor] <- -sum_batch dQ/da da/dtheta (model_actor_fit_step, SkillshotLearner.py:395-410)."""
        obs = _f32(obs, self.device).reshape(-1, 12 * self.frames)
        g = self._slice(self.grads, "actor")
        ws = self._workspace_for(obs.shape[0])
        tail = (obs.data_ptr(), obs.shape[0], None if slices_only else g.data_ptr(), self.stats[1:2].data_ptr(), ws.data_ptr(),
                ws.numel(), _stream(self.device))
        with torch.cuda.device(self.device):
            if self.update_precision == "bf16":
                rc = lib.ss_actor_grad_tc(self.actor.data_ptr(), self.critic.data_ptr(), *tail)
            else:
                rc = lib.ss_actor_grad_frames(self.actor.data_ptr(), self.critic.data_ptr(), self.frames, *tail)
        if slices_only and rc > 0:
            return rc
        check(rc, "ss_actor_grad")
        return g

    def apply_adam(self, which: str, grad_scale: float = 1.0, from_peers: bool = False):
        """tf.keras Adam.apply_gradients on one network (+ soft target update with self.tau).
        from_peers: the gradient is the rank-ordered sum of this rank's peer inbox (fused exchange)."""
        n = self.a_n if which == "actor" else self.c_n
        if which == "actor":
            self.step_actor += 1
            step, lr = self.step_actor, self.lr_actor
        else:
            self.step_critic += 1
            step, lr = self.step_critic, self.lr_critic
        if from_peers:
            px = self.peer
            with torch.cuda.device(self.device):
                check(lib.ss_peer_adam_tf(px.bases[px.rank], px.world, px.capacity, px.epoch,
                                          self._slice(self.params, which).data_ptr(), self._slice(self.adam_m, which).data_ptr(),
                                          self._slice(self.adam_v, which).data_ptr(), self._slice(self.target, which).data_ptr(),
                                          self._slice(self.grads, which).data_ptr(), n, step, lr, self.beta1, self.beta2,
                                          self.eps, self.tau, float(grad_scale), px.status.data_ptr(), _stream(self.device)),
                      "ss_peer_adam_tf")
            return
        with torch.cuda.device(self.device):
            check(lib.ss_adam_tf(self._slice(self.params, which).data_ptr(), self._slice(self.grads, which).data_ptr(),
                                 self._slice(self.adam_m, which).data_ptr(), self._slice(self.adam_v, which).data_ptr(),
                                 self._slice(self.target, which).data_ptr(), n, step, lr, self.beta1, self.beta2,
                                 self.eps, self.tau, float(grad_scale), _stream(self.device)), "ss_adam_tf")

    def reduce_adam(self, which: str, parts: int, aux: Optional[torch.Tensor], grad_scale: float = 1.0):
        """Single GPU: fixed-order sum of the `parts` gradient slices in the workspace + Adam + soft update, one kernel."""
        n = self.a_n if which == "actor" else self.c_n
        if which == "actor":
            self.step_actor += 1
            step, lr = self.step_actor, self.lr_actor
        else:
            self.step_critic += 1
            step, lr = self.step_critic, self.lr_critic
        with torch.cuda.device(self.device):
            check(lib.ss_reduce_adam_tf(self.workspace.data_ptr(), int(parts), n, _ptr(aux), self._slice(self.grads, which).data_ptr(),
                                        self._slice(self.params, which).data_ptr(), self._slice(self.adam_m, which).data_ptr(),
                                        self._slice(self.adam_v, which).data_ptr(), self._slice(self.target, which).data_ptr(),
                                        step, lr, self.beta1, self.beta2, self.eps, self.tau, float(grad_scale),
                                        _stream(self.device)), "ss_reduce_adam_tf")

    def critic_step(self, obs, act, y, keep=None):
        """One batch of model_critic.fit (SkillshotLearner.py:434): gradient of the batch-mean
        squared error, summed over the ranks when sharded, then Adam.  Returns the (device) sum of squared errors."""
        _, n_global, row_offset = shard_info(int(np.prod(y.shape)), self.group)
        if self.group is not None and self.peer is None:      # NCCL: gradient -> all-reduce -> Adam
            g = self.critic_grad(obs, act, y, keep, n_global=n_global, row_offset=row_offset)
            self._allreduce(g)
            self.apply_adam("critic")
            return self.stats[0]
        # per-CTA slices -> one kernel that sums them and applies Adam (single GPU), or pushes the sum into every
        # rank's inbox for the peers' Adam kernel (fused exchange)
        parts = self.critic_grad(obs, act, y, keep, n_global=n_global, row_offset=row_offset, slices_only=True)
        if self.peer is not None:
            self.peer.reduce_push(self.workspace, parts, self.c_n, self.stats[0:1])
            self.apply_adam("critic", from_peers=True)
        else:
            self.reduce_adam("critic", parts, self.stats[0:1])
        return self.stats[0]

    def actor_step(self, obs):
        """model_actor_fit_step (SkillshotLearner.py:386-417).  Returns the (device) sum of q."""
        if self.group is not None and self.peer is None:
            g = self.actor_grad(obs)
            self._allreduce(g)
            self.apply_adam("actor")
            return self.stats[1]
        parts = self.actor_grad(obs, slices_only=True)
        if self.peer is not None:
            self.peer.reduce_push(self.workspace, parts, self.a_n, self.stats[1:2])
            self.apply_adam("actor", from_peers=True)
        else:
            self.reduce_adam("actor", parts, self.stats[1:2])
        return self.stats[1]

    def update_from_ring(self, ring: "ReplayRing", batch: int, out: Optional[dict] = None, sample_early: bool = False):
        """One whole update step from ONE library call (ss_ddpg_update): sample `batch` rows of `ring`,
        TD target, critic step, actor step -- the launches critic_step / actor_step make, without the
        interpreter time between them.  Single GPU or the fused peer exchange (an NCCL all-reduce needs
        the host between the gradient and Adam: use the separate steps).  Returns the minibatch dict.
        sample_early: the caller guarantees that the previous launch on the stream neither writes the ring nor touches
        `out` (ss_ddpg_update_args.sample_early): the minibatch is gathered while that launch still runs."""
        if self.group is not None and self.peer is None:
            raise ValueError("update_from_ring needs collective='peer' when the update is sharded")
        if self.frames > 1:
            raise ValueError("the single-call update is built for the reference's 12-input networks: use the separate steps")
        if ring.size == 0:
            raise ValueError("sampling from an empty ring")
        dev = self.device
        if out is None or out["reward"].shape[0] != batch or "y" not in out:
            out = dict(obs=torch.empty((batch, 12), dtype=torch.float32, device=dev),
                       act=torch.empty((batch, 2), dtype=torch.float32, device=dev),
                       reward=torch.empty(batch, dtype=torch.float32, device=dev),
                       next_obs=torch.empty((batch, 12), dtype=torch.float32, device=dev),
                       done=torch.empty(batch, dtype=torch.uint8, device=dev),
                       indices=torch.empty(batch, dtype=torch.int64, device=dev),
                       y=torch.empty(batch, dtype=torch.float32, device=dev))
        _, n_global, row_offset = shard_info(batch, self.group)
        ws = self._workspace_for(batch)
        key = (ring.obs.data_ptr(), ring.next_obs.data_ptr(), ring.capacity, out["obs"].data_ptr(), out["y"].data_ptr(), ws.data_ptr(),
               ws.numel(), batch, id(self.peer))
        if self._update_args is None:
            self._update_args = {}
        a = self._update_args.get(key)      # one cached argument block per set of minibatch buffers (callers alternate two)
        if a is None:
            if len(self._update_args) >= 4:
                self._update_args.clear()
            a = self._update_args[key] = _lib.DdpgUpdateArgs()
            a.ring_obs, a.ring_act, a.ring_reward = ring.obs.data_ptr(), ring.act.data_ptr(), ring.reward.data_ptr()
            a.ring_next_obs, a.ring_done, a.capacity = ring.next_obs.data_ptr(), ring.done.data_ptr(), ring.capacity
            a.batch = batch
            for k in ("obs", "act", "reward", "next_obs", "done", "indices", "y"):
                setattr(a, k, out[k].data_ptr())
            for k in ("actor", "critic"):
                setattr(a, k, self._slice(self.params, k).data_ptr())
                setattr(a, "target_" + k, self._slice(self.target, k).data_ptr())
                setattr(a, "m_" + k, self._slice(self.adam_m, k).data_ptr())
                setattr(a, "v_" + k, self._slice(self.adam_v, k).data_ptr())
                setattr(a, "grad_" + k, self._slice(self.grads, k).data_ptr())
            a.stats = self.stats.data_ptr()
            a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
            # the actor -> critic forward pairs (TD target, actor step) as one launch each (ss_actor_critic_forward_tc), for
            # minibatches small enough that a forward launch is mostly set-up and pipeline fill; the mailbox is filled with its
            # empty mark here once and left so by every call (SS_UPDATE_PAIR=0: two launches each, for A/B measurements)
            # (measured, profiles/r2_update_pair_ab.txt: 162.5 -> 158.9 us at 65,536 rows, 236 -> 252 us at 131,072)
            if batch <= 65536 and os.environ.get("SS_UPDATE_PAIR", "1") != "0":
                if self._pair_mail is None or self._pair_mail.shape[0] != batch:
                    self._pair_mail = torch.full((batch, 2), _lib.PAIR_MAIL_EMPTY, dtype=torch.int32, device=dev)
                a.pair_mail = self._pair_mail.data_ptr()
            else:
                self._pair_mail = None
            if self.peer is not None:
                px = self.peer
                a.world, a.rank, a.peer_capacity = px.world, px.rank, px.capacity
                a.peer_bases = ctypes.cast(px.bases, ctypes.c_void_p)
                a.done_counter, a.status = px.done_counter.data_ptr(), px.status.data_ptr()
        a.size, a.replay_seed, a.replay_counter = ring.size, ring.seed, ring.counter
        a.gamma, a.tau, a.lr_actor, a.lr_critic = self.gamma, self.tau, self.lr_actor, self.lr_critic
        a.beta1, a.beta2, a.eps, a.dropout_rate = self.beta1, self.beta2, self.eps, self.dropout
        a.seed, a.counter = self.seed, self.counter
        a.step_critic, a.step_actor = self.step_critic + 1, self.step_actor + 1
        a.n_global, a.row_offset = n_global, row_offset
        a.tensor_cores = 1 if self.update_precision == "bf16" else 0
        a.sample_early = 1 if sample_early else 0
        if self.peer is not None:
            a.epoch = self.peer.epoch + 1
        with torch.cuda.device(dev):
            check(lib.ss_ddpg_update(ctypes.byref(a), _stream(dev)), "ss_ddpg_update")
        ring.counter += 1
        self.counter += 1
        self.step_critic += 1
        self.step_actor += 1
        if self.peer is not None:
            self.peer.epoch += 2
        return out

    # -- checkpoint ------------------------------------------------------------
    def state_dict(self):
        return dict(params=self.params.cpu(), target=self.target.cpu(), adam_m=self.adam_m.cpu(),
                    adam_v=self.adam_v.cpu(), step_actor=self.step_actor, step_critic=self.step_critic,
                    counter=self.counter, seed=self.seed)

    def load_state_dict(self, sd):
        for k in ("params", "target", "adam_m", "adam_v"):
            getattr(self, k).copy_(sd[k].to(self.device))
        self.step_actor, self.step_critic = int(sd["step_actor"]), int(sd["step_critic"])
        self.counter, self.seed = int(sd["counter"]), int(sd["seed"])


class ReplayRing:
    """Device-resident replay ring (structure of arrays).  The reference keeps one episode in
    Python lists and uses it once (SkillshotLearner.py:292-361); a ring sized to the episode and
    read back in order is that buffer."""

    def __init__(self, capacity: int, device="cuda", seed: int = 0, frames: int = 1):
        self.capacity, self.device, self.seed = int(capacity), torch.device(device), int(seed)
        self.frames = int(frames)          # observation rows of 12 * frames floats (the frame-stacked networks' inputs)
        self.width = 12 * self.frames
        c, dev = self.capacity, self.device
        self.obs = torch.zeros((c, self.width), dtype=torch.float32, device=dev)
        self.act = torch.zeros((c, 2), dtype=torch.float32, device=dev)
        self.reward = torch.zeros(c, dtype=torch.float32, device=dev)
        self.next_obs = torch.zeros((c, self.width), dtype=torch.float32, device=dev)
        self.done = torch.zeros(c, dtype=torch.uint8, device=dev)
        self.pos = 0
        self.size = 0
        self.counter = 0

    def push(self, obs, act, reward, next_obs, done=None, done_div: int = 1):
        obs, next_obs = _f32(obs, self.device).reshape(-1, self.width), _f32(next_obs, self.device).reshape(-1, self.width)
        act, reward = _f32(act, self.device).reshape(-1, 2), _f32(reward, self.device).reshape(-1)
        n = obs.shape[0]
        if n > self.capacity:
            raise ValueError("push of %d rows into a ring of %d" % (n, self.capacity))
        if done is not None:
            done = done.to(device=self.device, dtype=torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            check(lib.ss_replay_push_frames(self.obs.data_ptr(), self.act.data_ptr(), self.reward.data_ptr(),
                                            self.next_obs.data_ptr(), self.done.data_ptr(), self.capacity, self.pos, self.frames,
                                            obs.data_ptr(), act.data_ptr(), reward.data_ptr(), next_obs.data_ptr(),
                                            _ptr(done), int(done_div), n, _stream(self.device)), "ss_replay_push")
        self.pos = (self.pos + n) % self.capacity
        self.size = min(self.capacity, self.size + n)

    def state_dict(self):
        return dict(capacity=self.capacity, seed=self.seed, pos=self.pos, size=self.size, counter=self.counter, frames=self.frames,
                    **{k: getattr(self, k).cpu() for k in ("obs", "act", "reward", "next_obs", "done")})

    def load_state_dict(self, sd):
        if int(sd["capacity"]) != self.capacity:
            raise ValueError("checkpoint ring holds %d rows, this ring %d" % (sd["capacity"], self.capacity))
        for k in ("obs", "act", "reward", "next_obs", "done"):
            getattr(self, k).copy_(sd[k].to(self.device))
        self.seed, self.pos, self.size, self.counter = int(sd["seed"]), int(sd["pos"]), int(sd["size"]), int(sd["counter"])

    def sample(self, batch: int, indices=None, out: Optional[dict] = None):
        """dict(obs, act, reward, next_obs, done, indices) of `batch` rows: the rows `indices` or a
        uniform draw with replacement from the filled part of the ring."""
        if self.size == 0:
            raise ValueError("sampling from an empty ring")
        dev = self.device
        if out is None or out["reward"].shape[0] != batch:
            out = dict(obs=torch.empty((batch, self.width), dtype=torch.float32, device=dev),
                       act=torch.empty((batch, 2), dtype=torch.float32, device=dev),
                       reward=torch.empty(batch, dtype=torch.float32, device=dev),
                       next_obs=torch.empty((batch, self.width), dtype=torch.float32, device=dev),
                       done=torch.empty(batch, dtype=torch.uint8, device=dev),
                       indices=torch.empty(batch, dtype=torch.int64, device=dev))
        if indices is not None:
            indices = torch.as_tensor(indices).to(device=dev, dtype=torch.int64).contiguous()
            if indices.numel() != batch:
                raise ValueError("indices must have `batch` entries")
        with torch.cuda.device(dev):
            check(lib.ss_replay_sample_frames(self.obs.data_ptr(), self.act.data_ptr(), self.reward.data_ptr(),
                                       self.next_obs.data_ptr(), self.done.data_ptr(), self.capacity, self.size, self.frames,
                                       _ptr(indices), self.seed, self.counter, batch, out["obs"].data_ptr(),
                                       out["act"].data_ptr(), out["reward"].data_ptr(), out["next_obs"].data_ptr(),
                                       out["done"].data_ptr(), out["indices"].data_ptr(), _stream(dev)),
                  "ss_replay_sample")
        self.counter += 1
        return out


class SkillshotLearner:
    """The reference SkillshotLearner surface (SkillshotLearner.py:12-693) over the CUDA
    library: one SkillshotGame, two players sharing one actor, critic fitted on the
    immediate reward, actor stepped along dQ/da, multiplicative parameter noise.

    Hyper-parameters are plain attributes as in the reference
    (`learner.model_param_game_tick_limit = 200`, SkillshotLearner.py:688).
    """
    game_state_features = list(FEATURE_KEYS)          # SkillshotLearner.py:17-36

    def __init__(self, device="cuda", seed: Optional[int] = None):
        seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
        # environment (SkillshotLearner.py:41-44)
        self.game_environment = SkillshotGame(device=device)
        self.player_ids = (1, 2)
        self.max_dist_normaliser = (2 * (250 ** 2)) ** 0.5
        self.use_random_start = True
        # dir locations (SkillshotLearner.py:47-51)
        self.save_location = "training_models"
        self.actor_dir_name = "actor"
        self.critic_dir_name = "critic"
        self.training_progress_dir_name = "training_progress"
        self.training_boards_dir_name = "training_boards"
        # model (SkillshotLearner.py:54-58)
        self.dim_state_space = 12
        self.dim_action_space = 2
        self.dim_reward_space = 1
        self.networks = ActorCritic(device=device, seed=seed)
        # model hyper-params (SkillshotLearner.py:61-64)
        self.model_param_batch_size = 16
        self.model_param_game_tick_limit = 2000
        self.action_noise_sd = 0.15
        self.param_noise_sd = 0.5
        self._rng = np.random.RandomState(seed)      # the shuffles the reference takes from np.random
        self.last_fit = {}

    # -- acting (SkillshotLearner.py:206-281) ------------------------------------
    def do_actions(self, player_id, predictions):
        player = self.game_environment.get_player_by_id(player_id)
        player.move_direction_float(float(predictions[0]))
        player.move_look_float(float(predictions[1]))
        player.move_shoot_projectile()           # attempted every time, SkillshotLearner.py:212-213

    def _predict(self, game_state, player_id, **noise):
        features = np.asarray(self.prepare_states([game_state], player_id)[0], dtype=np.float32)
        return self.networks.actor_forward(features[None, :], **noise).cpu().numpy()

    def model_act(self, game_state, player_id):
        predictions = self._predict(game_state, player_id)
        self.do_actions(player_id, predictions[0])
        return predictions

    def model_act_action_noise(self, game_state, player_id):
        predictions = self._predict(game_state, player_id, action_noise_sd=self.action_noise_sd)
        self.do_actions(player_id, predictions[0])
        return predictions

    def model_act_param_noise(self, game_state, player_id):
        predictions = self._predict(game_state, player_id, param_noise_sd=self.param_noise_sd, noise_group=1)
        self.do_actions(player_id, predictions[0])
        return predictions

    # -- dataset preparation (SkillshotLearner.py:512-573) ------------------------
    def prepare_states(self, game_states, player_id):
        board = self.game_environment.board_size
        cooldown_max = self.game_environment.get_player_by_id(player_id).projectile.cooldown_max
        norm = self.max_dist_normaliser
        rows = []
        for state in game_states:
            p = state.get(player_id)
            rows.append([
                p.get("player_path_dist_opponent") / norm,
                p.get("player_dist_opponent") / norm,
                p.get("player_pos_x") / board[0],
                p.get("player_pos_y") / board[1],
                (p.get("player_rotation") % 2 * np.pi) / 2 * np.pi,        # precedence as written, :529
                p.get("projectile_cooldown") / cooldown_max,
                p.get("projectile_dist_opponent") / norm,
                p.get("projectile_pos_x") / board[0],
                p.get("projectile_pos_y") / board[1],
                (p.get("projectile_rotation") % 2 * np.pi) / 2 * np.pi,
                p.get("projectile_path_dist_opponent") / norm,
                int(p.get("projectile_future_collision_opponent")),
            ])
        return rows

    @staticmethod
    def prepare_actions(actions, player_id):
        return [a[0] for a in actions.get(player_id)]

    @staticmethod
    def prepare_rewards(rewards, player_id):
        return list(rewards.get(player_id))

    # -- rewards (SkillshotLearner.py:575-603) -------------------------------------
    def calculate_rewards_looking(self, game_states):
        size = self.game_environment.board_size[0]
        return [{pid: -s[pid]["player_path_dist_opponent"] / size for pid in self.player_ids} for s in game_states]

    def calculate_rewards_simple(self, game_states):
        out = []
        for s in game_states:
            out.append({pid: s[pid]["projectile_dist_opponent"] - s[opp]["projectile_dist_opponent"]
                        for pid, opp in zip(self.player_ids, self.player_ids[::-1])})
        return out

    def calculate_rewards(self, game_states, on_target_multiplier_reduction=0.25, loss_reward_multiplier=2,
                          base_reward_multiplier=0.75):
        """The distance-shaped reward of SkillshotLearner.py:605-661, behaviour kept as written:
        reward = (opponent projectile's distance to me - multiplier * my projectile's distance to the
        opponent) / max_dist, multiplier 0.75, 0.5 while my projectile is on target, 2.75 for the player
        that was not `game_winner`; the `game_winner` player's entry at its projectile's firing tick is
        overwritten with 1.  The reference reads "projectile_cooldown" / "projectile_age" from the
        OUTER state dict where they do not exist, so its best-distance bonus is always 0 (:643-648)."""
        p1, p2 = self.player_ids
        rewards = []
        for index, state in enumerate(game_states):
            dist = {p1: state[p1]["projectile_dist_opponent"], p2: state[p2]["projectile_dist_opponent"]}
            loser_id = 0
            winner_id = state.get("game_winner")
            if winner_id != 0:
                rewards[index - state[winner_id]["projectile_age"]][winner_id] = 1    # Python indexing, as written
                loser_id = p2 if winner_id == p1 else p1
            entry = {}
            for pid, opp in ((p1, p2), (p2, p1)):
                multi = base_reward_multiplier
                if state[pid]["projectile_future_collision_opponent"]:
                    multi = base_reward_multiplier - on_target_multiplier_reduction
                if pid == loser_id:
                    multi = base_reward_multiplier + loss_reward_multiplier
                entry[pid] = ((dist[opp] - dist[pid] * multi) + 0 * 2) / self.max_dist_normaliser
            rewards.append(entry)
        return rewards

    # -- fitting (SkillshotLearner.py:386-443) --------------------------------------
    def model_actor_fit_step(self, state_tensor):
        """One deterministic-policy-gradient step of the actor on a batch of states."""
        self.networks.actor_step(np.asarray(state_tensor, dtype=np.float32))

    def models_fit(self, states, actions, rewards):
        assert len(states) == len(actions) == len(rewards)
        states = np.asarray(states, np.float32)
        actions = np.asarray(actions, np.float32).reshape(len(states), -1)
        rewards = np.asarray(rewards, np.float32)
        assert states.shape[0] == actions.shape[0] == rewards.shape[0]
        indices = np.arange(states.shape[0])
        self._rng.shuffle(indices)                                   # SkillshotLearner.py:426-431
        dev = self.networks.device
        s, a, r = (torch.from_numpy(x[indices]).to(dev) for x in (states, actions, rewards))
        bs = self.model_param_batch_size
        # critic.fit: one epoch, Keras reshuffles, batches of 16 with the short last batch kept
        order = torch.from_numpy(self._rng.permutation(len(indices))).to(dev)
        sse = torch.zeros((), device=dev)
        for b in range(0, len(indices), bs):
            idx = order[b:b + bs]
            sse += self.networks.critic_step(s[idx], a[idx], r[idx])
        # then the actor, consecutive batches of the shuffled states (SkillshotLearner.py:440-443)
        qsum = torch.zeros((), device=dev)
        for b in range(0, len(indices), bs):
            qsum += self.networks.actor_step(s[b:b + bs])
        self.last_fit = dict(critic_loss=float(sse) / len(indices), mean_q=float(qsum) / len(indices))

    # -- training loop (SkillshotLearner.py:283-384) -----------------------------------
    def model_train(self, epochs, save_progress=False, save_boards=False):
        total = dict(epoch_ticks=[], epoch_winner=[], epoch_board_sequences=[])
        game = self.game_environment
        for epoch in range(epochs):
            game.game_reset(random_positions=self.use_random_start)
            states, boards = [game.get_state()], []
            actions = {pid: [] for pid in self.player_ids}
            while game.game_live and game.ticks < self.model_param_game_tick_limit:
                game_state = states[-1]
                for pid in self.player_ids:      # both act from the same pre-tick state
                    actions[pid].append(self.model_act_param_noise(game_state, pid))
                game.game_tick()
                states.append(game.get_state())
                if save_boards:
                    boards.append(game.get_board())
            print("Begin Fitting for Epoch:", epoch)
            rewards_per_state = self.calculate_rewards_looking(states[1:])   # reward of action t = state t+1
            rewards = {pid: [r[pid] for r in rewards_per_state] for pid in self.player_ids}
            ts, ta, tr = [], [], []
            for pid in self.player_ids:                                      # P1 rows then P2 rows
                ts += self.prepare_states(states[:-1], pid)
                ta += self.prepare_actions(actions, pid)
                tr += self.prepare_rewards(rewards, pid)
            ts, ta, tr = np.array(ts), np.array(ta), np.array(tr)
            assert ts.shape == ((len(states) - 1) * 2, self.dim_state_space)
            assert ta.shape == (ts.shape[0], self.dim_action_space) and tr.shape == (ts.shape[0],)
            self.models_fit(ts, ta, tr)
            total["epoch_ticks"].append(game.ticks)
            total["epoch_winner"].append(game.winner_id)
            if save_boards:
                total["epoch_board_sequences"].append(boards)
            print("Epoch {} Completed, ticks taken: {}, game winner: {}".format(epoch, game.ticks, game.winner_id))
        print("All Epochs Completed")
        if save_progress:                                    # SkillshotLearner.py:379-384
            self.save_actor_critic_models(epochs)
            self.save_training_progress(total)
        if save_boards:
            self.save_training_boards(total["epoch_board_sequences"])
        self.training_progress = total
        return total

    # -- persistence (SkillshotLearner.py:123-162, naming kept, bugs not) ------------------
    def save_actor_critic_models(self, epochs):
        """training_models/{actor,critic}/{start}_{end}_model.npz (the reference writes .h5 through
        Keras; same directory layout and epoch-range naming, arrays in get_weights() order), and next to them
        training_models/learner_state/{start}_{end}_state.pt: targets, Adam moments and step counters of THAT range."""
        name = None
        for which, dir_name in (("actor", self.actor_dir_name), ("critic", self.critic_dir_name)):
            d = os.path.join(self.save_location, dir_name)
            os.makedirs(d, exist_ok=True)
            ends = [int(f.split("_")[1]) for f in os.listdir(d) if f.endswith("_model.npz")]
            start = max(ends) + 1 if ends else 0                  # SkillshotLearner.py:153-157
            name = "%d_%d" % (start, start + epochs)
            np.savez(os.path.join(d, name + "_model.npz"), *self.networks.get_weights(which))
        d = os.path.join(self.save_location, "learner_state")
        os.makedirs(d, exist_ok=True)
        torch.save(self.networks.state_dict(), os.path.join(d, name + "_state.pt"))
        print("Actor and Critic Saved.")

    def save_training_progress(self, total_progress):
        """training_models/training_progress/training_progress.csv, appended (SkillshotLearner.py:164-173):
        one row per epoch with its tick count and winner (the board rasters go to save_training_boards)."""
        import pandas as pd
        d = os.path.join(self.save_location, self.training_progress_dir_name)
        os.makedirs(d, exist_ok=True)
        frame = pd.DataFrame(dict(epoch_ticks=total_progress["epoch_ticks"], epoch_winner=total_progress["epoch_winner"]))
        path = os.path.join(d, "training_progress.csv")
        frame.to_csv(path, mode="a", header=not os.path.exists(path))
        print("Training Progress Saved")

    def load_training_progress(self):
        import pandas as pd
        return pd.read_csv(os.path.join(self.save_location, self.training_progress_dir_name, "training_progress.csv"))

    def save_training_boards(self, epoch_board_list):
        """training_models/training_boards/training_boards.npy: per epoch the list of 250 x 250 rasters
        (SkillshotLearner.py:182-193), the input of SkillshotGameDisplay.display_sequence."""
        d = os.path.join(self.save_location, self.training_boards_dir_name)
        os.makedirs(d, exist_ok=True)
        arr = np.empty(len(epoch_board_list), dtype=object)          # epochs differ in length: ragged
        for k, boards in enumerate(epoch_board_list):
            arr[k] = np.asarray(boards)
        np.save(os.path.join(d, "training_boards"), arr, allow_pickle=True)
        print("Training Boards Saved")

    def load_training_boards(self):
        return np.load(os.path.join(self.save_location, self.training_boards_dir_name, "training_boards.npy"), allow_pickle=True)

    def load_actor_critic_models(self, load_index=-1):
        """Loads the `load_index`-th saved actor and critic (sorted by their end epoch, SkillshotLearner.py:123-137) and, when
        it exists, the optimiser state saved WITH THAT epoch range.  Returns True, or False when a directory holds no model
        (as the reference does; its assignment to a loop variable, which loaded nothing, is not reproduced)."""
        flat, names = {}, {}
        for which, dir_name in (("actor", self.actor_dir_name), ("critic", self.critic_dir_name)):
            d = os.path.join(self.save_location, dir_name)
            files = sorted((f for f in os.listdir(d) if f.endswith("_model.npz")), key=lambda x: int(x.split("_")[1])) if os.path.isdir(d) else []
            if len(files) == 0:
                print("Failed to load: ", d)
                return False
            try:
                chosen = files[load_index]
            except IndexError:
                print("Failed to load: ", d)
                return False
            z = np.load(os.path.join(d, chosen))
            flat[which] = np.concatenate([z[k].ravel() for k in z.files])
            names[which] = chosen[:-len("_model.npz")]
        self.networks.set_weights(flat["actor"], flat["critic"])
        state = os.path.join(self.save_location, "learner_state", names["actor"] + "_state.pt")
        if names["actor"] == names["critic"] and os.path.exists(state):
            sd = torch.load(state, map_location="cpu", weights_only=False)
            # the weights of the chosen files stay; the optimiser moments, targets and counters of the same range come back
            for k in ("target", "adam_m", "adam_v"):
                getattr(self.networks, k).copy_(sd[k].to(self.networks.device))
            self.networks.step_actor, self.networks.step_critic = int(sd["step_actor"]), int(sd["step_critic"])
            self.networks.counter = int(sd["counter"])
        return True


class FrameStackActor:
    """The "planning" actor of readme.md:18-20 (BASELINE.json configs[4]; no reference code, parity unpinned): the
    actor of model_define_actor with a first layer that reads the last `frames` observations of its player,
    12 * frames -> 256 relu -> 128 relu -> 2 tanh.  frames = 1 is the reference actor.  precision "f32": the exact
    float32 kernels; "bf16": the tensor-core kernels (ss_frames_tc.cu).  Exploration by parameter noise with one
    perturbed parameter vector per noise group.
    """

    def __init__(self, n_rows: int, frames: int = 20, device="cuda", seed: int = 0, precision: str = "f32",
                 params: Optional[torch.Tensor] = None, learner_stack: bool = False):
        """params: the actor's flat parameter vector to act with (e.g. ActorCritic(frames=...).actor, so that the learner's
        updates are what acts), default a fresh RandomNormal(0, 0.05) vector.  learner_stack: with the tensor-core path
        (whose history is fp16 operand tiles) also keep the float32 history the learner's dense input rows are cut from."""
        if precision not in ("f32", "bf16"):
            raise ValueError("precision must be 'f32' (exact path) or 'bf16' (tensor cores)")
        self.precision = precision
        self._ws = None
        if not torch.cuda.is_available():
            raise RuntimeError("skillshot_learning_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device)
        self.n_rows, self.frames, self.seed = int(n_rows), int(frames), int(seed)
        self.n_params = int(lib.ss_actor_frames_params(self.frames))
        if self.n_params < 0 or self.frames > 20:
            raise ValueError("frames must be 1..20")
        g = torch.Generator(device="cpu").manual_seed(self.seed)
        shapes = [(12 * self.frames, 256), (256,), (256, 128), (128,), (128, 2), (2,)]
        parts = [torch.randn(sh, generator=g) * 0.05 if len(sh) == 2 else torch.zeros(sh) for sh in shapes]   # RandomNormal(0, 0.05)
        self.shapes = shapes
        if params is not None:
            if params.numel() != self.n_params or params.dtype != torch.float32 or not params.is_contiguous():
                raise ValueError("params must be a contiguous float32 vector of %d values" % self.n_params)
            self.params = params
        else:
            self.params = torch.cat([p.reshape(-1) for p in parts]).to(self.device)
        self.stack32 = None
        if precision == "bf16" and learner_stack:
            self.stack32 = torch.zeros((self.n_rows, self.frames, 12), dtype=torch.float32, device=self.device)
        self._ordered = None
        if precision == "bf16":
            # the tensor-core path keeps the history as fp16 tiles in the MMA operand layout (ss_obs_stack_push_tc)
            self.stack = torch.zeros(int(lib.ss_obs_stack_tc_bytes(self.n_rows, self.frames)), dtype=torch.uint8, device=self.device)
        else:
            self.stack = torch.zeros((self.n_rows, self.frames, 12), dtype=torch.float32, device=self.device)
        self.head = -1                 # slot of the newest frame = head % frames
        self.counter = 0
        self._noisy = None

    def push(self, obs, done=None, done_div: int = 2):
        """Append the newest observation [n_rows,12]; rows whose game restarted (done, one flag per done_div rows)
        are refilled with it.  The first push fills every slot."""
        obs = _f32(obs, self.device).reshape(self.n_rows, 12)
        first = self.head < 0
        self.head += 1
        if first:
            done, done_div = torch.ones(self.n_rows, dtype=torch.uint8, device=self.device), 1
        elif done is not None:
            done = done.to(device=self.device, dtype=torch.uint8).contiguous()
        fn = lib.ss_obs_stack_push_tc if self.precision == "bf16" else lib.ss_obs_stack_push
        with torch.cuda.device(self.device):
            check(fn(self.stack.data_ptr(), self.n_rows, self.frames, self.head, obs.data_ptr(), _ptr(done), int(done_div),
                     _stream(self.device)), "ss_obs_stack_push")
            if self.stack32 is not None:
                check(lib.ss_obs_stack_push(self.stack32.data_ptr(), self.n_rows, self.frames, self.head, obs.data_ptr(), _ptr(done),
                                            int(done_div), _stream(self.device)), "ss_obs_stack_push")

    def ordered_rows(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The network input of every row as dense float32 rows [n_rows, 12 * frames], oldest frame first: what the learner
        of the stacked networks stores and trains on (ss_obs_stack_ordered)."""
        src = self.stack if self.precision == "f32" else self.stack32
        if src is None:
            raise ValueError("tensor-core history only: construct with learner_stack=True")
        if out is None:
            out = torch.empty((self.n_rows, 12 * self.frames), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.ss_obs_stack_ordered(src.data_ptr(), self.n_rows, self.frames, self.head, out.data_ptr(), _stream(self.device)),
                  "ss_obs_stack_ordered")
        return out

    def forward(self, param_noise_sd: float = 0.0, noise_group: int = 0, out: Optional[torch.Tensor] = None):
        """actions [n_rows,2] from the current stack; param_noise_sd > 0: rows i share the perturbed parameters of
        noise group i // noise_group (fresh draw per call)."""
        if self.head < 0:
            raise ValueError("push an observation first")
        if out is None:
            out = torch.empty((self.n_rows, 2), dtype=torch.float32, device=self.device)
        params, stride = self.params, 0
        with torch.cuda.device(self.device):
            if param_noise_sd > 0:
                if noise_group <= 0:
                    raise ValueError("noise_group must be positive")
                n_groups = (self.n_rows + noise_group - 1) // noise_group
                stride = (self.n_params + 3) // 4 * 4
                if self._noisy is None or self._noisy.numel() < n_groups * stride:
                    self._noisy = torch.empty(n_groups * stride, dtype=torch.float32, device=self.device)
                check(lib.ss_param_noise_groups(self.params.data_ptr(), self._noisy.data_ptr(), self.n_params, n_groups, stride,
                                                float(param_noise_sd), self.seed, self.counter, _stream(self.device)),
                      "ss_param_noise_groups")
                self.counter += 1
                params = self._noisy
            if self.precision == "bf16":
                need = int(lib.ss_actor_frames_tc_workspace_bytes(self.n_rows, int(noise_group) if stride else 0))
                if self._ws is None or self._ws.numel() < need:
                    self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
                check(lib.ss_actor_forward_frames_tc(params.data_ptr(), stride, int(noise_group), self.stack.data_ptr(), self.frames,
                                                     self.head, out.data_ptr(), self.n_rows, self._ws.data_ptr(), self._ws.numel(),
                                                     _stream(self.device)), "ss_actor_forward_frames_tc")
            else:
                check(lib.ss_actor_forward_frames(params.data_ptr(), stride, int(noise_group), self.stack.data_ptr(), self.frames,
                                                  self.head, out.data_ptr(), self.n_rows, _stream(self.device)),
                      "ss_actor_forward_frames")
        return out

    def ordered_stack(self) -> torch.Tensor:
        """[n_rows, frames * 12] network input, oldest frame first (introspection / tests)."""
        order = [(self.head + 1 + f) % self.frames for f in range(self.frames)]
        stack = self.stack
        if self.precision == "bf16":      # [tile][K / 8][128][8] fp16 -> [row][slot][12]
            tiles = (self.n_rows + 127) // 128
            x = stack.view(torch.float16).view(tiles, -1, 128, 8).permute(0, 2, 1, 3).reshape(tiles * 128, -1)
            stack = x[:self.n_rows, :self.frames * 12].float().reshape(self.n_rows, self.frames, 12)
        return stack[:, order, :].reshape(self.n_rows, self.frames * 12)


class SelfPlayTrainer:
    """Batched self-play: E games stepped together, both players driven by the shared actor
    with parameter-noise exploration, transitions kept in a device replay ring, critic and
    actor updated on sampled minibatches.  One rank per GPU; with a process group the only
    exchange is the all-reduce of the flat gradients (envs and replay shard naturally).
    """

    def __init__(self, n_envs: int, device="cuda", seed: int = 0, replay_capacity: Optional[int] = None,
                 batch_size: int = 4096, gamma: float = 0.0, tau: float = 1.0, param_noise_sd: float = 0.5,
                 noise_group: int = 128, reward_mode: str = "looking", tick_limit: int = 2000,
                 random_positions: bool = True, process_group=None, precision: str = "f32", collective: str = "nccl",
                 frames: int = 1):
        """frames > 1: the frame-stacked "planning" networks (readme.md:18-20, BASELINE.json configs[4]): both players act
        from their last `frames` observations (FrameStackActor; `precision` selects its float32 or tensor-core kernels),
        the replay ring stores the stacked observations, and the critic / actor update of SkillshotLearner.py:386-443 runs
        on the 12 * frames-input networks (float32 kernels, separate steps)."""
        self.device = torch.device(device)
        self.frames = int(frames)
        self.envs = SkillshotEnvs(n_envs, device=device, random_positions=random_positions, seed=seed,
                                  reward_mode=reward_mode, tick_limit=tick_limit, auto_reset=True)
        self.envs.collect_episode_stats = True     # the reference's per-episode ticks / winner log, reduced on the device
        self.networks = ActorCritic(device=device, seed=seed if process_group is None else 0, gamma=gamma, tau=tau,
                                    process_group=process_group, update_precision=precision if self.frames == 1 else "f32",
                                    collective=collective, frames=self.frames)
        self.networks.seed = seed          # exploration / dropout streams differ per rank, weights do not
        self.replay = ReplayRing(replay_capacity or 2 * n_envs * 8, device=device, seed=seed, frames=self.frames)
        self.batch_size, self.param_noise_sd, self.noise_group = int(batch_size), float(param_noise_sd), int(noise_group)
        self.precision = precision
        n = n_envs
        self.obs = self.envs.observe().contiguous()                       # [n,2,12] seen by the actor next
        self.prev_obs = torch.empty_like(self.obs)
        self.actions = torch.empty((n, 2, 2), dtype=torch.float32, device=self.device)
        self.stack = None
        if self.frames > 1:
            # the acting network IS the learner's actor vector; the history lives in the FrameStackActor
            self.stack = FrameStackActor(2 * n, frames=self.frames, device=device, seed=seed, precision=precision,
                                         params=self.networks.actor, learner_stack=True)
            self.stack.push(self.obs.reshape(-1, 12))
            self.rows = self.stack.ordered_rows()                         # [2n, 12 frames]: the stacked observation now
            self.prev_rows = torch.empty_like(self.rows)
        self._batch = None
        self._batches = [None, None]        # update(): two sets of minibatch buffers used in turn
        self._ring_clean = False            # no rollout (nothing of this object's that writes the ring) since the last update
        self.ticks = 0
        self.updates = 0
        self.exchange_check_every = 256     # updates between looks at the peer exchange's status word
        self._tile_ready, self._stream2 = None, None      # overlapped rollout (rollout()): allocated on first use

    def rollout_tick(self, store: bool = True):
        """One tick of every env: actor forward on both players' observations (fresh parameter
        noise per noise group), env step, transition push."""
        self._ring_clean = False
        if self.frames > 1:
            return self._rollout_tick_frames(store)
        self.prev_obs, self.obs = self.obs, self.prev_obs
        self.networks.actor_forward(self.prev_obs, param_noise_sd=self.param_noise_sd, noise_group=self.noise_group,
                                    out=self.actions.view(-1, 2), precision=self.precision)
        out = self.envs.step(self.actions, obs_out=self.obs)
        if store:
            # the ring's flag masks the TD bootstrap: only a hit (winner != 0) is a termination, a tick-limit restart is not
            self.replay.push(self.prev_obs, self.actions, out["reward"], self.obs, out["winner"], done_div=2)
        self.ticks += 1
        return out

    def _rollout_tick_frames(self, store: bool = True):
        """The same tick for the frame-stacked networks: the actor reads each player's last `frames` observations; the stored
        transition is (stacked obs, action, reward, stacked obs after the tick, hit flag).  A restarted game's history is
        refilled with its first observation (FrameStackActor.push)."""
        st = self.stack
        st.forward(param_noise_sd=self.param_noise_sd, noise_group=self.noise_group if self.param_noise_sd > 0 else 0,
                   out=self.actions.view(-1, 2))
        out = self.envs.step(self.actions, obs_out=self.obs)
        st.push(self.obs.reshape(-1, 12), out["done"], done_div=2)
        self.prev_rows, self.rows = self.rows, self.prev_rows
        st.ordered_rows(out=self.rows)
        if store:
            self.replay.push(self.prev_rows, self.actions, out["reward"], self.rows, out["winner"], done_div=2)
        self.ticks += 1
        return out

    def rollout(self, n_ticks: int, store: bool = True):
        """n_ticks rollout ticks enqueued by ONE library call (ss_selfplay_rollout): same kernels and Philox
        counters as n_ticks calls of rollout_tick, without the per-tick host work."""
        self._ring_clean = False
        if self.frames > 1:
            for _ in range(int(n_ticks)):
                out = self._rollout_tick_frames(store)
            return dict(obs=self.obs, reward=out["reward"], done=out["done"], winner=out["winner"])
        envs, net, rp = self.envs, self.networks, self.replay
        if not envs.auto_reset:
            raise ValueError("the batched rollout needs auto_reset envs")
        if store and 2 * envs.n_envs > rp.capacity:
            raise ValueError("replay ring smaller than one tick of transitions")
        out = envs._buffers(1)
        # the current observation is in self.obs; it becomes buffer A of the call
        a, b = self.obs, self.prev_obs
        # SS_ROLLOUT_OVERLAP=1 (experiment, off by default): the env step runs beside the forward kernel on a second stream
        # (ss_selfplay_rollout2), consuming its actions tile by tile.  Bit-exact, and it works when both grids fit the GPU at
        # once; at 262,144 envs the forward kernel's CTAs fill their SMs (223 KB shared memory, 3/4 of the registers), the env
        # CTAs are not co-resident and their bounded waits expire (SS_STATUS_ROLLOUT_TIMEOUT) -- DESIGN.md section 6.5
        overlap = self.precision == "bf16" and os.environ.get("SS_ROLLOUT_OVERLAP", "0") == "1"
        if overlap and self._tile_ready is None:
            tiles = (2 * envs.n_envs + 127) // 128 + max(self.noise_group, 128) // 128 + 1
            self._tile_ready = torch.zeros(tiles, dtype=torch.int32, device=self.device)
            self._stream2 = torch.cuda.Stream(self.device, priority=-1)      # HIGHER priority than the current stream: the forward kernels
        with torch.cuda.device(self.device):
            check(lib.ss_selfplay_rollout2(
                envs.state.data_ptr(), envs.n_envs, net.actor.data_ptr(), a.data_ptr(), b.data_ptr(), self.actions.data_ptr(),
                out["reward"].data_ptr(), out["done"].data_ptr(), out["winner"].data_ptr(),
                rp.obs.data_ptr() if store else None, rp.act.data_ptr(), rp.reward.data_ptr(), rp.next_obs.data_ptr(),
                rp.done.data_ptr(), rp.capacity, rp.pos, int(n_ticks), self.param_noise_sd, self.noise_group, 0.0,
                1 if self.precision == "bf16" else 0, REWARD_MODES[envs.reward_mode], envs.tick_limit,
                _lib.RESET_RANDOM if envs.random_positions else _lib.RESET_FIXED, envs.seed, envs.counter, net.seed,
                net.counter, _ptr(envs.speeds), envs.status.data_ptr(), _lib.STEP_EPISODE_STATS if envs.collect_episode_stats else 0,
                self._tile_ready.data_ptr() if overlap else None, self._stream2.cuda_stream if overlap else None,
                _stream(self.device)), "ss_selfplay_rollout")
        envs.counter += n_ticks
        net.counter += n_ticks
        if store:
            rp.pos = (rp.pos + 2 * envs.n_envs * n_ticks) % rp.capacity
            rp.size = min(rp.capacity, rp.size + 2 * envs.n_envs * n_ticks)
        rows = 2 * envs.n_envs
        in_place = (store and rp.capacity % rows == 0 and (rp.pos - rows * n_ticks) % rows == 0
                    and (rp.capacity // rows >= 2 or n_ticks == 1))       # the library's own condition (ss_selfplay_rollout)
        if (n_ticks & 1) and not in_place:        # (in place, the library leaves the current observation in buffer A)
            self.obs, self.prev_obs = b, a
        self.ticks += n_ticks
        return dict(obs=self.obs, reward=out["reward"][0], done=out["done"][0], winner=out["winner"][0])

    # -- checkpoint / resume (SkillshotLearner.py:123-162 keeps the two Keras models; a batched run also needs the
    #    optimiser moments, the targets, the games in flight, the replay ring and every Philox counter) ---------------
    def state_dict(self):
        sd = dict(envs=self.envs.state_dict(), networks=self.networks.state_dict(), replay=self.replay.state_dict(),
                  obs=self.obs.cpu(), actions=self.actions.cpu(), ticks=self.ticks, batch_size=self.batch_size,
                  param_noise_sd=self.param_noise_sd, noise_group=self.noise_group, precision=self.precision, frames=self.frames)
        if self.stack is not None:      # the players' observation histories and the noise counter of the stacked actor
            st = self.stack
            sd["stack"] = dict(stack=st.stack.cpu(), stack32=None if st.stack32 is None else st.stack32.cpu(), head=st.head,
                               counter=st.counter, rows=self.rows.cpu())
        return sd

    def load_state_dict(self, sd):
        self.envs.load_state_dict(sd["envs"])
        self.networks.load_state_dict(sd["networks"])
        self.replay.load_state_dict(sd["replay"])
        self.obs.copy_(sd["obs"].to(self.device))
        self.actions.copy_(sd["actions"].to(self.device))
        self.ticks, self.batch_size = int(sd["ticks"]), int(sd["batch_size"])
        self.param_noise_sd, self.noise_group = float(sd["param_noise_sd"]), int(sd["noise_group"])
        if self.stack is not None:
            st, ss = self.stack, sd["stack"]
            st.stack.copy_(ss["stack"].to(self.device))
            if st.stack32 is not None:
                st.stack32.copy_(ss["stack32"].to(self.device))
            st.head, st.counter = int(ss["head"]), int(ss["counter"])
            self.rows.copy_(ss["rows"].to(self.device))
        self._batch = None
        self._batches = [None, None]
        self._ring_clean = False

    def record_boards(self, env: int, n_ticks: int, path: Optional[str] = None, store: bool = True):
        """Play n_ticks rollout ticks and return the 250 x 250 rasters of game `env` after each of them
        (SkillshotGame.get_board, SkillshotGame.py:36-56, rebuilt on the host from the exported state row): the frames
        SkillshotGameDisplay.display_sequence shows.  With `path`, also written as training_boards.npy in the shape
        save_training_boards uses (one object-array entry = one sequence, SkillshotLearner.py:182-204)."""
        from .game import render_board
        boards = []
        for _ in range(int(n_ticks)):
            self.rollout_tick(store=store)
            boards.append(render_board(self.envs.export_state(int(env), 1), 0))
        boards = np.asarray(boards)
        if path is not None:
            os.makedirs(path, exist_ok=True)
            arr = np.empty(1, dtype=object)
            arr[0] = boards
            np.save(os.path.join(path, "training_boards"), arr, allow_pickle=True)
        return boards

    def progress(self, reset: bool = False) -> dict:
        """The training-progress log of the reference (per-episode ticks and winner, SkillshotLearner.py:164-180,
        365-366) for the games finished so far, reduced on the device by the step kernel (SkillshotEnvs.episode_summary)."""
        return self.envs.episode_summary(reset=reset)

    def check_exchange(self):
        """Raises if a rank's gradient failed to arrive in the fused peer exchange (the Adam kernels skip their writes from
        that point on, so the ranks still hold identical weights: reload or stop)."""
        if self.networks.peer is not None:
            self.networks.peer.check_status()

    def save(self, path: str):
        """Write a checkpoint the run can be resumed from bit-identically (one file per rank when sharded)."""
        self.check_exchange()
        torch.cuda.synchronize(self.device)
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        torch.save(self.state_dict(), path)

    def load(self, path: str):
        self.load_state_dict(torch.load(path, map_location="cpu", weights_only=False))

    def update(self):
        """One critic step and one actor step on a sampled minibatch: a single library call on one GPU or
        with the fused peer exchange, the separate steps (update_stepwise) around an NCCL all-reduce."""
        net = self.networks
        if (net.group is not None and net.peer is None) or self.frames > 1:
            return self.update_stepwise()
        # two sets of minibatch buffers in turn: when no rollout has touched the ring since the previous update, this
        # update's rows are drawn while that update's last kernels still run (ss_ddpg_update_args.sample_early)
        k = self.updates & 1
        # (measured, profiles/r2_update_sample_early_ab.txt: 158.1 -> 152.1 us at 65,536 rows; at 524,288 rows the gather
        #  competes with what it runs beside, 759 -> 764 us: small minibatches only)
        early = (self._ring_clean and self._batches[k] is not None and self.batch_size <= 131072 and
                 os.environ.get("SS_UPDATE_SAMPLE_EARLY", "1") != "0")
        self._batch = self._batches[k] = net.update_from_ring(self.replay, self.batch_size, out=self._batches[k], sample_early=early)
        self._ring_clean = True
        self.updates += 1
        if net.peer is not None and self.updates % self.exchange_check_every == 0:
            net.peer.check_status()         # one device-to-host read every few hundred updates
        return net.stats[0], net.stats[1]

    def update_stepwise(self):
        """The same update as one library call per step (sample, target, critic step, actor step)."""
        self._ring_clean = False
        self._batch = b = self.replay.sample(self.batch_size, out=self._batch)
        net = self.networks
        y = net.td_targets(b["reward"], b["next_obs"], b["done"])
        sse = net.critic_step(b["obs"], b["act"], y)
        q = net.actor_step(b["obs"])
        return sse, q
