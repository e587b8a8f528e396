"""Skillshot environments on the GPU.

`SkillshotEnvs` holds N independent games as a structure of arrays in HBM and
steps them with the fused sm_100a kernel behind `ss_env_step` (C ABI:
include/skillshot_b200.h).  `SkillshotGame`, `Player` and `Projectile` keep the
reference's object surface (SkillshotGame.py, Player.py, Projectile.py) as views
of one env, so `skillshot_playable.py`-style callers and
`SkillshotLearner.do_actions` keep working; rendering (`get_board`) stays on the
host.

There is no CPU path: every method ends in a kernel launch on a CUDA device.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import lib, check

REWARD_MODES = {"none": _lib.REWARD_NONE, "looking": _lib.REWARD_LOOKING,
                "terminal": _lib.REWARD_TERMINAL, "simple": _lib.REWARD_SIMPLE}

FEATURE_KEYS = [  # SkillshotGame.get_state key order, SkillshotGame.py:145-162
    "player_grad", "player_x_dir", "player_path_dist_opponent", "player_dist_opponent",
    "player_pos_x", "player_pos_y", "player_rotation", "projectile_cooldown",
    "projectile_grad", "projectile_x_dir", "projectile_path_dist_opponent",
    "projectile_pos_x", "projectile_pos_y", "projectile_rotation", "projectile_age",
    "projectile_valid", "projectile_dist_opponent", "projectile_future_collision_opponent",
]
_INT_KEYS = {"player_x_dir", "player_pos_x", "player_pos_y", "projectile_cooldown", "projectile_x_dir",
             "projectile_pos_x", "projectile_pos_y", "projectile_age"}
_BOOL_KEYS = {"projectile_valid", "projectile_future_collision_opponent"}

# column order of ss_env_export's int block
_EXPORT_COLS = ["px1", "px2", "py1", "py2", "qx1", "qx2", "qy1", "qy2", "cd1", "cd2",
                "age1", "age2", "valid1", "valid2", "ticks", "live", "winner"]


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class SkillshotEnvs:
    """N independent SkillshotGame instances stepped together on one GPU.

    One call to :meth:`step` is one iteration of the reference's rollout loop
    for every env (SkillshotLearner.py:304-315): both players act from the
    pre-tick state, the game ticks, and reward / done / winner / observation are
    produced from the post-tick state.
    """

    def __init__(self, n_envs: int, device="cuda", random_positions: bool = False, seed: int = 0,
                 reward_mode: str = "looking", tick_limit: int = 0, auto_reset: bool = False):
        if not torch.cuda.is_available():
            raise RuntimeError("skillshot_learning_b200 needs a CUDA device (no CPU fallback)")
        self.n_envs = int(n_envs)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("device must be a CUDA device")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.random_positions = bool(random_positions)
        self.seed = int(seed)
        self.reward_mode = reward_mode
        self.tick_limit = int(tick_limit)
        self.auto_reset = bool(auto_reset)
        self.counter = 0           # Philox counter base: advances with every tick / reset
        n = self.n_envs
        self.state = torch.zeros(_lib.STATE_BYTES_PER_ENV * n, dtype=torch.uint8, device=self.device)
        # {uint32 status, pad, uint64 episode statistics[72]}: the block SS_STEP_EPISODE_STATS describes
        self._status_block = torch.zeros(2 + 2 * _lib.EPISODE_STATS, dtype=torch.int32, device=self.device)
        self.status = self._status_block[0:1]
        self.episode_stats = self._status_block[2:].view(torch.int64)
        self.collect_episode_stats = False     # count every finished game on the device (episode_summary); opt-in
        self.speeds: Optional[torch.Tensor] = None
        self._obs = torch.empty((n, 2, _lib.NUM_OBS), dtype=torch.float32, device=self.device)
        self._out = {}
        self.reset()

    # -- configuration -----------------------------------------------------
    def set_speeds(self, speed_move, speed_look, proj_speed, cooldown_max):
        """Per-env game-speed constants (Player.py:14-15, Projectile.py:9-10; readme.md:22-23)."""
        n = self.n_envs
        buf = torch.empty(32 * n, dtype=torch.uint8, device=self.device)
        f = buf[:16 * n].view(torch.float64).view(n, 2)
        f[:, 0] = torch.as_tensor(speed_move, dtype=torch.float64, device=self.device)
        f[:, 1] = torch.as_tensor(speed_look, dtype=torch.float64, device=self.device)
        buf[16 * n:].view(torch.float64).view(n, 2)[:, 0] = torch.as_tensor(proj_speed, dtype=torch.float64, device=self.device)
        buf[16 * n:].view(torch.int64).view(n, 2)[:, 1] = torch.as_tensor(cooldown_max, dtype=torch.int64, device=self.device)
        self.speeds = buf

    # -- checkpoint --------------------------------------------------------
    def state_dict(self):
        """Everything a resumed run needs to continue bit-identically: the packed env state, the Philox counter, the
        per-env speed constants and the configuration."""
        return dict(n_envs=self.n_envs, state=self.state.cpu(), counter=self.counter, seed=self.seed,
                    episode_stats=self.episode_stats.cpu(),
                    speeds=None if self.speeds is None else self.speeds.cpu(), random_positions=self.random_positions,
                    reward_mode=self.reward_mode, tick_limit=self.tick_limit, auto_reset=self.auto_reset)

    def load_state_dict(self, sd):
        if int(sd["n_envs"]) != self.n_envs:
            raise ValueError("checkpoint holds %d envs, this batch %d" % (sd["n_envs"], self.n_envs))
        self.state.copy_(sd["state"].to(self.device))
        self.counter, self.seed = int(sd["counter"]), int(sd["seed"])
        if sd.get("episode_stats") is not None:
            self.episode_stats.copy_(sd["episode_stats"].to(self.device))
        self.speeds = None if sd["speeds"] is None else sd["speeds"].to(self.device)
        self.random_positions, self.reward_mode = bool(sd["random_positions"]), sd["reward_mode"]
        self.tick_limit, self.auto_reset = int(sd["tick_limit"]), bool(sd["auto_reset"])

    # -- reset -------------------------------------------------------------
    def reset(self, mask: Optional[torch.Tensor] = None, positions=None, random_positions: Optional[bool] = None):
        """SkillshotGame.game_reset for the envs selected by `mask` (all by default)."""
        rnd = self.random_positions if random_positions is None else bool(random_positions)
        mode, pos = (_lib.RESET_RANDOM if rnd else _lib.RESET_FIXED), None
        if positions is not None:
            pos = torch.as_tensor(np.asarray(positions), dtype=torch.int32).reshape(self.n_envs, 4).contiguous().to(self.device)
            if int(pos.min()) < 0 or int(pos.max()) > 255:
                raise ValueError("positions must lie in 0..255")
            mode = _lib.RESET_GIVEN
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            check(lib.ss_env_reset(self.state.data_ptr(), self.n_envs, _ptr(mask), mode, _ptr(pos),
                                   self.seed, self.counter, _stream(self.device)), "ss_env_reset")
        self.counter += 1

    # -- step --------------------------------------------------------------
    def _buffers(self, n_ticks: int):
        buf = self._out.get(n_ticks)
        if buf is None:
            n = self.n_envs
            buf = dict(reward=torch.zeros((n_ticks, n, 2), dtype=torch.float32, device=self.device),
                       done=torch.zeros((n_ticks, n), dtype=torch.uint8, device=self.device),
                       winner=torch.zeros((n_ticks, n), dtype=torch.uint8, device=self.device))
            self._out = {n_ticks: buf}
        return buf

    def _launch(self, a, obs, reward, done, winner, K, flags):
        with torch.cuda.device(self.device):
            check(lib.ss_env_step(self.state.data_ptr(), self.n_envs, a.data_ptr(), _ptr(obs),
                                  _ptr(reward), _ptr(done), _ptr(winner),
                                  K, REWARD_MODES[self.reward_mode], self.tick_limit, int(self.auto_reset),
                                  _lib.RESET_RANDOM if self.random_positions else _lib.RESET_FIXED,
                                  self.seed, self.counter, _ptr(self.speeds), self.status.data_ptr(),
                                  flags | (_lib.STEP_EPISODE_STATS if self.collect_episode_stats else 0), _stream(self.device)),
                  "ss_env_step")
        self.counter += K

    def episode_summary(self, reset: bool = False) -> dict:
        """(with collect_episode_stats = True)  The per-episode log of the reference (ticks and winner of every finished
        game, SkillshotLearner.py:164-180, 365-366) reduced on the device: counts of the games that ended since the last reset of the statistics, by how
        they ended, their mean length and a 64-bin histogram of their lengths."""
        st = self.episode_stats.cpu().numpy()
        if reset:
            self.episode_stats.zero_()
        n = int(st[0])
        width = (self.tick_limit + 63) // 64 if self.tick_limit > 0 else 32
        return dict(episodes=n, player1_hit=int(st[1]), player2_hit=int(st[2]), tick_limit=int(st[3]),
                    mean_ticks=float(st[4]) / n if n else float("nan"), histogram=st[8:72].copy(), bin_ticks=width)

    def step(self, actions: torch.Tensor, want_obs: bool = True, obs_every_tick: bool = False,
             obs_out: Optional[torch.Tensor] = None):
        """actions float32 [n,2,2] (one tick) or [K,n,2,2] (K ticks fused in one launch).

        Returns dict(obs [n,2,12] | [K,n,2,12] | None, reward [K,n,2], done [K,n], winner [K,n]);
        the leading K axis is dropped for one-tick input.  The tensors are reused
        by the next call.
        """
        if not torch.is_tensor(actions):
            actions = torch.from_numpy(np.ascontiguousarray(actions, dtype=np.float32))
        single = actions.dim() == 3
        a = actions.reshape((-1, self.n_envs, 2, 2))
        if a.dtype != torch.float32 or not a.is_contiguous() or a.device != self.device:
            a = a.to(device=self.device, dtype=torch.float32).contiguous()
        K = a.shape[0]
        buf = self._buffers(K)
        obs = None
        flags = 0
        if want_obs:
            if obs_every_tick and K > 1:
                obs = torch.empty((K, self.n_envs, 2, _lib.NUM_OBS), dtype=torch.float32, device=self.device)
                flags |= _lib.STEP_OBS_EVERY_TICK
            elif obs_out is not None:      # caller-owned float32 [n,2,12] (the rollout's double buffer)
                if obs_out.dtype != torch.float32 or not obs_out.is_contiguous() or obs_out.numel() != self.n_envs * 24:
                    raise ValueError("obs_out must be a contiguous float32 [n,2,12] tensor")
                obs = obs_out
            else:
                obs = self._obs
        self._launch(a, obs, buf["reward"], buf["done"], buf["winner"], K, flags)
        if single:
            return dict(obs=obs, reward=buf["reward"][0], done=buf["done"][0], winner=buf["winner"][0])
        return dict(obs=obs, reward=buf["reward"], done=buf["done"], winner=buf["winner"])

    # -- host-buffer API (end-to-end path) -----------------------------------
    def alloc_host_outputs(self, n_ticks: int, outputs: str = "full"):
        """Pinned host buffers for step_host: reward [T,n,2] f32, done [T,n] u8, winner [T,n] u8; outputs="flags": one packed
        byte per env and tick, flags [T,n] u8 (unpack_flags)."""
        n = self.n_envs
        if outputs == "flags":
            return dict(flags=torch.empty((n_ticks, n), dtype=torch.uint8, pin_memory=True))
        return dict(reward=torch.empty((n_ticks, n, 2), dtype=torch.float32, pin_memory=True),
                    done=torch.empty((n_ticks, n), dtype=torch.uint8, pin_memory=True),
                    winner=torch.empty((n_ticks, n), dtype=torch.uint8, pin_memory=True))

    def step_host(self, host_actions: torch.Tensor, host_out: dict, ticks_per_launch: int = 32,
                  n_buffers: int = 3):
        """T ticks with HOST buffers: host_actions (pinned) float32 [T,n,2,2] in, reward / done /
        winner written to the pinned tensors of `host_out`.  The ticks are played in launches of
        `ticks_per_launch`; host->device copies, kernels and device->host copies of successive
        chunks overlap on `n_buffers` streams.  Returns after everything has landed on the host."""
        T, n = host_actions.shape[0], self.n_envs
        KF = min(ticks_per_launch, T)
        if T % KF:
            raise ValueError("T must be a multiple of ticks_per_launch")
        if "flags" in host_out:
            return self._step_host_flags(host_actions, host_out, KF, n_buffers)
        key = ("host", KF, n_buffers)
        ring = self._host_ring.get(key) if hasattr(self, "_host_ring") else None
        if ring is None:
            ring = []
            for _ in range(n_buffers):
                ring.append(dict(stream=torch.cuda.Stream(self.device),
                                 act=torch.empty((KF, n, 2, 2), dtype=torch.float32, device=self.device),
                                 reward=torch.zeros((KF, n, 2), dtype=torch.float32, device=self.device),
                                 done=torch.zeros((KF, n), dtype=torch.uint8, device=self.device),
                                 winner=torch.zeros((KF, n), dtype=torch.uint8, device=self.device)))
            self._host_ring = {key: ring}
        cur = torch.cuda.current_stream(self.device)
        prev = cur.record_event()
        for c in range(T // KF):
            b = ring[c % n_buffers]
            sl = slice(c * KF, (c + 1) * KF)
            with torch.cuda.stream(b["stream"]):
                b["act"].copy_(host_actions[sl], non_blocking=True)
                b["stream"].wait_event(prev)                 # the state is carried from chunk to chunk
                self._launch(b["act"], None, b["reward"], b["done"], b["winner"], KF, 0)
                prev = b["stream"].record_event()
                host_out["reward"][sl].copy_(b["reward"], non_blocking=True)
                host_out["done"][sl].copy_(b["done"], non_blocking=True)
                host_out["winner"][sl].copy_(b["winner"], non_blocking=True)
        for b in ring:
            b["stream"].synchronize()
        cur.wait_event(prev)
        return host_out

    def _step_host_flags(self, host_actions, host_out, KF, n_buffers):
        """step_host with the packed one-byte-per-env-step output (ss_env_step_packed): terminal reward mode, reference speed
        constants.  16 bytes per env-step go up the bus, 1 comes back."""
        if self.reward_mode != "terminal" or self.speeds is not None or self.collect_episode_stats:
            raise ValueError("packed flags: terminal reward mode, reference speeds, no episode statistics")
        T, n = host_actions.shape[0], self.n_envs
        key = ("flags", KF, n_buffers)
        ring = self._host_ring.get(key) if hasattr(self, "_host_ring") else None
        if ring is None:
            ring = [dict(stream=torch.cuda.Stream(self.device),
                         act=torch.empty((KF, n, 2, 2), dtype=torch.float32, device=self.device),
                         flags=torch.zeros((KF, n), dtype=torch.uint8, device=self.device)) for _ in range(n_buffers)]
            self._host_ring = {key: ring}
        cur = torch.cuda.current_stream(self.device)
        prev = cur.record_event()
        mode = _lib.RESET_RANDOM if self.random_positions else _lib.RESET_FIXED
        for c in range(T // KF):
            b = ring[c % n_buffers]
            sl = slice(c * KF, (c + 1) * KF)
            with torch.cuda.stream(b["stream"]), torch.cuda.device(self.device):
                b["act"].copy_(host_actions[sl], non_blocking=True)
                b["stream"].wait_event(prev)                 # the state is carried from chunk to chunk
                check(lib.ss_env_step_packed(self.state.data_ptr(), n, b["act"].data_ptr(), b["flags"].data_ptr(), KF,
                                             self.tick_limit, int(self.auto_reset), mode, self.seed, self.counter,
                                             self.status.data_ptr(), b["stream"].cuda_stream), "ss_env_step_packed")
                self.counter += KF
                prev = b["stream"].record_event()
                host_out["flags"][sl].copy_(b["flags"], non_blocking=True)
        for b in ring:
            b["stream"].synchronize()
        cur.wait_event(prev)
        return host_out

    @staticmethod
    def unpack_flags(flags):
        """Packed step outputs -> dict(done, winner, reward [...,2]) as numpy arrays (include/skillshot_b200.h,
        ss_env_step_packed): done = bit 0, winner_id = bits 1-2, bit 3 = the hit happened on this tick, which is when the
        terminal reward is paid: -1 to the player that was hit, +1 to the shooter."""
        f = flags.numpy() if torch.is_tensor(flags) else np.asarray(flags)
        done, winner, hit = f & 1, (f >> 1) & 3, (f >> 3) & 1
        reward = np.zeros(f.shape + (2,), np.float32)
        reward[..., 0] = np.where(hit == 1, np.where(winner == 1, -1.0, 1.0), 0.0)
        reward[..., 1] = np.where(hit == 1, np.where(winner == 2, -1.0, 1.0), 0.0)
        return dict(done=done.astype(np.uint8), winner=winner.astype(np.uint8), reward=reward)

    def observe(self) -> torch.Tensor:
        """float32 [n,2,12] observation of the current state (prepare_states of get_state)."""
        _, obs, _ = self.features(want_feat=False)
        return obs.to(torch.float32)

    def check_status(self):
        """Raises where the reference would have raised during the steps so far."""
        s = int(self.status.item())
        if s & _lib.STATUS_ROLLOUT_TIMEOUT:
            self.status.zero_()
            raise RuntimeError("overlapped rollout: the env step never received a tile of actions from the forward kernel")
        if s & _lib.STATUS_NAN:
            self.status.zero_()
            raise ValueError("cannot convert float NaN to integer")   # int(round(nan)), Player.py:63

    # -- features ----------------------------------------------------------
    def features(self, want_feat: bool = True, want_obs: bool = True):
        """(feat f64 [n,2,18] | None, obs f64 [n,2,12] | None, general i32 [n,3]) of the current state."""
        n = self.n_envs
        feat = torch.empty((n, 2, _lib.NUM_FEATURES), dtype=torch.float64, device=self.device) if want_feat else None
        obs = torch.empty((n, 2, _lib.NUM_OBS), dtype=torch.float64, device=self.device) if want_obs else None
        gen = torch.empty((n, 3), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.ss_env_features(self.state.data_ptr(), n, _ptr(feat), _ptr(obs), gen.data_ptr(),
                                      _ptr(self.speeds), _stream(self.device)), "ss_env_features")
        return feat, obs, gen

    # -- state import / export --------------------------------------------
    def export_state(self, first: int = 0, count: Optional[int] = None):
        """dict of numpy arrays with the reference-natural fields of envs [first, first+count)."""
        count = self.n_envs - first if count is None else count
        ints = torch.empty((count, _lib.EXPORT_INTS), dtype=torch.int32, device=self.device)
        rots = torch.empty((count, 4), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.ss_env_export(self.state.data_ptr(), self.n_envs, first, count, ints.data_ptr(),
                                    rots.data_ptr(), _stream(self.device)), "ss_env_export")
        return unpack_export(ints.cpu().numpy(), rots.cpu().numpy())

    def import_state(self, fields: dict, first: int = 0):
        ints, rots = pack_import(fields)
        count = ints.shape[0]
        ti = torch.from_numpy(ints).to(self.device)
        tr = torch.from_numpy(rots).to(self.device)
        with torch.cuda.device(self.device):
            check(lib.ss_env_import(self.state.data_ptr(), self.n_envs, first, count, ti.data_ptr(),
                                    tr.data_ptr(), _stream(self.device)), "ss_env_import")

    def apply(self, env: int, player: int, op: int, value: float = 0.0):
        """One reference method call on one env (ss_env_apply)."""
        with torch.cuda.device(self.device):
            check(lib.ss_env_apply(self.state.data_ptr(), self.n_envs, int(env), int(player), int(op),
                                   float(value), _ptr(self.speeds), self.status.data_ptr(),
                                   _stream(self.device)), "ss_env_apply")

    def game(self, env: int) -> "SkillshotGame":
        """A reference-surface view of env `env`."""
        return SkillshotGame(_envs=self, _index=env)


def unpack_export(ints: np.ndarray, rots: np.ndarray) -> dict:
    """ss_env_export blocks -> fields named like the reference attributes, [count,2] per player."""
    c = {k: ints[:, j] for j, k in enumerate(_EXPORT_COLS)}
    two = lambda a, b: np.stack([c[a], c[b]], axis=1).astype(np.int64)
    return dict(px=two("px1", "px2"), py=two("py1", "py2"), qx=two("qx1", "qx2"), qy=two("qy1", "qy2"),
                cd=two("cd1", "cd2"), age=two("age1", "age2"), valid=two("valid1", "valid2"),
                ticks=c["ticks"].astype(np.int64), live=c["live"].astype(np.int64),
                winner=c["winner"].astype(np.int64),
                prot=rots[:, 0:2].copy(), qrot=rots[:, 2:4].copy())


def pack_import(f: dict):
    n = len(f["ticks"])
    ints = np.zeros((n, _lib.EXPORT_INTS), np.int32)
    for j, (key, p) in enumerate([("px", 0), ("px", 1), ("py", 0), ("py", 1), ("qx", 0), ("qx", 1),
                                   ("qy", 0), ("qy", 1), ("cd", 0), ("cd", 1), ("age", 0), ("age", 1),
                                   ("valid", 0), ("valid", 1)]):
        ints[:, j] = np.asarray(f[key])[:, p]
    ints[:, 14], ints[:, 15], ints[:, 16] = f["ticks"], f["live"], f["winner"]
    pos = ints[:, :8]
    if pos.min() < 0 or pos.max() > 255:
        raise ValueError("positions must lie in 0..255")
    rots = np.concatenate([np.asarray(f["prot"], np.float64), np.asarray(f["qrot"], np.float64)], axis=1)
    return np.ascontiguousarray(ints), np.ascontiguousarray(rots)


# ---------------------------------------------------------------------------
# Reference object surface (one env)
# ---------------------------------------------------------------------------
class _Pos:
    """Player.pos / Projectile.pos: a 2-element mutable sequence living on the device."""

    def __init__(self, game, kx, ky, p):
        self._g, self._kx, self._ky, self._p = game, kx, ky, p

    def _get(self):
        row = self._g._row()
        return [int(row[self._kx][0, self._p]), int(row[self._ky][0, self._p])]

    def __getitem__(self, i):
        return self._get()[i]

    def __setitem__(self, i, v):
        self._g._set_field(self._kx if i == 0 else self._ky, self._p, int(v))

    def __iter__(self):
        return iter(self._get())

    def __len__(self):
        return 2

    def __repr__(self):
        return repr(self._get())

    def __eq__(self, other):
        return list(self) == list(other)


class Projectile:
    """Projectile.py surface: pos, rotation, cooldown_current, age, valid."""
    shape_image = [[1, 0, 1], [0, 1, 0], [1, 0, 1]]      # Projectile.py:5-7
    cooldown_max = 15                                    # Projectile.py:9
    speed_move = 5                                       # Projectile.py:10

    def __init__(self, game, p):
        self._g, self._p = game, p
        self.board_dim = game.board_size
        self.shape_size = (3, 3)

    pos = property(lambda s: _Pos(s._g, "qx", "qy", s._p),
                   lambda s, v: (s._g._set_field("qx", s._p, int(v[0])), s._g._set_field("qy", s._p, int(v[1]))) and None)
    rotation = property(lambda s: float(s._g._row()["qrot"][0, s._p]), lambda s, v: s._g._set_field("qrot", s._p, float(v)))
    cooldown_current = property(lambda s: int(s._g._row()["cd"][0, s._p]), lambda s, v: s._g._set_field("cd", s._p, int(v)))
    age = property(lambda s: int(s._g._row()["age"][0, s._p]), lambda s, v: s._g._set_field("age", s._p, int(v)))
    valid = property(lambda s: bool(s._g._row()["valid"][0, s._p]), lambda s, v: s._g._set_field("valid", s._p, int(bool(v))))

    def set_position(self, location):
        self.pos = location

    def set_rotation(self, rotation):
        self.rotation = rotation


class Player:
    """Player.py surface: pos, rotation, projectile and the move methods."""
    shape_image = [[0, 0, 0, 0, 0], [0, 1, 1, 1, 0], [0, 1, 1, 1, 0], [0, 1, 1, 1, 0], [0, 0, 0, 0, 0]]
    speed_move = 3          # Player.py:14
    speed_look = 0.25       # Player.py:15

    def __init__(self, game, player_id):
        self._g, self._p = game, player_id - 1
        self.id = player_id
        self.board_dim = game.board_size
        self.shape_size = (5, 5)
        self.projectile = Projectile(game, self._p)

    pos = property(lambda s: _Pos(s._g, "px", "py", s._p),
                   lambda s, v: (s._g._set_field("px", s._p, int(v[0])), s._g._set_field("py", s._p, int(v[1]))) and None)
    rotation = property(lambda s: float(s._g._row()["prot"][0, s._p]), lambda s, v: s._g._set_field("prot", s._p, float(v)))

    def _op(self, op, value=0.0):
        self._g._apply(self._p, op, value)

    def move_look_left(self):
        self._op(_lib.OP_LOOK_LEFT)

    def move_look_right(self):
        self._op(_lib.OP_LOOK_RIGHT)

    def move_look_float(self, angle):
        self._op(_lib.OP_MOVE_LOOK_FLOAT, angle)

    def move_forwards(self):
        self._op(_lib.OP_MOVE_FORWARDS)

    def move_backwards(self):
        self._op(_lib.OP_MOVE_BACKWARDS)

    def move_direction_float(self, speed):
        self._op(_lib.OP_MOVE_DIRECTION_FLOAT, speed)
        self._g._envs.check_status()      # the reference raises on NaN here (Player.py:63)

    def move_shoot_projectile(self):
        self._op(_lib.OP_SHOOT)

    def check_pos_valid(self, check_x, check_y):     # Player.py:70-76
        return (check_x + 5 <= 250 and check_x >= 0 and check_y + 5 <= 250 and check_y >= 0)


class SkillshotGame:
    """The reference SkillshotGame surface (SkillshotGame.py:8-169) over one device env."""

    def __init__(self, random_positions=False, _envs: Optional[SkillshotEnvs] = None, _index: int = 0,
                 device="cuda"):
        self.board_size = (250, 250)
        self.board = np.zeros(self.board_size, dtype=int)
        if _envs is None:
            _envs = SkillshotEnvs(1, device=device, random_positions=random_positions,
                                  seed=int(np.random.randint(0, 2 ** 31 - 1)), reward_mode="none")
        self._envs, self._i = _envs, int(_index)
        self._cache = None
        self.player1 = Player(self, 1)
        self.player2 = Player(self, 2)

    # -- device row cache ---------------------------------------------------
    def _row(self):
        if self._cache is None:
            self._cache = self._envs.export_state(self._i, 1)
        return self._cache

    def _set_field(self, key, p, value):
        row = {k: np.array(v, copy=True) for k, v in self._row().items()}
        if p is None:
            row[key][0] = value
        else:
            row[key][0, p] = value
        self._envs.import_state(row, self._i)
        self._cache = None

    def _apply(self, p, op, value=0.0):
        self._envs.apply(self._i, p, op, value)
        self._cache = None

    ticks = property(lambda s: int(s._row()["ticks"][0]), lambda s, v: s._set_field("ticks", None, int(v)))
    game_live = property(lambda s: bool(s._row()["live"][0]), lambda s, v: s._set_field("live", None, int(bool(v))))
    winner_id = property(lambda s: int(s._row()["winner"][0]), lambda s, v: s._set_field("winner", None, int(v)))

    # -- reference methods ----------------------------------------------------
    def get_player_by_id(self, player_id):           # SkillshotGame.py:27-34
        if player_id == 1:
            return self.player1
        if player_id == 2:
            return self.player2
        return None

    def game_tick(self):                             # SkillshotGame.py:115-122
        was_live = self.game_live
        self._apply(0, _lib.OP_GAME_TICK)
        self._envs.check_status()
        if was_live and not self.game_live:
            print("Player", self.winner_id, "loss")   # SkillshotGame.py:76

    def game_reset(self, random_positions=False):    # SkillshotGame.py:168-169
        mask = torch.zeros(self._envs.n_envs, dtype=torch.uint8)
        mask[self._i] = 1
        self._envs.reset(mask=mask, random_positions=random_positions)
        self._cache = None

    def get_state(self):                             # SkillshotGame.py:136-166
        feat, _, gen = self._envs.features(want_obs=False)
        f = feat[self._i].cpu().numpy()
        g = gen[self._i].cpu().numpy()
        out = dict(game_live=bool(g[0]), ticks=int(g[1]), game_winner=int(g[2]))
        for p in (1, 2):
            d = {}
            for j, key in enumerate(FEATURE_KEYS):
                v = float(f[p - 1, j])
                d[key] = int(v) if key in _INT_KEYS else (bool(v) if key in _BOOL_KEYS else v)
            out[p] = d
        return out

    def get_board(self):
        """Host-side raster of SkillshotGame.get_board (SkillshotGame.py:36-56); rendering stays on the host."""
        return render_board(self._row(), 0)

    @staticmethod
    def get_dist_point_point(point1, point2):        # SkillshotGame.py:132-134
        return ((point1[0] - point2[0]) ** 2 + (point1[1] - point2[1]) ** 2) ** 0.5


def render_board(fields: dict, j: int) -> np.ndarray:
    """250x250 int raster of env row j of an export_state() dict (SkillshotGame.py:36-56)."""
    board = np.zeros((250, 250), dtype=int)
    for p, (colour, pointer) in enumerate(((1, 3), (2, 4))):
        x, y, rot = int(fields["px"][j, p]), int(fields["py"][j, p]), float(fields["prot"][j, p])
        board[x + 1:x + 4, y + 1:y + 4] = colour                       # the 3x3 body of the 5x5 shape
        ix = math.floor(-math.sin(rot) * 5 / 2 + 5 / 2)                # SkillshotGame.py:47-48
        iy = math.floor(-math.cos(rot) * 5 / 2 + 5 / 2)
        if 0 <= ix < 5 and 0 <= iy < 5:
            board[ix + x, iy + y] = pointer
        if fields["valid"][j, p]:
            qx, qy = int(fields["qx"][j, p]), int(fields["qy"][j, p])
            for dx, dy in ((0, 0), (2, 0), (1, 1), (0, 2), (2, 2)):    # Projectile.shape_image
                board[qx + dx, qy + dy] = pointer
    return board
