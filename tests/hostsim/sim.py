"""numpy driver of the TEST-ONLY host build of the kernel core (hostsim.cpp).
Mirrors the SkillshotEnvs interface so that the parity checks in tests/parity.py
run unchanged against either the host simulation (here, no GPU) or the CUDA
library (on the B200)."""
import ctypes

import numpy as np

from skillshot_learning_b200 import _lib as L
from skillshot_learning_b200.game import REWARD_MODES, pack_import, unpack_export
from tests.hostsim.build import build

_hs = None


def hs():
    global _hs
    if _hs is None:
        _hs = ctypes.CDLL(build())
        vp, i64, i32, u64 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_uint64
        _hs.hs_env_reset.argtypes = [vp, i64, vp, i32, vp, u64, u64]
        _hs.hs_env_step.argtypes = [vp, i64, vp, vp, vp, vp, vp, i32, i32, i64, i32, i32, u64, u64, vp, vp, i32]
        _hs.hs_env_features.argtypes = [vp, i64, vp, vp, vp, vp]
        _hs.hs_env_export.argtypes = [vp, i64, i64, i64, vp, vp]
        _hs.hs_env_import.argtypes = [vp, i64, i64, i64, vp, vp]
    return _hs


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class HostSimEnvs:
    def __init__(self, n_envs, random_positions=False, seed=0, reward_mode="looking", tick_limit=0,
                 auto_reset=False):
        self.n_envs = n_envs
        self.random_positions, self.seed = random_positions, seed
        self.reward_mode, self.tick_limit, self.auto_reset = reward_mode, tick_limit, auto_reset
        self.counter = 0
        self.state = np.zeros(64 * n_envs, np.uint8)
        self.status = np.zeros(1, np.uint32)
        self.speeds = None
        self.reset()

    def set_speeds(self, speed_move, speed_look, proj_speed, cooldown_max):
        n = self.n_envs
        buf = np.zeros(32 * n, np.uint8)
        f = buf[:16 * n].view(np.float64).reshape(n, 2)
        f[:, 0], f[:, 1] = speed_move, speed_look
        buf[16 * n:].view(np.float64).reshape(n, 2)[:, 0] = proj_speed
        buf[16 * n:].view(np.int64).reshape(n, 2)[:, 1] = cooldown_max
        self.speeds = buf

    def reset(self, mask=None, positions=None, random_positions=None):
        rnd = self.random_positions if random_positions is None else random_positions
        mode, pos = (L.RESET_RANDOM if rnd else L.RESET_FIXED), None
        if positions is not None:
            pos = np.ascontiguousarray(positions, np.int32).reshape(self.n_envs, 4)
            mode = L.RESET_GIVEN
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
        hs().hs_env_reset(_p(self.state), self.n_envs, _p(mask), mode, _p(pos), self.seed, self.counter)
        self.counter += 1

    def step(self, actions, want_obs=True, obs_every_tick=False):
        a = np.ascontiguousarray(np.asarray(actions, np.float32))
        single = a.ndim == 3
        a = a.reshape(-1, self.n_envs, 2, 2)
        K, n = a.shape[0], self.n_envs
        rew = np.zeros((K, n, 2), np.float32)
        done = np.zeros((K, n), np.uint8)
        win = np.zeros((K, n), np.uint8)
        obs, flags = None, 0
        if want_obs:
            if obs_every_tick and K > 1:
                obs = np.zeros((K, n, 2, 12), np.float32); flags |= L.STEP_OBS_EVERY_TICK
            else:
                obs = np.zeros((n, 2, 12), np.float32)
        hs().hs_env_step(_p(self.state), n, _p(a), _p(obs), _p(rew), _p(done), _p(win), K,
                         REWARD_MODES[self.reward_mode], self.tick_limit, int(self.auto_reset),
                         L.RESET_RANDOM if self.random_positions else L.RESET_FIXED, self.seed, self.counter,
                         _p(self.speeds), _p(self.status), flags)
        self.counter += K
        if single:
            return dict(obs=obs, reward=rew[0], done=done[0], winner=win[0])
        return dict(obs=obs, reward=rew, done=done, winner=win)

    def features(self, want_feat=True, want_obs=True):
        n = self.n_envs
        feat = np.zeros((n, 2, 18)) if want_feat else None
        obs = np.zeros((n, 2, 12)) if want_obs else None
        gen = np.zeros((n, 3), np.int32)
        hs().hs_env_features(_p(self.state), n, _p(feat), _p(obs), _p(gen), _p(self.speeds))
        return feat, obs, gen

    def export_state(self, first=0, count=None):
        count = self.n_envs - first if count is None else count
        ints = np.zeros((count, L.EXPORT_INTS), np.int32)
        rots = np.zeros((count, 4))
        hs().hs_env_export(_p(self.state), self.n_envs, first, count, _p(ints), _p(rots))
        return unpack_export(ints, rots)

    def import_state(self, fields, first=0):
        ints, rots = pack_import(fields)
        hs().hs_env_import(_p(self.state), self.n_envs, first, ints.shape[0], _p(ints), _p(rots))

    def status_bits(self):
        return int(self.status[0])
