"""ctypes front-end of the CPU oracle (oracle/skillshot_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of skillshot_oracle.c.  Imported by
tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference),
never by skillshot_learning_b200/.

The oracle restates the reference game (Projectile.py, Player.py,
SkillshotGame.py) and the pure-Python learner helpers (do_actions,
prepare_states, calculate_rewards_looking/_simple of SkillshotLearner.py); each
C function cites the reference lines it follows.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libskillshot_oracle.so")

NFEAT = 18
NOBS = 12

FEATURE_KEYS = [  # SkillshotGame.get_state key order, SkillshotGame.py:145-162
    "player_grad", "player_x_dir", "player_path_dist_opponent", "player_dist_opponent",
    "player_pos_x", "player_pos_y", "player_rotation", "projectile_cooldown",
    "projectile_grad", "projectile_x_dir", "projectile_path_dist_opponent",
    "projectile_pos_x", "projectile_pos_y", "projectile_rotation", "projectile_age",
    "projectile_valid", "projectile_dist_opponent", "projectile_future_collision_opponent",
]

ENV_DTYPE = np.dtype(
    [
        ("px", "<i8", (2,)), ("py", "<i8", (2,)), ("prot", "<f8", (2,)),
        ("qx", "<i8", (2,)), ("qy", "<i8", (2,)), ("qrot", "<f8", (2,)),
        ("cd", "<i8", (2,)), ("age", "<i8", (2,)), ("valid", "<i4", (2,)),
        ("ticks", "<i8"), ("live", "<i4"), ("winner", "<i4"),
        ("speed_move", "<f8"), ("speed_look", "<f8"), ("proj_speed", "<f8"),
        ("cooldown_max", "<i8"), ("np_pos", "<i4"), ("pad_", "<i4"),
    ],
    align=True,
)


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "skillshot_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        vp, i64, i32, f64 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double
        L.ss_oracle_env_size.restype = i64
        assert L.ss_oracle_env_size() == ENV_DTYPE.itemsize, (L.ss_oracle_env_size(), ENV_DTYPE.itemsize)
        L.ss_oracle_reset_batch.argtypes = [vp, i64, vp]
        L.ss_oracle_step_batch.argtypes = [vp, i64, vp, vp, vp, vp, vp, i32, i64, i32, vp, i32]
        L.ss_oracle_step_batch.restype = i32
        L.ss_oracle_features_batch.argtypes = [vp, i64, vp, vp, vp]
        L.ss_oracle_move_direction_float.argtypes = [vp, i32, f64]
        L.ss_oracle_move_direction_float.restype = i32
        L.ss_oracle_move_look_float.argtypes = [vp, i32, f64]
        L.ss_oracle_move_shoot.argtypes = [vp, i32]
        L.ss_oracle_move_step.argtypes = [vp, i32, i32]
        L.ss_oracle_move_step.restype = i32
        L.ss_oracle_look_step.argtypes = [vp, i32, i32]
        L.ss_oracle_game_tick.argtypes = [vp]
        L.ss_oracle_game_tick.restype = i32
        L.ss_oracle_do_actions.argtypes = [vp, i32, f64, f64]
        L.ss_oracle_do_actions.restype = i32
        L.ss_oracle_max_threads.restype = i32
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class OracleEnvs:
    """n independent reference-semantics games held as a C struct array."""

    def __init__(self, n: int, positions: np.ndarray | None = None):
        self.n = int(n)
        self.envs = np.zeros(self.n, dtype=ENV_DTYPE)
        self.reset(positions)

    def reset(self, positions: np.ndarray | None = None):
        """positions: None = fixed start (SkillshotGame.py:17-18) or int64 [n,4] = p1x,p1y,p2x,p2y."""
        pos = None
        if positions is not None:
            pos = np.ascontiguousarray(positions, dtype=np.int64).reshape(self.n, 4)
        lib().ss_oracle_reset_batch(_p(self.envs), self.n, _p(pos))

    def set_speeds(self, speed_move, speed_look, proj_speed, cooldown_max):
        self.envs["speed_move"] = speed_move
        self.envs["speed_look"] = speed_look
        self.envs["proj_speed"] = proj_speed
        self.envs["cooldown_max"] = cooldown_max

    def step(self, actions: np.ndarray, want_obs=True, reward_mode=1, tick_limit=0,
             auto_reset=False, reset_pos=None, nthreads=1):
        """One model_train tick for every env (SkillshotLearner.py:304-315).

        actions float32 [n,2,2] = (player, (move, look)).  Returns a dict with
        obs f32 [n,2,12] (or None), reward f32 [n,2], done u8 [n], winner u8 [n],
        errors = number of envs where the reference would have raised.
        """
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.n, 2, 2)
        obs = np.empty((self.n, 2, NOBS), np.float32) if want_obs else None
        rew = np.zeros((self.n, 2), np.float32)
        done = np.zeros(self.n, np.uint8)
        win = np.zeros(self.n, np.uint8)
        rp = None
        if reset_pos is not None:
            rp = np.ascontiguousarray(reset_pos, dtype=np.int64).reshape(self.n, 4)
        err = lib().ss_oracle_step_batch(_p(self.envs), self.n, _p(a), _p(obs), _p(rew), _p(done),
                                         _p(win), int(reward_mode), int(tick_limit),
                                         int(bool(auto_reset)), _p(rp), int(nthreads))
        return dict(obs=obs, reward=rew, done=done, winner=win, errors=err)

    def features(self):
        """(feat f64 [n,2,18], obs f64 [n,2,12], general i64 [n,3]) of the current state."""
        feat = np.empty((self.n, 2, NFEAT), np.float64)
        obs = np.empty((self.n, 2, NOBS), np.float64)
        gen = np.empty((self.n, 3), np.int64)
        lib().ss_oracle_features_batch(_p(self.envs), self.n, _p(feat), _p(obs), _p(gen))
        return feat, obs, gen

    # --- single-env object-style calls (env index i, player id 1 or 2) ---
    def _e(self, i):
        return ctypes.c_void_p(self.envs.ctypes.data + i * ENV_DTYPE.itemsize)

    def move_direction_float(self, i, pid, speed):
        if lib().ss_oracle_move_direction_float(self._e(i), pid - 1, float(speed)):
            raise ValueError("cannot convert float NaN to integer")

    def move_look_float(self, i, pid, angle):
        lib().ss_oracle_move_look_float(self._e(i), pid - 1, float(angle))

    def move_shoot_projectile(self, i, pid):
        lib().ss_oracle_move_shoot(self._e(i), pid - 1)

    def move_step(self, i, pid, direction):
        lib().ss_oracle_move_step(self._e(i), pid - 1, int(direction))

    def look_step(self, i, pid, direction):
        lib().ss_oracle_look_step(self._e(i), pid - 1, int(direction))

    def game_tick(self, i):
        if lib().ss_oracle_game_tick(self._e(i)):
            raise ValueError("cannot convert float NaN to integer")

    def do_actions(self, i, pid, a_move, a_look):
        if lib().ss_oracle_do_actions(self._e(i), pid - 1, float(a_move), float(a_look)):
            raise ValueError("cannot convert float NaN to integer")

    def snapshot(self):
        """Discrete + rotation state as plain arrays (for comparisons)."""
        e = self.envs
        return dict(
            px=e["px"].copy(), py=e["py"].copy(), prot=e["prot"].copy(),
            qx=e["qx"].copy(), qy=e["qy"].copy(), qrot=e["qrot"].copy(),
            cd=e["cd"].copy(), age=e["age"].copy(), valid=e["valid"].copy(),
            ticks=e["ticks"].copy(), live=e["live"].copy(), winner=e["winner"].copy(),
        )
