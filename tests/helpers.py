"""Shared helpers for the parity tests."""
import os

import numpy as np

from oracle.oracle import OracleEnvs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_FILES = ("kat", "lockstep_fixed", "lockstep_random", "close_hits", "speeds")   # "boards": tests/test_boards_golden.py
INT_FIELDS = ("px", "py", "qx", "qy", "cd", "age", "valid")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def oracle_from_golden(g):
    n = g["actions"].shape[0]
    o = OracleEnvs(n, g["positions"])
    o.envs["prot"] = g["rotations"]
    o.envs["np_pos"] = g["np_pos"]
    if "speed_move" in g:          # per-env Player / Projectile speed constants (Player.py:14-15, Projectile.py:9-10)
        o.set_speeds(g["speed_move"], g["speed_look"], g["proj_speed"], g["cooldown_max"])
    return o
