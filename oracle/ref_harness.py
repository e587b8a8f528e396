"""Runs the UNMODIFIED reference (imported from /root/reference) -- only usable
in the authoring container, where /root/reference is mounted.  Used by
oracle/gen_golden.py to produce tests/golden/*.npz and by the
`requires_reference` tests that cross-check the C oracle live.

TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

The reference's learner module imports tensorflow at module scope
(SkillshotLearner.py:5-8); TensorFlow is not installed, so a stub module that
provides only the names the module body touches is registered first.  The
pure-Python learner methods (do_actions, prepare_states, calculate_rewards*)
then run verbatim on an instance made with object.__new__.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types
import warnings

import numpy as np

REFERENCE_DIR = os.environ.get("SKILLSHOT_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.exists(os.path.join(REFERENCE_DIR, "SkillshotGame.py"))


def _install_tf_stub():
    if "tensorflow" in sys.modules:
        return
    tf = types.ModuleType("tensorflow")
    tf.function = lambda f: f
    keras = types.ModuleType("tensorflow.keras")
    backend = types.ModuleType("tensorflow.keras.backend")
    layers = types.ModuleType("tensorflow.keras.layers")
    for name in ("Input", "Model"):
        setattr(keras, name, type(name, (), {}))
    for name in ("Dense", "GaussianNoise", "concatenate", "Dropout"):
        setattr(layers, name, type(name, (), {}))
    keras.backend = backend
    keras.layers = layers
    tf.keras = keras
    sys.modules.update({
        "tensorflow": tf, "tensorflow.keras": keras,
        "tensorflow.keras.backend": backend, "tensorflow.keras.layers": layers,
    })


_mods = None


def modules():
    """(SkillshotGame class, SkillshotLearner class) of the reference."""
    global _mods
    if _mods is None:
        if not available():
            raise RuntimeError("reference not mounted at %s" % REFERENCE_DIR)
        if REFERENCE_DIR not in sys.path:
            sys.path.insert(0, REFERENCE_DIR)
        _install_tf_stub()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")  # `is not 0` SyntaxWarning, SkillshotGame.py:44,54
            from SkillshotGame import SkillshotGame
            from SkillshotLearner import SkillshotLearner
        _mods = (SkillshotGame, SkillshotLearner)
    return _mods


def make_learner(game):
    """A reference SkillshotLearner without its Keras models (see module doc)."""
    _, SkillshotLearner = modules()
    skl = object.__new__(SkillshotLearner)
    skl.game_environment = game
    skl.player_ids = (1, 2)
    skl.max_dist_normaliser = (2 * (250 ** 2)) ** 0.5   # SkillshotLearner.py:43
    return skl


def make_game(positions=None, rotations=None):
    """positions: None (fixed start) or (p1x, p1y, p2x, p2y) ints."""
    SkillshotGame, _ = modules()
    g = SkillshotGame()
    if positions is not None:
        # the reference's random start stores numpy int64 rows (SkillshotGame.py:15)
        arr = np.array([[positions[0], positions[1]], [positions[2], positions[3]]], dtype=np.int64)
        g.player1.pos, g.player2.pos = arr
    if rotations is not None:
        g.player1.rotation, g.player2.rotation = float(rotations[0]), float(rotations[1])
    return g


STATE_INT_FIELDS = ("px", "py", "qx", "qy", "cd", "age", "valid")


def read_state(g):
    """All mutable fields of one reference game as plain Python numbers."""
    p = (g.player1, g.player2)
    return dict(
        px=[int(q.pos[0]) for q in p], py=[int(q.pos[1]) for q in p],
        prot=[float(q.rotation) for q in p],
        qx=[int(q.projectile.pos[0]) for q in p], qy=[int(q.projectile.pos[1]) for q in p],
        qrot=[float(q.projectile.rotation) for q in p],
        cd=[int(q.projectile.cooldown_current) for q in p],
        age=[int(q.projectile.age) for q in p],
        valid=[int(bool(q.projectile.valid)) for q in p],
        ticks=int(g.ticks), live=int(bool(g.game_live)), winner=int(g.winner_id),
    )


def features_of(state_dict):
    """get_state() dict -> float64 [2,18] in key order."""
    from oracle.oracle import FEATURE_KEYS
    out = np.empty((2, 18), np.float64)
    for p in (1, 2):
        for k, key in enumerate(FEATURE_KEYS):
            out[p - 1, k] = float(state_dict[p][key])
    return out


def run_episode(actions, positions=None, rotations=None):
    """Step one reference game through `actions` (float32 [T,2,2]) exactly as
    model_train does per tick (SkillshotLearner.py:304-315), continuing past the
    terminal tick.  Returns per-tick records, index 0 = initial state."""
    g = make_game(positions, rotations)
    skl = make_learner(g)
    T = actions.shape[0]
    rec = dict((k, np.zeros((T + 1, 2), np.int64)) for k in STATE_INT_FIELDS)
    rec.update(prot=np.zeros((T + 1, 2)), qrot=np.zeros((T + 1, 2)),
               ticks=np.zeros(T + 1, np.int64), live=np.zeros(T + 1, np.int64),
               winner=np.zeros(T + 1, np.int64),
               feat=np.zeros((T + 1, 2, 18)), obs=np.zeros((T + 1, 2, 12)),
               rew_looking=np.zeros((T + 1, 2)), rew_simple=np.zeros((T + 1, 2)))
    sink = io.StringIO()

    def record(t):
        with contextlib.redirect_stdout(sink):
            sd = g.get_state()
            st = read_state(g)
            for k in STATE_INT_FIELDS + ("prot", "qrot"):
                rec[k][t] = st[k]
            rec["ticks"][t], rec["live"][t], rec["winner"][t] = st["ticks"], st["live"], st["winner"]
            rec["feat"][t] = features_of(sd)
            for p in (1, 2):
                rec["obs"][t, p - 1] = np.array(skl.prepare_states([sd], p)[0], dtype=np.float64)
            rl = skl.calculate_rewards_looking([sd])[0]
            rs = skl.calculate_rewards_simple([sd])[0]
            rec["rew_looking"][t] = [rl[1], rl[2]]
            rec["rew_simple"][t] = [rs[1], rs[2]]

    record(0)
    for t in range(T):
        with contextlib.redirect_stdout(sink):
            for p in (1, 2):
                # float32 values handed over as Python floats (SURVEY hard part 1)
                skl.do_actions(p, (float(actions[t, p - 1, 0]), float(actions[t, p - 1, 1])))
            g.game_tick()
        record(t + 1)
    return rec
