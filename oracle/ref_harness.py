"""Runs the UNMODIFIED reference: imported from /root/reference where that is
mounted (the authoring container), else from the sourceless byte-code that
oracle/build_ref.py compiled from it into oracle/_ref/ (a git-ignored build
output that travels to the GPU box).  Used by oracle/gen_golden.py to produce
tests/golden/*.npz, by the `requires_reference` tests that cross-check the C
oracle and the device facade live, and by bench.py's CPU legs.

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

SOURCE_DIR = os.environ.get("SKILLSHOT_REFERENCE_DIR", "/root/reference")
COMPILED_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def source() -> str | None:
    """"source" (the mounted reference), "compiled" (oracle/_ref byte-code of it) or None."""
    if os.path.exists(os.path.join(SOURCE_DIR, "SkillshotGame.py")):
        return "source"
    from oracle import build_ref
    if build_ref.built():
        return "compiled"
    return None


def available() -> bool:
    return source() is not None


def reference_dir() -> str:
    return SOURCE_DIR if source() == "source" else COMPILED_DIR


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


KEYS = dict(K_w=119, K_s=115, K_a=97, K_d=100, K_SPACE=32, K_UP=1073741906, K_DOWN=1073741905, K_LEFT=1073741904,
            K_RIGHT=1073741903, K_PERIOD=46, K_0=48)


def _install_pygame_stub():
    """InputHandler.py imports pygame for its key constants only (InputHandler.py:1, 10-52); pygame is not installed."""
    if "pygame" not in sys.modules:
        pg = types.ModuleType("pygame")
        for k, v in KEYS.items():
            setattr(pg, k, v)
        sys.modules["pygame"] = pg
    return sys.modules["pygame"]


def input_handler():
    """A reference InputHandler (InputHandler.py) driven through input_start / input_stop with the stub's key codes."""
    modules()
    _install_pygame_stub()
    from InputHandler import InputHandler
    return InputHandler()


def playable_tick():
    """Code object of skillshot_playable.py:51-64 (key states -> Player.move_* calls -> game_tick), to be exec'd with the
    globals `inputHandler` and `skillshotGame`: cut from the mounted source, or loaded from oracle/_ref."""
    from oracle import build_ref
    if source() == "source":
        return build_ref.playable_tick_code(SOURCE_DIR)
    import marshal
    with open(os.path.join(COMPILED_DIR, build_ref.PLAYABLE_TICK), "rb") as f:
        return marshal.load(f)


_mods = None


def modules():
    """(SkillshotGame class, SkillshotLearner class) of the reference."""
    global _mods
    if _mods is None:
        if not available():
            raise RuntimeError("reference neither mounted at %s nor compiled into %s" % (SOURCE_DIR, COMPILED_DIR))
        _install_tf_stub()
        _install_pygame_stub()
        if source() == "source":
            if SOURCE_DIR not in sys.path:
                sys.path.insert(0, SOURCE_DIR)
        else:
            from oracle import build_ref
            for m in build_ref.MODULES:
                build_ref.load_module(m)
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


def set_speeds(g, speed_move, speed_look, proj_speed, cooldown_max):
    """Per-game speed constants: the reference keeps them as class attributes (Player.py:14-15, Projectile.py:9-10);
    instance attributes of one game's objects shadow them for that game only."""
    for p in (g.player1, g.player2):
        p.speed_move, p.speed_look = speed_move, speed_look
        p.projectile.speed_move, p.projectile.cooldown_max = proj_speed, cooldown_max


BOARD_CELLS = 32      # a board holds at most 2 x (9 body cells incl. pointer) + 2 x 5 projectile cells non-zero


def sparse_board(board):
    """get_board() raster (SkillshotGame.py:36-56) as its non-zero cells: int16 [BOARD_CELLS, 3] = (x, y, value), -1 padded,
    in row-major order of the raster."""
    xs, ys = np.nonzero(board)
    assert len(xs) <= BOARD_CELLS
    out = np.full((BOARD_CELLS, 3), -1, np.int16)
    out[:len(xs), 0], out[:len(xs), 1], out[:len(xs), 2] = xs, ys, board[xs, ys]
    return out


def run_episode(actions, positions=None, rotations=None, speeds=None, boards=False, features=True):
    """Step one reference game through `actions` (float32 [T,2,2]) exactly as
    model_train does per tick (SkillshotLearner.py:304-315), continuing past the
    terminal tick.  Returns per-tick records, index 0 = initial state.
    speeds: (speed_move, speed_look, proj_speed, cooldown_max) of this game, or None (class constants);
    boards: also record get_board() per tick (sparse_board); features=False records the raw state only."""
    g = make_game(positions, rotations)
    if speeds is not None:
        set_speeds(g, *speeds)
    skl = make_learner(g)
    T = actions.shape[0]
    rec = dict((k, np.zeros((T + 1, 2), np.int64)) for k in STATE_INT_FIELDS)
    rec.update(prot=np.zeros((T + 1, 2)), qrot=np.zeros((T + 1, 2)),
               ticks=np.zeros(T + 1, np.int64), live=np.zeros(T + 1, np.int64),
               winner=np.zeros(T + 1, np.int64))
    if features:
        rec.update(feat=np.zeros((T + 1, 2, 18)), obs=np.zeros((T + 1, 2, 12)),
                   rew_looking=np.zeros((T + 1, 2)), rew_simple=np.zeros((T + 1, 2)))
    if boards:
        rec["board"] = np.full((T + 1, BOARD_CELLS, 3), -1, np.int16)
    sink = io.StringIO()

    def record(t):
        with contextlib.redirect_stdout(sink), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            st = read_state(g)
            for k in STATE_INT_FIELDS + ("prot", "qrot"):
                rec[k][t] = st[k]
            rec["ticks"][t], rec["live"][t], rec["winner"][t] = st["ticks"], st["live"], st["winner"]
            if features:
                sd = g.get_state()
                rec["feat"][t] = features_of(sd)
                for p in (1, 2):
                    rec["obs"][t, p - 1] = np.array(skl.prepare_states([sd], p)[0], dtype=np.float64)
                rl = skl.calculate_rewards_looking([sd])[0]
                rs = skl.calculate_rewards_simple([sd])[0]
                rec["rew_looking"][t] = [rl[1], rl[2]]
                rec["rew_simple"][t] = [rs[1], rs[2]]
            if boards:
                rec["board"][t] = sparse_board(g.get_board())

    record(0)
    for t in range(T):
        with contextlib.redirect_stdout(sink):
            for p in (1, 2):
                # float32 values handed over as Python floats (SURVEY hard part 1)
                skl.do_actions(p, (float(actions[t, p - 1, 0]), float(actions[t, p - 1, 1])))
            g.game_tick()
        sink.seek(0); sink.truncate(0)
        record(t + 1)
    return rec
