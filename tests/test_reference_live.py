"""Live cross-checks against the UNMODIFIED reference (`requires_reference`: /root/reference, or its byte-code compiled
into oracle/_ref by oracle/build_ref.py).  The golden files pin ~7,000 env-ticks; here the C oracle (the checker every GPU
parity test relies on) is stepped in lockstep with fresh reference games on > 2e5 env-ticks -- fixed and random starts,
per-env speed constants (Player.py:14-15, Projectile.py:9-10), structured and out-of-range actions, play continuing
after the terminal tick -- and render_board is compared with the reference's own get_board() on every tick of a batch."""
import numpy as np
import pytest

from oracle import ref_harness
from oracle.oracle import OracleEnvs
from tests.helpers import INT_FIELDS

pytestmark = pytest.mark.requires_reference


def _actions(rng, n, T):
    a = (rng.uniform(-1, 1, size=(n, T, 2, 2)) * 1.3).astype(np.float32)
    special = np.array([0.0, 1.0, -1.0, 0.5, -0.5, 2.0, -3.0], np.float32)
    m = rng.uniform(size=a.shape) < 0.06
    a[m] = special[rng.integers(0, len(special), size=int(m.sum()))]
    return a


def _lockstep(n, T, seed, start, speeds, features_every=0):
    """n reference games x T ticks against the oracle; returns the number of env-ticks compared and of terminal games."""
    rng = np.random.default_rng(seed)
    actions = _actions(rng, n, T)
    pos = np.tile(np.array([50, 50, 200, 200], np.int64), (n, 1))
    if start == "random":
        pos = rng.integers(25, 225, size=(n, 4))
    elif start == "close":
        p1 = rng.integers(40, 200, size=(n, 2))
        pos = np.concatenate([p1, np.clip(p1 + rng.integers(-40, 41, size=(n, 2)), 0, 245)], axis=1)
    sp = None
    if speeds:
        sc = lambda: rng.uniform(0.5, 2.0, n)
        sp = (3.0 * sc(), 0.25 * sc(), 5.0 * sc(), np.maximum(1, np.rint(15.0 * sc())).astype(np.int64))
    fixed = start == "fixed"
    recs = [ref_harness.run_episode(actions[i], None if fixed else tuple(int(v) for v in pos[i]), None,
                                    speeds=None if sp is None else (float(sp[0][i]), float(sp[1][i]), float(sp[2][i]), int(sp[3][i])),
                                    features=bool(features_every)) for i in range(n)]
    ref = {k: np.stack([r[k] for r in recs]) for k in recs[0]}
    orc = OracleEnvs(n, None if fixed else pos)
    orc.envs["np_pos"] = 0 if fixed else 1
    if sp is not None:
        orc.set_speeds(*sp)
    for t in range(T + 1):
        if t:
            out = orc.step(actions[:, t - 1], want_obs=bool(features_every), reward_mode=1)
            assert out["errors"] == 0
        s = orc.snapshot()
        for k in INT_FIELDS + ("ticks", "live", "winner"):
            np.testing.assert_array_equal(s[k], ref[k][:, t], err_msg=f"t={t} {k}")
        assert s["prot"].tobytes() == np.ascontiguousarray(ref["prot"][:, t]).tobytes(), f"t={t} prot"
        assert s["qrot"].tobytes() == np.ascontiguousarray(ref["qrot"][:, t]).tobytes(), f"t={t} qrot"
        if features_every and t % features_every == 0:
            feat, obs, _ = orc.features()
            np.testing.assert_array_equal(feat, ref["feat"][:, t], err_msg=f"t={t} get_state")          # float64, bit for bit
            np.testing.assert_array_equal(obs, ref["obs"][:, t], err_msg=f"t={t} prepare_states")
            if t:
                np.testing.assert_array_equal(out["reward"], ref["rew_looking"][:, t].astype(np.float32))
    return n * T, int((ref["live"][:, -1] == 0).sum())


def test_oracle_matches_the_live_reference_on_2e5_env_ticks():
    total = terminals = 0
    for seed, (n, T, start, speeds) in enumerate([(40, 2000, "random", False), (30, 2000, "fixed", False),
                                                  (40, 1000, "random", True), (120, 200, "close", False),
                                                  (120, 200, "close", True)]):
        ticks, dead = _lockstep(n, T, 100 + seed, start, speeds)
        total += ticks
        terminals += dead
    assert total >= 200000 and terminals >= 100


def test_oracle_features_match_the_live_reference():
    """get_state / prepare_states / calculate_rewards_looking of the live reference, float64 bit for bit, speeds included."""
    ticks = 0
    for seed, (start, speeds) in enumerate([("random", False), ("close", True), ("fixed", False)]):
        ticks += _lockstep(24, 160, 200 + seed, start, speeds, features_every=1)[0]
    assert ticks >= 10000


def test_render_board_matches_the_live_reference_get_board():
    """render_board (the host-side rebuild of get_board, SkillshotGame.py:36-56) against the reference's own raster on
    every tick of live games, including pointer cells at the extremes and projectiles leaving the board."""
    from skillshot_learning_b200.game import render_board
    rng = np.random.default_rng(7)
    n, T = 30, 120
    actions = _actions(rng, n, T)
    pos = rng.integers(0, 246, size=(n, 4))
    boards = 0
    for i in range(n):
        rec = ref_harness.run_episode(actions[i], tuple(int(v) for v in pos[i]), rng.uniform(-4, 4, 2), boards=True, features=False)
        for t in range(T + 1):
            fields = {k: rec[k][t:t + 1] for k in ("px", "py", "qx", "qy", "valid", "prot")}
            want = np.zeros((250, 250), dtype=int)
            for x, y, v in rec["board"][t]:
                if v >= 0:
                    want[x, y] = v
            assert np.array_equal(render_board(fields, 0), want), (i, t)
            boards += 1
    assert boards >= 3000
