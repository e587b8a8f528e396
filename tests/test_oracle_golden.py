"""Pins the C oracle to the reference: every field the unmodified reference
produced (tests/golden/*.npz, written by oracle/gen_golden.py) must be
reproduced bit-for-bit -- ints, rotations AND the float64 features /
observations / rewards, since the oracle evaluates the same glibc libm calls in
the same order as CPython does."""
import numpy as np
import pytest

from tests.helpers import GOLDEN_FILES, INT_FIELDS, load_golden, oracle_from_golden


def _check_tick(o, g, t, name):
    s = o.snapshot()
    for k in INT_FIELDS:
        np.testing.assert_array_equal(s[k], g[k][:, t], err_msg=f"{name} t={t} {k}")
    for k in ("ticks", "live", "winner"):
        np.testing.assert_array_equal(s[k], g[k][:, t], err_msg=f"{name} t={t} {k}")
    # rotations: exact float64 equality
    assert s["prot"].tobytes() == g["prot"][:, t].tobytes(), f"{name} t={t} prot"
    assert s["qrot"].tobytes() == g["qrot"][:, t].tobytes(), f"{name} t={t} qrot"
    feat, obs, gen = o.features()
    np.testing.assert_array_equal(feat, g["feat"][:, t], err_msg=f"{name} t={t} feat")
    np.testing.assert_array_equal(obs, g["obs"][:, t], err_msg=f"{name} t={t} obs")
    np.testing.assert_array_equal(gen[:, 0], g["live"][:, t])
    np.testing.assert_array_equal(gen[:, 1], g["ticks"][:, t])
    np.testing.assert_array_equal(gen[:, 2], g["winner"][:, t])


@pytest.mark.parametrize("name", GOLDEN_FILES)
def test_oracle_reproduces_reference_bit_exact(name):
    g = load_golden(name)
    o = oracle_from_golden(g)
    T = g["actions"].shape[1]
    _check_tick(o, g, 0, name)
    for t in range(T):
        out = o.step(g["actions"][:, t], want_obs=True, reward_mode=1)
        assert out["errors"] == 0
        _check_tick(o, g, t + 1, name)
        # step outputs: float32 casts of the float64 golden values
        np.testing.assert_array_equal(out["obs"], g["obs"][:, t + 1].astype(np.float32))
        np.testing.assert_array_equal(out["reward"], g["rew_looking"][:, t + 1].astype(np.float32))
        np.testing.assert_array_equal(out["winner"], g["winner"][:, t + 1].astype(np.uint8))
        np.testing.assert_array_equal(out["done"], (g["live"][:, t + 1] == 0).astype(np.uint8))


@pytest.mark.parametrize("name", GOLDEN_FILES)
def test_oracle_reward_simple(name):
    g = load_golden(name)
    o = oracle_from_golden(g)
    for t in range(g["actions"].shape[1]):
        out = o.step(g["actions"][:, t], want_obs=False, reward_mode=3)
        np.testing.assert_array_equal(out["reward"], g["rew_simple"][:, t + 1].astype(np.float32))


def test_kat_values_from_survey():
    """The hand-recorded known answers of SURVEY.md section 4."""
    g = load_golden("kat")
    # KAT-A lifecycle: P1 projectile (50,45) valid cd14 age1 at tick 1; (50,0) tick 10;
    # invalid from tick 11, frozen; cd 0 age 15 at tick 15; respawn at tick 16.
    A = 0
    assert (g["qx"][A, 1, 0], g["qy"][A, 1, 0], g["valid"][A, 1, 0], g["cd"][A, 1, 0], g["age"][A, 1, 0]) == (50, 45, 1, 14, 1)
    assert (g["qx"][A, 10, 0], g["qy"][A, 10, 0], g["valid"][A, 10, 0]) == (50, 0, 1)
    assert (g["qx"][A, 11, 0], g["qy"][A, 11, 0], g["valid"][A, 11, 0]) == (50, 0, 0)
    assert (g["cd"][A, 15, 0], g["age"][A, 15, 0]) == (0, 15)
    assert (g["qx"][A, 16, 0], g["qy"][A, 16, 0], g["cd"][A, 16, 0], g["age"][A, 16, 0]) == (50, 45, 14, 1)
    assert (g["qx"][A, 15, 1], g["qy"][A, 15, 1], g["valid"][A, 15, 1]) == (200, 125, 1)
    # KAT-B vertical hit: terminal at tick 5, winner_id 1, P2 projectile at (100,105)
    B = 1
    assert g["live"][B, 4] == 1 and g["live"][B, 5] == 0 and g["winner"][B, 5] == 1
    assert (g["qx"][B, 5, 1], g["qy"][B, 5, 1]) == (100, 105)
    # KAT-C double hit: only id 1 recorded; ticks frozen afterwards
    C = 2
    assert g["live"][C, 5] == 0 and g["winner"][C, 5] == 1 and g["ticks"][C, 20] == 5
    assert (g["qx"][C, 5, 0], g["qy"][C, 5, 0], g["qx"][C, 5, 1], g["qy"][C, 5, 1]) == (125, 100, 103, 100)
    # KAT-D wall: whole move rejected
    D = 3
    assert (g["px"][D, 20, 0], g["py"][D, 20, 0]) == (1, 100)
    # KAT-E features
    E = 4
    assert (g["px"][E, 3, 0], g["py"][E, 3, 0], g["prot"][E, 3, 0]) == (51, 47, 0.40625)
    assert (g["qx"][E, 3, 0], g["qy"][E, 3, 0], g["qrot"][E, 3, 0], g["cd"][E, 3, 0], g["age"][E, 3, 0]) == (47, 32, 0.125, 12, 3)
    assert (g["px"][E, 3, 1], g["py"][E, 3, 1], g["prot"][E, 3, 1]) == (201, 200, -0.28125)
    np.testing.assert_allclose(g["obs"][E, 3, 0], [0.218724, 0.60603, 0.204, 0.188, 2.004763, 0.8, 0.644608,
                                                   0.188, 0.128, 0.61685, 0.372937, 0], atol=1e-6)
    np.testing.assert_allclose(g["rew_looking"][E, 1:4, 0], [-0.517522, -0.331216, -0.309323], atol=1e-6)
    np.testing.assert_allclose(g["rew_looking"][E, 1:4, 1], [-0.673116, -0.785394, -0.746290], atol=1e-6)
    # initial get_state
    assert g["feat"][A, 0, 0, 0] == 1.633123935319537e16
    assert g["feat"][A, 0, 0, 2] == 150.0
    assert g["feat"][A, 0, 0, 3] == 212.13203435596427
    assert g["feat"][A, 0, 0, 16] == 282.842712474619


def test_oracle_nan_action_is_an_error():
    """int(round(nan)) raises in the reference (Player.py:63); the oracle reports it."""
    from oracle.oracle import OracleEnvs
    o = OracleEnvs(2)
    a = np.zeros((2, 2, 2), np.float32)
    a[1, 0, 0] = np.nan
    assert o.step(a)["errors"] == 1


def test_oracle_tick_limit_and_auto_reset():
    from oracle.oracle import OracleEnvs
    o = OracleEnvs(3)
    a = np.zeros((3, 2, 2), np.float32)
    for t in range(4):
        out = o.step(a, tick_limit=4, auto_reset=True)
    assert out["done"].tolist() == [1, 1, 1]
    s = o.snapshot()
    assert s["ticks"].tolist() == [0, 0, 0] and s["px"][:, 0].tolist() == [50, 50, 50]
    feat0, obs0, _ = OracleEnvs(3).features()
    np.testing.assert_array_equal(out["obs"], obs0.astype(np.float32))
