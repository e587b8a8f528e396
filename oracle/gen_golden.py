"""Generates tests/golden/*.npz from the UNMODIFIED reference.

Run in the authoring container (needs /root/reference):
    python -m oracle.gen_golden

The reference has no tests or golden vectors of its own,
so these files are the pinned known answers: per-tick dumps of every mutable
field of the reference SkillshotGame plus get_state(), prepare_states() and
calculate_rewards_looking/_simple outputs, for seeded float32 action streams.

TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os

import numpy as np

from oracle import ref_harness

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _stack(recs):
    return {k: np.stack([r[k] for r in recs]) for k in recs[0]}


def _save(name, actions, positions, rotations, recs, np_pos, extra=None):
    d = _stack(recs)
    d.update(extra or {})
    # 1 where Player.pos was a numpy int64 row (as after a random start), 0 where it
    # was the fixed start's Python list: selects sqrt vs pow in get_dist_point_point
    d["np_pos"] = np.asarray(np_pos, np.int32)
    d["actions"] = actions.astype(np.float32)
    d["positions"] = positions.astype(np.int64)
    d["rotations"] = rotations.astype(np.float64)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    live_end = d["live"][:, -1]
    print("%-22s envs=%d ticks=%d terminals=%d winners(1/2)=%d/%d  %.0f KB" % (
        name, actions.shape[0], actions.shape[1], int((live_end == 0).sum()),
        int((d["winner"][:, -1] == 1).sum()), int((d["winner"][:, -1] == 2).sum()),
        os.path.getsize(path) / 1024))


def scenario(name, n, T, seed, start="fixed", action_scale=1.0, close=False):
    rng = np.random.default_rng(seed)
    actions = (rng.uniform(-1, 1, size=(n, T, 2, 2)) * action_scale).astype(np.float32)
    # sprinkle exact structured values: 0, +-1, +-0.5 (ties of round-half-even)
    special = np.array([0.0, 1.0, -1.0, 0.5, -0.5, 2.0, -3.0], np.float32)
    mask = rng.uniform(size=actions.shape) < 0.08
    actions[mask] = special[rng.integers(0, len(special), size=int(mask.sum()))]
    positions = np.tile(np.array([50, 50, 200, 200], np.int64), (n, 1))
    rotations = np.zeros((n, 2))
    if start == "random":   # SkillshotGame.py:15  randint(25, 225, (2, 2))
        positions = rng.integers(25, 225, size=(n, 4))
    if close:               # close starts so that projectiles actually hit
        p1 = rng.integers(40, 200, size=(n, 2))
        off = rng.integers(-30, 31, size=(n, 2))
        positions = np.concatenate([p1, np.clip(p1 + off, 0, 245)], axis=1)
        # aim roughly at each other: bearing of -sin/-cos motion convention
        d = (positions[:, 2:] - positions[:, :2]).astype(np.float64)
        aim = np.arctan2(-d[:, 0], -d[:, 1])
        rotations = np.stack([aim, aim + np.pi], axis=1) + rng.normal(0, 0.15, size=(n, 2))
        actions[:, :, :, 1] *= 0.2   # small turns keep them on target
    recs = []
    fixed = (start == "fixed" and not close)
    for i in range(n):
        recs.append(ref_harness.run_episode(
            actions[i], None if fixed else tuple(int(v) for v in positions[i]),
            None if not close else rotations[i]))
    _save(name, actions, positions, rotations, recs, [0 if fixed else 1] * n)


def speeds_scenario(name, n, T, seed):
    """Per-env game-speed constants (the readme.md:22-23 sweep): Player.speed_move / speed_look and
    Projectile.speed_move / cooldown_max (class attributes, Player.py:14-15, Projectile.py:9-10) overridden per game,
    U(0.5, 2) x the defaults; close random starts so that projectiles of every speed get to hit."""
    rng = np.random.default_rng(seed)
    actions = (rng.uniform(-1, 1, size=(n, T, 2, 2)) * 1.2).astype(np.float32)
    p1 = rng.integers(40, 200, size=(n, 2))
    positions = np.concatenate([p1, np.clip(p1 + rng.integers(-60, 61, size=(n, 2)), 0, 245)], axis=1)
    d = (positions[:, 2:] - positions[:, :2]).astype(np.float64)
    aim = np.arctan2(-d[:, 0], -d[:, 1])
    rotations = np.stack([aim, aim + np.pi], axis=1) + rng.normal(0, 0.3, size=(n, 2))
    actions[:, :, :, 1] *= 0.3
    scale = lambda: rng.uniform(0.5, 2.0, n)
    sm, sl, ps = 3.0 * scale(), 0.25 * scale(), 5.0 * scale()
    cm = np.maximum(1, np.rint(15.0 * scale())).astype(np.int64)
    recs = [ref_harness.run_episode(actions[i], tuple(int(v) for v in positions[i]), rotations[i],
                                    speeds=(float(sm[i]), float(sl[i]), float(ps[i]), int(cm[i]))) for i in range(n)]
    _save(name, actions, positions, rotations, recs, [1] * n,
          extra=dict(speed_move=sm, speed_look=sl, proj_speed=ps, cooldown_max=cm))


def boards_scenario(name, n, T, seed):
    """get_board() rasters per tick (SkillshotGame.py:36-56) as sparse cell lists, with the state they were drawn from:
    random starts and turning players (the direction pointer visits every cell it can), shots fired throughout."""
    rng = np.random.default_rng(seed)
    actions = (rng.uniform(-1, 1, size=(n, T, 2, 2)) * 1.2).astype(np.float32)
    positions = rng.integers(25, 225, size=(n, 4))
    positions[0] = (0, 0, 245, 245)            # corners of the board: the raster's edges
    rotations = rng.uniform(-np.pi, np.pi, size=(n, 2))
    rotations[1] = (np.pi / 2, -np.pi / 2)     # pointer index at the extremes of floor(-sin * 2.5 + 2.5)
    recs = [ref_harness.run_episode(actions[i], tuple(int(v) for v in positions[i]), rotations[i], boards=True, features=False)
            for i in range(n)]
    _save(name, actions, positions, rotations, recs, [1] * n)


def kats():
    """KAT-A..E of SURVEY.md section 4 as one file (zero/explicit actions)."""
    T = 20
    cases = []
    z = np.zeros((T, 2, 2), np.float32)
    cases.append(("A_lifecycle", z, None, None))
    cases.append(("B_vertical_hit", z, (100, 100, 100, 130), (0.0, 0.0)))
    cases.append(("C_double_hit", z, (100, 100, 128, 100), (-np.pi / 2, np.pi / 2)))
    a = z.copy(); a[:, 0, 0] = 1.0
    cases.append(("D_wall", a, (1, 100, 200, 200), (np.pi / 4, 0.0)))
    e = z.copy()
    e[0] = [[1, .5], [-1, -.5]]; e[1] = [[.25, 1], [.75, -1]]; e[2] = [[-.5, .125], [.5, .375]]
    cases.append(("E_features", e, None, None))
    recs, acts, poss, rots = [], [], [], []
    for name, a, pos, rot in cases:
        recs.append(ref_harness.run_episode(a, pos, rot))
        acts.append(a)
        poss.append(pos if pos is not None else (50, 50, 200, 200))
        rots.append(rot if rot is not None else (0.0, 0.0))
    _save("kat", np.stack(acts), np.array(poss), np.array(rots), recs,
          [0 if c[2] is None else 1 for c in cases])


def main():
    os.makedirs(OUT, exist_ok=True)
    kats()
    scenario("lockstep_fixed", 24, 64, seed=1)
    scenario("lockstep_random", 24, 64, seed=2, start="random", action_scale=1.3)
    scenario("close_hits", 48, 40, seed=3, close=True)
    speeds_scenario("speeds", 32, 48, seed=4)
    boards_scenario("boards", 12, 40, seed=5)


if __name__ == "__main__":
    main()
