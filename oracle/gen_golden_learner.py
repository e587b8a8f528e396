"""Generates tests/golden/learner_host.npz from the UNMODIFIED reference (authoring container only):
the pure-Python dataset / reward helpers of SkillshotLearner (prepare_states, prepare_actions,
prepare_rewards, calculate_rewards_looking / _simple / calculate_rewards, SkillshotLearner.py:512-661)
on constructed episodes, one of which ends in a hit so the winner branch of calculate_rewards runs.

    python oracle/gen_golden_learner.py
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_harness  # noqa: E402

KEYS = None


def episode(actions, positions, rotations):
    g = ref_harness.make_game(positions, rotations)
    skl = ref_harness.make_learner(g)
    sink = io.StringIO()
    states = []
    with contextlib.redirect_stdout(sink):
        states.append(g.get_state())
        for t in range(actions.shape[0]):
            if not g.game_live:
                break
            for p in (1, 2):
                skl.do_actions(p, (float(actions[t, p - 1, 0]), float(actions[t, p - 1, 1])))
            g.game_tick()
            states.append(g.get_state())
        post = states[1:]
        out = dict(
            looking=np.array([[r[1], r[2]] for r in skl.calculate_rewards_looking(post)]),
            simple=np.array([[r[1], r[2]] for r in skl.calculate_rewards_simple(post)]),
            shaped=np.array([[r[1], r[2]] for r in skl.calculate_rewards(post)], dtype=np.float64),
            prepared=np.array([skl.prepare_states(states, p) for p in (1, 2)]),
        )
    feat = np.array([ref_harness.features_of(s) for s in states])
    gen = np.array([[int(s["game_live"]), s["ticks"], s["game_winner"]] for s in states])
    return dict(feat=feat, general=gen, **out)


def main():
    rng = np.random.default_rng(3)
    eps = {
        # vertical shoot-out: P2 is hit at tick 5 (KAT-B of SURVEY.md 4)
        "hit": episode(np.zeros((8, 2, 2), np.float32), (100, 100, 100, 130), (0.0, 0.0)),
        # facing each other, some turning: future-collision flag toggles
        "duel": episode(rng.uniform(-0.2, 0.2, (40, 2, 2)).astype(np.float32), (60, 120, 190, 120), (-np.pi / 2, np.pi / 2)),
        "random": episode(rng.uniform(-1, 1, (60, 2, 2)).astype(np.float32), (50, 50, 200, 200), (0.0, 0.0)),
    }
    flat = {}
    for name, e in eps.items():
        for k, v in e.items():
            flat["%s_%s" % (name, k)] = v
    out = os.path.join(os.path.dirname(HERE), "tests", "golden", "learner_host.npz")
    np.savez_compressed(out, **flat)
    print(out, {k: v.shape for k, v in flat.items()})


if __name__ == "__main__":
    main()
