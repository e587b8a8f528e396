"""The host-side dataset / reward helpers of the SkillshotLearner facade against golden vectors
generated from the unmodified reference (oracle/gen_golden_learner.py ->
tests/golden/learner_host.npz): prepare_states, calculate_rewards_looking / _simple and the shaped
calculate_rewards (SkillshotLearner.py:512-661).  Pure Python: runs without a GPU."""
import os

import numpy as np
import pytest

from tests.helpers import GOLDEN

INT_KEYS = {"player_x_dir", "player_pos_x", "player_pos_y", "projectile_cooldown", "projectile_x_dir",
            "projectile_pos_x", "projectile_pos_y", "projectile_age"}
BOOL_KEYS = {"projectile_valid", "projectile_future_collision_opponent"}


class _Proj:
    cooldown_max = 15


class _Player:
    projectile = _Proj()


class _Game:
    board_size = (250, 250)

    def get_player_by_id(self, _):
        return _Player()


def _facade():
    from skillshot_learning_b200.learner import SkillshotLearner
    skl = object.__new__(SkillshotLearner)          # no device: only the host helpers are exercised
    skl.game_environment = _Game()
    skl.player_ids = (1, 2)
    skl.max_dist_normaliser = (2 * (250 ** 2)) ** 0.5
    return skl


def _states(g, name):
    from skillshot_learning_b200.game import FEATURE_KEYS
    out = []
    for f, gen in zip(g[name + "_feat"], g[name + "_general"]):
        st = dict(game_live=bool(gen[0]), ticks=int(gen[1]), game_winner=int(gen[2]))
        for p in (1, 2):
            d = {}
            for j, key in enumerate(FEATURE_KEYS):
                v = float(f[p - 1, j])
                d[key] = int(v) if key in INT_KEYS else (bool(v) if key in BOOL_KEYS else v)
            st[p] = d
        out.append(st)
    return out


@pytest.mark.parametrize("name", ["hit", "duel", "random"])
def test_host_helpers_match_the_reference(name):
    g = dict(np.load(os.path.join(GOLDEN, "learner_host.npz")))
    skl = _facade()
    states = _states(g, name)
    for p in (1, 2):
        got = np.array(skl.prepare_states(states, p), dtype=np.float64)
        assert got.tobytes() == g[name + "_prepared"][p - 1].tobytes()          # bit-exact
    post = states[1:]
    for fn, key in ((skl.calculate_rewards_looking, "looking"), (skl.calculate_rewards_simple, "simple"),
                    (skl.calculate_rewards, "shaped")):
        got = np.array([[r[1], r[2]] for r in fn(post)], dtype=np.float64)
        assert got.tobytes() == g["%s_%s" % (name, key)].tobytes(), key
    if name == "hit":
        # game_winner is the id of the player that was HIT (SkillshotGame.py:77); the +1 lands on its
        # row at its own projectile's firing tick (:624-626)
        assert states[-1]["game_winner"] == 1 and g["hit_shaped"][3, 0] == 1.0
    rewards = {1: [1.0, 2.0], 2: [3.0]}
    assert skl.prepare_rewards(rewards, 2) == [3.0]
    acts = {1: [np.array([[0.1, 0.2]])], 2: [np.array([[0.3, 0.4]])]}
    assert np.array_equal(skl.prepare_actions(acts, 2)[0], [0.3, 0.4])
