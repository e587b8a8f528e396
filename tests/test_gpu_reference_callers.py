"""The drop-in proof: the reference's OWN caller code, executed verbatim, on this package's objects.

`requires_reference` + `gpu`: the unmodified reference modules come from /root/reference or from their byte-code in
oracle/_ref (which travels to the GPU box).  Two callers exist in the reference:
  * SkillshotLearner.do_actions / prepare_states / calculate_rewards_looking / calculate_rewards_simple /
    calculate_rewards (SkillshotLearner.py:206-213, 512-661) -- bound methods of a reference SkillshotLearner whose
    game_environment is a skillshot_learning_b200.SkillshotGame (a view of one device env);
  * the per-frame key-press block of skillshot_playable.py:51-64, compiled from the reference source, with the reference's
    InputHandler, driving that same facade.
Each runs in lockstep on a reference SkillshotGame and on the facade; every mutable field must be EQUAL after every call,
get_state features agree to 1e-12 (device libm vs glibc), prepare_states rows to 1e-12, boards cell for cell."""
import contextlib
import io

import numpy as np
import pytest

from oracle import ref_harness
from tests import parity

pytestmark = [pytest.mark.gpu, pytest.mark.requires_reference]

EXACT_KEYS = ("player_x_dir", "player_pos_x", "player_pos_y", "player_rotation", "projectile_cooldown", "projectile_x_dir",
              "projectile_pos_x", "projectile_pos_y", "projectile_rotation", "projectile_age", "projectile_valid")
CLOSE_KEYS = ("player_grad", "player_path_dist_opponent", "player_dist_opponent", "projectile_grad",
              "projectile_path_dist_opponent", "projectile_dist_opponent")


def _pair(positions=None, rotations=None):
    """(reference game, facade game) in the same state."""
    from skillshot_learning_b200 import SkillshotGame
    ref = ref_harness.make_game(positions, rotations)
    fac = SkillshotGame(device="cuda:0")
    if positions is not None:
        fac.player1.pos = [positions[0], positions[1]]
        fac.player2.pos = [positions[2], positions[3]]
    if rotations is not None:
        fac.player1.rotation, fac.player2.rotation = float(rotations[0]), float(rotations[1])
    return ref, fac


def _assert_same_state(ref, fac, msg):
    a, b = ref_harness.read_state(ref), ref_harness.read_state(fac)
    assert a == b, (msg, a, b)          # ints equal, float64 rotations bit-equal (== on Python floats)


def _ambiguous(sd, pid):
    """The future-collision flag of this view is decided by rounding noise in the reference itself (tests/parity.py)."""
    opp = 3 - pid
    st = dict(qx=np.array([[sd[1]["projectile_pos_x"], sd[2]["projectile_pos_x"]]]), qy=np.array([[sd[1]["projectile_pos_y"], sd[2]["projectile_pos_y"]]]),
              px=np.array([[sd[1]["player_pos_x"], sd[2]["player_pos_x"]]]), py=np.array([[sd[1]["player_pos_y"], sd[2]["player_pos_y"]]]),
              valid=np.array([[int(sd[1]["projectile_valid"]), int(sd[2]["projectile_valid"])]]),
              qrot=np.array([[sd[1]["projectile_rotation"], sd[2]["projectile_rotation"]]]))
    grad = np.array([[sd[1]["projectile_grad"], sd[2]["projectile_grad"]]])
    return bool(parity.ambiguous_future_collision(st, grad, count=False)[0, pid - 1])


def _assert_same_get_state(sr, sf, msg):
    assert (sr["game_live"], sr["ticks"], sr["game_winner"]) == (sf["game_live"], sf["ticks"], sf["game_winner"]), msg
    for pid in (1, 2):
        for k in EXACT_KEYS:
            assert sr[pid][k] == sf[pid][k], (msg, pid, k, sr[pid][k], sf[pid][k])
        for k in CLOSE_KEYS:
            np.testing.assert_allclose(sf[pid][k], sr[pid][k], rtol=1e-12, atol=1e-9, err_msg=f"{msg} {pid} {k}")
        if not _ambiguous(sr, pid):
            k = "projectile_future_collision_opponent"
            assert sr[pid][k] == sf[pid][k], (msg, pid, k)


@pytest.mark.parametrize("case", [("fixed", None, None), ("random", (131, 77, 160, 92), (0.4, 3.3)),
                                  ("close", (100, 100, 104, 121), (3.1, 0.05))])
def test_reference_learner_methods_run_verbatim_on_the_facade(case):
    name, positions, rotations = case
    ref, fac = _pair(positions, rotations)
    skl_ref, skl_fac = ref_harness.make_learner(ref), ref_harness.make_learner(fac)   # the reference class, twice
    rng = np.random.default_rng(len(name))
    T = 48
    actions = (rng.uniform(-1.2, 1.2, size=(T, 2, 2))).astype(np.float32)
    actions[5] = [[1.0, -1.0], [0.5, 0.0]]
    if name == "close":
        actions[..., 1] *= 0.2
    sink = io.StringIO()
    states_ref, states_fac = [], []
    with contextlib.redirect_stdout(sink):
        for t in range(T):
            for pid in (1, 2):
                pred = (float(actions[t, pid - 1, 0]), float(actions[t, pid - 1, 1]))
                skl_ref.do_actions(pid, pred)                                          # SkillshotLearner.py:206-213, verbatim
                skl_fac.do_actions(pid, pred)
                _assert_same_state(ref, fac, f"{name} t={t} after do_actions({pid})")
            ref.game_tick()
            fac.game_tick()
            _assert_same_state(ref, fac, f"{name} t={t} after game_tick")
            sr, sf = ref.get_state(), fac.get_state()
            _assert_same_get_state(sr, sf, f"{name} t={t}")
            states_ref.append(sr)
            states_fac.append(sf)
        for pid in (1, 2):                                                             # SkillshotLearner.py:512-543, verbatim
            rows_ref = np.array(skl_ref.prepare_states(states_ref, pid), dtype=np.float64)
            rows_fac = np.array(skl_fac.prepare_states(states_fac, pid), dtype=np.float64)
            amb = np.array([_ambiguous(s, pid) for s in states_ref])
            rows_fac[amb, 11] = rows_ref[amb, 11]
            np.testing.assert_array_equal(rows_fac[:, 11], rows_ref[:, 11])
            np.testing.assert_allclose(rows_fac, rows_ref, rtol=1e-12, atol=1e-12)
        # the three reward functions (SkillshotLearner.py:575-661), verbatim on the facade's state dicts
        for fn in ("calculate_rewards_looking", "calculate_rewards_simple"):
            for a, b in zip(getattr(skl_ref, fn)(states_ref), getattr(skl_fac, fn)(states_fac)):
                for pid in (1, 2):
                    np.testing.assert_allclose(b[pid], a[pid], rtol=1e-12, atol=1e-9)
        if not any(_ambiguous(s, pid) for s in states_ref for pid in (1, 2)):
            for a, b in zip(skl_ref.calculate_rewards(states_ref), skl_fac.calculate_rewards(states_fac)):
                for pid in (1, 2):
                    np.testing.assert_allclose(b[pid], a[pid], rtol=1e-12, atol=1e-9)
    if name == "close":
        assert not ref.game_live and ref.winner_id == fac.winner_id != 0               # the scenario ends in a hit


def test_playable_key_press_block_runs_verbatim_on_the_facade():
    """skillshot_playable.py:51-64 with the reference's InputHandler: scripted key-down / key-up events for both players,
    60 frames; state equal after every frame and get_board() (the raster the script draws, :66-69) cell for cell."""
    K = ref_harness.KEYS
    tick = ref_harness.playable_tick()
    ref, fac = _pair()
    ih_ref, ih_fac = ref_harness.input_handler(), ref_harness.input_handler()
    script = {0: [("down", "K_w"), ("down", "K_UP"), ("down", "K_SPACE")], 4: [("down", "K_a")], 9: [("up", "K_a"), ("down", "K_PERIOD")],
              12: [("up", "K_w"), ("down", "K_s"), ("down", "K_RIGHT")], 20: [("up", "K_s"), ("down", "K_d"), ("down", "K_w")],
              27: [("up", "K_RIGHT"), ("down", "K_LEFT"), ("up", "K_SPACE")], 33: [("down", "K_DOWN"), ("up", "K_UP")],
              40: [("up", "K_d"), ("down", "K_SPACE")], 50: [("up", "K_DOWN"), ("up", "K_LEFT"), ("down", "K_UP")]}
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        for frame in range(60):
            for kind, key in script.get(frame, []):
                for ih in (ih_ref, ih_fac):
                    (ih.input_start if kind == "down" else ih.input_stop)(K[key])
            exec(tick, {"inputHandler": ih_ref, "skillshotGame": ref})
            exec(tick, {"inputHandler": ih_fac, "skillshotGame": fac})
            _assert_same_state(ref, fac, f"frame {frame}")
            assert np.array_equal(np.asarray(ref.get_board()), fac.get_board()), frame
    assert ref.ticks == 60 and ref.player1.projectile.age < 60          # shots were fired and re-fired
