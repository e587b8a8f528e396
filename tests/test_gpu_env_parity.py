"""GPU parity tests proper: the CUDA library (through the C ABI, via the
SkillshotEnvs facade) against the golden vectors and the oracle.  Same checks as
tests/test_hostsim_parity.py, plus full-size property tests."""
import numpy as np
import pytest

from tests import parity
from tests.helpers import GOLDEN_FILES

pytestmark = pytest.mark.gpu


def make(n, **kw):
    from skillshot_learning_b200.game import SkillshotEnvs
    return SkillshotEnvs(n, device="cuda:0", **kw)


@pytest.mark.parametrize("name", GOLDEN_FILES)
def test_golden_lockstep(name):
    parity.check_golden_lockstep(make, name)


@pytest.mark.parametrize("name", ["lockstep_random", "close_hits"])
def test_golden_fused_ticks(name):
    parity.check_golden_fused(make, name, K=8)


def test_oracle_lockstep_random():
    parity.check_oracle_lockstep(make, n=4096, T=256, seed=1, compare_every=4)


def test_oracle_lockstep_hits_terminal_reward():
    hits = parity.check_oracle_lockstep(make, n=4096, T=64, seed=2, close=True, reward_mode="terminal")
    assert hits > 400


def test_oracle_lockstep_fused_chunks():
    parity.check_oracle_lockstep(make, n=2048, T=128, seed=3, close=True, reward_mode="simple", chunk=16)


def test_oracle_lockstep_ragged_size():
    # n not a multiple of the warp / CTA size exercises the tail lanes of the coalesced obs store
    parity.check_oracle_lockstep(make, n=1000 + 37, T=32, seed=4)


def test_oracle_lockstep_fused_ragged_sizes():
    """Fused physics ticks (the one-thread-per-player kernel) on env counts that leave a warp / CTA partly filled: lanes past
    the end replay the last env, and nothing may leak into its outputs."""
    for n, seed in ((1, 21), (5, 22), (1000 + 37, 23)):
        parity.check_oracle_lockstep(make, n=n, T=48, seed=seed, close=True, reward_mode="terminal", chunk=8)
    parity.check_oracle_lockstep(make, n=333, T=33 * 3, seed=24, close=True, reward_mode="none", chunk=3)   # odd fused count


def test_single_env():
    parity.check_oracle_lockstep(make, n=1, T=64, seed=6)


def test_auto_reset():
    parity.check_auto_reset(make)


def test_random_reset():
    parity.check_random_reset_properties(make)


def test_speeds():
    parity.check_speeds(make)


def test_nan_action_raises_like_reference():
    e = make(4)
    a = np.zeros((4, 2, 2), np.float32)
    a[2, 1, 0] = np.nan
    import torch
    e.step(torch.from_numpy(a))
    with pytest.raises(ValueError):
        e.check_status()


def test_nan_action_raises_like_reference_fused():
    """The same through the fused physics-only kernel (one thread per player, rotations turned one tick ahead): a NaN move
    raises on its own tick; a NaN look makes the rotation NaN and raises on the NEXT tick's move (int(round(nan)),
    Player.py:63) -- so not at all when it arrives on the last tick of the run (no projectile is fired on tick 1: the
    cooldown set on tick 0 is still running, so Projectile.move_forwards does not see the NaN either)."""
    import torch
    for component, n_ticks, raises in ((0, 4, True), (1, 4, True), (1, 2, False)):
        e = make(6, reward_mode="terminal")
        a = np.zeros((n_ticks, 6, 2, 2), np.float32)
        a[1, 3, 1, component] = np.nan                  # tick 1, env 3, player 2: move (0) or look (1)
        e.step(torch.from_numpy(a), want_obs=False)
        if raises:
            with pytest.raises(ValueError):
                e.check_status()
        else:
            e.check_status()


def test_tan_half_pi_matches_libm():
    """get_gradient_dir at rotation 0 is tan(pi/2) = 1.633123935319537e16 in the reference
    (initial get_state); the projectile's future-collision flag depends on these bits."""
    e = make(2)
    feat, _, _ = e.features()
    f = feat.cpu().numpy()
    assert f[0, 0, 0] == 1.633123935319537e16 and f[0, 0, 8] == 1.633123935319537e16
    assert f[0, 0, 2] == 150.0 and f[0, 0, 3] == 212.13203435596427


def test_full_size_physics_matches_oracle_256_ticks():
    """BASELINE config 2: 65,536 envs, random starts; the first 256 ticks of 1,024 envs
    spread over the batch are compared exactly with the oracle, and every env's final
    discrete state is checked against the oracle after 64 ticks (sizes the oracle
    finishes in seconds)."""
    import torch
    from oracle.oracle import OracleEnvs
    n, T = 65536, 64
    rng = np.random.default_rng(123)
    pos = rng.integers(25, 225, size=(n, 4))
    envs = make(n, reward_mode="terminal")
    envs.reset(positions=pos)
    orc = OracleEnvs(n, pos)
    for t in range(T):
        a = parity.random_actions(rng, (n, 2, 2))
        out = envs.step(torch.from_numpy(a), want_obs=False)
        ro = orc.step(a, want_obs=False, reward_mode=2, nthreads=0)
        np.testing.assert_array_equal(out["winner"].cpu().numpy(), ro["winner"])
        np.testing.assert_array_equal(out["reward"].cpu().numpy(), ro["reward"])
    parity.assert_state_equal(envs.export_state(), orc.snapshot(), "65536 envs")


def test_bench_shape_full_size_matches_oracle():
    """bench.py's timed configuration as it is timed: 65,536 envs x 2,048 ticks in launches of 1,024 fused ticks (one
    896-thread CTA per SM), Philox random starts and auto-reset at the 2,000-tick limit, terminal reward.  Every winner,
    done flag and reward of all 1.3e8 env-steps and the state after every launch are compared with the oracle for equality."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(99)

    def actions(K):
        return (torch.rand((K, 65536, 2, 2), device="cuda", generator=g) * 2.4 - 1.2).contiguous()

    import bench
    episodes, hits = parity.check_bench_shape(make, n=65536, T=2048, K=bench.TICKS_PER_LAUNCH, action_source=actions)
    assert episodes >= 65536 and hits > 1000          # every env restarts at least once (the 2,000-tick limit), many by a hit


def test_bench_shape_short_episodes():
    """The same shape with a 40-tick limit: ~50 Philox restarts per env inside the fused launches."""
    episodes, hits = parity.check_bench_shape(make, n=8192, T=256, K=128, tick_limit=40, seed=77)
    assert episodes > 8192 * 5


def test_fused_physics_kernel_is_identical_for_every_cta_size(monkeypatch):
    """step_pp_kernel picks its CTA size from the env count (896 threads = one CTA of 28 warps per SM when that still fills
    the GPU, else 448, else 64; SS_STEP_BLK forces one).  The three sizes must produce the same rewards, flags, packed bytes
    and state on a ragged env count with short episodes (restarts inside the fused launch), and the 64-thread form is the
    one every small oracle test above runs through."""
    import torch
    n, K = 70001, 48
    kw = dict(random_positions=True, seed=31, reward_mode="terminal", tick_limit=19, auto_reset=True)
    g = torch.Generator(device="cuda").manual_seed(8)
    actions = (torch.rand((K, n, 2, 2), device="cuda", generator=g) * 2.4 - 1.2).contiguous()
    host_actions = actions.cpu().pin_memory()
    got = {}
    for blk in ("64", "448", "896"):
        monkeypatch.setenv("SS_STEP_BLK", blk)
        e = make(n, **kw)
        out = e.step(actions, want_obs=False)
        p = make(n, **kw)
        flags = p.step_host(host_actions, p.alloc_host_outputs(K, outputs="flags"), ticks_per_launch=K)["flags"].clone()
        got[blk] = (out["reward"].clone(), out["done"].clone(), out["winner"].clone(), e.export_state(), flags, p.export_state())
        e.check_status()
    assert int(got["64"][1].sum()) > n
    for blk in ("448", "896"):
        for a, b in zip(got["64"][:3], got[blk][:3]):
            assert torch.equal(a, b), blk
        parity.assert_state_equal(got[blk][3], got["64"][3], "state, CTA size " + blk)
        assert torch.equal(got[blk][4], got["64"][4]), blk
        parity.assert_state_equal(got[blk][5], got["64"][5], "packed state, CTA size " + blk)
    parity.assert_state_equal(got["64"][5], got["64"][3], "packed vs three arrays")
    # the one-tick physics kernel has two CTA sizes as well (448 threads when the batch is exactly one wave of one CTA per SM)
    monkeypatch.delenv("SS_STEP_BLK")
    one = {}
    for blk in ("64", "448"):
        monkeypatch.setenv("SS_STEP1_BLK", blk)
        e = make(n, **kw)
        outs = []
        for t in range(40):
            o = e.step(actions[t], want_obs=False)
            outs.append((o["reward"].clone(), o["done"].clone(), o["winner"].clone()))
        one[blk] = (outs, e.export_state())
        e.check_status()
    for (ra, da, wa), (rb, db, wb) in zip(one["64"][0], one["448"][0]):
        assert torch.equal(ra, rb) and torch.equal(da, db) and torch.equal(wa, wb)
    parity.assert_state_equal(one["448"][1], one["64"][1], "one-tick kernel, CTA size 448")
    for t in range(40):                                       # ... and the one-tick launches equal the fused launch's first ticks
        assert torch.equal(one["64"][0][t][2], got["64"][2][t]) and torch.equal(one["64"][0][t][0], got["64"][0][t])


def test_step_host_full_and_packed_outputs_agree():
    """The host-buffer API: pinned host actions in, outputs to pinned host memory, chunks pipelined over streams.  The full
    outputs (reward / done / winner) equal a device-resident run; the packed one-byte output (ss_env_step_packed) decodes to
    the same three arrays and leaves the same state."""
    import torch
    from skillshot_learning_b200.game import SkillshotEnvs
    n, T = 3000, 96
    kw = dict(random_positions=True, seed=21, reward_mode="terminal", tick_limit=25, auto_reset=True)
    ref, full, packed = (make(n, **kw) for _ in range(3))
    g = torch.Generator().manual_seed(4)
    actions = (torch.rand((T, n, 2, 2), generator=g) * 2.4 - 1.2).pin_memory()
    want = ref.step(actions.cuda(), want_obs=False)
    out_full = full.step_host(actions, full.alloc_host_outputs(T), ticks_per_launch=32)
    out_flags = packed.step_host(actions, packed.alloc_host_outputs(T, outputs="flags"), ticks_per_launch=32)
    dec = SkillshotEnvs.unpack_flags(out_flags["flags"])
    for k in ("reward", "done", "winner"):
        np.testing.assert_array_equal(out_full[k].numpy(), want[k].cpu().numpy(), err_msg=k)
        np.testing.assert_array_equal(dec[k], want[k].cpu().numpy(), err_msg="packed " + k)
    assert int(want["done"].sum()) > n and int((want["winner"] != 0).sum()) > 0
    parity.assert_state_equal(full.export_state(), ref.export_state(), "step_host state")
    parity.assert_state_equal(packed.export_state(), ref.export_state(), "packed state")
    packed.check_status()


def test_facade_matches_reference_surface():
    """The SkillshotGame / Player / Projectile object surface on one device env (KAT-A, KAT-D)."""
    from skillshot_learning_b200.game import SkillshotGame
    g = SkillshotGame()
    assert list(g.player1.pos) == [50, 50] and list(g.player2.pos) == [200, 200]
    assert g.ticks == 0 and g.game_live and g.winner_id == 0
    for _ in range(3):
        for pid in (1, 2):
            p = g.get_player_by_id(pid)
            p.move_direction_float(0.0); p.move_look_float(0.0); p.move_shoot_projectile()
        g.game_tick()
    assert list(g.player1.projectile.pos) == [50, 35] and g.player1.projectile.valid
    assert g.player1.projectile.cooldown_current == 12 and g.player1.projectile.age == 3
    st = g.get_state()
    assert st["ticks"] == 3 and st[1]["projectile_pos_y"] == 35 and st[1]["player_grad"] == 1.633123935319537e16
    # KAT-D: wall rejection through attribute writes, as reference callers do
    g.player1.pos[0] = 1; g.player1.pos[1] = 100; g.player1.rotation = np.pi / 4
    g.player1.move_direction_float(1.0)
    assert list(g.player1.pos) == [1, 100]
    g.player1.move_backwards()
    assert list(g.player1.pos) == [3, 102]
    g.player1.move_look_left()
    assert g.player1.rotation == np.pi / 4 + 0.25
    board = g.get_board()
    assert board.shape == (250, 250) and board[4, 103] == 1
    g.game_reset()
    assert g.ticks == 0 and list(g.player1.pos) == [50, 50]


@pytest.mark.parametrize("auto_reset,fused", [(True, 1), (True, 8), (False, 4)])
def test_episode_statistics_match_the_per_tick_done_log(auto_reset, fused):
    """SS_STEP_EPISODE_STATS: the device-side reduction of the reference's per-episode log (ticks and winner of every
    finished game, SkillshotLearner.py:164-180) against the same figures rebuilt on the host from the per-tick done /
    winner outputs.  Without auto-reset a finished game stays done and must be counted once."""
    import torch
    n, T, limit = 3000, 96, 40
    envs = make(n, random_positions=True, seed=3, reward_mode="terminal", tick_limit=limit, auto_reset=auto_reset)
    envs.collect_episode_stats = True
    g = torch.Generator(device="cuda").manual_seed(1)
    age = np.zeros(n, np.int64)
    over = np.zeros(n, bool)
    want = dict(episodes=0, hit1=0, hit2=0, limit=0, ticks=0, hist=np.zeros(64, np.int64))
    width = (limit + 63) // 64
    for _ in range(T // fused):
        a = torch.rand((fused, n, 2, 2), device="cuda", generator=g) * 2.4 - 1.2
        out = envs.step(a, want_obs=False)
        done, winner = out["done"].cpu().numpy().reshape(fused, n), out["winner"].cpu().numpy().reshape(fused, n)
        for t in range(fused):
            age += ~over
            ended = (done[t] != 0) & ~over
            want["episodes"] += int(ended.sum())
            want["hit1"] += int((ended & (winner[t] == 1)).sum())
            want["hit2"] += int((ended & (winner[t] == 2)).sum())
            want["limit"] += int((ended & (winner[t] == 0)).sum())
            want["ticks"] += int(age[ended].sum())
            np.add.at(want["hist"], np.minimum(63, age[ended] // width), 1)
            if auto_reset:
                age[ended] = 0
            else:
                over |= ended
    got = envs.episode_summary()
    assert want["episodes"] > n // 2
    assert (got["episodes"], got["player1_hit"], got["player2_hit"], got["tick_limit"]) == (
        want["episodes"], want["hit1"], want["hit2"], want["limit"])
    assert abs(got["mean_ticks"] * got["episodes"] - want["ticks"]) < 0.5
    assert np.array_equal(got["histogram"], want["hist"]) and got["bin_ticks"] == width
    envs.episode_summary(reset=True)
    assert envs.episode_summary()["episodes"] == 0
