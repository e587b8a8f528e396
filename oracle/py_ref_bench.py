"""Times the UNMODIFIED Python reference on the host cores:
one SkillshotGame per worker process, `multiprocessing` over the cores, float32 random actions
handed over as Python floats, SkillshotLearner.do_actions for both players + game_tick per env-step,
game_reset(random_positions=True) at a hit or at the 2,000-tick limit (the physics-only workload of
bench.py), and the same with get_state + prepare_states for both players (the rollout's env side).

    python -m oracle.py_ref_bench [--seconds S] [--procs P]     -> one JSON line

TEST / BENCH INFRASTRUCTURE ONLY (bench.py's cpu_baseline and --impl reference legs run it as a
subprocess so that no CUDA context is forked).  Needs /root/reference or oracle/_ref (build_ref.py).
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TICK_LIMIT = 2000          # SkillshotLearner.py:62


def _worker(args):
    seed, seconds, with_obs = args
    import numpy as np
    from oracle import ref_harness
    SkillshotGame, _ = ref_harness.modules()
    rng = np.random.default_rng(seed)
    actions = rng.uniform(-1.2, 1.2, size=(4096, 2, 2)).astype(np.float32).tolist()   # Python floats of float32 values
    sink = io.StringIO()
    steps = 0
    with contextlib.redirect_stdout(sink):
        np.random.seed(seed)
        g = SkillshotGame(random_positions=True)
        skl = ref_harness.make_learner(g)
        t0 = time.perf_counter()
        deadline = t0 + seconds
        while True:
            for a in actions:
                skl.do_actions(1, a[0])
                skl.do_actions(2, a[1])
                g.game_tick()
                if with_obs:
                    st = g.get_state()
                    skl.prepare_states([st], 1)
                    skl.prepare_states([st], 2)
                if not g.game_live or g.ticks >= TICK_LIMIT:
                    g.game_reset(random_positions=True)
                    skl.game_environment = g
            steps += len(actions)
            sink.seek(0); sink.truncate(0)
            if time.perf_counter() >= deadline:
                break
        dt = time.perf_counter() - t0
    return steps, dt


def measure(seconds: float, procs: int):
    from oracle import ref_harness
    out = {"source": ref_harness.source(), "procs": procs}
    ctx = mp.get_context("fork")
    for key, with_obs in (("tick", False), ("tick_get_state_prepare_states", True)):
        with ctx.Pool(procs) as pool:
            res = pool.map(_worker, [(1000 + i, seconds, with_obs) for i in range(procs)])
        rates = [s / dt for s, dt in res]
        out[key] = {"env_steps_per_sec": sum(rates), "env_steps_per_sec_per_core": sum(rates) / procs,
                    "seconds": max(dt for _, dt in res)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=2.0)
    ap.add_argument("--procs", type=int, default=len(os.sched_getaffinity(0)))
    a = ap.parse_args()
    print(json.dumps(measure(a.seconds, a.procs)))


if __name__ == "__main__":
    main()
