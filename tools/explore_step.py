"""Sweep of ss_env_step launch times on one GPU (not a bench: exploration for tuning).
Prints one line per configuration: envs, ticks/launch, obs, reward mode, us/launch,
env-steps/s, algorithmic GB/s (202 or 298 B per env-step) and moved GB/s."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from skillshot_learning_b200 import SkillshotEnvs

dev = torch.device("cuda:0")


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters   # us


def main():
    print("envs ticks obs reward us_per_launch env_steps_per_s algo_GBs moved_GBs")
    for E in (65536, 1048576):
        for K in (1, 32):
            if E * K * 16 > 3e9:
                continue
            for obs, reward in ((False, "terminal"), (True, "looking"), (False, "looking"), (True, "terminal")):
                envs = SkillshotEnvs(E, device=dev, random_positions=True, seed=1, reward_mode=reward,
                                     tick_limit=2000, auto_reset=True)
                nbuf = max(1, min(8, int(3e8 // (E * K * 16)) ))     # rotate > L2 worth of actions when cheap
                acts = [(torch.rand((K, E, 2, 2), device=dev) * 2.4 - 1.2) for _ in range(nbuf)]
                i = [0]

                def fn():
                    envs.step(acts[i[0] % nbuf], want_obs=obs)
                    i[0] += 1
                iters = max(5, min(200, int(2e8 // (E * K))))
                us = timeit(fn, iters)
                steps = E * K / (us * 1e-6)
                algo = (298 if obs else 202)
                moved = 16 + 10 + 128.0 / K + (96.0 / K if obs else 0)
                print(E, K, int(obs), reward, "%.2f" % us, "%.3e" % steps, "%.0f" % (steps * algo / 1e9),
                      "%.0f" % (steps * moved / 1e9), flush=True)
                del envs, acts


if __name__ == "__main__":
    main()
