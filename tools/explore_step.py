"""Sweep of ss_env_step launch times on one GPU (not a bench: exploration for tuning).
Launches are captured in a CUDA graph (20 per replay) so that Python/ctypes overhead
is not in the numbers.  One line per configuration."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from skillshot_learning_b200 import SkillshotEnvs

dev = torch.device("cuda:0")
PER_GRAPH = 20


def main():
    Es = [int(x) for x in os.environ.get("SS_E", "65536,262144,1048576").split(",")]
    Ks = [int(x) for x in os.environ.get("SS_K", "1,32").split(",")]
    print("envs ticks obs reward us_per_launch env_steps_per_s algo_GBs frac_of_6556 moved_GBs")
    for E in Es:
        for K in Ks:
            combos = ((False, "terminal"), (True, "looking"), (False, "looking"), (True, "terminal"))
            if os.environ.get("SS_ONLY") == "physics":
                combos = ((False, "terminal"),)
            for obs, reward in combos:
                envs = SkillshotEnvs(E, device=dev, random_positions=True, seed=1, reward_mode=reward,
                                     tick_limit=2000, auto_reset=True)
                nbuf = max(1, min(PER_GRAPH, int(4e8 // (E * K * 16))))
                acts = [(torch.rand((K, E, 2, 2), device=dev) * 2.4 - 1.2) for _ in range(nbuf)]
                s = torch.cuda.Stream()
                with torch.cuda.stream(s):
                    for j in range(3):
                        envs.step(acts[j % nbuf], want_obs=obs)
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=s):
                        for j in range(PER_GRAPH):
                            envs.step(acts[j % nbuf], want_obs=obs)
                    reps = max(2, min(50, int(4e8 // (E * K * PER_GRAPH))))
                    g.replay()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(s)
                    for _ in range(reps):
                        g.replay()
                    e1.record(s)
                    torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / (reps * PER_GRAPH)
                steps = E * K / (us * 1e-6)
                algo = (298 if obs else 202)
                moved = 16 + 10 + 128.0 / K + (96.0 / K if obs else 0)
                print(E, K, int(obs), reward, "%.2f" % us, "%.3e" % steps, "%.0f" % (steps * algo / 1e9),
                      "%.2f" % (steps * algo / 6556.5e9), "%.0f" % (steps * moved / 1e9), flush=True)
                del envs, acts, g


if __name__ == "__main__":
    main()
