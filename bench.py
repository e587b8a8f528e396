#!/usr/bin/env python
"""bench.py -- env-steps/s of the Skillshot hot path on B200 (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): physics-only, 65,536 SkillshotGame envs per
GPU, U(-1.2,1.2) float32 random actions, random starts, terminal +1/-1/0 reward,
2,000-tick limit with auto-reset.  One bench "step" = TICKS ticks of all envs
(TICKS * 65,536 env-steps per GPU), played by ss_env_step in launches of
TICKS_PER_LAUNCH fused ticks.  The action stream of a step ([TICKS, E, 2, 2]
float32 = 2.1 GB) is larger than the 126 MB L2, so every timed iteration streams
it from HBM ("inputs larger than L2"); the 4 MB game state is L2-resident by
design.

  value     whole-job env-steps/s, actions resident in HBM, CUDA-event timed
  e2e       the same through SkillshotEnvs.step_host: pinned HOST actions in,
            reward/done/winner back to pinned HOST memory, copies inside the timing
  roofline  HBM: algorithmic 202 B per env-step x env-steps per
            launch / average launch time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the C oracle port (oracle/skillshot_oracle.c, OpenMP over envs)
            on the host cores, on a bounded sample of the same workload

  learner   the other half of BASELINE.json's metric, on the same GPUs in the same run:
            rollout (configs[2]: 262,144 envs per GPU, tensor-core actor forward with
            parameter noise + env step with observations + replay push) in env-steps/s and
            samples/s, the DDPG update (sample, TD targets, critic step, actor step, one
            gradient all-reduce each) in update samples/s, and the tensor roofline of the
            actor-forward kernel (72,192 algorithmic FLOP per row against bf16_tflops)

--impl reference times that CPU port alone (the Python reference cannot travel
to the GPU box) and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

ENVS_PER_GPU = 65536
TICKS = 2048                # ticks per bench step (one full 2,000-tick episode plus the auto-reset)
TICKS_PER_LAUNCH = 128      # fused ticks per ss_env_step launch (32: 4.8e10, 128: 5.1e10 env-steps/s)
E2E_TICKS_PER_LAUNCH = 32   # the host-buffer leg pipelines copy in / kernel / copy out per chunk: finer chunks overlap better
TICK_LIMIT = 2000           # SkillshotLearner.py:62
ALGO_BYTES_PER_ENV_STEP = 202   # SURVEY.md 8(d), physics-only
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
WORKLOAD = "physics-only: 65,536 SkillshotGame envs per GPU, random actions, terminal reward, 2000-tick auto-reset"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def ncu_traffic(ticks_per_launch=None):
    """dram bytes per launch of the dominant kernel from the committed ncu capture (taken at the default number of fused
    ticks per launch), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            t = json.load(open(p))
            if ticks_per_launch is None or t.get("ticks_per_launch", ticks_per_launch) == ticks_per_launch:
                return t.get("step_kernel_physics_bytes_per_launch")
        except Exception:
            pass
    return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index: int, period: float = 0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------
# CPU port (oracle) timing -- cpu_baseline leg and --impl reference arm
# ---------------------------------------------------------------------------
def cpu_port_run(n_envs: int, ticks: int, nthreads: int, seed: int = 0, envs=None, actions=None):
    """Times `ticks` ticks of n_envs oracle envs; returns (seconds, envs, actions)."""
    from oracle.oracle import OracleEnvs
    rng = np.random.default_rng(seed)
    if envs is None:
        envs = OracleEnvs(n_envs, rng.integers(25, 225, size=(n_envs, 4)))
    if actions is None:
        actions = rng.uniform(-1.2, 1.2, size=(8, n_envs, 2, 2)).astype(np.float32)
    t0 = time.perf_counter()
    for t in range(ticks):
        envs.step(actions[t % actions.shape[0]], want_obs=False, reward_mode=2, tick_limit=TICK_LIMIT,
                  auto_reset=True, nthreads=nthreads)
    return time.perf_counter() - t0, envs, actions


def cpu_baseline(target_seconds: float = 4.0):
    """Oracle port on all host cores, bounded sample (~10-30 s of CPU work)."""
    from oracle.oracle import lib as olib
    cores = int(olib().ss_oracle_max_threads())
    dt, envs, actions = cpu_port_run(ENVS_PER_GPU, 4, cores)            # warm-up + calibration
    ticks = int(max(8, min(4096, target_seconds / max(dt / 4, 1e-6))))
    dt, _, _ = cpu_port_run(ENVS_PER_GPU, ticks, cores, envs=envs, actions=actions)
    out = {"value": ENVS_PER_GPU * ticks / dt, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": "%d envs x %d ticks (%.1f s wall, %d OpenMP threads) of the same workload, C port of the "
                     "Python reference (oracle/skillshot_oracle.c)" % (ENVS_PER_GPU, ticks, dt, cores)}
    # SURVEY.md 8(d) config 1, second figure: tick + get_state + prepare_states (the rollout's env side)
    rng = np.random.default_rng(1)
    t_obs = max(4, ticks // 8)
    t0 = time.perf_counter()
    for t in range(t_obs):
        envs.step(actions[t % actions.shape[0]], want_obs=True, reward_mode=1, tick_limit=TICK_LIMIT, auto_reset=True, nthreads=cores)
    out["with_observations_env_steps_per_sec"] = ENVS_PER_GPU * t_obs / (time.perf_counter() - t0)
    out["learner_update_batch16"] = cpu_learner_baseline()
    return out


def cpu_learner_baseline(target_seconds: float = 3.0):
    """The reference's own update on the host: Keras-semantics restatement in torch-CPU (oracle/learner_oracle.py; TensorFlow is
    not installable here), batches of 16 as model_param_batch_size (SkillshotLearner.py:61): critic fit step + actor fit step."""
    import torch
    from oracle import learner_oracle as lo
    rng = np.random.default_rng(0)
    L = lo.LearnerOracle(lo.init_actor(rng), lo.init_critic(rng))
    n = 16 * 64
    s = rng.uniform(0, 1, (n, 12)).astype(np.float32); a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    y = -rng.uniform(0, 1, n).astype(np.float32); keep = (rng.uniform(size=(n, 256)) >= 0.2).astype(np.float32)
    order = np.arange(n)
    L.critic_fit(s[:32], a[:32], y[:32], order[:32], keep[:32]); L.actor_fit(s[:32])          # warm-up
    done, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < target_seconds:
        L.critic_fit(s, a, y, order, keep)
        L.actor_fit(s)
        done += n
    dt = time.perf_counter() - t0
    return {"samples_per_sec": done / dt, "threads": int(torch.get_num_threads()),
            "sample": "%d rows in batches of 16 (%.1f s): critic fit step + actor fit step per batch, torch-CPU restatement of "
                      "the Keras update (the reference's TensorFlow is not installable here)" % (done, dt)}


ROLLOUT_ENVS = 262144          # BASELINE.json configs[2]
TRAIN_BATCH = 65536            # update rows per GPU per step
SM_BATCH = 148 * 512           # the same sized to the machine: 4 tiles of 128 rows per SM, no partial round
ACTOR_FLOP_PER_ROW = 72192     # 2 * (12*256 + 256*128 + 128*2)
UPDATE_FLOP_PER_ROW = 638976   # full DDPG update with target actor + critic forward on s'


def learner_legs(dev, rank, world, seed, peaks, collective, ticks=64, updates=20):
    """Rollout, DDPG update and the actor-forward tensor roofline on this rank's GPU.
    Returns per-rank times in ms: (rollout per tick, update per step, actor forward per launch)."""
    import torch
    import torch.distributed as dist
    from skillshot_learning_b200 import SelfPlayTrainer

    E = ROLLOUT_ENVS
    group = -(-(2 * E // 128) // 148) * 128       # one parameter-noise draw per SM-sized slice of the batch (28 tiles)
    tr = SelfPlayTrainer(E, device=dev, seed=seed, replay_capacity=2 * E * 4, batch_size=TRAIN_BATCH,
                         gamma=0.99, tau=0.005, param_noise_sd=0.5, noise_group=group, reward_mode="looking",
                         tick_limit=TICK_LIMIT, process_group=True if world > 1 else None, precision="bf16",
                         collective=collective)

    def timed(fn, iters, warm=3):
        for _ in range(warm):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / iters

    chunk = 16                                     # ticks enqueued per library call (ss_selfplay_rollout)
    t_roll = timed(lambda: tr.rollout(chunk), max(1, ticks // chunk)) / chunk
    t_upd = timed(tr.update, updates)                 # gradient GEMMs on tcgen05 (bf16 operands, f32 accumulate)
    tr.networks.update_precision = "f32"
    t_upd32 = timed(tr.update, max(3, updates // 4))  # the exact float32 kernels, same schedule
    tr.networks.update_precision = "bf16"
    tr.batch_size, tr._batch = 2 * E, None            # one update on as many rows as one rollout tick produces
    t_upd_big = timed(tr.update, max(3, updates // 2))
    tr.batch_size, tr._batch = SM_BATCH, None         # a whole number of 128-row tiles on every SM
    t_upd_sm = timed(tr.update, updates)
    tr.batch_size, tr._batch = TRAIN_BATCH, None
    obs, act = tr.obs.view(-1, 12), tr.actions.view(-1, 2)
    t_fwd = timed(lambda: tr.networks.actor_forward(obs, out=act, precision="bf16"), 50)
    tr.envs.check_status()

    # ---- BASELINE.json configs[4]: 20-frame planning actor + randomised per-env game speeds ----
    from skillshot_learning_b200 import FrameStackActor, SkillshotEnvs
    E5, F5 = 65536, 20
    envs5 = SkillshotEnvs(E5, device=dev, random_positions=True, seed=seed + 7, reward_mode="looking", tick_limit=TICK_LIMIT,
                          auto_reset=True)
    g5 = torch.Generator(device=dev).manual_seed(seed + 8)
    scale = lambda: 0.5 + 1.5 * torch.rand(E5, device=dev, generator=g5)          # U(0.5, 2) x the reference constants
    envs5.set_speeds(3.0 * scale(), 0.25 * scale(), 5.0 * scale(), (15.0 * scale()).round().clamp(min=1))
    actor5 = FrameStackActor(2 * E5, frames=F5, device=dev, seed=seed, precision="bf16")
    actor5.push(envs5.observe().reshape(-1, 12))
    act5 = torch.empty((E5, 2, 2), dtype=torch.float32, device=dev)

    def tick5():
        actor5.forward(param_noise_sd=0.5, noise_group=1024, out=act5.view(-1, 2))
        out = envs5.step(act5)
        actor5.push(out["obs"].reshape(-1, 12), out["done"], done_div=2)

    t_cfg5 = timed(tick5, 32)
    envs5.check_status()

    # ---- BASELINE.json configs[3]: full self-play training, 1,048,576 envs in total sharded over the GPUs ----
    del tr, envs5, actor5
    torch.cuda.empty_cache()
    E4 = 1048576 // world
    tr4 = SelfPlayTrainer(E4, device=dev, seed=seed + 11, replay_capacity=2 * E4 * 2, batch_size=TRAIN_BATCH, process_group=True if world > 1 else None,
                          gamma=0.99, tau=0.005, param_noise_sd=0.5, noise_group=-(-(2 * E4 // 128) // 148) * 128,
                          reward_mode="looking", tick_limit=TICK_LIMIT, precision="bf16", collective=collective)

    def iteration():
        tr4.rollout(1)
        tr4.update()

    t_cfg4 = timed(iteration, 16)
    tr4.envs.check_status()
    return t_roll, t_upd, t_fwd, t_upd32, t_upd_big, t_cfg5, t_upd_sm, t_cfg4


def learner_report(t_roll, t_upd, t_fwd, t_upd32, t_upd_big, t_cfg5, t_upd_sm, t_cfg4, world, peaks, peak_kind, collective="nccl"):
    rows = 2 * ROLLOUT_ENVS
    tf = ACTOR_FLOP_PER_ROW * rows / (t_fwd * 1e-3) / 1e12
    return {
        "rollout": {"workload": "262,144 envs per GPU: bf16 tensor-core actor forward on 524,288 observations with "
                                "parameter noise (sd 0.5) + env step with observations and looking reward, transitions produced "
                                "in place in the replay ring; 16 ticks per ss_selfplay_rollout call",
                    "env_steps_per_sec": world * ROLLOUT_ENVS / (t_roll * 1e-3),
                    "samples_per_sec": world * rows / (t_roll * 1e-3), "ms_per_tick": t_roll},
        "train": {"workload": "DDPG update, %d rows per GPU: replay sample, TD targets (gamma 0.99), critic step "
                              "(dropout 0.2), actor step, Adam + soft update (tau 0.005), two flat-gradient exchanges (%s)"
                              % (TRAIN_BATCH, "single GPU: none" if world == 1 else
                                 ("fused NVLink peer-memory reduce-push + Adam" if collective == "peer" else "NCCL all-reduce")),
                  "samples_per_sec": world * TRAIN_BATCH / (t_upd * 1e-3), "ms_per_update": t_upd,
                  "dtype": "bf16 operands, f32 accumulate (tcgen05); Adam and parameters f32",
                  "algorithmic_tflops": world * TRAIN_BATCH * UPDATE_FLOP_PER_ROW / (t_upd * 1e-3) / 1e12,
                  "f32_path_samples_per_sec": world * TRAIN_BATCH / (t_upd32 * 1e-3),
                  "rows_75776_per_gpu": {"note": "148 SMs x 512 rows: every SM gets 4 whole tiles (65,536 rows are 3.46 per SM, "
                                                 "i.e. 4 rounds with a partial one)",
                                         "samples_per_sec": world * SM_BATCH / (t_upd_sm * 1e-3), "ms_per_update": t_upd_sm,
                                         "algorithmic_tflops": world * SM_BATCH * UPDATE_FLOP_PER_ROW / (t_upd_sm * 1e-3) / 1e12},
                  "rows_524288_per_gpu": {"samples_per_sec": world * rows / (t_upd_big * 1e-3), "ms_per_update": t_upd_big,
                                          "algorithmic_tflops": world * rows * UPDATE_FLOP_PER_ROW / (t_upd_big * 1e-3) / 1e12}},
        "selfplay_training": {
            "workload": "BASELINE.json configs[3]: 1,048,576 envs in total (%d per GPU), one iteration = one rollout tick of every env "
                        "(tensor-core actor, parameter noise, transitions into the replay ring) + one %d-row-per-GPU DDPG update; "
                        "strong scaling (the total is fixed)" % (1048576 // world, TRAIN_BATCH),
            "env_steps_per_sec": 1048576 / (t_cfg4 * 1e-3), "update_samples_per_sec": world * TRAIN_BATCH / (t_cfg4 * 1e-3),
            "ms_per_iteration": t_cfg4},
        "planning_actor_speed_sweep": {
            "workload": "BASELINE.json configs[4] (no reference code): 65,536 envs per GPU with per-env speed constants "
                        "U(0.5, 2) x the reference's, 20-frame stacked-observation actor (240 -> 256 -> 128 -> 2, tcgen05 "
                        "kernels) with parameter noise (sd 0.5, one draw per 1,024 rows) + env step + frame-stack push",
            "env_steps_per_sec": world * 65536 / (t_cfg5 * 1e-3), "samples_per_sec": world * 131072 / (t_cfg5 * 1e-3),
            "ms_per_tick": t_cfg5},
        "actor_forward_roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                                   "frac": tf / peaks["bf16_tflops"], "traffic": None, "peak_source": peak_kind,
                                   "kernel": "actor_fwd_tc_kernel", "rows_per_launch": rows,
                                   "algorithmic_flop_per_row": ACTOR_FLOP_PER_ROW, "launch_us": t_fwd * 1e3,
                                   "dtype": "bf16 operands, f32 accumulate"},
    }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle.oracle import lib as olib
    cores = int(olib().ss_oracle_max_threads())
    ticks_per_step = 64
    _, envs, actions = cpu_port_run(ENVS_PER_GPU, 2, cores)
    for _ in range(args.warmup):
        cpu_port_run(ENVS_PER_GPU, ticks_per_step, cores, envs=envs, actions=actions)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_run(ENVS_PER_GPU, ticks_per_step, cores, envs=envs, actions=actions)
    dt = time.perf_counter() - t0
    value = ENVS_PER_GPU * ticks_per_step * args.steps / dt
    sample = "each step = %d envs x %d ticks of the workload on %d OpenMP threads" % (ENVS_PER_GPU, ticks_per_step, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_gpu": ENVS_PER_GPU, "ticks_per_step": ticks_per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def bind_to_gpu_numa(index: int):
    """Pin this process to the CPUs NVML reports as local to GPU `index`, BEFORE the pinned host buffers of the
    end-to-end leg are allocated (first touch then lands on the GPU's own NUMA node).  With 8 ranks streaming
    3.5 GB per step each through host memory, remote-socket buffers halve the end-to-end rate."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from skillshot_learning_b200 import SkillshotEnvs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the GPU arm has no CPU fallback)")
    numa_cpus = bind_to_gpu_numa(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    E, T, KF = ENVS_PER_GPU, args.ticks, args.ticks_per_launch
    assert T % KF == 0
    launches_per_step = T // KF
    envs = SkillshotEnvs(E, device=dev, random_positions=True, seed=1234 + rank, reward_mode="terminal",
                         tick_limit=TICK_LIMIT, auto_reset=True)
    gen = torch.Generator(device=dev)
    gen.manual_seed(99 + rank)
    actions = (torch.rand((T, E, 2, 2), device=dev, generator=gen) * 2.4 - 1.2).contiguous()

    def device_step():
        for c in range(launches_per_step):
            envs.step(actions[c * KF:(c + 1) * KF], want_obs=False)

    # ---- device-resident throughput (value) ----
    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        device_step()
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    envs.check_status()

    # ---- end to end through the host-buffer API (e2e) ----
    e2e_steps, e2e_s = max(3, min(args.steps, 10)), float("nan")
    if not args.no_e2e:
        host_actions = torch.empty((T, E, 2, 2), dtype=torch.float32, pin_memory=True)
        host_actions.copy_(actions)
        host_out = envs.alloc_host_outputs(T)
        KE = E2E_TICKS_PER_LAUNCH if T % E2E_TICKS_PER_LAUNCH == 0 else KF
        for _ in range(2):
            envs.step_host(host_actions, host_out, ticks_per_launch=KE)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            envs.step_host(host_actions, host_out, ticks_per_launch=KE)   # synchronises before returning
        barrier()
        e2e_s = time.perf_counter() - t0

    # ---- learner legs: rollout, DDPG update, tensor roofline of the actor forward ----
    lt = (float("nan"),) * 8
    if not args.no_learner:
        del actions
        torch.cuda.empty_cache()
        lt = learner_legs(dev, rank, world, 4321 + rank, measured_peaks()[0], args.collective)

    if world > 1:
        t = torch.tensor([ms, e2e_s, *lt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s, lt = float(t[0]), float(t[1]), tuple(float(x) for x in t[2:])

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        total_env_steps = world * E * T * args.steps
        value = total_env_steps / (ms * 1e-3)
        launch_s = (ms * 1e-3) / (args.steps * launches_per_step)
        achieved = ALGO_BYTES_PER_ENV_STEP * E * KF / launch_s / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": E, "ticks_per_step": T, "ticks_per_launch": KF,
                       "l2": "action stream per step (%.0f MB) exceeds the 126 MB L2; game state (4 MB) is L2-resident by design"
                             % (T * E * 16 / 1e6)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": ncu_traffic(KF), "peak_source": peak_kind,
                         "kernel": "step_kernel<OBS=false,SPEEDS=false>",
                         "algorithmic_bytes_per_env_step": ALGO_BYTES_PER_ENV_STEP,
                         "moved_bytes_per_env_step": 16 + 8 + 2 + 128.0 / KF,
                         "launch_us": launch_s * 1e6},
            "e2e": {"value": None if args.no_e2e else world * E * T * e2e_steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": T * E * 16, "d2h_bytes_per_step": T * E * 10,
                    "ticks_per_launch": E2E_TICKS_PER_LAUNCH if T % E2E_TICKS_PER_LAUNCH == 0 else KF,
                    "note": "PCIe-bound: 26 B per env-step cross the bus (float32 actions in; reward, done, winner out); "
                            "rank processes pinned to their GPU's NUMA node (%d CPUs) before the pinned buffers are allocated"
                            % numa_cpus},
            "gpu_launches": args.steps * launches_per_step,
            "clocks": clocks,
        }
        if not args.no_learner:
            line["learner"] = learner_report(*lt, world, peaks, peak_kind, args.collective)
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ticks", type=int, default=TICKS)
    ap.add_argument("--ticks-per-launch", type=int, default=TICKS_PER_LAUNCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    ap.add_argument("--no-learner", action="store_true", help="skip the rollout / update / actor-forward legs")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="gradient exchange of the update at N > 1: fused NVLink peer-memory kernels, or an NCCL all-reduce")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
