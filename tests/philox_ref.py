"""Host restatement (numpy) of the library's counter-based draws
(skillshot_learning_b200/csrc/ss_rng.cuh, ss_env_core.cuh): Philox4x32-10, the
(seed, tag, a, b, counter) keying, the open-interval uniform and Box-Muller.
Test infrastructure: the GPU tests compare kernels that draw on the device with it."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
TAG_PARAM_NOISE, TAG_DROPOUT, TAG_ACTION_NOISE, TAG_REPLAY = 1, 2, 3, 4
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy arrays of uint32 counters; scalar key."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & MASK, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & MASK, lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def draw4(seed, tag, a, b, counter):
    seed, counter = int(seed), int(counter)
    return philox4x32_10(a, b, counter & 0xFFFFFFFF, tag ^ (counter >> 32), seed & 0xFFFFFFFF, seed >> 32)


def unit_open(x):
    return ((x >> np.uint32(8)).astype(np.float64) + 0.5) / 16777216.0


def normal4(seed, tag, a, b, counter):
    """[..., 4] standard normals (float64 evaluation of the device's float32 Box-Muller)."""
    x, y, z, w = draw4(seed, tag, a, b, counter)
    out = []
    for u, v in ((x, y), (z, w)):
        r = np.sqrt(-2.0 * np.log(unit_open(u)))
        th = 2.0 * np.pi * unit_open(v)
        out += [r * np.cos(th), r * np.sin(th)]
    return np.stack(out, axis=-1)


def param_noise_eps(n_params, seed, group, counter):
    """eps_p for p < n_params as ss_param_noise / ss_actor_forward draw them."""
    quads = (n_params + 3) // 4
    z = normal4(seed, TAG_PARAM_NOISE, np.arange(quads, dtype=np.uint64), np.uint64(group), counter)
    return z.reshape(-1)[:n_params]


def dropout_keep(n, seed, counter, rate, row_offset=0):
    """keep mask uint8 [n, 256] of the Philox dropout of ss_critic_grad / ss_critic_grad_tc: one draw per
    (global row, 8 units), 16 bits per unit, kept when the 16-bit value >= floor(rate * 65536)."""
    thresh = np.uint32(int(np.float32(rate) * np.float32(65536.0)))
    rows = (np.arange(n) + row_offset).astype(np.uint64)
    chunks = np.arange(32, dtype=np.uint64)
    u = draw4(seed, TAG_DROPOUT, rows[:, None], chunks[None, :], counter)          # 4 x [n, 32]
    keep = np.zeros((n, 256), np.uint8)
    for e in range(8):
        w = u[e >> 1]
        v = (w >> np.uint32(16)) if (e & 1) else (w & np.uint32(0xFFFF))
        keep[:, e::8] = (v >= thresh).astype(np.uint8)
    return keep


def replay_indices(batch, size, seed, counter):
    b = np.arange(batch)
    u = draw4(seed, TAG_REPLAY, (b // 4).astype(np.uint64), np.uint64(0), counter)
    lane = b % 4
    x = np.where(lane == 0, u[0], np.where(lane == 1, u[1], np.where(lane == 2, u[2], u[3]))).astype(np.uint64)
    return ((x * np.uint64(size)) >> np.uint64(32)).astype(np.int64)


def reset_positions(n_envs, seed, counter):
    """int64 [n_envs, 4] = (p1x, p1y, p2x, p2y) of SS_RESET_RANDOM for Philox counter `counter` (ss_env_core.cuh
    reset_random: counter words (env lo, env hi, counter lo, counter hi), key = seed; coordinate = 25 + mulhi(u, 200),
    the library's draw for np.random.randint(25, 225), SkillshotGame.py:15)."""
    env = np.arange(n_envs, dtype=np.uint64)
    seed, counter = int(seed), int(counter)
    u = philox4x32_10(env & MASK, env >> np.uint64(32), counter & 0xFFFFFFFF, counter >> 32, seed & 0xFFFFFFFF, seed >> 32)
    return np.stack([25 + ((x.astype(np.uint64) * np.uint64(200)) >> np.uint64(32)).astype(np.int64) for x in u], axis=1)
