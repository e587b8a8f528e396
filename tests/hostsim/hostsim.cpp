// hostsim.cpp -- TEST-ONLY host build of skillshot_learning_b200/csrc/ss_env_core.cuh.
//
// The authoring container has no GPU; this file compiles the very same
// __host__ __device__ game logic the sm_100a kernels inline, over the same packed
// 64-byte SoA state layout, so the logic (operation order, packing, reset,
// reward / done / auto-reset rules) can be checked against the oracle on the CPU.
// It is NOT a fallback: nothing in skillshot_learning_b200/ loads it, and the
// product fails loudly when its CUDA library is missing.
#include <stdint.h>
#include <string.h>

#include "../../include/skillshot_b200.h"
#include "../../skillshot_learning_b200/csrc/ss_env_core.cuh"

using namespace ss;

namespace {
struct Planes { double *rot, *qrot; Int4 *ia, *ib; };
Planes planes_of(void *state, int64_t n) {
    char *b = (char *)state;
    return Planes{(double *)b, (double *)(b + 16 * n), (Int4 *)(b + 32 * n), (Int4 *)(b + 48 * n)};
}
void load_env(const Planes &s, int64_t i, Env &e) {
    unpack(e, s.rot[2 * i], s.rot[2 * i + 1], s.qrot[2 * i], s.qrot[2 * i + 1], s.ia[i], s.ib[i]);
}
void store_env(const Planes &s, int64_t i, const Env &e) {
    pack(e, s.ia[i], s.ib[i]);
    s.rot[2 * i] = e.prot[0]; s.rot[2 * i + 1] = e.prot[1];
    s.qrot[2 * i] = e.qrot[0]; s.qrot[2 * i + 1] = e.qrot[1];
}
Speeds load_speeds(const void *speeds, int64_t n, int64_t i) {
    if (!speeds) return default_speeds();
    const char *b = (const char *)speeds;
    const double *a = (const double *)b + 2 * i;
    const double *p1 = (const double *)(b + 16 * n) + 2 * i;
    long long cm = ((const long long *)(b + 16 * n))[2 * i + 1];
    return Speeds{a[0], a[1], p1[0], (int)cm, 1.0f / (float)cm};
}
struct HostSink {
    float *obs;
    void put(int j, float a, float b, float c, float d) { float *o = obs + 4 * j; o[0] = a; o[1] = b; o[2] = c; o[3] = d; }
};
}  // namespace

extern "C" {

// the kernel core's sincos, for tests/test_sincos.py
void hs_sincos(const double *x, int64_t n, double *s, double *c) {
    for (int64_t i = 0; i < n; ++i) sincos_d(x[i], s + i, c + i);
}

int hs_env_reset(void *state, int64_t n, const uint8_t *mask, int reset_mode, const int32_t *positions,
                 uint64_t seed, uint64_t counter) {
    Planes S = planes_of(state, n);
    for (int64_t i = 0; i < n; ++i) {
        if (mask && !mask[i]) continue;
        Env e;
        if (reset_mode == SS_RESET_GIVEN) reset_env(e, positions[4 * i], positions[4 * i + 1], positions[4 * i + 2], positions[4 * i + 3]);
        else if (reset_mode == SS_RESET_RANDOM) reset_random(e, seed, (uint64_t)i, counter);
        else reset_env(e, 50, 50, 200, 200);
        store_env(S, i, e);
    }
    return 0;
}

int hs_env_step(void *state, int64_t n, const float *actions, float *obs_out, float *reward_out,
                uint8_t *done_out, uint8_t *winner_out, int n_ticks, int reward_mode, int64_t tick_limit,
                int auto_reset, int reset_mode, uint64_t seed, uint64_t counter, const void *speeds,
                uint32_t *status_out, int flags) {
    Planes S = planes_of(state, n);
    TickParams P;
    P.tick_limit = tick_limit; P.seed = seed; P.counter = counter;
    P.reward_mode = reward_mode; P.auto_reset = auto_reset ? 1 : 0; P.reset_mode = reset_mode;
    const bool every = flags & SS_STEP_OBS_EVERY_TICK;
    uint32_t status = 0;
    const bool shaped = (reward_mode == SS_REWARD_LOOKING || reward_mode == SS_REWARD_SIMPLE) && reward_out;
    const bool carry = obs_out || n_ticks > 1 || shaped;       // same dispatch as ss_env_step
    for (int64_t i = 0; i < n; ++i) {
        Env e;
        load_env(S, i, e);
        Speeds k = load_speeds(speeds, n, i);
        Trig tr;
        if (carry) trig_of(e, tr);
        for (int t = 0; t < n_ticks; ++t) {
            const int64_t row = (int64_t)t * n + i;
            const float *a = actions + row * 4;
            const bool want_obs = obs_out && (every || t == n_ticks - 1);
            float r[2], obs[2 * kNumObs];
            int done, winner;
            HostSink sink{obs};
            if (obs_out) tick_env<true, true>(e, a[0], a[1], a[2], a[3], k, P, (uint64_t)i, t, want_obs, status, tr, r, done, winner, sink);
            else if (carry) tick_env<false, true>(e, a[0], a[1], a[2], a[3], k, P, (uint64_t)i, t, false, status, tr, r, done, winner, sink);
            else tick_env<false, false>(e, a[0], a[1], a[2], a[3], k, P, (uint64_t)i, t, false, status, tr, r, done, winner, sink);
            if (reward_out && reward_mode != SS_REWARD_NONE) { reward_out[row * 2] = r[0]; reward_out[row * 2 + 1] = r[1]; }
            if (done_out) done_out[row] = (uint8_t)done;
            if (winner_out) winner_out[row] = (uint8_t)winner;
            if (want_obs) memcpy(obs_out + ((every ? (int64_t)t * n : 0) + i) * 2 * kNumObs, obs, sizeof obs);
        }
        store_env(S, i, e);
    }
    if (status_out) *status_out |= status;
    return 0;
}

int hs_env_features(const void *state, int64_t n, double *feat_out, double *obs_out, int32_t *general_out,
                    const void *speeds) {
    Planes S = planes_of((void *)state, n);
    for (int64_t i = 0; i < n; ++i) {
        Env e;
        load_env(S, i, e);
        Speeds k = load_speeds(speeds, n, i);
        if (general_out) { general_out[3 * i] = e.live; general_out[3 * i + 1] = e.ticks; general_out[3 * i + 2] = e.winner; }
        if (feat_out) { features_of<0>(e, feat_out + (2 * i) * kNumFeat); features_of<1>(e, feat_out + (2 * i + 1) * kNumFeat); }
        if (obs_out) {
            View v0 = view_of<0, false>(e), v1 = view_of<1, false>(e);
            obs_of<0>(e, v0, k, obs_out + (2 * i) * kNumObs);
            obs_of<1>(e, v1, k, obs_out + (2 * i + 1) * kNumObs);
        }
    }
    return 0;
}

int hs_env_export(const void *state, int64_t n, int64_t first, int64_t count, int32_t *ints, double *rots) {
    Planes S = planes_of((void *)state, n);
    for (int64_t j = 0; j < count; ++j) {
        Env e;
        load_env(S, first + j, e);
        int32_t *o = ints + j * SS_EXPORT_INTS;
        o[0] = e.px[0]; o[1] = e.px[1]; o[2] = e.py[0]; o[3] = e.py[1];
        o[4] = e.qx[0]; o[5] = e.qx[1]; o[6] = e.qy[0]; o[7] = e.qy[1];
        o[8] = e.cd[0]; o[9] = e.cd[1]; o[10] = e.age[0]; o[11] = e.age[1];
        o[12] = e.valid[0]; o[13] = e.valid[1]; o[14] = e.ticks; o[15] = e.live; o[16] = e.winner;
        double *r = rots + j * 4;
        r[0] = e.prot[0]; r[1] = e.prot[1]; r[2] = e.qrot[0]; r[3] = e.qrot[1];
    }
    return 0;
}

int hs_env_import(void *state, int64_t n, int64_t first, int64_t count, const int32_t *ints, const double *rots) {
    Planes S = planes_of(state, n);
    for (int64_t j = 0; j < count; ++j) {
        Env e;
        const int32_t *o = ints + j * SS_EXPORT_INTS;
        e.px[0] = o[0] & 255; e.px[1] = o[1] & 255; e.py[0] = o[2] & 255; e.py[1] = o[3] & 255;
        e.qx[0] = o[4] & 255; e.qx[1] = o[5] & 255; e.qy[0] = o[6] & 255; e.qy[1] = o[7] & 255;
        e.cd[0] = o[8]; e.cd[1] = o[9]; e.age[0] = o[10]; e.age[1] = o[11];
        e.valid[0] = o[12] != 0; e.valid[1] = o[13] != 0; e.ticks = o[14]; e.live = o[15] != 0; e.winner = o[16] & 3;
        const double *r = rots + j * 4;
        e.prot[0] = r[0]; e.prot[1] = r[1]; e.qrot[0] = r[2]; e.qrot[1] = r[3];
        store_env(S, first + j, e);
    }
    return 0;
}

}  // extern "C"
