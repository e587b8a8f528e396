// ss_env_core.cuh -- one SkillshotGame instance held in registers.
//
// Every function here is __host__ __device__ so that the identical logic can be
// (a) inlined into the sm_100a kernels of ss_env.cu and (b) compiled for the
// host by tests/hostsim (a test-only build used to check the game logic in the
// GPU-less authoring container; the product never loads it).
//
// Numerics contract (SURVEY.md "hard parts" 1-4): actions are float32 values
// promoted to float64; all physics is float64 with NO fused multiply-add (the
// reference is CPython: every * and - rounds separately), rounding to integer
// positions is round-half-to-even (Python round()).  The translation unit is
// compiled with -fmad=false and the products that feed a rounding use the
// explicit _rn intrinsics as well.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define SS_HD __host__ __device__ __forceinline__
#else
#define SS_HD inline
#endif

namespace ss {

constexpr int kBoard = 250;        // SkillshotGame.py:11
constexpr int kPlayerSize = 5;     // Player.py:9-13, 23
constexpr int kProjSize = 3;       // Projectile.py:5-7, 20
constexpr int kNumFeat = 18;
constexpr int kNumObs = 12;
constexpr double kPi = 3.141592653589793;          // math.pi / np.pi
constexpr double kHalfPi = 1.5707963267948966;     // math.pi / 2 (exact halving)
constexpr double kMaxDist = 353.5533905932738;     // (2 * 250**2) ** 0.5, SkillshotLearner.py:43

constexpr uint32_t kStatusNaN = 1u;

// ---- exact (unfused) float64 arithmetic --------------------------------
SS_HD double mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
SS_HD double add(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
SS_HD double sub(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
SS_HD double divd(double a, double b) {
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
// Python round() -> int: round-half-to-even.  Caller has excluded NaN/inf.
SS_HD int rint_i(double v) {
#ifdef __CUDA_ARCH__
    return __double2int_rn(v);
#else
    return (int)rint(v);
#endif
}
// ---- sin/cos ------------------------------------------------------------
// Library fallback (huge or non-finite arguments only).
SS_HD void sincos_lib(double r, double *s, double *c) {
#ifdef __CUDA_ARCH__
    sincos(r, s, c);
#else
    *s = sin(r); *c = cos(r);
#endif
}
SS_HD int lo_word(double v) {
#ifdef __CUDA_ARCH__
    return __double2loint(v);
#else
    uint64_t b; memcpy(&b, &v, sizeof b); return (int)(uint32_t)b;
#endif
}
SS_HD double flip_sign_if(double v, int cond) {    // v or -v without a branch
#ifdef __CUDA_ARCH__
    return __hiloint2double(__double2hiint(v) ^ (cond ? (int)0x80000000 : 0), __double2loint(v));
#else
    return cond ? -v : v;
#endif
}
// sin and cos of one float64 rotation: Cody-Waite reduction by pi/2 in three FMA
// steps (each FMA rounds once, relative to its own result, so the reduced argument
// keeps full precision), then the fdlibm minimax kernels on [-pi/4, pi/4].
// Measured max error vs 80-bit libm: 1.52 ulp, mean 0.3 ulp (tests/test_sincos.py; the
// CUDA library's documented bound is 2 ulp; the kernels alone stay under 0.75 ulp).  Pure IEEE
// arithmetic with explicit FMAs, so host and device builds return identical bits.
// About a third of the instructions of the CUDA library sincos (no slow-path call,
// no local-memory out-parameters), which is what the step kernel's issue rate needs.
// The 17 float64 constants.  On the device the multi-tick kernels read them from constant memory: an FMA can take
// one operand straight from the constant bank, whereas a literal costs two uniform-register moves in front of every use
// (345 UMOVs of the physics kernel's 2,112 instructions before this): +7.6 % on the issue-bound fused kernels.  The
// one-tick physics kernel keeps the literals (sincos_d_lit): it is memory-bound and short, and the constant-cache
// latency in front of each thread's only tick cost it 14 %.
#define SS_C0 6755399441055744.0           /* 1.5 * 2^52: round-to-nearest-int trick */
#define SS_C1 0.6366197723675814           /* 2/pi */
#define SS_C2 -1.5707963267948966          /* -pi/2 in three parts */
#define SS_C3 -6.123233995736766e-17
#define SS_C4 1.4973849048591698e-33
#define SS_C5 1.58969099521155010221e-10   /* sin kernel */
#define SS_C6 -2.50507602534068634195e-08
#define SS_C7 2.75573137070700676789e-06
#define SS_C8 -1.98412698298579493134e-04
#define SS_C9 8.33333333332248946124e-03
#define SS_C10 -1.66666666666666324348e-01
#define SS_C11 -1.13596475577881948265e-11  /* cos kernel */
#define SS_C12 2.08757232129817482790e-09
#define SS_C13 -2.75573143513906633035e-07
#define SS_C14 2.48015872894767294178e-05
#define SS_C15 -1.38888888888741095749e-03
#define SS_C16 4.16666666666666019037e-02
#define SS_SINCOS_CONSTANTS SS_C0, SS_C1, SS_C2, SS_C3, SS_C4, SS_C5, SS_C6, SS_C7, SS_C8, SS_C9, SS_C10, SS_C11, SS_C12, SS_C13, \
                            SS_C14, SS_C15, SS_C16
#ifdef __CUDACC__
static __constant__ double kSinCosDev[17] = {SS_SINCOS_CONSTANTS};
#endif
#ifdef __CUDA_ARCH__
#define SS_SC(i) kSinCosDev[i]
#else
#define SS_SC(i) SS_C##i
#endif
#define SS_LIT(i) SS_C##i

// Every multiply / add / subtract below goes through the explicit round-to-nearest wrappers (mul / add / sub): the
// result is then the same whether or not the translation unit lets the compiler contract a * b + c (ss_env.cu is built
// with -fmad=false, the tensor-core units that inline this for the fused rollout kernel are not), and the same as the
// host build's.  The fma() calls are the algorithm's own.
#define SS_SINCOS_BODY(C, CHECKED)                                                                          \
    {                                                                                                       \
        if (CHECKED && !(fabs(x) < 1.0e5)) { sincos_lib(x, sp, cp); return; }                               \
        const double kMagic = C(0);                                                                         \
        double q = fma(x, C(1), kMagic); /* x * 2/pi */                                                     \
        const int k = lo_word(q);                                                                           \
        q = sub(q, kMagic);                                                                                 \
        double t = fma(q, C(2), x);                                                                         \
        t = fma(q, C(3), t);                                                                                \
        t = fma(q, C(4), t);                                                                                \
        const double t2 = mul(t, t);                                                                        \
        double ps = fma(C(5), t2, C(6));                                                                    \
        ps = fma(ps, t2, C(7));                                                                             \
        ps = fma(ps, t2, C(8));                                                                             \
        ps = fma(ps, t2, C(9));                                                                             \
        ps = fma(ps, t2, C(10));                                                                            \
        const double sn = fma(mul(t, t2), ps, t);                                                           \
        double pc = fma(C(11), t2, C(12));                                                                  \
        pc = fma(pc, t2, C(13));                                                                            \
        pc = fma(pc, t2, C(14));                                                                            \
        pc = fma(pc, t2, C(15));                                                                            \
        pc = fma(pc, t2, C(16));                                                                            \
        /* 1 - t2/2 + t2^2*pc with the rounding error of (1 - t2/2) fed back (fdlibm/musl __cos form) */    \
        const double hz = mul(0.5, t2), w = sub(1.0, hz);                                                   \
        const double cs = add(w, add(sub(sub(1.0, w), hz), mul(mul(t2, t2), pc)));                          \
        /* quadrant k mod 4: (s,c) = (sn,cs), (cs,-sn), (-sn,-cs), (-cs,sn) */                              \
        const int swap = k & 1;                                                                             \
        const double a = swap ? cs : sn, b = swap ? sn : cs;                                                \
        *sp = flip_sign_if(a, (k & 2) != 0);                                                                \
        *cp = flip_sign_if(b, ((k + 1) & 2) != 0);                                                          \
    }
SS_HD void sincos_d(double x, double *sp, double *cp) SS_SINCOS_BODY(SS_SC, true)
SS_HD void sincos_d_lit(double x, double *sp, double *cp) SS_SINCOS_BODY(SS_LIT, true)
// the same without the range check, for callers that have established |x| < 1e5 themselves (the per-player step kernel
// checks once per launch: a rotation moves by at most 0.25 per tick)
SS_HD void sincos_d_unchecked(double x, double *sp, double *cp) SS_SINCOS_BODY(SS_SC, false)
SS_HD bool finite_d(double v) { return v - v == 0.0; }

// Python float % 2 (Objects/floatobject.c float_rem): fmod, result takes the
// divisor's sign.
SS_HD double py_mod2(double v) {
    double m = fmod(v, 2.0);
    if (m != 0.0) { if (m < 0) m += 2.0; } else { m = 0.0; }
    return m;
}

// ---- per-env constants (class attributes in the reference) --------------
struct Speeds {
    double speed_move;   // Player.speed_move = 3        Player.py:14
    double speed_look;   // Player.speed_look = 0.25     Player.py:15
    double proj_speed;   // Projectile.speed_move = 5    Projectile.py:10
    int cooldown_max;    // Projectile.cooldown_max = 15 Projectile.py:9
    float inv_cooldown;  // 1 / cooldown_max, for the float32 observation
};
SS_HD Speeds default_speeds() { return Speeds{3.0, 0.25, 5.0, 15, 1.0f / 15.0f}; }

// ---- one game in registers ----------------------------------------------
struct Env {
    double prot[2], qrot[2];
    int px[2], py[2], qx[2], qy[2];
    int cd[2], age[2], valid[2];
    int ticks, live, winner;
};

// the two 16-byte integer planes of the HBM layout (include/skillshot_b200.h)
struct Int4 { int x, y, z, w; };

SS_HD void unpack(Env &e, double r0, double r1, double q0, double q1, Int4 a, Int4 b) {
    e.prot[0] = r0; e.prot[1] = r1; e.qrot[0] = q0; e.qrot[1] = q1;
    uint32_t pp = (uint32_t)a.x, qq = (uint32_t)a.y;
    e.px[0] = pp & 255; e.py[0] = (pp >> 8) & 255; e.px[1] = (pp >> 16) & 255; e.py[1] = pp >> 24;
    e.qx[0] = qq & 255; e.qy[0] = (qq >> 8) & 255; e.qx[1] = (qq >> 16) & 255; e.qy[1] = qq >> 24;
    e.cd[0] = a.z; e.cd[1] = a.w;
    e.age[0] = b.x; e.age[1] = b.y; e.ticks = b.z;
    uint32_t f = (uint32_t)b.w;
    e.valid[0] = f & 1; e.valid[1] = (f >> 1) & 1; e.live = (f >> 2) & 1; e.winner = (f >> 4) & 3;
}
SS_HD void pack(const Env &e, Int4 &a, Int4 &b) {
    a.x = (int)((uint32_t)e.px[0] | ((uint32_t)e.py[0] << 8) | ((uint32_t)e.px[1] << 16) | ((uint32_t)e.py[1] << 24));
    a.y = (int)((uint32_t)e.qx[0] | ((uint32_t)e.qy[0] << 8) | ((uint32_t)e.qx[1] << 16) | ((uint32_t)e.qy[1] << 24));
    a.z = e.cd[0]; a.w = e.cd[1];
    b.x = e.age[0]; b.y = e.age[1]; b.z = e.ticks;
    b.w = (int)((uint32_t)e.valid[0] | ((uint32_t)e.valid[1] << 1) | ((uint32_t)e.live << 2) | ((uint32_t)e.winner << 4));
}

// SkillshotGame.__init__ (SkillshotGame.py:10-25) with given player positions.
SS_HD void reset_env(Env &e, int p1x, int p1y, int p2x, int p2y) {
    e.px[0] = p1x; e.py[0] = p1y; e.px[1] = p2x; e.py[1] = p2y;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        e.prot[p] = 0.0;                 // Player.py:21
        e.qx[p] = 0; e.qy[p] = 0;        // Player.py:25  Projectile((0, 0), ...)
        e.qrot[p] = 0.0;                 // Projectile.py:14
        e.cd[p] = 0; e.age[p] = 0;       // Projectile.py:16-17
        e.valid[p] = 0;                  // Projectile.py:18
    }
    e.ticks = 0; e.live = 1; e.winner = 0;   // SkillshotGame.py:23-25
}

// Player.check_pos_valid (Player.py:70-76)
SS_HD bool player_pos_valid(int x, int y) {
    return x + kPlayerSize <= kBoard && x >= 0 && y + kPlayerSize <= kBoard && y >= 0;
}
// Projectile.check_pos_valid (Projectile.py:30-36)
SS_HD bool proj_pos_valid(int x, int y) {
    return x + kProjSize <= kBoard && x >= 0 && y + kProjSize <= kBoard && y >= 0;
}

SS_HD double clip_unit(double v) {       // Player.py:36-37, 60-61 (NaN falls through)
    v = (v >= 1.0) ? 1.0 : v;
    v = (v <= -1.0) ? -1.0 : v;
    return v;
}

// Player.move_direction_float (Player.py:57-68).  (s, c) = sin/cos of the
// player's rotation BEFORE this tick's turn.
template <int P>
SS_HD void move_direction_float(Env &e, double speed, double s, double c, const Speeds &k, uint32_t &status) {
    speed = clip_unit(speed);
    double vx = sub((double)e.px[P], mul(mul(s, k.speed_move), speed));   // Player.py:63
    double vy = sub((double)e.py[P], mul(mul(c, k.speed_move), speed));   // Player.py:64
    if (!(finite_d(vx) && finite_d(vy))) { status |= kStatusNaN; return; }   // int(round(nan)) raises
    int nx = rint_i(vx), ny = rint_i(vy);
    if (player_pos_valid(nx, ny)) { e.px[P] = nx; e.py[P] = ny; }        // Player.py:66-68
}

// Player.move_look_float (Player.py:33-39)
template <int P>
SS_HD void move_look_float(Env &e, double angle, const Speeds &k) {
    e.prot[P] = add(e.prot[P], mul(clip_unit(angle), k.speed_look));
}

// Player.move_shoot_projectile (Player.py:78-89)
template <int P>
SS_HD bool move_shoot(Env &e, const Speeds &k) {
    if (e.cd[P] <= 0) {
        e.qx[P] = e.px[P]; e.qy[P] = e.py[P];
        e.qrot[P] = e.prot[P];
        e.valid[P] = 1;
        e.cd[P] = k.cooldown_max;
        e.age[P] = 0;
        return true;
    }
    return false;
}

// Projectile.tick (Projectile.py:38-53).  (s, c) = sin/cos of the projectile
// rotation.
template <int P>
SS_HD void proj_tick(Env &e, double s, double c, const Speeds &k, uint32_t &status) {
    double vx = sub((double)e.qx[P], mul(s, k.proj_speed));   // Projectile.py:40
    double vy = sub((double)e.qy[P], mul(c, k.proj_speed));   // Projectile.py:41
    if (!(finite_d(vx) && finite_d(vy))) { status |= kStatusNaN; vx = -1.0; vy = -1.0; }
    int nx = rint_i(vx), ny = rint_i(vy);
    if (e.valid[P] && proj_pos_valid(nx, ny)) { e.qx[P] = nx; e.qy[P] = ny; }   // :43-45
    else e.valid[P] = 0;                                                        // :47
    e.cd[P] -= 1;                                                               // :52
    e.age[P] += 1;                                                              // :53
}

// One (player P, enemy projectile 1-P) pair of SkillshotGame.check_collision
// (SkillshotGame.py:58-94).
template <int P>
SS_HD bool hit_test(const Env &e) {
    constexpr int Q = 1 - P;
    if (!e.valid[Q]) return false;
    int pl = e.px[P], pr = e.px[P] + kPlayerSize, pt = e.py[P], pb = e.py[P] + kPlayerSize;
    int jl = e.qx[Q], jr = e.qx[Q] + kProjSize, jt = e.qy[Q], jb = e.qy[Q] - kProjSize;   // :72 minus
    bool in_x = (pl <= jr && jr <= pr) || (pl <= jl && jl <= pr);
    bool in_y = (pt <= jt && jt <= pb) || (pt <= jb && jb <= pb);
    return in_x && in_y;
}

// SkillshotGame.check_collision: pair (P1, P2's projectile) first; the first hit
// breaks, so a double hit records id 1 only.  winner_id = the player that was hit.
SS_HD void check_collision(Env &e) {
    if (hit_test<0>(e)) { e.winner = 1; e.live = 0; }
    else if (hit_test<1>(e)) { e.winner = 2; e.live = 0; }
}

// get_gradient_dir + get_dist_line_point + get_dist_point_point +
// check_future_collision for player P, as written in the reference
// (Player.py:91-100, SkillshotGame.py:96-134).
struct View {
    double player_grad, player_path_dist, player_dist;
    double proj_grad, proj_path_dist, proj_dist;
    int player_x_dir, proj_x_dir, future_collision;
};

SS_HD double dist_line_point(double g, int lx, int ly, int cx, int cy) {   // SkillshotGame.py:124-130
    double c = sub((double)ly, mul(g, (double)lx));
    double num = fabs(add(sub(mul(g, (double)cx), (double)cy), c));
    return divd(num, sqrt(add(mul(g, g), 1.0)));
}
SS_HD double dist_point_point(int ax, int ay, int bx, int by) {             // SkillshotGame.py:132-134
    int dx = ax - bx, dy = ay - by;
    return sqrt((double)(dx * dx + dy * dy));
}

template <int P, bool FULL>
SS_HD View view_of(const Env &e) {
    constexpr int O = 1 - P;
    View v;
    v.player_grad = tan(add(-e.prot[P], kHalfPi));                            // Player.py:94
    v.proj_grad = tan(add(-e.qrot[P], kHalfPi));                              // Projectile.py:58
    if (FULL) {
        v.player_x_dir = (-sin(e.prot[P]) >= 0.0) ? 1 : -1;                   // Player.py:96
        v.proj_x_dir = (-sin(e.qrot[P]) >= 0.0) ? 1 : -1;
    } else {
        v.player_x_dir = 1; v.proj_x_dir = 1;
    }
    v.player_path_dist = dist_line_point(v.player_grad, e.px[P], e.py[P], e.px[O], e.py[O]);
    v.player_dist = dist_point_point(e.px[P], e.py[P], e.px[O], e.py[O]);
    v.proj_path_dist = dist_line_point(v.proj_grad, e.qx[P], e.qy[P], e.px[O], e.py[O]);
    v.proj_dist = dist_point_point(e.qx[P], e.qy[P], e.px[O], e.py[O]);
    // check_future_collision (SkillshotGame.py:96-113).  The direction guard at
    // :109 passes for the projectile's own left bound whatever x_dir is, and the
    // second bound re-tests the same two opponent bounds, so the result is the
    // OR over the opponent's two x bounds.
    int fc = 0;
    if (e.valid[P]) {
        double yint = sub((double)e.qy[P], mul(v.proj_grad, (double)e.qx[P]));   // Projectile.py:62
        double lo = (double)e.py[O], hi = (double)(e.py[O] + kPlayerSize);
        double v0 = add(mul(v.proj_grad, (double)e.px[O]), yint);
        double v1 = add(mul(v.proj_grad, (double)(e.px[O] + kPlayerSize)), yint);
        fc = ((lo <= v0 && v0 <= hi) || (lo <= v1 && v1 <= hi)) ? 1 : 0;
    }
    v.future_collision = fc;
    return v;
}

// get_state per-player feature vector in the dict's key order (SkillshotGame.py:145-162)
template <int P>
SS_HD void features_of(const Env &e, double *f) {
    View v = view_of<P, true>(e);
    f[0] = v.player_grad; f[1] = (double)v.player_x_dir; f[2] = v.player_path_dist; f[3] = v.player_dist;
    f[4] = (double)e.px[P]; f[5] = (double)e.py[P]; f[6] = e.prot[P]; f[7] = (double)e.cd[P];
    f[8] = v.proj_grad; f[9] = (double)v.proj_x_dir; f[10] = v.proj_path_dist;
    f[11] = (double)e.qx[P]; f[12] = (double)e.qy[P]; f[13] = e.qrot[P]; f[14] = (double)e.age[P];
    f[15] = (double)e.valid[P]; f[16] = v.proj_dist; f[17] = (double)v.future_collision;
}

// prepare_states (SkillshotLearner.py:525-539), divisions as written.
SS_HD double rot_term(double rot) {   // (rot % 2 * np.pi) / 2 * np.pi, literal precedence
    return mul(divd(mul(py_mod2(rot), kPi), 2.0), kPi);
}
template <int P>
SS_HD void obs_of(const Env &e, const View &v, const Speeds &k, double *o) {
    o[0] = divd(v.player_path_dist, kMaxDist);
    o[1] = divd(v.player_dist, kMaxDist);
    o[2] = divd((double)e.px[P], (double)kBoard);
    o[3] = divd((double)e.py[P], (double)kBoard);
    o[4] = rot_term(e.prot[P]);
    o[5] = divd((double)e.cd[P], (double)k.cooldown_max);
    o[6] = divd(v.proj_dist, kMaxDist);
    o[7] = divd((double)e.qx[P], (double)kBoard);
    o[8] = divd((double)e.qy[P], (double)kBoard);
    o[9] = rot_term(e.qrot[P]);
    o[10] = divd(v.proj_path_dist, kMaxDist);
    o[11] = (double)v.future_collision;
}

// ---- Philox4x32-10 (Salmon et al., SC'11), counter-based RNG -------------
struct U4 { uint32_t x, y, z, w; };
SS_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
SS_HD U4 philox4x32_10(U4 ctr, uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = mulhi32(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = mulhi32(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = U4{hi1 ^ ctr.y ^ k0, lo1, hi0 ^ ctr.w ^ k1, lo0};
        k0 += W0; k1 += W1;
    }
    return ctr;
}
// np.random.randint(25, 225) equivalent draw (SkillshotGame.py:15); the MT19937
// stream itself is not part of the contract (SURVEY hard part 10).
SS_HD int rand_coord(uint32_t u) { return 25 + (int)mulhi32(u, 200u); }

SS_HD void reset_random(Env &e, uint64_t seed, uint64_t env, uint64_t counter) {
    U4 r = philox4x32_10(U4{(uint32_t)env, (uint32_t)(env >> 32), (uint32_t)counter, (uint32_t)(counter >> 32)},
                         (uint32_t)seed, (uint32_t)(seed >> 32));
    reset_env(e, rand_coord(r.x), rand_coord(r.y), rand_coord(r.z), rand_coord(r.w));
}

// ---- one model_train tick (SkillshotLearner.py:304-315) -----------------
// sin/cos of the four rotations of the CURRENT state.  A projectile's rotation
// only changes when it is (re)spawned, to its owner's rotation, so in the fused
// multi-tick loop the four pairs are carried in registers and only the two player
// pairs are re-evaluated per tick (after the turn).  sincos is a pure function, so
// carried values are bit-identical to recomputed ones.
struct Trig { double ps[2], pc[2], qs[2], qc[2]; };

SS_HD void trig_of(const Env &e, Trig &tr) {
    sincos_d(e.prot[0], &tr.ps[0], &tr.pc[0]);
    sincos_d(e.prot[1], &tr.ps[1], &tr.pc[1]);
    sincos_d(e.qrot[0], &tr.qs[0], &tr.qc[0]);
    sincos_d(e.qrot[1], &tr.qs[1], &tr.qc[1]);
}
SS_HD void trig_zero(Trig &tr) {       // all rotations 0 after a reset: sin 0 = 0, cos 0 = 1
    tr.ps[0] = tr.ps[1] = tr.qs[0] = tr.qs[1] = 0.0;
    tr.pc[0] = tr.pc[1] = tr.qc[0] = tr.qc[1] = 1.0;
}

template <int P>
SS_HD void player_acts(Env &e, float a_move, float a_look, const Speeds &k, uint32_t &status, Trig &tr) {
    // do_actions(P+1): SkillshotLearner.py:206-213 -- not gated on game_live
    move_direction_float<P>(e, (double)a_move, tr.ps[P], tr.pc[P], k, status);   // rotation BEFORE the turn
    move_look_float<P>(e, (double)a_look, k);
    sincos_d(e.prot[P], &tr.ps[P], &tr.pc[P]);
    if (move_shoot<P>(e, k)) { tr.qs[P] = tr.ps[P]; tr.qc[P] = tr.pc[P]; }       // projectile takes the new rotation
}

// Actions of both players + game_tick, carrying Trig (valid for the pre-tick state on
// entry, for the post-tick state on exit).  a = (p1 move, p1 look, p2 move, p2 look).
SS_HD void act_and_tick_carry(Env &e, float a0, float a1, float a2, float a3, const Speeds &k,
                              uint32_t &status, Trig &tr) {
    player_acts<0>(e, a0, a1, k, status, tr);
    player_acts<1>(e, a2, a3, k, status, tr);
    if (e.live) {                                   // game_tick, SkillshotGame.py:115-122
        e.ticks += 1;
        proj_tick<0>(e, tr.qs[0], tr.qc[0], k, status);
        proj_tick<1>(e, tr.qs[1], tr.qc[1], k, status);
        check_collision(e);
    }
}

// The same without carried state (one physics-only tick per launch): 4 sincos.
SS_HD void act_and_tick(Env &e, float a0, float a1, float a2, float a3, const Speeds &k, uint32_t &status) {
    double s, c;
    sincos_d_lit(e.prot[0], &s, &c);
    move_direction_float<0>(e, (double)a0, s, c, k, status);
    move_look_float<0>(e, (double)a1, k);
    move_shoot<0>(e, k);
    sincos_d_lit(e.prot[1], &s, &c);
    move_direction_float<1>(e, (double)a2, s, c, k, status);
    move_look_float<1>(e, (double)a3, k);
    move_shoot<1>(e, k);
    if (e.live) {
        e.ticks += 1;
        sincos_d_lit(e.qrot[0], &s, &c);
        proj_tick<0>(e, s, c, k, status);
        sincos_d_lit(e.qrot[1], &s, &c);
        proj_tick<1>(e, s, c, k, status);
        check_collision(e);
    }
}

// ---- float32 observation / reward path -----------------------------------
// The step kernel's observation and shaped rewards are float32 with a 1e-6
// tolerance (BASELINE.json north_star), so they are evaluated from the carried
// sin/cos instead of the reference's tan/sqrt/divide chain:
//   get_dist_line_point(tan(pi/2 - r), l, c) == |cos r * (cx-lx) - sin r * (cy-ly)|
// (same line, |direction| = 1; differs from the literal float64 value by ~1e-13).
// ss_env_features keeps the literal float64 evaluation (view_of / features_of).
// The future-collision flag is a comparison, so it keeps the reference's own
// expression g*x + (y - g*x) with g = cos/sin (tan itself where sin r ~ 0, the
// g = 1.6e16 case of an unturned projectile).
constexpr double kInvMaxDist = 1.0 / 353.5533905932738;
constexpr double kHalfPiSq = 4.934802200544679;     // pi * pi / 2

struct FastView { double player_path_dist, proj_path_dist, player_dist, proj_dist; int future_collision; };

SS_HD double dist_i(int dx, int dy, bool precise) {
    int d2 = dx * dx + dy * dy;
    return precise ? sqrt((double)d2) : (double)sqrtf((float)d2);
}
// rot % 2 (Python): rot - 2*floor(rot/2) is the same single rounding as fmod + fix-up
SS_HD double py_mod2_fast(double v) { return v - 2.0 * floor(v * 0.5); }

template <int P>
SS_HD FastView fast_view(const Env &e, const Trig &tr, bool precise) {
    constexpr int O = 1 - P;
    FastView v;
    int dx = e.px[O] - e.px[P], dy = e.py[O] - e.py[P];
    v.player_path_dist = fabs(sub(mul(tr.pc[P], (double)dx), mul(tr.ps[P], (double)dy)));
    v.player_dist = dist_i(dx, dy, precise);
    int ex = e.px[O] - e.qx[P], ey = e.py[O] - e.qy[P];
    v.proj_path_dist = fabs(sub(mul(tr.qc[P], (double)ex), mul(tr.qs[P], (double)ey)));
    v.proj_dist = dist_i(ex, ey, precise);
    int fc = 0;
    if (e.valid[P]) {                                    // check_future_collision, SkillshotGame.py:96-113
        const double s = tr.qs[P], c = tr.qc[P];
        const int d0 = e.px[O] - e.qx[P], d1 = d0 + kPlayerSize;     // opponent x bounds relative to the projectile
        const int l = e.py[O] - e.qy[P], h = l + kPlayerSize;        // opponent y bounds relative to the projectile
        if (fmin(fabs(s), fabs(c)) < 1e-6) {
            // axis-aligned shots (sin or cos ~ 0: unturned projectiles, quarter turns) keep the
            // reference's own expression with g = tan(-rot + pi/2), whose argument rounding decides them
            double g = tan(add(-e.qrot[P], kHalfPi));
            double yint = sub((double)e.qy[P], mul(g, (double)e.qx[P]));
            double lo = (double)e.py[O], hi = (double)(e.py[O] + kPlayerSize);
            double v0 = add(mul(g, (double)e.px[O]), yint);
            double v1 = add(mul(g, (double)(e.px[O] + kPlayerSize)), yint);
            fc = ((lo <= v0 && v0 <= hi) || (lo <= v1 && v1 <= hi)) ? 1 : 0;
        } else {
            // lo <= y + (c/s)(x_b - x) <= hi  <=>  l*s <= c*d <= h*s (s > 0; reversed for s < 0):
            // the same line test without the divide.  It can differ from the reference only
            // where the reference's own value is rounding noise (the line through a box corner).
            const double ls = (double)l * s, hs = (double)h * s;
            const double lo = fmin(ls, hs), hi = fmax(ls, hs);
            const double c0 = c * (double)d0, c1 = c * (double)d1;
            fc = ((lo <= c0 && c0 <= hi) || (lo <= c1 && c1 <= hi)) ? 1 : 0;
        }
    }
    v.future_collision = fc;
    return v;
}

// The 12 observation floats of player P, handed to `sink` four at a time (slots
// 3P .. 3P+2 of the env's six float4) so that no more than four are live at once.
template <int P, class Sink>
SS_HD void fast_obs(const Env &e, const FastView &v, const Speeds &k, Sink &sink) {   // SkillshotLearner.py:525-539
    constexpr float kInvBoard = 1.0f / 250.0f;
    sink.put(3 * P + 0,
             (float)(v.player_path_dist * kInvMaxDist), (float)(v.player_dist * kInvMaxDist),
             (float)e.px[P] * kInvBoard, (float)e.py[P] * kInvBoard);
    sink.put(3 * P + 1,
             (float)(py_mod2_fast(e.prot[P]) * kHalfPiSq), (float)e.cd[P] * k.inv_cooldown,
             (float)(v.proj_dist * kInvMaxDist), (float)e.qx[P] * kInvBoard);
    sink.put(3 * P + 2,
             (float)e.qy[P] * kInvBoard, (float)(py_mod2_fast(e.qrot[P]) * kHalfPiSq),
             (float)(v.proj_path_dist * kInvMaxDist), (float)v.future_collision);
}

SS_HD void fast_rewards(int reward_mode, const FastView &v0, const FastView &v1, float *r) {
    if (reward_mode == 1) {            // calculate_rewards_looking, SkillshotLearner.py:584
        r[0] = (float)(v0.player_path_dist * -0.004);
        r[1] = (float)(v1.player_path_dist * -0.004);
    } else {                           // calculate_rewards_simple, SkillshotLearner.py:600
        r[0] = (float)sub(v0.proj_dist, v1.proj_dist);
        r[1] = (float)sub(v1.proj_dist, v0.proj_dist);
    }
}

struct TickParams {
    int64_t tick_limit;      // <= 0: none (model_param_game_tick_limit, SkillshotLearner.py:62, 302)
    uint64_t seed, counter;  // Philox key / counter base for random resets
    int reward_mode, auto_reset, reset_mode;
};

// One tick of one env as the batched step defines it: both players act, the game
// ticks, reward / done / winner are taken from the post-tick state, a done env is
// optionally reset, and the observation (if wanted) is that of the state the
// actor will see next.  obs = 24 floats (player 1's 12, then player 2's).
// CARRY: tr holds the sin/cos of the current rotations across calls (required
// for OBS and for the shaped rewards).
template <bool OBS, bool CARRY, class Sink>
SS_HD void tick_env(Env &e, float a0, float a1, float a2, float a3, const Speeds &k,
                    const TickParams &P, uint64_t env_id, int t, bool want_obs,
                    uint32_t &status, Trig &tr, float *r, int &done, int &winner, Sink &sink) {
    const int was_live = e.live;
    if (CARRY) act_and_tick_carry(e, a0, a1, a2, a3, k, status, tr);
    else act_and_tick(e, a0, a1, a2, a3, k, status);
    done = ((!e.live) || (P.tick_limit > 0 && e.ticks >= P.tick_limit)) ? 1 : 0;
    winner = e.winner;
    const bool will_reset = P.auto_reset && done;
    const bool reward_view = CARRY && (P.reward_mode == 1 || P.reward_mode == 3);
    const bool precise = (P.reward_mode == 3);
    r[0] = 0.f; r[1] = 0.f;
    if (P.reward_mode == 2 && was_live && !e.live && e.winner != 0) {   // readme.md:10
        r[e.winner - 1] = -1.f;       // the player that was hit
        r[2 - e.winner] = 1.f;        // the shooter
    }
    if (CARRY) {
        FastView v0, v1;
        if (reward_view && will_reset) {  // rare: the reward belongs to the pre-reset state
            v0 = fast_view<0>(e, tr, precise); v1 = fast_view<1>(e, tr, precise);
            fast_rewards(P.reward_mode, v0, v1, r);
        }
        if (will_reset) {
            if (P.reset_mode == 1) reset_random(e, P.seed, env_id, P.counter + (uint64_t)t);
            else reset_env(e, 50, 50, 200, 200);
            trig_zero(tr);
        }
        const bool need_view = (OBS && want_obs) || (reward_view && !will_reset);
        if (need_view) { v0 = fast_view<0>(e, tr, precise); v1 = fast_view<1>(e, tr, precise); }
        if (reward_view && !will_reset) fast_rewards(P.reward_mode, v0, v1, r);
        if (OBS && want_obs) {
            fast_obs<0>(e, v0, k, sink);
            fast_obs<1>(e, v1, k, sink);
        }
    } else if (will_reset) {
        if (P.reset_mode == 1) reset_random(e, P.seed, env_id, P.counter + (uint64_t)t);
        else reset_env(e, 50, 50, 200, 200);
    }
}

}  // namespace ss
