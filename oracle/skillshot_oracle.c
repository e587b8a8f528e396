/*
 * skillshot_oracle.c -- CPU restatement of the Skillshot_Learning game path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package
 * (skillshot_learning_b200/) may import, link or call this file.  It is used by
 * tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline /
 * --impl reference legs as the checker and as the timed CPU baseline.
 *
 * Parity status: PINNED.  oracle/gen_golden.py runs the unmodified reference
 * (imported from /root/reference in the authoring container) on seeded action
 * streams and writes tests/golden/.npz; tests/test_oracle_golden.py checks
 * every field of this restatement against those vectors bit-for-bit (the
 * reference's float arithmetic is CPython floats over glibc libm, which is
 * exactly what this file evaluates: same libm, same operation order, no FMA
 * contraction -- compile with -ffp-contract=off).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference repository root).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SS_BOARD 250        /* SkillshotGame.py:11  board_size = (250, 250) */
#define SS_PLAYER_SIZE 5    /* Player.py:9-13,23    5x5 shape_image        */
#define SS_PROJ_SIZE 3      /* Projectile.py:5-7,20 3x3 shape_image        */
#define SS_PLAYER_SPEED 3   /* Player.py:14         speed_move = 3         */
#define SS_LOOK_SPEED 0.25  /* Player.py:15         speed_look = 0.25      */
#define SS_PROJ_SPEED 5     /* Projectile.py:10     speed_move = 5         */
#define SS_COOLDOWN_MAX 15  /* Projectile.py:9      cooldown_max = 15      */

#define SS_NFEAT 18         /* per-player keys of get_state, SkillshotGame.py:145-162 */
#define SS_NOBS 12          /* prepare_states, SkillshotLearner.py:525-539 */

/* One game instance: SkillshotGame + 2 Player + 2 Projectile objects
 * (SkillshotGame.py:10-25, Player.py:17-25, Projectile.py:12-20). */
typedef struct {
    int64_t px[2], py[2];    /* Player.pos                       */
    double  prot[2];         /* Player.rotation                  */
    int64_t qx[2], qy[2];    /* Projectile.pos                   */
    double  qrot[2];         /* Projectile.rotation              */
    int64_t cd[2];           /* Projectile.cooldown_current      */
    int64_t age[2];          /* Projectile.age                   */
    int32_t valid[2];        /* Projectile.valid                 */
    int64_t ticks;           /* SkillshotGame.ticks              */
    int32_t live;            /* SkillshotGame.game_live          */
    int32_t winner;          /* SkillshotGame.winner_id          */
    /* per-env speed constants (class attributes in the reference; per-env
     * here so the readme.md:22-23 speed sweep can be expressed) */
    double  speed_move, speed_look, proj_speed;
    int64_t cooldown_max;
    /* 1 when Player.pos holds numpy int64 (random start, SkillshotGame.py:15):
     * `np.int64 ** 0.5` is evaluated by numpy as sqrt(), whereas the fixed
     * start's Python `int ** 0.5` is libm pow(x, 0.5); the two differ by 1 ulp
     * in ~0.08 % of arguments (probed), so get_dist_point_point depends on it. */
    int32_t np_pos;
    int32_t pad_;
} ss_oracle_env;

/* Python round() on a float with ndigits=None is round-half-to-even
 * (Objects/floatobject.c float___round___impl); rint() in the default
 * rounding mode is the same function. */
static inline int64_t py_int_round(double v) { return (int64_t)rint(v); }

/* Python float %: fmod, then move the result to the divisor's sign
 * (Objects/floatobject.c float_rem). */
static inline double py_float_mod(double v, double w)
{
    double m = fmod(v, w);
    if (m != 0.0) {
        if ((w < 0) != (m < 0)) m += w;
    } else {
        m = copysign(0.0, w);
    }
    return m;
}

/* SkillshotGame.__init__ / game_reset (SkillshotGame.py:10-25, 168-169).
 * pos == NULL: fixed start P1 [50,50], P2 [200,200]; otherwise pos holds
 * {p1x, p1y, p2x, p2y} (the np.random.randint(25,225,(2,2)) draw is injected
 * by the caller, the MT19937 stream is not part of the contract). */
void ss_oracle_reset(ss_oracle_env *e, const int64_t *pos)
{
    if (pos) {
        e->px[0] = pos[0]; e->py[0] = pos[1];
        e->px[1] = pos[2]; e->py[1] = pos[3];
        e->np_pos = 1;
    } else {
        e->np_pos = 0;
        e->px[0] = 50;  e->py[0] = 50;     /* SkillshotGame.py:17 */
        e->px[1] = 200; e->py[1] = 200;    /* SkillshotGame.py:18 */
    }
    for (int i = 0; i < 2; ++i) {
        e->prot[i] = 0.0;                  /* Player.py:21 */
        e->qx[i] = 0; e->qy[i] = 0;        /* Player.py:25 Projectile((0, 0), ...) */
        e->qrot[i] = 0.0;                  /* Projectile.py:14 */
        e->cd[i] = 0; e->age[i] = 0;       /* Projectile.py:16-17 */
        e->valid[i] = 0;                   /* Projectile.py:18 */
    }
    e->ticks = 0; e->live = 1; e->winner = 0;  /* SkillshotGame.py:23-25 */
    e->speed_move = SS_PLAYER_SPEED;
    e->speed_look = SS_LOOK_SPEED;
    e->proj_speed = SS_PROJ_SPEED;
    e->cooldown_max = SS_COOLDOWN_MAX;
}

void ss_oracle_set_speeds(ss_oracle_env *e, double speed_move, double speed_look,
                          double proj_speed, int64_t cooldown_max)
{
    e->speed_move = speed_move; e->speed_look = speed_look;
    e->proj_speed = proj_speed; e->cooldown_max = cooldown_max;
}

/* Player.check_pos_valid (Player.py:70-76). */
static inline int player_pos_valid(int64_t x, int64_t y)
{
    return x + SS_PLAYER_SIZE <= SS_BOARD && x >= 0 &&
           y + SS_PLAYER_SIZE <= SS_BOARD && y >= 0;
}

/* Projectile.check_pos_valid (Projectile.py:30-36). */
static inline int proj_pos_valid(int64_t x, int64_t y)
{
    return x + SS_PROJ_SIZE <= SS_BOARD && x >= 0 &&
           y + SS_PROJ_SIZE <= SS_BOARD && y >= 0;
}

/* Player.move_direction_float (Player.py:57-68).  Returns -1 where the
 * reference raises ValueError (int(round(nan))), else 0.  p is 0 or 1. */
int ss_oracle_move_direction_float(ss_oracle_env *e, int p, double speed)
{
    if (speed >= 1) speed = 1;             /* Player.py:60 */
    if (speed <= -1) speed = -1;           /* Player.py:61 */
    /* Player.py:63-64: pos - sin(rot) * speed_move * speed, left to right */
    double vx = (double)e->px[p] - sin(e->prot[p]) * e->speed_move * speed;
    double vy = (double)e->py[p] - cos(e->prot[p]) * e->speed_move * speed;
    if (isnan(vx) || isnan(vy) || isinf(vx) || isinf(vy)) return -1;
    int64_t nx = py_int_round(vx), ny = py_int_round(vy);
    if (player_pos_valid(nx, ny)) {        /* Player.py:66-68: both or neither */
        e->px[p] = nx; e->py[p] = ny;
    }
    return 0;
}

/* Player.move_look_float (Player.py:33-39). */
void ss_oracle_move_look_float(ss_oracle_env *e, int p, double angle)
{
    if (angle >= 1) angle = 1;
    if (angle <= -1) angle = -1;
    e->prot[p] += angle * e->speed_look;
}

/* Player.move_shoot_projectile (Player.py:78-89). */
void ss_oracle_move_shoot(ss_oracle_env *e, int p)
{
    if (e->cd[p] <= 0) {
        e->qx[p] = e->px[p]; e->qy[p] = e->py[p];   /* set_position(self.pos): a copy */
        e->qrot[p] = e->prot[p];
        e->valid[p] = 1;
        e->cd[p] = e->cooldown_max;
        e->age[p] = 0;
    }
}

/* Discrete key-press moves (Player.py:27-31, 41-55), used by
 * skillshot_playable.py:51-61.  dir = +1 forwards, -1 backwards. */
int ss_oracle_move_step(ss_oracle_env *e, int p, int dir)
{
    double vx, vy;
    if (dir > 0) {                                  /* Player.py:41-47 */
        vx = (double)e->px[p] - sin(e->prot[p]) * e->speed_move;
        vy = (double)e->py[p] - cos(e->prot[p]) * e->speed_move;
    } else {                                        /* Player.py:49-55 */
        vx = (double)e->px[p] + sin(e->prot[p]) * e->speed_move;
        vy = (double)e->py[p] + cos(e->prot[p]) * e->speed_move;
    }
    if (isnan(vx) || isnan(vy)) return -1;
    int64_t nx = py_int_round(vx), ny = py_int_round(vy);
    if (player_pos_valid(nx, ny)) { e->px[p] = nx; e->py[p] = ny; }
    return 0;
}

/* Player.move_look_left / move_look_right (Player.py:27-31). dir=+1 left. */
void ss_oracle_look_step(ss_oracle_env *e, int p, int dir)
{
    if (dir > 0) e->prot[p] += e->speed_look; else e->prot[p] -= e->speed_look;
}

/* Projectile.tick = move_forwards + counters (Projectile.py:38-53).
 * Returns -1 where the reference raises (NaN rotation). */
static int proj_tick(ss_oracle_env *e, int p)
{
    double vx = (double)e->qx[p] - sin(e->qrot[p]) * e->proj_speed;   /* :40 */
    double vy = (double)e->qy[p] - cos(e->qrot[p]) * e->proj_speed;   /* :41 */
    if (isnan(vx) || isnan(vy)) return -1;
    int64_t nx = py_int_round(vx), ny = py_int_round(vy);
    if (e->valid[p] && proj_pos_valid(nx, ny)) {                      /* :43-45 */
        e->qx[p] = nx; e->qy[p] = ny;
    } else {
        e->valid[p] = 0;                                              /* :47 */
    }
    e->cd[p] -= 1;                                                    /* :52 */
    e->age[p] += 1;                                                   /* :53 */
    return 0;
}

/* SkillshotGame.check_collision (SkillshotGame.py:58-94): pair (P1, P2's
 * projectile) first, then (P2, P1's); the first hit breaks out, so a double
 * hit records id 1 only.  winner_id is the id of the player that was HIT. */
static void check_collision(ss_oracle_env *e)
{
    for (int p = 0; p < 2; ++p) {
        int q = 1 - p;                       /* enemy projectile owner */
        if (!e->valid[q]) continue;          /* :62 */
        int64_t pl = e->px[p], pr = e->px[p] + SS_PLAYER_SIZE;   /* :64-65 */
        int64_t pt = e->py[p], pb = e->py[p] + SS_PLAYER_SIZE;   /* :66-67 */
        int64_t jl = e->qx[q], jr = e->qx[q] + SS_PROJ_SIZE;     /* :69-70 */
        int64_t jt = e->qy[q], jb = e->qy[q] - SS_PROJ_SIZE;     /* :71-72 (minus) */
        int in_r = pl <= jr && jr <= pr, in_l = pl <= jl && jl <= pr;
        int in_t = pt <= jt && jt <= pb, in_b = pt <= jb && jb <= pb;
        if ((in_r && in_t) || (in_r && in_b) || (in_l && in_t) || (in_l && in_b)) { /* :75-94 */
            e->winner = p + 1;
            e->live = 0;
            break;
        }
    }
}

/* SkillshotGame.game_tick (SkillshotGame.py:115-122). */
int ss_oracle_game_tick(ss_oracle_env *e)
{
    if (e->live) {
        e->ticks += 1;
        if (proj_tick(e, 0)) return -1;
        if (proj_tick(e, 1)) return -1;
        check_collision(e);
    }
    return 0;
}

/* SkillshotLearner.do_actions (SkillshotLearner.py:206-213): move, look,
 * always attempt to shoot.  Not gated on game_live (KAT-C). */
int ss_oracle_do_actions(ss_oracle_env *e, int p, double a_move, double a_look)
{
    if (ss_oracle_move_direction_float(e, p, a_move)) return -1;
    ss_oracle_move_look_float(e, p, a_look);
    ss_oracle_move_shoot(e, p);
    return 0;
}

/* get_gradient_dir (Player.py:91-100, Projectile.py:55-64; same body). */
static inline void gradient_dir(double rot, int64_t x, int64_t y,
                                double *grad, int *x_dir, double *y_int)
{
    *grad = tan(-rot + M_PI / 2);
    *x_dir = (-sin(rot) >= 0) ? 1 : -1;
    *y_int = (double)y - *grad * (double)x;
}

/* SkillshotGame.get_dist_line_point (SkillshotGame.py:124-130).  g**2 on a
 * Python float goes through libm pow(g, 2.0). */
static inline double dist_line_point(double g, int64_t lx, int64_t ly, int64_t cx, int64_t cy)
{
    double c = (double)ly - g * (double)lx;
    return fabs(g * (double)cx - (double)cy + c) / sqrt(pow(g, 2.0) + 1);
}

/* SkillshotGame.get_dist_point_point (SkillshotGame.py:132-134): integer
 * squares, then int ** 0.5 == libm pow(float(n), 0.5) for Python ints, sqrt
 * for numpy ints (see ss_oracle_env.np_pos). */
static inline double dist_point_point(int64_t ax, int64_t ay, int64_t bx, int64_t by, int np_pos)
{
    int64_t d2 = (ax - bx) * (ax - bx) + (ay - by) * (ay - by);
    return np_pos ? sqrt((double)d2) : pow((double)d2, 0.5);
}

/* SkillshotGame.check_future_collision (SkillshotGame.py:96-113).  The
 * direction guard at :109 is kept as written (it is vacuous for the first
 * projectile bound). */
static int future_collision(const ss_oracle_env *e, int p, int o)
{
    if (!e->valid[p]) return 0;
    double g, yint; int xd;
    gradient_dir(e->qrot[p], e->qx[p], e->qy[p], &g, &xd, &yint);
    int64_t xbp[2] = { e->qx[p], e->qx[p] + SS_PROJ_SIZE };
    int64_t xbo[2] = { e->px[o], e->px[o] + SS_PLAYER_SIZE };
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            if ((xbp[i] - e->qx[p]) * xd >= 0) {
                double v = g * (double)xbo[j] + yint;
                if ((double)e->py[o] <= v && v <= (double)(e->py[o] + SS_PLAYER_SIZE))
                    return 1;
            }
    return 0;
}

/* SkillshotGame.get_state (SkillshotGame.py:136-166).  feat is [2][18] in the
 * dict's key order; general = {game_live, ticks, game_winner}. */
void ss_oracle_get_state(const ss_oracle_env *e, double *feat, int64_t *general)
{
    general[0] = e->live; general[1] = e->ticks; general[2] = e->winner;
    for (int p = 0; p < 2; ++p) {
        int o = 1 - p;
        double *f = feat + p * SS_NFEAT;
        double pg, py_, qg, qy_; int pxd, qxd;
        gradient_dir(e->prot[p], e->px[p], e->py[p], &pg, &pxd, &py_);
        gradient_dir(e->qrot[p], e->qx[p], e->qy[p], &qg, &qxd, &qy_);
        f[0] = pg;                                                             /* player_grad */
        f[1] = pxd;                                                            /* player_x_dir */
        f[2] = dist_line_point(pg, e->px[p], e->py[p], e->px[o], e->py[o]);    /* player_path_dist_opponent */
        f[3] = dist_point_point(e->px[p], e->py[p], e->px[o], e->py[o], e->np_pos);       /* player_dist_opponent */
        f[4] = (double)e->px[p];
        f[5] = (double)e->py[p];
        f[6] = e->prot[p];
        f[7] = (double)e->cd[p];
        f[8] = qg;
        f[9] = qxd;
        f[10] = dist_line_point(qg, e->qx[p], e->qy[p], e->px[o], e->py[o]);   /* projectile_path_dist_opponent */
        f[11] = (double)e->qx[p];
        f[12] = (double)e->qy[p];
        f[13] = e->qrot[p];
        f[14] = (double)e->age[p];
        f[15] = (double)e->valid[p];
        f[16] = dist_point_point(e->qx[p], e->qy[p], e->px[o], e->py[o], e->np_pos);      /* projectile_dist_opponent */
        f[17] = (double)future_collision(e, p, o);
    }
}

/* SkillshotLearner.prepare_states for one state dict (SkillshotLearner.py:
 * 512-543).  obs is [2][12] doubles (the reference hands float64 lists to
 * Keras, which casts to float32 at predict/fit). */
void ss_oracle_prepare_obs(const ss_oracle_env *e, const double *feat, double *obs)
{
    /* SkillshotLearner.py:43  (2 * (250 ** 2)) ** 0.5 -- int ** 0.5 -> pow */
    const double D = pow((double)(2 * (SS_BOARD * SS_BOARD)), 0.5);
    for (int p = 0; p < 2; ++p) {
        const double *f = feat + p * SS_NFEAT;
        double *o = obs + p * SS_NOBS;
        o[0] = f[2] / D;                                             /* :525 */
        o[1] = f[3] / D;                                             /* :526 */
        o[2] = f[4] / SS_BOARD;                                      /* :527 */
        o[3] = f[5] / SS_BOARD;                                      /* :528 */
        o[4] = (py_float_mod(f[6], 2.0) * M_PI) / 2 * M_PI;          /* :529 literal precedence */
        o[5] = f[7] / (double)e->cooldown_max;                       /* :532-533 */
        o[6] = f[16] / D;                                            /* :534 */
        o[7] = f[11] / SS_BOARD;                                     /* :535 */
        o[8] = f[12] / SS_BOARD;                                     /* :536 */
        o[9] = (py_float_mod(f[13], 2.0) * M_PI) / 2 * M_PI;         /* :537 */
        o[10] = f[10] / D;                                           /* :538 */
        o[11] = (double)(int)f[17];                                  /* :539 */
    }
}

/* Reward functions over one post-tick state.
 *  mode 1: calculate_rewards_looking (SkillshotLearner.py:575-588, the active one)
 *  mode 2: terminal +1/-1/0 (readme.md:10; no reference code -- defined here as
 *          -1 for the player that was hit (winner_id), +1 for the other, only
 *          on the tick the game ends; parity unpinned)
 *  mode 3: calculate_rewards_simple (SkillshotLearner.py:590-603) */
void ss_oracle_rewards(const ss_oracle_env *e, const double *feat, int mode,
                       int just_ended, double *r)
{
    r[0] = r[1] = 0.0;
    if (mode == 1) {
        for (int p = 0; p < 2; ++p) r[p] = -feat[p * SS_NFEAT + 2] / SS_BOARD;
    } else if (mode == 2) {
        if (just_ended && e->winner != 0) {
            r[e->winner - 1] = -1.0;
            r[2 - e->winner] = 1.0;
        }
    } else if (mode == 3) {
        for (int p = 0; p < 2; ++p)
            r[p] = feat[p * SS_NFEAT + 16] - feat[(1 - p) * SS_NFEAT + 16];
    }
}

/* ---------------------------------------------------------------------- *
 * Batched driver: the same per-env sequence SkillshotLearner.model_train
 * runs per tick (SkillshotLearner.py:304-315): both players act from the
 * pre-tick state (P1 then P2), game_tick, get_state, prepare_states,
 * reward of the post-tick state.  One env per loop iteration; OpenMP over
 * envs for the CPU baseline.
 *
 *   actions  [n][2][2] float32 (promoted to double, SURVEY hard part 1)
 *   obs      [n][2][12] float32 or NULL
 *   reward   [n][2] float32 or NULL
 *   done     [n] uint8  (1 when !live or ticks >= tick_limit after the tick)
 *   winner   [n] uint8
 *   auto_reset: a done env is reset to the fixed start after its outputs
 *   are taken (obs is then the post-reset observation), or to
 *   reset_pos[n][4] if given.
 * Returns the number of envs that hit a reference exception (NaN).
 * ---------------------------------------------------------------------- */
int ss_oracle_step_batch(ss_oracle_env *envs, int64_t n, const float *actions,
                         float *obs, float *reward, uint8_t *done, uint8_t *winner,
                         int reward_mode, int64_t tick_limit, int auto_reset,
                         const int64_t *reset_pos, int nthreads)
{
    int errors = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static) reduction(+:errors) if (n >= 256)
#endif
    for (int64_t i = 0; i < n; ++i) {
        ss_oracle_env *e = &envs[i];
        const float *a = actions + i * 4;
        int was_live = e->live;
        int err = 0;
        err |= ss_oracle_do_actions(e, 0, (double)a[0], (double)a[1]);
        err |= ss_oracle_do_actions(e, 1, (double)a[2], (double)a[3]);
        err |= ss_oracle_game_tick(e);
        if (err) errors += 1;
        int is_done = (!e->live) || (tick_limit > 0 && e->ticks >= tick_limit);
        double feat[2 * SS_NFEAT]; int64_t general[3];
        int will_reset = auto_reset && is_done;
        int need_feat = (reward && (reward_mode == 1 || reward_mode == 3)) || (obs && !will_reset);
        if (need_feat) ss_oracle_get_state(e, feat, general);
        if (reward) {
            double r[2];
            ss_oracle_rewards(e, feat, reward_mode, was_live && !e->live, r);
            reward[i * 2 + 0] = (float)r[0];
            reward[i * 2 + 1] = (float)r[1];
        }
        if (done) done[i] = (uint8_t)is_done;
        if (winner) winner[i] = (uint8_t)e->winner;
        if (will_reset) {
            double sm = e->speed_move, sl = e->speed_look, ps = e->proj_speed;
            int64_t cm = e->cooldown_max;
            ss_oracle_reset(e, reset_pos ? reset_pos + i * 4 : 0);
            ss_oracle_set_speeds(e, sm, sl, ps, cm);
            if (obs) ss_oracle_get_state(e, feat, general);
        }
        if (obs) {
            double o[2 * SS_NOBS];
            ss_oracle_prepare_obs(e, feat, o);
            for (int k = 0; k < 2 * SS_NOBS; ++k) obs[i * 2 * SS_NOBS + k] = (float)o[k];
        }
    }
    return errors;
}

/* Features / observations of the current state for n envs, float64 out. */
void ss_oracle_features_batch(const ss_oracle_env *envs, int64_t n, double *feat,
                              double *obs, int64_t *general)
{
    for (int64_t i = 0; i < n; ++i) {
        double f[2 * SS_NFEAT]; int64_t g[3];
        ss_oracle_get_state(&envs[i], f, g);
        if (feat) memcpy(feat + i * 2 * SS_NFEAT, f, sizeof f);
        if (general) memcpy(general + i * 3, g, sizeof g);
        if (obs) ss_oracle_prepare_obs(&envs[i], f, obs + i * 2 * SS_NOBS);
    }
}

void ss_oracle_reset_batch(ss_oracle_env *envs, int64_t n, const int64_t *pos)
{
    for (int64_t i = 0; i < n; ++i) ss_oracle_reset(&envs[i], pos ? pos + i * 4 : 0);
}

int64_t ss_oracle_env_size(void) { return (int64_t)sizeof(ss_oracle_env); }
int ss_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
