/*
 * nmpc_oracle.c — CPU restatement (FP64, plain C) of the NMPC solve the reference
 * shells out to OpEn for.  TEST INFRASTRUCTURE ONLY: this file is the parity
 * checker and the CPU baseline; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (mpc_trajectory_generator_b200/) never links or calls it.
 *
 * PARITY UNPINNED.  The reference holds no solver source, no tests and no golden
 * vectors for this path.  The arithmetic lives in third-party
 * packages that are absent from /root/reference and from this image:
 *   opengen==0.6.4 (env/environment.yml:14) -> Rust crate `optimization_engine`
 *   (PANOC + ALM/PM; version chosen by opengen, not pinned by the reference),
 *   crate `lbfgs`, and CasADi (unpinned, env/environment.yml:15) for psi/grad psi.
 * What IS in the reference, and is followed line by line here:
 *   - the problem definition  src/mpc/mpc_generator.py:66-175   (stage(), eval_psi())
 *   - the parameter layout    src/mpc/mpc_generator.py:71-79,93-104,
 *                             src/path_generator.py:378-379       (stage())
 *   - the solver settings     src/mpc/mpc_generator.py:184-186    (tolerance 1e-4)
 * The algorithm (PANOC, L-BFGS, ALM/PM outer loop, TCP-server warm start) is
 * restated from OpEn's published algorithm / public source as recalled
 * (optimization_engine: core/panoc/panoc_engine.rs, panoc_cache.rs,
 * panoc_optimizer.rs, lipschitz_estimator.rs, alm/alm_optimizer.rs; lbfgs crate
 * lib.rs; opengen templates tcp_server.rs / optimizer.rs) — a spec to re-verify
 * against OpEn when it can be installed, not a cited fact.  Each function names
 * the OpEn routine it restates.
 *
 * TWO BUILDS OF THIS FILE (oracle/Makefile):
 *
 *  libnmpc_oracle.so  — the ARITHMETIC CONTRACT shared with the CUDA kernel:
 *    IEEE binary64, no implicit contraction (-ffp-contract=off), explicit fma() exactly where
 *    written, own sincos (Cody-Waite + fdlibm kernels, <= 2 ulp), reciprocals of per-problem
 *    constants precomputed once, and every sum over the horizon in the order the kernel's
 *    evaluation GROUP produces it: the horizon is cut into G chunks of S consecutive steps
 *    (layout_for(): G = 8 lanes for N <= 24, else 16; lane i owns steps S*i .. S*i+S-1);
 *    a sum is the serial sum of each chunk followed by an xor-butterfly over the G chunk
 *    partials; prefix / suffix sums along the horizon are the serial scan of each chunk plus
 *    a Kogge-Stone scan of the chunk totals.  These orders differ from a serial loop only in
 *    the last bits, and they let the GPU result be compared BIT FOR BIT with this file.
 *
 *  libnmpc_oracle_serial.so (-DNMPC_ORACLE_SERIAL) — the same control flow with the
 *    REFERENCE's arithmetic: libm sin/cos, true divisions, separate multiply and add (no fma),
 *    the rollout as the literal recurrence of src/mpc/mpc_generator.py:88-90, every sum a
 *    serial loop in ascending t.  It shares no reduction order, no sincos and no reciprocal
 *    with the kernel: the solve-level pin that is independent of the kernel's arithmetic
 *    (tests/test_serial_pin.py compares full solves at the north_star tolerance 1e-4).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/nmpc_b200.h"

#define MAXG 16
#define MAXT NMPC_MAX_HORIZON
#define MEMP1 (NMPC_LBFGS_MAX + 1)

#ifdef NMPC_ORACLE_SERIAL
#define FMA(a, b, c) ((a) * (b) + (c))
#define RDIV(x, d, inv) ((x) / (d))
#else
#define FMA(a, b, c) fma((a), (b), (c))
#define RDIV(x, d, inv) ((x) * (inv))
#endif

/* ------------------------------------------------------------------------- */
/* constants of OpEn's PANOC engine (panoc_engine.rs, panoc_cache.rs)          */
#define MIN_L_ESTIMATE 1e-10
#define GAMMA_L_COEFF 0.95
#define DELTA_LIPSCHITZ 1e-12
#define EPSILON_LIPSCHITZ 1e-6
#define LIPSCHITZ_UPDATE_EPSILON 1e-6
#define MAX_LIPSCHITZ_UPDATE_ITERATIONS 10
#define MAX_LIPSCHITZ_CONSTANT 1e9
#define MAX_LINESEARCH_ITERATIONS 10
#define CBFGS_ALPHA 1.0 /* DEFAULT_CBFGS_ALPHA: |g|^1 */
#define CBFGS_EPSILON 1e-8
#define SY_EPSILON 1e-10
#define DBL_EPS 2.220446049250313e-16 /* f64::EPSILON */
#define Y_SET_BOUND 1e12              /* opengen SetYCalculator.LARGE_NUM */

static int g_trace = 0; /* NMPC_ORACLE_TRACE, read once per batch call */
/* NMPC_ORACLE_VARIANT (bit mask, read once per batch call; 0 = the restatement everything else is compared with).
 * Each bit flips ONE behaviour that was restated from recollection of OpEn's source to what a reader of OpEn's
 * documentation might expect instead; tools/sensitivity.py reports how far the iteration profile and the replies move
 * (DESIGN.md §3).  Never set by the tests or the benchmark.
 *   1  AKKT residual with the gradient of the PREVIOUS iterate (not the degenerate |gamma*fpr| / gamma)
 *   2  exhausted line search falls back to u_half (not: keeps the last trial point)
 *   4  the Lipschitz estimate restores u (not: leaves it perturbed by h)
 *   8  ALM criterion 1 may hold in the first outer iteration (not: needs a second one)
 *  16  the penalty parameter may grow after the first outer iteration (not: iteration 0 always counts as a stall) */
static int g_variant = 0;

/* ------------------------------------------------------------------------- */
/* sincos: Cody-Waite reduction by pi/2 (three fma terms), fdlibm kernel
 * polynomials evaluated by Horner with fma.  Replaces the libm sin/cos calls in
 * CasADi's generated C for cs.cos/cs.sin (src/mpc/mpc_generator.py:88-89,118). */
static const double TWO_OVER_PI = 6.36619772367581382433e-01;
static const double PIO2_HI = 1.57079632679489655800e+00;
static const double PIO2_MD = 6.12323399573676603587e-17;
static const double PIO2_LO = -1.49738490485916983294e-33;
static const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                    S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                    S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
static const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                    C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                    C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;

void nmpc_oracle_sincos(double x, double* s, double* c) {
#ifdef NMPC_ORACLE_SERIAL
    *s = sin(x);
    *c = cos(x);
#else
    if (!(fabs(x) < 1.0e8)) { /* also catches NaN/inf: the solve reports NotFinite */
        *s = NAN;
        *c = NAN;
        return;
    }
    double kf = rint(x * TWO_OVER_PI);
    double r = fma(-kf, PIO2_HI, x);
    r = fma(-kf, PIO2_MD, r);
    r = fma(-kf, PIO2_LO, r);
    int k = (int)kf;
    double z = r * r;
    double ps = fma(z, S6, S5);
    ps = fma(z, ps, S4);
    ps = fma(z, ps, S3);
    ps = fma(z, ps, S2);
    ps = fma(z, ps, S1);
    double sr = fma(r * z, ps, r);
    double pc = fma(z, C6, C5);
    pc = fma(z, pc, C4);
    pc = fma(z, pc, C3);
    pc = fma(z, pc, C2);
    pc = fma(z, pc, C1);
    double cr = fma(z * z, pc, fma(-0.5, z, 1.0));
    switch (k & 3) {
        case 0: *s = sr; *c = cr; break;
        case 1: *s = cr; *c = -sr; break;
        case 2: *s = -sr; *c = -cr; break;
        default: *s = -cr; *c = sr; break;
    }
#endif
}

/* min/max written as compare-selects (the kernel uses the same forms): a NaN operand yields the
 * constant, like C's fmax/fmin, and the sign of a zero result is fixed (+0). */
static inline double sel_clamp01(double t) {
    double r = (t > 0.0) ? t : 0.0;
    return (r < 1.0) ? r : 1.0;
}
/* z - Proj_[lo,hi](z):  max(z - hi, 0) + min(z - lo, 0) */
static inline double sel_excess(double z, double lo, double hi) { return (z > hi) ? z - hi : ((z < lo) ? z - lo : 0.0); }

/* ------------------------------------------------------------------------- */
/* horizon layout of one evaluation group of the kernel: G lanes x S consecutive steps */
void nmpc_oracle_layout(int N, int* G, int* S) {
    if (N <= 16) { *G = 8; *S = 2; }
    else if (N <= 24) { *G = 8; *S = 3; }
    else if (N <= 32) { *G = 16; *S = 2; }
    else if (N <= 48) { *G = 16; *S = 3; }
    else if (N <= 64) { *G = 16; *S = 4; }
    else { *G = 16; *S = 6; }
}

/* sum over the horizon of per-step values e[t] */
static double hsum(const double* e, int N, int G, int S) {
#ifdef NMPC_ORACLE_SERIAL
    (void)G; (void)S;
    double a = 0.0;
    for (int t = 0; t < N; t++) a = a + e[t];
    return a;
#else
    double p[MAXG];
    for (int i = 0; i < G; i++) {
        double a = (S * i < N) ? e[S * i] : 0.0;
        for (int s = 1; s < S; s++) {
            int t = S * i + s;
            a = a + ((t < N) ? e[t] : 0.0);
        }
        p[i] = a;
    }
    for (int off = G / 2; off; off >>= 1)
        for (int i = 0; i < off; i++) p[i] = p[i] + p[i + off];
    return p[0];
#endif
}

/* inclusive / exclusive prefix sums along the horizon */
static void prefix_scan(const double* x, int N, int G, int S, double* incl, double* excl) {
#ifdef NMPC_ORACLE_SERIAL
    (void)G; (void)S;
    double a = 0.0;
    for (int t = 0; t < N; t++) { excl[t] = a; a = a + x[t]; incl[t] = a; }
#else
    double c[MAXT], T[MAXG], E[MAXG];
    for (int i = 0; i < G; i++) {
        double a = 0.0;
        for (int s = 0; s < S; s++) {
            int t = S * i + s;
            double v = (t < N) ? x[t] : 0.0;
            a = (s == 0) ? v : a + v;
            c[t] = a;
        }
        T[i] = a;
    }
    for (int off = 1; off < G; off <<= 1)
        for (int i = G - 1; i >= off; i--) T[i] = T[i] + T[i - off];
    for (int i = 0; i < G; i++) E[i] = i ? T[i - 1] : 0.0;
    for (int t = 0; t < N; t++) {
        int i = t / S, s = t % S;
        incl[t] = E[i] + c[t];
        excl[t] = s ? E[i] + c[t - 1] : E[i];
    }
#endif
}

/* inclusive suffix sums (from the end of the horizon) */
static void suffix_scan(const double* x, int N, int G, int S, double* suf) {
#ifdef NMPC_ORACLE_SERIAL
    (void)G; (void)S;
    double a = 0.0;
    for (int t = N - 1; t >= 0; t--) { a = a + x[t]; suf[t] = a; }
#else
    double d[MAXT], T[MAXG], E[MAXG];
    for (int i = 0; i < G; i++) {
        double a = 0.0;
        for (int s = S - 1; s >= 0; s--) {
            int t = S * i + s;
            double v = (t < N) ? x[t] : 0.0;
            a = (s == S - 1) ? v : a + v;
            d[t] = a;
        }
        T[i] = a;
    }
    for (int off = 1; off < G; off <<= 1)
        for (int i = 0; i + off < G; i++) T[i] = T[i] + T[i + off];
    for (int i = 0; i < G; i++) E[i] = (i + 1 < G) ? T[i + 1] : 0.0;
    for (int t = 0; t < N; t++) suf[t] = E[t / S] + d[t];
#endif
}

/* ------------------------------------------------------------------------- */
/* staged problem: the parameter vector unpacked once per solve                */
typedef struct {
    int N, Nobs, Nd, G, S, mem;
    double ts, inv_ts;
    double vmin, vmax, wmax, amin, amax, aamax;
    double x0, y0, th0, vinit, winit, xref, yref, thref;
    double q, qv, qth, rv, rw, qN, qthN, qcte, ap, wp;
    double vref[MAXT];
    double s1x[MAXT], s1y[MAXT], sdx[MAXT], sdy[MAXT], sden[MAXT], sinv[MAXT]; /* segment i = 1..N-1 */
    double *cx, *cy, *cr2;                                     /* [Nobs]   */
    double *ex, *ey, *eca, *esa, *erx2, *ery2, *eirx2, *eiry2; /* [Nd * N] index k*N + t */
    double* buf;
    size_t buf_len;
    int n_cost, n_grad;
} staged;

/* parameter layout: src/mpc/mpc_generator.py:71-79 (scalars), :93-95 (circles),
 * :98-104 (ellipses: obstacle-major, time, 5 values), :79,126-131 (reference points) */
static int stage(staged* S, const nmpc_config* cfg, const double* p) {
    int N = cfg->N_hor, Nobs = cfg->Nobs, Nd = cfg->Ndynobs;
    if (N < 2 || N > MAXT || Nobs < 0 || Nd < 0) return 1;
    if (cfg->lbfgs_memory < 1 || cfg->lbfgs_memory > NMPC_LBFGS_MAX) return 1;
    S->N = N; S->Nobs = Nobs; S->Nd = Nd; S->mem = cfg->lbfgs_memory;
    nmpc_oracle_layout(N, &S->G, &S->S);
    S->ts = cfg->ts; S->inv_ts = 1.0 / cfg->ts;
    S->vmin = cfg->lin_vel_min; S->vmax = cfg->lin_vel_max; S->wmax = cfg->ang_vel_max;
    S->amin = cfg->lin_acc_min; S->amax = cfg->lin_acc_max; S->aamax = cfg->ang_acc_max;
    S->x0 = p[0]; S->y0 = p[1]; S->th0 = p[2]; S->vinit = p[3]; S->winit = p[4];
    S->xref = p[5]; S->yref = p[6]; S->thref = p[7];
    S->q = p[10]; S->qv = p[11]; S->qth = p[12]; S->rv = p[13]; S->rw = p[14];
    S->qN = p[15]; S->qthN = p[16]; S->qcte = p[17]; S->ap = p[18]; S->wp = p[19];
    for (int t = 0; t < N; t++) S->vref[t] = p[NMPC_NZ + t];
    size_t nd = (size_t)3 * Nobs + (size_t)8 * Nd * N + 8;
    if (nd > S->buf_len) {
        free(S->buf);
        S->buf = (double*)malloc(nd * sizeof(double));
        S->buf_len = S->buf ? nd : 0;
        if (!S->buf) return 3;
    }
    double* b = S->buf;
    S->cx = b; b += Nobs; S->cy = b; b += Nobs; S->cr2 = b; b += Nobs;
    S->ex = b; b += Nd * N; S->ey = b; b += Nd * N; S->eca = b; b += Nd * N; S->esa = b; b += Nd * N;
    S->erx2 = b; b += Nd * N; S->ery2 = b; b += Nd * N; S->eirx2 = b; b += Nd * N; S->eiry2 = b; b += Nd * N;
    const double* pc = p + NMPC_NZ + N;
    for (int k = 0; k < Nobs; k++) {
        S->cx[k] = pc[3 * k]; S->cy[k] = pc[3 * k + 1];
        S->cr2[k] = pc[3 * k + 2] * pc[3 * k + 2]; /* rs_static**2, :112 */
    }
    const double* pe = pc + 3 * Nobs;
    for (int k = 0; k < Nd; k++)
        for (int t = 0; t < N; t++) {
            const double* e = pe + (size_t)k * 5 * N + 5 * t;
            int i = k * N + t;
            S->ex[i] = e[0]; S->ey[i] = e[1];
            S->erx2[i] = e[2] * e[2]; S->ery2[i] = e[3] * e[3]; /* / x_radius**2, :118 */
            S->eirx2[i] = 1.0 / S->erx2[i];
            S->eiry2[i] = 1.0 / S->ery2[i];
            nmpc_oracle_sincos(e[4], &S->esa[i], &S->eca[i]);
        }
    const double* pr = pe + (size_t)5 * Nd * N;
    for (int i = 1; i < N; i++) { /* segments (ref[i-1], ref[i]), :126-133 */
        double ax = pr[3 * (i - 1)], ay = pr[3 * (i - 1) + 1];
        double dx = pr[3 * i] - ax, dy = pr[3 * i + 1] - ay;
        S->s1x[i] = ax; S->s1y[i] = ay; S->sdx[i] = dx; S->sdy[i] = dy;
        S->sden[i] = FMA(dx, dx, dy * dy) + 1e-16; /* |s2-s1|^2 + 1e-16, :135 */
        S->sinv[i] = 1.0 / S->sden[i];
    }
    S->n_cost = S->n_grad = 0;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* psi(u; c, y, p) and grad psi.
 *   f    : src/mpc/mpc_generator.py:81-148 (loop over t), :157-171 (acceleration cost)
 *   F1   : :157-162  F1 = [acc; omega_acc], set C :164-168
 *   F2   : :106-119  F2_k = sum_t max(0, inside_k(t))   (sum over t BEFORE squaring)
 *   psi  = f + c/2 * ( dist^2_C(F1 + y/max(c,1)) + |F2|^2 )   (opengen builder
 *          __construct_function_psi; with_penalty_constraints -> F2,
 *          with_aug_lagrangian_constraints -> F1/C, :173-175)
 * grad psi is the hand-written adjoint of that graph with CasADi's sub-gradient
 * conventions: d fmax(a,b) = [a>=b, !(a>=b)], d fmin(a,b) = [a<=b, !(a<=b)],
 * mmin = left fold of fmin (first minimal segment takes the gradient).
 * y is in F1 order [acc(N); omega_acc(N)].  Any output pointer may be NULL. */
static double eval_psi(staged* S, const double* u, double c, const double* y, double* grad,
                       double* F1, double* F2) {
    const int N = S->N, G = S->G, SS = S->S;
    const double ts = S->ts, inv_ts = S->inv_ts;
    const double hc = 0.5 * c, cden = fmax(c, 1.0), inv_c = 1.0 / cden;
    double sn[MAXT], cs[MAXT];
    double X[MAXT], Y[MAXT], TH[MAXT], thpre[MAXT], xpre[MAXT], ypre[MAXT];
    double gX[MAXT], gY[MAXT], mind2[MAXT], cl[MAXT], Aa[MAXT], Aw[MAXT];
    double h[MAXT], hp[MAXT], hdx[MAXT], hdy[MAXT];
    if (grad) S->n_grad++; else S->n_cost++;

    /* rollout x += ts*(v*cos th); y += ts*(v*sin th); th += ts*w   (:88-90) */
#ifdef NMPC_ORACLE_SERIAL
    {
        double x = S->x0, yy = S->y0, th = S->th0;
        for (int t = 0; t < N; t++) {
            xpre[t] = x; ypre[t] = yy; thpre[t] = th;
            nmpc_oracle_sincos(th, &sn[t], &cs[t]);
            x = x + ts * (u[2 * t] * cs[t]);
            yy = yy + ts * (u[2 * t] * sn[t]);
            th = th + ts * u[2 * t + 1];
            X[t] = x; Y[t] = yy; TH[t] = th;
        }
    }
#else
    {
        double tw[MAXT], a[MAXT], b[MAXT], inclT[MAXT], exclT[MAXT];
        double inclA[MAXT], exclA[MAXT], inclB[MAXT], exclB[MAXT];
        for (int t = 0; t < N; t++) tw[t] = ts * u[2 * t + 1];
        prefix_scan(tw, N, G, SS, inclT, exclT);
        for (int t = 0; t < N; t++) {
            thpre[t] = S->th0 + exclT[t];
            TH[t] = S->th0 + inclT[t];
            nmpc_oracle_sincos(thpre[t], &sn[t], &cs[t]);
            a[t] = ts * (u[2 * t] * cs[t]);
            b[t] = ts * (u[2 * t] * sn[t]);
        }
        prefix_scan(a, N, G, SS, inclA, exclA);
        prefix_scan(b, N, G, SS, inclB, exclB);
        for (int t = 0; t < N; t++) {
            xpre[t] = S->x0 + exclA[t]; ypre[t] = S->y0 + exclB[t];
            X[t] = S->x0 + inclA[t];    Y[t] = S->y0 + inclB[t];
        }
    }
#endif

    /* cross-track error: min over segments of squared distance (:122-144) */
    for (int t = 0; t < N; t++) {
        double best = INFINITY, bex = 0.0, bey = 0.0, bth = 0.0;
        int bi = 1;
        for (int i = 1; i < N; i++) {
            double px = X[t] - S->s1x[i], py = Y[t] - S->s1y[i];
            double that = RDIV(FMA(px, S->sdx[i], py * S->sdy[i]), S->sden[i], S->sinv[i]);
            double tst = sel_clamp01(that); /* fmin(fmax(t_hat, 0), 1), :138 */
#ifdef NMPC_ORACLE_SERIAL
            double ex = S->s1x[i] + tst * S->sdx[i] - X[t], ey = S->s1y[i] + tst * S->sdy[i] - Y[t]; /* temp_vec, :140 */
#else
            double ex = fma(tst, S->sdx[i], -px), ey = fma(tst, S->sdy[i], -py);
#endif
            double d2 = FMA(ex, ex, ey * ey);
            if (d2 < best) { best = d2; bi = i; bex = ex; bey = ey; bth = that; }
        }
        mind2[t] = best;
        if (grad) {
            double ed = (bth >= 0.0 && bth <= 1.0) ? RDIV(FMA(bex, S->sdx[bi], bey * S->sdy[bi]), S->sden[bi], S->sinv[bi]) : 0.0;
            double k2 = 2.0 * S->qcte;
            gX[t] = k2 * FMA(ed, S->sdx[bi], -bex);
            gY[t] = k2 * FMA(ed, S->sdy[bi], -bey);
        }
    }

    /* obstacle penalty F2 (:106-119): circles then ellipses; F2_k = sum over the horizon of max(0, inside) */
    double pen = 0.0;
    for (int k = 0; k < S->Nobs; k++) {
        for (int t = 0; t < N; t++) {
            hdx[t] = X[t] - S->cx[k]; hdy[t] = Y[t] - S->cy[k];
            h[t] = FMA(-hdy[t], hdy[t], FMA(-hdx[t], hdx[t], S->cr2[k]));
            hp[t] = (h[t] > 0.0) ? h[t] : 0.0;
        }
        double g = hsum(hp, N, G, SS);
        if (F2) F2[k] = g;
        pen = FMA(g, g, pen);
        if (grad && g > 0.0) {
            double cg = c * g;
            for (int t = 0; t < N; t++)
                if (h[t] > 0.0) {
                    gX[t] = FMA(cg, -2.0 * hdx[t], gX[t]);
                    gY[t] = FMA(cg, -2.0 * hdy[t], gY[t]);
                }
        }
    }
    for (int k = 0; k < S->Nd; k++) {
        double ta[MAXT], tb[MAXT];
        for (int t = 0; t < N; t++) {
            int i = k * N + t;
            double dx = X[t] - S->ex[i], dy = Y[t] - S->ey[i];
            double ea = FMA(dx, S->eca[i], dy * S->esa[i]);
            double eb = FMA(dx, S->esa[i], -(dy * S->eca[i]));
#ifdef NMPC_ORACLE_SERIAL
            h[t] = 1.0 - (ea * ea) / S->erx2[i] - (eb * eb) / S->ery2[i];
            ta[t] = ea / S->erx2[i]; tb[t] = eb / S->ery2[i];
#else
            h[t] = fma(-(eb * eb), S->eiry2[i], fma(-(ea * ea), S->eirx2[i], 1.0));
            ta[t] = ea * S->eirx2[i]; tb[t] = eb * S->eiry2[i];
#endif
            hp[t] = (h[t] > 0.0) ? h[t] : 0.0;
        }
        double g = hsum(hp, N, G, SS);
        if (F2) F2[S->Nobs + k] = g;
        pen = FMA(g, g, pen);
        if (grad && g > 0.0) {
            double cg = c * g;
            for (int t = 0; t < N; t++)
                if (h[t] > 0.0) {
                    int i = k * N + t;
                    double hX = -2.0 * FMA(ta[t], S->eca[i], tb[t] * S->esa[i]);
                    double hY = -2.0 * FMA(ta[t], S->esa[i], -(tb[t] * S->eca[i]));
                    gX[t] = FMA(cg, hX, gX[t]);
                    gY[t] = FMA(cg, hY, gY[t]);
                }
        }
    }

    /* stage cost (:84-86), acceleration cost and ALM term (:157-171) */
    for (int t = 0; t < N; t++) {
        double v = u[2 * t], w = u[2 * t + 1];
        double vp = t ? u[2 * t - 2] : S->vinit, wp_ = t ? u[2 * t - 1] : S->winit;
        double c0 = S->rv * (v * v);
        c0 = FMA(S->rw, w * w, c0);
        double dv = v - S->vref[t];
        c0 = FMA(S->qv, dv * dv, c0);
        double ex = xpre[t] - S->xref, ey = ypre[t] - S->yref, et = thpre[t] - S->thref;
        c0 = FMA(S->q, FMA(ex, ex, ey * ey), c0);
        c0 = FMA(S->qth, et * et, c0);
        c0 = FMA(S->qcte, mind2[t], c0);
        double acc = RDIV(v - vp, ts, inv_ts), aac = RDIV(w - wp_, ts, inv_ts);
        c0 = FMA(S->ap, acc * acc, c0);
        c0 = FMA(S->wp, aac * aac, c0);
#ifdef NMPC_ORACLE_SERIAL
        double za = acc + (y ? y[t] : 0.0) / cden, zw = aac + (y ? y[N + t] : 0.0) / cden;
#else
        double za = fma(y ? y[t] : 0.0, inv_c, acc), zw = fma(y ? y[N + t] : 0.0, inv_c, aac);
#endif
        double da = sel_excess(za, S->amin, S->amax);   /* z - Proj_C(z) */
        double dw = sel_excess(zw, -S->aamax, S->aamax);
        c0 = FMA(hc, FMA(da, da, dw * dw), c0);
        cl[t] = c0;
        Aa[t] = RDIV(FMA(c, da, (2.0 * S->ap) * acc), ts, inv_ts);
        Aw[t] = RDIV(FMA(c, dw, (2.0 * S->wp) * aac), ts, inv_ts);
        if (F1) { F1[t] = acc; F1[N + t] = aac; }
    }
    (void)inv_c;
    /* terminal cost (:148) */
    double eXN = X[N - 1] - S->xref, eYN = Y[N - 1] - S->yref, eTN = TH[N - 1] - S->thref;
    double term = FMA(S->qN, FMA(eXN, eXN, eYN * eYN), S->qthN * (eTN * eTN));
    double psi = FMA(hc, pen, hsum(cl, N, G, SS) + term);
    if (!grad) return psi;

    /* backward sweep */
    double mth[MAXT], LX[MAXT], LY[MAXT], nn[MAXT], rr[MAXT], TT[MAXT];
    for (int t = 0; t < N; t++) {
        double qq = (t + 1 < N) ? S->q : S->qN, qt = (t + 1 < N) ? S->qth : S->qthN;
        gX[t] = FMA(2.0 * qq, X[t] - S->xref, gX[t]);
        gY[t] = FMA(2.0 * qq, Y[t] - S->yref, gY[t]);
        mth[t] = (2.0 * qt) * (TH[t] - S->thref);
    }
    suffix_scan(gX, N, G, SS, LX);
    suffix_scan(gY, N, G, SS, LY);
    for (int t = 0; t < N; t++) nn[t] = (ts * u[2 * t]) * FMA(cs[t], LY[t], -(sn[t] * LX[t]));
    for (int t = 0; t < N; t++) rr[t] = mth[t] + ((t + 1 < N) ? nn[t + 1] : 0.0);
    suffix_scan(rr, N, G, SS, TT);
    for (int t = 0; t < N; t++) {
        double v = u[2 * t], w = u[2 * t + 1];
        double An = (t + 1 < N) ? Aa[t + 1] : 0.0, Wn = (t + 1 < N) ? Aw[t + 1] : 0.0;
        double lv = FMA(2.0 * S->rv, v, (2.0 * S->qv) * (v - S->vref[t])) + (Aa[t] - An);
        double lw = (2.0 * S->rw) * w + (Aw[t] - Wn);
        grad[2 * t] = FMA(ts, FMA(cs[t], LX[t], sn[t] * LY[t]), lv);
        grad[2 * t + 1] = FMA(ts, TT[t], lw);
    }
    return psi;
}

/* ------------------------------------------------------------------------- */
/* vector helpers on interleaved 2N vectors: per-step partial, then the horizon sum */
static double vdot(const staged* S, const double* a, const double* b) {
    double e[MAXT];
    for (int t = 0; t < S->N; t++) e[t] = FMA(a[2 * t + 1], b[2 * t + 1], a[2 * t] * b[2 * t]);
    return hsum(e, S->N, S->G, S->S);
}
static double vdiff2(const staged* S, const double* a, const double* b) { /* |a-b|^2 */
    double e[MAXT];
    for (int t = 0; t < S->N; t++) {
        double d0 = a[2 * t] - b[2 * t], d1 = a[2 * t + 1] - b[2 * t + 1];
        e[t] = FMA(d1, d1, d0 * d0);
    }
    return hsum(e, S->N, S->G, S->S);
}
static int all_finite(const double* a, int n) {
    for (int i = 0; i < n; i++)
        if (!isfinite(a[i])) return 0;
    return 1;
}

/* ------------------------------------------------------------------------- */
/* L-BFGS buffer — restates crate `lbfgs` (lib.rs: Lbfgs::new/reset/apply_hessian/
 * update_hessian/new_s_and_y_valid) as configured by PANOCCache::new
 * (cbfgs alpha 1, cbfgs epsilon 1e-8, sy epsilon 1e-10).  Slot 0 is the newest
 * pair; the extra slot is the staging area that rotate_right(1) moves to the front. */
typedef struct {
    const staged* S;
    int n2, mem, active, first_old, head;
    double lgamma;
    double s[MEMP1][2 * MAXT], y[MEMP1][2 * MAXT];
    double rho[MEMP1], alpha[NMPC_LBFGS_MAX];
    double syd[MEMP1]; /* by physical slot p: <s_p, y_q> with q the pair accepted right after p (contract build) */
    double old_state[2 * MAXT], old_g[2 * MAXT];
} lbfgs_t;

static int lb_slot(const lbfgs_t* L, int i) { return (L->head + i) % (L->mem + 1); }
static void lb_reset(lbfgs_t* L) { L->active = 0; L->first_old = 1; }

static void lb_update(lbfgs_t* L, const double* g, const double* state) {
    const int n2 = L->n2;
    if (L->first_old) {
        L->first_old = 0;
        memcpy(L->old_state, state, n2 * sizeof(double));
        memcpy(L->old_g, g, n2 * sizeof(double));
        return;
    }
    int tmp = lb_slot(L, L->mem);
    double* s = L->s[tmp];
    double* y = L->y[tmp];
    for (int i = 0; i < n2; i++) { s[i] = state[i] - L->old_state[i]; y[i] = g[i] - L->old_g[i]; }
    double ys = vdot(L->S, s, y);
    double ss = vdot(L->S, s, s);
    L->rho[tmp] = 1.0 / ys;
    if (ss <= DBL_EPS || ys <= SY_EPSILON) return; /* rejection */
    double lhs = ys / ss;
    double rhs = CBFGS_EPSILON * sqrt(vdot(L->S, g, g)); /* eps * |g|^alpha, alpha = 1 */
    if (!(lhs > rhs && isfinite(lhs) && isfinite(rhs))) return;
    memcpy(L->old_state, state, n2 * sizeof(double));
    memcpy(L->old_g, g, n2 * sizeof(double));
    double yy = vdot(L->S, y, y);
    if (L->active > 0) { /* Gram entry of the pair that was the newest so far against the new one (see lb_apply) */
        int prev = lb_slot(L, 0);
        L->syd[prev] = vdot(L->S, L->s[prev], y);
    }
    L->head = (L->head + L->mem) % (L->mem + 1); /* rotate_right(1): staging slot becomes slot 0 */
    L->lgamma = (1.0 / L->rho[tmp]) / yy;
    L->active = (L->active + 1 < L->mem) ? L->active + 1 : L->mem;
}

/* Two-loop recursion of the lbfgs crate (Lbfgs::apply_hessian).
 *
 * Serial build: the literal recursion — 2*active sequential reductions.
 *
 * Contract build (shared with the CUDA kernel): the same recursion taken TWO steps at a time.  For consecutive
 * pairs k (newer) and k+1 of the forward loop
 *     alpha_k     = rho_k     <s_k, q>
 *     alpha_{k+1} = rho_{k+1} <s_{k+1}, q - alpha_k y_k> = rho_{k+1} ( <s_{k+1}, q> - alpha_k <s_{k+1}, y_k> ),
 * so both inner products are taken against the SAME q (one interleaved warp reduction instead of two dependent
 * ones) and <s_{k+1}, y_k> is a Gram entry that only changes when a pair is accepted (syd[], one extra inner
 * product per update).  The backward loop pairs k and k-1 the same way with <y_{k-1}, s_k> = syd[slot k].
 * Algebraically identical to the literal recursion; it differs in rounding only. */
static void lb_apply(lbfgs_t* L, double* q) {
    const int n2 = L->n2;
    if (L->active == 0) return; /* empty buffer: H = I */
#ifdef NMPC_ORACLE_SERIAL
    for (int k = 0; k < L->active; k++) {
        int sl = lb_slot(L, k);
        double al = L->rho[sl] * vdot(L->S, L->s[sl], q);
        L->alpha[k] = al;
        for (int i = 0; i < n2; i++) q[i] = FMA(-al, L->y[sl][i], q[i]);
    }
    for (int i = 0; i < n2; i++) q[i] = q[i] * L->lgamma;
    for (int k = L->active - 1; k >= 0; k--) {
        int sl = lb_slot(L, k);
        double beta = L->rho[sl] * vdot(L->S, L->y[sl], q);
        double co = L->alpha[k] - beta;
        for (int i = 0; i < n2; i++) q[i] = FMA(co, L->s[sl][i], q[i]);
    }
#else
    const int m = L->active;
    int k = 0;
    for (; k + 1 < m; k += 2) {
        int s0 = lb_slot(L, k), s1 = lb_slot(L, k + 1);
        double pa = vdot(L->S, L->s[s0], q), pb = vdot(L->S, L->s[s1], q);
        double al0 = L->rho[s0] * pa;
        double al1 = L->rho[s1] * fma(-al0, L->syd[s1], pb);
        L->alpha[k] = al0;
        L->alpha[k + 1] = al1;
        for (int i = 0; i < n2; i++) q[i] = fma(-al1, L->y[s1][i], fma(-al0, L->y[s0][i], q[i]));
    }
    if (k < m) {
        int sl = lb_slot(L, k);
        double al = L->rho[sl] * vdot(L->S, L->s[sl], q);
        L->alpha[k] = al;
        for (int i = 0; i < n2; i++) q[i] = fma(-al, L->y[sl][i], q[i]);
    }
    for (int i = 0; i < n2; i++) q[i] = q[i] * L->lgamma;
    k = m - 1;
    for (; k >= 1; k -= 2) {
        int s0 = lb_slot(L, k), s1 = lb_slot(L, k - 1);
        double qa = vdot(L->S, L->y[s0], q), qb = vdot(L->S, L->y[s1], q);
        double c0 = L->alpha[k] - L->rho[s0] * qa;
        double c1 = L->alpha[k - 1] - L->rho[s1] * fma(c0, L->syd[s0], qb);
        for (int i = 0; i < n2; i++) q[i] = fma(c1, L->s[s1][i], fma(c0, L->s[s0][i], q[i]));
    }
    if (k == 0) {
        int sl = lb_slot(L, 0);
        double beta = L->rho[sl] * vdot(L->S, L->y[sl], q);
        double co = L->alpha[0] - beta;
        for (int i = 0; i < n2; i++) q[i] = fma(co, L->s[sl][i], q[i]);
    }
#endif
}

/* ------------------------------------------------------------------------- */
/* PANOC — restates PANOCEngine::{init, step, ...} (panoc_engine.rs),
 * PANOCCache::{exit_condition, akkt_residual} (panoc_cache.rs),
 * LipschitzEstimator::estimate_local_lipschitz (lipschitz_estimator.rs) and
 * PANOCOptimizer::solve (panoc_optimizer.rs). */
typedef struct {
    staged* S;
    double c;
    const double* y;
    int n2, iteration;
    double gamma, inv_gamma, sigma, lip, cost, norm_fpr, tau, tol, akkt_tol;
    double grad[2 * MAXT], uhalf[2 * MAXT], fpr[2 * MAXT], dir[2 * MAXT], gstep[2 * MAXT], uplus[2 * MAXT];
    double grad_prev[2 * MAXT]; /* variant 1 only */
    lbfgs_t lb;
} panoc_t;

/* Rectangle::project (constraints/rectangle.rs): comparison-based, so a NaN stays a NaN */
static inline double clampd(double x, double lo, double hi) { return (x < lo) ? lo : ((x > hi) ? hi : x); }
static void project_U(const staged* S, double* v) { /* Rectangle U, src/mpc/mpc_generator.py:151-153 */
    for (int t = 0; t < S->N; t++) {
        v[2 * t] = clampd(v[2 * t], S->vmin, S->vmax);
        v[2 * t + 1] = clampd(v[2 * t + 1], -S->wmax, S->wmax);
    }
}
static void grad_step_half(panoc_t* C, const double* u) { /* gradient_step() + half_step() */
    for (int i = 0; i < C->n2; i++) { C->gstep[i] = FMA(-C->gamma, C->grad[i], u[i]); C->uhalf[i] = C->gstep[i]; }
    project_U(C->S, C->uhalf);
}
static void compute_fpr(panoc_t* C, const double* u) {
    double e[MAXT];
    for (int t = 0; t < C->S->N; t++) {
        double d0 = u[2 * t] - C->uhalf[2 * t], d1 = u[2 * t + 1] - C->uhalf[2 * t + 1];
        C->fpr[2 * t] = d0; C->fpr[2 * t + 1] = d1;
        e[t] = FMA(d1, d1, d0 * d0);
    }
    C->norm_fpr = sqrt(hsum(e, C->S->N, C->S->G, C->S->S));
}
static void set_gamma(panoc_t* C, double g) { C->gamma = g; C->inv_gamma = 1.0 / g; }

static void panoc_init(panoc_t* C, double* u) {
    staged* S = C->S;
    const int N = S->N, n2 = C->n2;
    lb_reset(&C->lb);
    C->tau = 1.0; C->iteration = 0;
    /* cost and gradient at u; estimate_loc_lip perturbs u by h and LEAVES it perturbed */
    C->cost = eval_psi(S, u, C->c, C->y, C->grad, 0, 0);
    double hv[2 * MAXT], gh[2 * MAXT], e[MAXT];
    for (int i = 0; i < n2; i++) {
        double e_ = EPSILON_LIPSCHITZ * u[i];
        hv[i] = (e_ > DELTA_LIPSCHITZ) ? e_ : DELTA_LIPSCHITZ; /* max{delta, epsilon*u} */
    }
    for (int t = 0; t < N; t++) e[t] = FMA(hv[2 * t + 1], hv[2 * t + 1], hv[2 * t] * hv[2 * t]);
    double norm_h = sqrt(hsum(e, N, S->G, S->S));
    for (int i = 0; i < n2; i++) u[i] = u[i] + hv[i];
    eval_psi(S, u, C->c, C->y, gh, 0, 0);
    C->lip = sqrt(vdiff2(S, gh, C->grad)) / norm_h;
    if (g_variant & 4)
        for (int i = 0; i < n2; i++) u[i] = u[i] - hv[i];
    memset(C->grad_prev, 0, sizeof(C->grad_prev));
    set_gamma(C, GAMMA_L_COEFF / fmax(C->lip, MIN_L_ESTIMATE));
    C->sigma = (1.0 - GAMMA_L_COEFF) / (4.0 * C->gamma);
    grad_step_half(C, u);
}

/* returns 1 to continue, 0 when the exit condition holds */
static int panoc_step(panoc_t* C, double* u) {
    staged* S = C->S;
    const int N = S->N, n2 = C->n2;
    compute_fpr(C, u);
    /* exit_condition(): |gamma*fpr| < tol  AND  akkt residual < eps_nu.
     * akkt_residual = | fpr/gamma + df - df_prev | where cache_previous_gradient()
     * has just copied df into df_prev for iteration >= 1 (zeros at iteration 0). */
    if (C->norm_fpr < C->tol) {
        double e[MAXT];
        for (int t = 0; t < N; t++) {
            double g0 = C->grad[2 * t], g1 = C->grad[2 * t + 1];
            double p0 = C->iteration ? g0 : 0.0, p1 = C->iteration ? g1 : 0.0;
            if (g_variant & 1) { p0 = C->grad_prev[2 * t]; p1 = C->grad_prev[2 * t + 1]; }
#ifdef NMPC_ORACLE_SERIAL
            double r0 = C->fpr[2 * t] / C->gamma + g0 - p0;
            double r1 = C->fpr[2 * t + 1] / C->gamma + g1 - p1;
#else
            double r0 = fma(C->fpr[2 * t], C->inv_gamma, g0) - p0;
            double r1 = fma(C->fpr[2 * t + 1], C->inv_gamma, g1) - p1;
#endif
            e[t] = FMA(r1, r1, r0 * r0);
        }
        if (sqrt(hsum(e, N, S->G, S->S)) < C->akkt_tol) return 0;
    }
    /* update_lipschitz_constant() */
    double cost_half = eval_psi(S, C->uhalf, C->c, C->y, 0, 0, 0);
    C->cost = eval_psi(S, u, C->c, C->y, 0, 0, 0);
    int it = 0;
    for (;;) {
        double ip = vdot(S, C->grad, C->fpr);
        double rhs = C->cost + LIPSCHITZ_UPDATE_EPSILON * fabs(C->cost) - ip +
                     (GAMMA_L_COEFF * 0.5 * C->inv_gamma) * (C->norm_fpr * C->norm_fpr);
        if (!(cost_half > rhs && it < MAX_LIPSCHITZ_UPDATE_ITERATIONS && C->lip < MAX_LIPSCHITZ_CONSTANT)) break;
        lb_reset(&C->lb);
        C->lip *= 2.0;
        set_gamma(C, C->gamma / 2.0);
        grad_step_half(C, u);
        cost_half = eval_psi(S, C->uhalf, C->c, C->y, 0, 0, 0);
        compute_fpr(C, u);
        it++;
    }
    C->sigma = (1.0 - GAMMA_L_COEFF) / (4.0 * C->gamma);
    if (g_variant & 1) memcpy(C->grad_prev, C->grad, n2 * sizeof(double));
    /* lbfgs_direction() */
    lb_update(&C->lb, C->fpr, u);
    if (C->iteration > 0) {
        memcpy(C->dir, C->fpr, n2 * sizeof(double));
        lb_apply(&C->lb, C->dir);
    }
    if (C->iteration == 0) { /* update_no_linesearch() */
        memcpy(u, C->uhalf, n2 * sizeof(double));
        C->cost = eval_psi(S, u, C->c, C->y, C->grad, 0, 0);
        grad_step_half(C, u);
    } else { /* linesearch(): FBE decrease; up to MAX+1 trial points, the last is kept */
        double dist2 = vdiff2(S, C->gstep, C->uhalf);
        double fbe = C->cost - (0.5 * C->gamma) * vdot(S, C->grad, C->grad) + (0.5 * dist2) * C->inv_gamma;
        double rhs_ls = fbe - C->sigma * (C->norm_fpr * C->norm_fpr);
        C->tau = 1.0;
        int nls = 0;
        for (;;) {
            double om = 1.0 - C->tau;
            for (int i = 0; i < n2; i++) C->uplus[i] = FMA(-C->tau, C->dir[i], FMA(-om, C->fpr[i], u[i]));
            C->cost = eval_psi(S, C->uplus, C->c, C->y, C->grad, 0, 0);
            grad_step_half(C, C->uplus);
            double d2 = vdiff2(S, C->gstep, C->uhalf);
            double lhs = C->cost - (0.5 * C->gamma) * vdot(S, C->grad, C->grad) + (0.5 * d2) * C->inv_gamma;
            if (!(lhs > rhs_ls && nls < MAX_LINESEARCH_ITERATIONS)) break;
            C->tau /= 2.0;
            nls++;
        }
        if ((g_variant & 2) && nls == MAX_LINESEARCH_ITERATIONS) {
            /* fallback: u <- the current half step (the last trial's projected gradient step); cost / gradient there */
            double tmpu[2 * MAXT];
            memcpy(tmpu, C->uhalf, n2 * sizeof(double));
            memcpy(C->uplus, tmpu, n2 * sizeof(double));
            C->cost = eval_psi(S, C->uplus, C->c, C->y, C->grad, 0, 0);
            grad_step_half(C, C->uplus);
        }
        memcpy(u, C->uplus, n2 * sizeof(double));
    }
    if (g_trace)
        fprintf(stderr, "it %d cost %.12e nfpr %.3e gamma %.3e tau %.3e active %d lip %.3e\n", C->iteration, C->cost,
                C->norm_fpr, C->gamma, C->tau, C->lb.active, C->lip);
    C->iteration++;
    return 1;
}

/* PANOCOptimizer::solve: note `step` runs once more after the iteration budget is
 * spent and the status is decided by the budget flag alone. */
static int panoc_solve(panoc_t* C, double* u, int max_iter, int* iters) {
    panoc_init(C, u);
    int num_iter = 0, cont = 1;
    int flag = panoc_step(C, u);
    while (flag && cont) {
        num_iter++;
        cont = num_iter < max_iter;
        flag = panoc_step(C, u);
    }
    *iters = num_iter;
    if (!all_finite(u, C->n2)) return NMPC_NOT_FINITE;
    memcpy(u, C->uhalf, C->n2 * sizeof(double));
    return cont ? NMPC_CONVERGED : NMPC_NOT_CONVERGED_ITERATIONS;
}

/* ------------------------------------------------------------------------- */
/* per-thread workspace: nothing is allocated inside a solve */
typedef struct {
    staged S;
    panoc_t C;
    double* F2;
    size_t F2_len;
} workspace;

static workspace* ws_new(void) { return (workspace*)calloc(1, sizeof(workspace)); }
static void ws_free(workspace* w) {
    if (!w) return;
    free(w->S.buf);
    free(w->F2);
    free(w);
}

/* ALM / penalty outer loop — restates AlmOptimizer::{solve, step,
 * update_lagrange_multipliers, is_exit_criterion_satisfied,
 * is_penalty_stall_criterion, final_cache_update} (alm/alm_optimizer.rs) with the
 * settings the generated optimizer.rs passes (opengen 0.6.4 defaults). */
static int solve_ws(workspace* W, const nmpc_config* cfg, const double* p, double* u, double* y, nmpc_stats* st) {
    staged* S = &W->S;
    int rc = stage(S, cfg, p);
    if (rc) return -rc;
    const int N = S->N, n2 = 2 * N;
    panoc_t* C = &W->C;
    double yp[2 * MAXT], w[2 * MAXT], ybuf[2 * MAXT];
    size_t nf2 = (size_t)S->Nobs + S->Nd + 1;
    if (nf2 > W->F2_len) {
        free(W->F2);
        W->F2 = (double*)malloc(nf2 * sizeof(double));
        W->F2_len = W->F2 ? nf2 : 0;
        if (!W->F2) return -3;
    }
    double* F2 = W->F2;
    if (!y) { memset(ybuf, 0, sizeof(ybuf)); y = ybuf; }
    C->S = S; C->n2 = n2; C->y = y;
    C->lb.S = S; C->lb.n2 = n2; C->lb.mem = S->mem; C->lb.head = 0;
    C->tol = cfg->tolerance;
    C->c = cfg->initial_penalty;
    C->akkt_tol = cfg->initial_tolerance;
    C->norm_fpr = 0.0;
    int iteration = 0, inner_total = 0, num_outer = 0, status = NMPC_CONVERGED, done = 0;
    double f2n = 0.0, f2np = 0.0, dyn = 0.0, dynp = 0.0;
    for (int outer = 0; outer < cfg->max_outer_iterations; outer++) {
        num_outer++;
        for (int i = 0; i < n2; i++) y[i] = clampd(y[i], -Y_SET_BOUND, Y_SET_BOUND); /* project_on_set_y */
        int iters = 0;
        int inner = panoc_solve(C, u, cfg->max_inner_iterations, &iters);
        inner_total += iters;
        if (inner == NMPC_NOT_FINITE) { status = NMPC_NOT_FINITE; done = 2; break; }
        status = inner;
        /* y+ = y + c*(F1(u) - Proj_C(F1(u) + y/c)) ; F2(u) */
        eval_psi(S, u, 0.0, 0, 0, w, F2);
        S->n_cost--; /* F1/F2 mappings, not a psi evaluation */
        double e[MAXT];
        for (int t = 0; t < N; t++) {
            double za = w[t] + y[t] / C->c, zw = w[N + t] + y[N + t] / C->c;
            za = clampd(za, S->amin, S->amax);
            zw = clampd(zw, -S->aamax, S->aamax);
            yp[t] = FMA(C->c, w[t] - za, y[t]);
            yp[N + t] = FMA(C->c, w[N + t] - zw, y[N + t]);
            double d0 = yp[t] - y[t], d1 = yp[N + t] - y[N + t];
            e[t] = FMA(d1, d1, d0 * d0);
        }
        dynp = sqrt(hsum(e, N, S->G, S->S));
        double acc = 0.0;
        for (int k = 0; k < S->Nobs + S->Nd; k++) acc = FMA(F2[k], F2[k], acc);
        f2np = sqrt(acc);
        int crit1 = (iteration > 0 || (g_variant & 8)) && dynp <= C->c * cfg->delta_tolerance + DBL_EPS;
        int crit2 = (S->Nobs + S->Nd == 0) || f2np <= cfg->delta_tolerance + DBL_EPS;
        int crit3 = C->akkt_tol <= cfg->tolerance + DBL_EPS;
        if (crit1 && crit2 && crit3) { done = 1; break; }
        int stall;
        if (iteration == 0 && !(g_variant & 16)) stall = 1;
        else {
            int ca = dynp <= cfg->sufficient_decrease_coeff * dyn + DBL_EPS;
            int cp = f2np <= cfg->sufficient_decrease_coeff * f2n + DBL_EPS;
            stall = (S->Nobs + S->Nd > 0) ? (ca && cp) : ca;
        }
        if (!stall) C->c *= cfg->penalty_update_factor;
        C->akkt_tol = fmax(C->akkt_tol * cfg->inner_tolerance_update, cfg->tolerance);
        iteration++;
        dyn = dynp; f2n = f2np;
        memcpy(y, yp, n2 * sizeof(double));
    }
    if (done != 2 && num_outer == cfg->max_outer_iterations) status = NMPC_NOT_CONVERGED_ITERATIONS;
    if (st) {
        st->exit_status = status; st->outer_iterations = num_outer; st->inner_iterations = inner_total;
        st->last_norm_fpr = C->norm_fpr; st->delta_y_norm_over_c = dynp / C->c; st->f2_norm = f2np;
        st->penalty = C->c;
        st->cost = (status == NMPC_NOT_FINITE) ? NAN : eval_psi(S, u, 0.0, 0, 0, 0, 0);
        if (status != NMPC_NOT_FINITE) S->n_cost--;
        st->n_cost_evals = S->n_cost; st->n_grad_evals = S->n_grad; st->reserved = 0;
    }
    return status;
}

int nmpc_oracle_solve(const nmpc_config* cfg, const double* p, double* u, double* y, nmpc_stats* st) {
    workspace* W = ws_new();
    if (!W) return -3;
    int rc = solve_ws(W, cfg, p, u, y, st);
    ws_free(W);
    return rc;
}

int nmpc_oracle_eval(const nmpc_config* cfg, const double* p, const double* u, double c, const double* y,
                     double* psi, double* grad, double* F1, double* F2) {
    workspace* W = ws_new();
    if (!W) return -3;
    int rc = stage(&W->S, cfg, p);
    if (!rc) {
        double gtmp[2 * MAXT];
        double v = eval_psi(&W->S, u, c, y, grad ? grad : gtmp, F1, F2);
        if (psi) *psi = v;
    }
    ws_free(W);
    return rc ? -rc : 0;
}

/* batch drivers: one problem per OpenMP thread (nthreads <= 0: all cores) */
int nmpc_oracle_solve_batch(const nmpc_config* cfg, int32_t B, const double* Pm, double* U, double* Y,
                            int32_t* status, nmpc_stats* stats, int nthreads) {
    const int np = nmpc_param_len(cfg), n2 = 2 * cfg->N_hor;
    int bad = 0;
    g_trace = getenv("NMPC_ORACLE_TRACE") != NULL;
    g_variant = getenv("NMPC_ORACLE_VARIANT") ? atoi(getenv("NMPC_ORACLE_VARIANT")) : 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
#endif
    {
        workspace* W = ws_new();
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int32_t b = 0; b < B; b++) {
            nmpc_stats st;
            memset(&st, 0, sizeof(st));
            int s = W ? solve_ws(W, cfg, Pm + (size_t)b * np, U + (size_t)b * n2, Y ? Y + (size_t)b * n2 : 0, &st) : -3;
            if (s < 0) bad = 1;
            if (status) status[b] = s;
            if (stats) stats[b] = st;
        }
        ws_free(W);
    }
    (void)nthreads;
    return bad ? NMPC_ERR_INVALID : NMPC_OK;
}

int nmpc_oracle_eval_batch(const nmpc_config* cfg, int32_t B, const double* Pm, const double* U, const double* c,
                           const double* Y, double* psi, double* grad, double* F1, double* F2) {
    const int np = nmpc_param_len(cfg), n2 = 2 * cfg->N_hor, nf2 = cfg->Nobs + cfg->Ndynobs;
    for (int32_t b = 0; b < B; b++) {
        int rc = nmpc_oracle_eval(cfg, Pm + (size_t)b * np, U + (size_t)b * n2, c[b], Y ? Y + (size_t)b * n2 : 0,
                                  psi ? psi + b : 0, grad ? grad + (size_t)b * n2 : 0, F1 ? F1 + (size_t)b * n2 : 0,
                                  F2 ? F2 + (size_t)b * nf2 : 0);
        if (rc) return NMPC_ERR_INVALID;
    }
    return NMPC_OK;
}

int nmpc_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int nmpc_oracle_is_serial(void) {
#ifdef NMPC_ORACLE_SERIAL
    return 1;
#else
    return 0;
#endif
}

/* host copies of the config helpers so the oracle is self-contained */
void nmpc_default_config(nmpc_config* cfg) {
    memset(cfg, 0, sizeof(*cfg));
    cfg->N_hor = 20; cfg->Nobs = 10; cfg->Ndynobs = 3; /* configs/default.yaml:7,38,39 */
    cfg->lbfgs_memory = 10; cfg->max_inner_iterations = 500; cfg->max_outer_iterations = 10;
    cfg->ts = 0.2;                                                          /* :18 */
    cfg->lin_vel_min = -0.5; cfg->lin_vel_max = 1.5; cfg->ang_vel_max = 0.5; /* :8-9,12 */
    cfg->lin_acc_min = -1.0; cfg->lin_acc_max = 1.0; cfg->ang_acc_max = 3.0; /* :10-11,13 */
    cfg->tolerance = 1e-4; cfg->initial_tolerance = 1e-4; cfg->delta_tolerance = 1e-4;
    cfg->inner_tolerance_update = 0.1; cfg->penalty_update_factor = 5.0; cfg->initial_penalty = 1.0;
    cfg->sufficient_decrease_coeff = 0.1;
}
int32_t nmpc_param_len(const nmpc_config* cfg) {
    return NMPC_NZ + cfg->N_hor + 3 * cfg->Nobs + 5 * cfg->Ndynobs * cfg->N_hor + 3 * cfg->N_hor;
}
