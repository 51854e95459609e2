// nmpc_kernels.cu — B200 (sm_100a) batched NMPC solver: one persistent warp per problem.
//
// Replaces what sits behind `mng.call(parameters)` in the reference
// (src/mpc/mpc_generator.py:206): the OpEn-generated solver for the problem that
// MpcModule.build() defines (src/mpc/mpc_generator.py:66-193).  A warp owns one NMPC
// instance at a time and pulls the next one from an atomic queue:
//   * lane l owns horizon steps t = l + 32*j (P = ceil(N/32) register passes);
//   * the diff-drive rollout (src/mpc/mpc_generator.py:88-90) and the adjoint sweep are
//     Kogge-Stone scans over lanes (theta, then x/y; Lambda_x/Lambda_y, then Theta);
//   * each lane walks the N-1 reference segments (cross-track error, :122-144), the
//     static circles and its own time slice of the dynamic ellipses (:93-119) for its
//     own predicted point, with the per-problem constants staged in shared memory;
//   * PANOC vectors, the L-BFGS (s, y) ring and the staged problem live in the warp's
//     shared-memory arena; reductions are xor-butterflies; scalars are warp-uniform.
// The arithmetic contract (operation order, explicit fma, own sincos) is the one stated
// in DESIGN.md §4; compile with --fmad=false so nothing else is contracted.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <new>

#include "../../include/nmpc_b200.h"

#define FULL 0xffffffffu
#define MEMP1 (NMPC_LBFGS_MAX + 1)

// OpEn PANOC constants (panoc_engine.rs) — see oracle/nmpc_oracle.c for the restatement notes
#define MIN_L_ESTIMATE 1e-10
#define GAMMA_L_COEFF 0.95
#define DELTA_LIPSCHITZ 1e-12
#define EPSILON_LIPSCHITZ 1e-6
#define LIPSCHITZ_UPDATE_EPSILON 1e-6
#define MAX_LIPSCHITZ_UPDATE_ITERATIONS 10
#define MAX_LIPSCHITZ_CONSTANT 1e9
#define MAX_LINESEARCH_ITERATIONS 10
#define CBFGS_EPSILON 1e-8
#define SY_EPSILON 1e-10
#define DBL_EPS 2.220446049250313e-16
#define Y_SET_BOUND 1e12

extern __shared__ __align__(16) double smem[];

// ---------------------------------------------------------------------------------
// per-warp shared-memory arena (offsets in doubles; every block is 16-byte aligned)
enum { V_GRAD = 0, V_UHALF, V_FPR, V_DIR, V_GSTEP, V_UPLUS, V_OLDS, V_OLDG, V_S, V_Y = V_S + MEMP1, V_END = V_Y + MEMP1 };
enum { H_X0 = 0, H_Y0, H_TH0, H_VINIT, H_WINIT, H_XREF, H_YREF, H_THREF, H_Q, H_QV, H_QTH, H_RV, H_RW, H_QN, H_QTHN,
       H_QCTE, H_AP, H_WP, H_INVTS, H_COUNT = 20 };

struct Lay {
    int n2, s1, sd, sinv, circ, ell, rho, alpha, hdr, vref, total;
};
__host__ __device__ inline int even_up(int x) { return (x + 1) & ~1; }
__host__ __device__ inline Lay make_layout(int N, int Nobs, int Nd) {
    Lay L;
    L.n2 = 2 * N;
    int o = V_END * 2 * N;
    L.s1 = o; o += 2 * N;
    L.sd = o; o += 2 * N;
    L.sinv = o; o += even_up(N);
    L.circ = o; o += 4 * Nobs;
    L.ell = o; o += even_up(6 * Nd * N);
    L.rho = o; o += 12;
    L.alpha = o; o += 12;
    L.hdr = o; o += H_COUNT;
    L.vref = o; o += even_up(N);
    L.total = o;
    return L;
}

struct KArgs {
    nmpc_config cfg;
    int B, np;
    const double* P;
    double* U;
    double* Y;
    int32_t* status;
    nmpc_stats* stats;
    unsigned int* counter;
    // eval kernel only
    const double* cvec;
    double *psi, *grad, *F1, *F2;
};

// ---------------------------------------------------------------------------------
// sincos: Cody-Waite by pi/2 with fma, fdlibm kernel polynomials (same as the oracle)
__device__ __forceinline__ void nm_sincos(double x, double& s, double& c) {
    if (!(fabs(x) < 1.0e8)) {
        s = CUDART_NAN;
        c = CUDART_NAN;
        return;
    }
    double kf = rint(x * 6.36619772367581382433e-01);
    double r = fma(-kf, 1.57079632679489655800e+00, x);
    r = fma(-kf, 6.12323399573676603587e-17, r);
    r = fma(-kf, -1.49738490485916983294e-33, r);
    int k = (int)kf;
    double z = r * r;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    double sr = fma(r * z, ps, r);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    double cr = fma(z * z, pc, fma(-0.5, z, 1.0));
    int q = k & 3;
    double s0 = (q & 1) ? cr : sr;
    double c0 = (q & 1) ? sr : cr;
    s = (q & 2) ? -s0 : s0;
    c = ((q + 1) & 2) ? -c0 : c0;
}

// Rectangle::project of OpEn is comparison-based: a NaN stays a NaN (and ends the solve as NotFinite)
__device__ __forceinline__ double clampd(double x, double lo, double hi) { return (x < lo) ? lo : ((x > hi) ? hi : x); }

// ---------------------------------------------------------------------------------
// warp-ordered reductions
__device__ __forceinline__ double butterfly(double a) {
#pragma unroll
    for (int off = 16; off; off >>= 1) a = a + __shfl_xor_sync(FULL, a, off);
    return a;
}
template <int P>
__device__ __forceinline__ double hsum(const double (&e)[P]) {
    double a = e[0];
#pragma unroll
    for (int j = 1; j < P; j++) a = a + e[j];
    return butterfly(a);
}
template <int P>
__device__ __forceinline__ void prefix_scan(const double (&x)[P], double (&incl)[P], double (&excl)[P], int lane) {
    double carry = 0.0;
#pragma unroll
    for (int j = 0; j < P; j++) {
        double l = x[j];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            double y = __shfl_up_sync(FULL, l, off);
            if (lane >= off) l = l + y;
        }
        double lm1 = __shfl_up_sync(FULL, l, 1);
        double g = (j == 0) ? l : carry + l;
        excl[j] = (lane == 0) ? carry : ((j == 0) ? lm1 : carry + lm1);
        incl[j] = g;
        carry = __shfl_sync(FULL, g, 31);
    }
}
template <int P>
__device__ __forceinline__ void prefix_scan2(const double (&xa)[P], const double (&xb)[P], double (&ia)[P],
                                             double (&ea)[P], double (&ib)[P], double (&eb)[P], int lane) {
    double ca = 0.0, cb = 0.0;
#pragma unroll
    for (int j = 0; j < P; j++) {
        double la = xa[j], lb = xb[j];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            double ya = __shfl_up_sync(FULL, la, off);
            double yb = __shfl_up_sync(FULL, lb, off);
            if (lane >= off) {
                la = la + ya;
                lb = lb + yb;
            }
        }
        double ma = __shfl_up_sync(FULL, la, 1), mb = __shfl_up_sync(FULL, lb, 1);
        double ga = (j == 0) ? la : ca + la, gb = (j == 0) ? lb : cb + lb;
        ea[j] = (lane == 0) ? ca : ((j == 0) ? ma : ca + ma);
        eb[j] = (lane == 0) ? cb : ((j == 0) ? mb : cb + mb);
        ia[j] = ga;
        ib[j] = gb;
        ca = __shfl_sync(FULL, ga, 31);
        cb = __shfl_sync(FULL, gb, 31);
    }
}
template <int P>
__device__ __forceinline__ void suffix_scan(const double (&x)[P], double (&suf)[P], int lane) {
    double carry = 0.0;
#pragma unroll
    for (int j = P - 1; j >= 0; j--) {
        double l = x[j];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            double y = __shfl_down_sync(FULL, l, off);
            if (lane + off < 32) l = l + y;
        }
        double g = (j == P - 1) ? l : carry + l;
        suf[j] = g;
        carry = __shfl_sync(FULL, g, 0);
    }
}
template <int P>
__device__ __forceinline__ void suffix_scan2(const double (&xa)[P], const double (&xb)[P], double (&sa)[P],
                                             double (&sb)[P], int lane) {
    double ca = 0.0, cb = 0.0;
#pragma unroll
    for (int j = P - 1; j >= 0; j--) {
        double la = xa[j], lb = xb[j];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            double ya = __shfl_down_sync(FULL, la, off);
            double yb = __shfl_down_sync(FULL, lb, off);
            if (lane + off < 32) {
                la = la + ya;
                lb = lb + yb;
            }
        }
        double ga = (j == P - 1) ? la : ca + la, gb = (j == P - 1) ? lb : cb + lb;
        sa[j] = ga;
        sb[j] = gb;
        ca = __shfl_sync(FULL, ga, 0);
        cb = __shfl_sync(FULL, gb, 0);
    }
}
template <int P>
__device__ __forceinline__ double wdot(const double2 (&a)[P], const double2 (&b)[P]) {
    double e[P];
#pragma unroll
    for (int j = 0; j < P; j++) e[j] = fma(a[j].y, b[j].y, a[j].x * b[j].x);
    return hsum<P>(e);
}
template <int P>
__device__ __forceinline__ double wdiff2(const double2 (&a)[P], const double2 (&b)[P]) {
    double e[P];
#pragma unroll
    for (int j = 0; j < P; j++) {
        double d0 = a[j].x - b[j].x, d1 = a[j].y - b[j].y;
        e[j] = fma(d1, d1, d0 * d0);
    }
    return hsum<P>(e);
}

// ---------------------------------------------------------------------------------
// stage one problem: unpack the parameter row (layout: include/nmpc_b200.h) into the arena
__device__ void stage_problem(const nmpc_config& cfg, const Lay& L, int wb, int lane, const double* __restrict__ p) {
    const int N = cfg.N_hor, Nobs = cfg.Nobs, Nd = cfg.Ndynobs;
    double* hdr = smem + wb + L.hdr;
    if (lane < 8) hdr[lane] = p[lane];
    if (lane >= 8 && lane < 18) hdr[lane] = p[lane + 2];
    if (lane == 18) hdr[H_INVTS] = 1.0 / cfg.ts;
    for (int t = lane; t < N; t += 32) smem[wb + L.vref + t] = p[NMPC_NZ + t];
    const double* pc = p + NMPC_NZ + N;
    for (int k = lane; k < Nobs; k += 32) {
        double r = pc[3 * k + 2];
        double* c4 = smem + wb + L.circ + 4 * k;
        c4[0] = pc[3 * k];
        c4[1] = pc[3 * k + 1];
        c4[2] = r * r;
        c4[3] = 0.0;
    }
    const double* pe = pc + 3 * Nobs;
    const int ne = Nd * N;
    double* el = smem + wb + L.ell;  // SoA: ex, ey, cosA, sinA, 1/rx^2, 1/ry^2, each [Nd*N] (index k*N + t)
    for (int i = lane; i < ne; i += 32) {
        const double* e = pe + 5 * i;  // obstacle-major then time: offset k*5N + 5t = 5*(k*N + t)
        double sa, ca;
        nm_sincos(e[4], sa, ca);
        el[i] = e[0];
        el[ne + i] = e[1];
        el[2 * ne + i] = ca;
        el[3 * ne + i] = sa;
        el[4 * ne + i] = 1.0 / (e[2] * e[2]);
        el[5 * ne + i] = 1.0 / (e[3] * e[3]);
    }
    const double* pr = pe + 5 * ne;
    double2* s1 = reinterpret_cast<double2*>(smem + wb + L.s1);
    double2* sd = reinterpret_cast<double2*>(smem + wb + L.sd);
    for (int i = lane; i < N; i += 32) {
        if (i >= 1) {
            double ax = pr[3 * (i - 1)], ay = pr[3 * (i - 1) + 1];
            double dx = pr[3 * i] - ax, dy = pr[3 * i + 1] - ay;
            s1[i] = make_double2(ax, ay);
            sd[i] = make_double2(dx, dy);
            smem[wb + L.sinv + i] = 1.0 / (fma(dx, dx, dy * dy) + 1e-16);
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------
// psi / grad psi for one problem (the warp's staged problem).  MODE: 0 cost only,
// 1 cost + gradient, 2 obstacle mapping F2 only (returns |F2|^2 in `pen`).
struct Pen {
    double c, hc, inv_c;
};
__device__ __forceinline__ Pen make_pen(double c) {
    Pen p;
    p.c = c;
    p.hc = 0.5 * c;
    p.inv_c = 1.0 / fmax(c, 1.0);
    return p;
}

template <int P, int MODE>
__device__ double eval_psi(const nmpc_config& cfg, const Lay& L, const int wb, const int lane, const double2 (&uv)[P],
                           const Pen pn, const double2 (&yl)[P], double2 (&gout)[P], double& pen_out,
                           double* __restrict__ F2g) {
    constexpr bool GRAD = (MODE == 1);
    const int N = cfg.N_hor;
    const double ts = cfg.ts;
    const double* hdr = smem + wb + L.hdr;
    const double2* S1 = reinterpret_cast<const double2*>(smem + wb + L.s1);
    const double2* SD = reinterpret_cast<const double2*>(smem + wb + L.sd);
    const double* SINV = smem + wb + L.sinv;
    bool act[P];
    double tw[P], inclT[P], exclT[P];
#pragma unroll
    for (int j = 0; j < P; j++) {
        act[j] = (lane + 32 * j) < N;
        tw[j] = act[j] ? ts * uv[j].y : 0.0;
    }
    prefix_scan<P>(tw, inclT, exclT, lane);
    double sn[P], cs[P], thpre[P], TH[P], a[P], b[P];
    const double th0 = hdr[H_TH0];
#pragma unroll
    for (int j = 0; j < P; j++) {
        thpre[j] = th0 + exclT[j];
        TH[j] = th0 + inclT[j];
        nm_sincos(thpre[j], sn[j], cs[j]);
        a[j] = act[j] ? ts * (uv[j].x * cs[j]) : 0.0;
        b[j] = act[j] ? ts * (uv[j].x * sn[j]) : 0.0;
    }
    double X[P], Y[P], xpre[P], ypre[P];
    {
        double ia[P], ea[P], ib[P], eb[P];
        prefix_scan2<P>(a, b, ia, ea, ib, eb, lane);
        const double x0 = hdr[H_X0], y0 = hdr[H_Y0];
#pragma unroll
        for (int j = 0; j < P; j++) {
            xpre[j] = x0 + ea[j];
            ypre[j] = y0 + eb[j];
            X[j] = x0 + ia[j];
            Y[j] = y0 + ib[j];
        }
    }

    double gX[P], gY[P], mind2[P];
#pragma unroll
    for (int j = 0; j < P; j++) gX[j] = gY[j] = mind2[j] = 0.0;

    if (MODE != 2) {
        // cross-track error: each lane scans the N-1 segments for its own predicted point
        double best[P], bex[P], bey[P], bth[P];
        int bi[P];
#pragma unroll
        for (int j = 0; j < P; j++) {
            best[j] = CUDART_INF;
            bex[j] = bey[j] = bth[j] = 0.0;
            bi[j] = 1;
        }
        for (int i = 1; i < N; i++) {
            const double2 s1 = S1[i], d = SD[i];
            const double inv = SINV[i];
#pragma unroll
            for (int j = 0; j < P; j++) {
                double px = X[j] - s1.x, py = Y[j] - s1.y;
                double that = fma(px, d.x, py * d.y) * inv;
                double tst = fmin(fmax(that, 0.0), 1.0);
                double ex = fma(tst, d.x, -px), ey = fma(tst, d.y, -py);
                double d2 = fma(ex, ex, ey * ey);
                if (d2 < best[j]) {
                    best[j] = d2;
                    bi[j] = i;
                    bex[j] = ex;
                    bey[j] = ey;
                    bth[j] = that;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < P; j++) {
            mind2[j] = best[j];
            if (GRAD) {
                const double2 d = SD[bi[j]];
                double ed = (bth[j] >= 0.0 && bth[j] <= 1.0) ? fma(bex[j], d.x, bey[j] * d.y) * SINV[bi[j]] : 0.0;
                double k2 = 2.0 * hdr[H_QCTE];
                gX[j] = k2 * fma(ed, d.x, -bex[j]);
                gY[j] = k2 * fma(ed, d.y, -bey[j]);
            }
        }
    }

    // obstacle penalty F2: circles, then this lane's time slice of each ellipse
    double pen = 0.0;
    {
        const double* CIRC = smem + wb + L.circ;
        for (int k = 0; k < cfg.Nobs; k++) {
            const double2 cxy = *reinterpret_cast<const double2*>(CIRC + 4 * k);
            const double r2 = CIRC[4 * k + 2];
            double h[P], dx[P], dy[P];
            unsigned any = 0;
            unsigned m[P];
#pragma unroll
            for (int j = 0; j < P; j++) {
                dx[j] = X[j] - cxy.x;
                dy[j] = Y[j] - cxy.y;
                h[j] = fma(-dy[j], dy[j], fma(-dx[j], dx[j], r2));
                m[j] = __ballot_sync(FULL, act[j] && h[j] > 0.0);
                any |= m[j];
            }
            double g = 0.0;
            if (any) {
#pragma unroll
                for (int j = 0; j < P; j++) {
                    unsigned mm = m[j];
                    while (mm) {
                        int src = __ffs(mm) - 1;
                        g = g + __shfl_sync(FULL, h[j], src);
                        mm &= mm - 1;
                    }
                }
            }
            if (F2g && lane == 0) F2g[k] = g;
            pen = fma(g, g, pen);
            if (GRAD && g > 0.0) {
                const double cg = pn.c * g;
#pragma unroll
                for (int j = 0; j < P; j++)
                    if (act[j] && h[j] > 0.0) {
                        gX[j] = fma(cg, -2.0 * dx[j], gX[j]);
                        gY[j] = fma(cg, -2.0 * dy[j], gY[j]);
                    }
            }
        }
        const int ne = cfg.Ndynobs * N;
        const double* EL = smem + wb + L.ell;
        for (int k = 0; k < cfg.Ndynobs; k++) {
            double h[P], ta[P], tb[P], eca[P], esa[P];
            unsigned any = 0;
            unsigned m[P];
#pragma unroll
            for (int j = 0; j < P; j++) {
                const int t = lane + 32 * j;
                const int i = act[j] ? k * N + t : k * N;
                double dx = X[j] - EL[i], dy = Y[j] - EL[ne + i];
                eca[j] = EL[2 * ne + i];
                esa[j] = EL[3 * ne + i];
                double irx2 = EL[4 * ne + i], iry2 = EL[5 * ne + i];
                double ea = fma(dx, eca[j], dy * esa[j]);
                double eb = fma(dx, esa[j], -(dy * eca[j]));
                h[j] = fma(-(eb * eb), iry2, fma(-(ea * ea), irx2, 1.0));
                ta[j] = ea * irx2;
                tb[j] = eb * iry2;
                m[j] = __ballot_sync(FULL, act[j] && h[j] > 0.0);
                any |= m[j];
            }
            double g = 0.0;
            if (any) {
#pragma unroll
                for (int j = 0; j < P; j++) {
                    unsigned mm = m[j];
                    while (mm) {
                        int src = __ffs(mm) - 1;
                        g = g + __shfl_sync(FULL, h[j], src);
                        mm &= mm - 1;
                    }
                }
            }
            if (F2g && lane == 0) F2g[cfg.Nobs + k] = g;
            pen = fma(g, g, pen);
            if (GRAD && g > 0.0) {
                const double cg = pn.c * g;
#pragma unroll
                for (int j = 0; j < P; j++)
                    if (act[j] && h[j] > 0.0) {
                        double hX = -2.0 * fma(ta[j], eca[j], tb[j] * esa[j]);
                        double hY = -2.0 * fma(ta[j], esa[j], -(tb[j] * eca[j]));
                        gX[j] = fma(cg, hX, gX[j]);
                        gY[j] = fma(cg, hY, gY[j]);
                    }
            }
        }
    }
    pen_out = pen;
    if (MODE == 2) return 0.0;

    // stage cost, acceleration cost, ALM term
    const double inv_ts = hdr[H_INVTS];
    const double xref = hdr[H_XREF], yref = hdr[H_YREF], thref = hdr[H_THREF];
    double cl[P], Aa[P], Aw[P];
#pragma unroll
    for (int j = 0; j < P; j++) {
        const int t = lane + 32 * j;
        const double v = uv[j].x, w = uv[j].y;
        double vp = __shfl_up_sync(FULL, v, 1), wp_ = __shfl_up_sync(FULL, w, 1);
        if (j > 0) {
            double v31 = __shfl_sync(FULL, uv[j > 0 ? j - 1 : 0].x, 31), w31 = __shfl_sync(FULL, uv[j > 0 ? j - 1 : 0].y, 31);
            if (lane == 0) {
                vp = v31;
                wp_ = w31;
            }
        } else if (lane == 0) {
            vp = hdr[H_VINIT];
            wp_ = hdr[H_WINIT];
        }
        double c0 = hdr[H_RV] * (v * v);
        c0 = fma(hdr[H_RW], w * w, c0);
        const double vref = act[j] ? smem[wb + L.vref + t] : 0.0;
        double dv = v - vref;
        c0 = fma(hdr[H_QV], dv * dv, c0);
        double ex = xpre[j] - xref, ey = ypre[j] - yref, et = thpre[j] - thref;
        c0 = fma(hdr[H_Q], fma(ex, ex, ey * ey), c0);
        c0 = fma(hdr[H_QTH], et * et, c0);
        c0 = fma(hdr[H_QCTE], mind2[j], c0);
        double acc = (v - vp) * inv_ts, aac = (w - wp_) * inv_ts;
        c0 = fma(hdr[H_AP], acc * acc, c0);
        c0 = fma(hdr[H_WP], aac * aac, c0);
        double za = fma(yl[j].x, pn.inv_c, acc), zw = fma(yl[j].y, pn.inv_c, aac);
        double da = fmax(za - cfg.lin_acc_max, 0.0) + fmin(za - cfg.lin_acc_min, 0.0);
        double dw = fmax(zw - cfg.ang_acc_max, 0.0) + fmin(zw + cfg.ang_acc_max, 0.0);
        c0 = fma(pn.hc, fma(da, da, dw * dw), c0);
        cl[j] = act[j] ? c0 : 0.0;
        Aa[j] = act[j] ? fma(pn.c, da, (2.0 * hdr[H_AP]) * acc) * inv_ts : 0.0;
        Aw[j] = act[j] ? fma(pn.c, dw, (2.0 * hdr[H_WP]) * aac) * inv_ts : 0.0;
    }
    // terminal cost at t = N-1
    const int lN = (N - 1) & 31, jN = (N - 1) >> 5;
    double XN = 0.0, YN = 0.0, TN = 0.0;
#pragma unroll
    for (int j = 0; j < P; j++)
        if (j == jN) {
            XN = __shfl_sync(FULL, X[j], lN);
            YN = __shfl_sync(FULL, Y[j], lN);
            TN = __shfl_sync(FULL, TH[j], lN);
        }
    const double eXN = XN - xref, eYN = YN - yref, eTN = TN - thref;
    const double term = fma(hdr[H_QN], fma(eXN, eXN, eYN * eYN), hdr[H_QTHN] * (eTN * eTN));
    const double psi = fma(pn.hc, pen, hsum<P>(cl) + term);
    if (!GRAD) return psi;

    // backward sweep
    double mth[P];
#pragma unroll
    for (int j = 0; j < P; j++) {
        const int t = lane + 32 * j;
        const bool last = !(t + 1 < N);
        const double qq = last ? hdr[H_QN] : hdr[H_Q], qt = last ? hdr[H_QTHN] : hdr[H_QTH];
        gX[j] = act[j] ? fma(2.0 * qq, X[j] - xref, gX[j]) : 0.0;
        gY[j] = act[j] ? fma(2.0 * qq, Y[j] - yref, gY[j]) : 0.0;
        mth[j] = act[j] ? (2.0 * qt) * (TH[j] - thref) : 0.0;
    }
    double LX[P], LY[P];
    suffix_scan2<P>(gX, gY, LX, LY, lane);
    double nn[P], rr[P], TT[P];
#pragma unroll
    for (int j = 0; j < P; j++) nn[j] = act[j] ? (ts * uv[j].x) * fma(cs[j], LY[j], -(sn[j] * LX[j])) : 0.0;
#pragma unroll
    for (int j = 0; j < P; j++) {
        double nx = __shfl_down_sync(FULL, nn[j], 1);
        double n0 = __shfl_sync(FULL, nn[(j + 1 < P) ? j + 1 : j], 0);
        if (lane == 31) nx = (j + 1 < P) ? n0 : 0.0;
        rr[j] = act[j] ? mth[j] + nx : 0.0;
    }
    suffix_scan<P>(rr, TT, lane);
#pragma unroll
    for (int j = 0; j < P; j++) {
        const int t = lane + 32 * j;
        const double v = uv[j].x, w = uv[j].y;
        double An = __shfl_down_sync(FULL, Aa[j], 1), Wn = __shfl_down_sync(FULL, Aw[j], 1);
        double A0 = __shfl_sync(FULL, Aa[(j + 1 < P) ? j + 1 : j], 0), W0 = __shfl_sync(FULL, Aw[(j + 1 < P) ? j + 1 : j], 0);
        if (lane == 31) {
            An = (j + 1 < P) ? A0 : 0.0;
            Wn = (j + 1 < P) ? W0 : 0.0;
        }
        const double vref = act[j] ? smem[wb + L.vref + t] : 0.0;
        double lv = fma(2.0 * hdr[H_RV], v, (2.0 * hdr[H_QV]) * (v - vref)) + (Aa[j] - An);
        double lw = (2.0 * hdr[H_RW]) * w + (Aw[j] - Wn);
        double gv = fma(ts, fma(cs[j], LX[j], sn[j] * LY[j]), lv);
        double gw = fma(ts, TT[j], lw);
        gout[j] = act[j] ? make_double2(gv, gw) : make_double2(0.0, 0.0);
    }
    return psi;
}

// ---------------------------------------------------------------------------------
// the solver: ALM/PM outer loop around PANOC (control flow = oracle/nmpc_oracle.c)
template <int P>
struct Solver {
    const nmpc_config& cfg;
    const Lay& L;
    const int wb, lane;
    bool act[P];
    int tix[P];
    // warp-uniform PANOC state
    double gamma, inv_gamma, sigma, lip, cost, norm_fpr, tau, akkt_tol;
    Pen pn;
    int iteration, n_cost, n_grad;
    // L-BFGS ring
    int lb_active, lb_first, lb_head;
    double lb_gamma;
    double2 yl[P];

    __device__ Solver(const nmpc_config& c, const Lay& l, int wb_, int lane_) : cfg(c), L(l), wb(wb_), lane(lane_) {
#pragma unroll
        for (int j = 0; j < P; j++) {
            tix[j] = lane + 32 * j;
            act[j] = tix[j] < cfg.N_hor;
        }
    }
    __device__ __forceinline__ double2* vec(int k) const { return reinterpret_cast<double2*>(smem + wb + k * L.n2); }
    __device__ __forceinline__ void ld(int k, double2 (&r)[P]) const {
        const double2* v = vec(k);
#pragma unroll
        for (int j = 0; j < P; j++) r[j] = act[j] ? v[tix[j]] : make_double2(0.0, 0.0);
    }
    __device__ __forceinline__ void st(int k, const double2 (&r)[P]) const {
        double2* v = vec(k);
#pragma unroll
        for (int j = 0; j < P; j++)
            if (act[j]) v[tix[j]] = r[j];
    }
    __device__ __forceinline__ int slot(int i) const { return (lb_head + i) % (cfg.lbfgs_memory + 1); }

    template <int MODE>
    __device__ __forceinline__ double eval(const double2 (&u)[P], double2 (&g)[P], double& pen) {
        if (MODE == 1) n_grad++;
        if (MODE == 0) n_cost++;
        return eval_psi<P, MODE>(cfg, L, wb, lane, u, pn, yl, g, pen, nullptr);
    }
    __device__ __forceinline__ void set_gamma(double g) {
        gamma = g;
        inv_gamma = 1.0 / g;
    }
    // gradient_step() + half_step(): gstep = u - gamma*grad ; uhalf = Proj_U(gstep)
    __device__ __forceinline__ void grad_step_half(const double2 (&u)[P], const double2 (&g)[P], double2 (&gs)[P],
                                                   double2 (&uh)[P]) {
#pragma unroll
        for (int j = 0; j < P; j++) {
            gs[j].x = fma(-gamma, g[j].x, u[j].x);
            gs[j].y = fma(-gamma, g[j].y, u[j].y);
            uh[j].x = act[j] ? clampd(gs[j].x, cfg.lin_vel_min, cfg.lin_vel_max) : 0.0;
            uh[j].y = act[j] ? clampd(gs[j].y, -cfg.ang_vel_max, cfg.ang_vel_max) : 0.0;
        }
        st(V_GSTEP, gs);
        st(V_UHALF, uh);
    }

    __device__ void lb_reset() {
        lb_active = 0;
        lb_first = 1;
    }
    __device__ void lb_update(const double2 (&g)[P], const double2 (&state)[P]) {
        if (lb_first) {
            lb_first = 0;
            st(V_OLDS, state);
            st(V_OLDG, g);
            return;
        }
        double2 os[P], og[P], s[P], y[P];
        ld(V_OLDS, os);
        ld(V_OLDG, og);
#pragma unroll
        for (int j = 0; j < P; j++) {
            s[j] = make_double2(state[j].x - os[j].x, state[j].y - os[j].y);
            y[j] = make_double2(g[j].x - og[j].x, g[j].y - og[j].y);
        }
        const int tmp = slot(cfg.lbfgs_memory);
        st(V_S + tmp, s);
        st(V_Y + tmp, y);
        const double ys = wdot<P>(s, y), ss = wdot<P>(s, s);
        double* rho = smem + wb + L.rho;
        const double rho_new = 1.0 / ys;
        if (ss <= DBL_EPS || ys <= SY_EPSILON) return;
        const double lhs = ys / ss, rhs = CBFGS_EPSILON * sqrt(wdot<P>(g, g));
        if (!(lhs > rhs && isfinite(lhs) && isfinite(rhs))) return;
        st(V_OLDS, state);
        st(V_OLDG, g);
        if (lane == 0) rho[tmp] = rho_new;
        lb_head = (lb_head + cfg.lbfgs_memory) % (cfg.lbfgs_memory + 1);
        lb_gamma = (1.0 / rho_new) / wdot<P>(y, y);
        lb_active = (lb_active + 1 < cfg.lbfgs_memory) ? lb_active + 1 : cfg.lbfgs_memory;
        __syncwarp();
    }
    __device__ void lb_apply(double2 (&q)[P]) {
        if (lb_active == 0) return;
        const double* rho = smem + wb + L.rho;
        double* alpha = smem + wb + L.alpha;
        for (int k = 0; k < lb_active; k++) {
            const int sl = slot(k);
            double2 s[P], y[P];
            ld(V_S + sl, s);
            ld(V_Y + sl, y);
            const double al = rho[sl] * wdot<P>(s, q);
            if (lane == 0) alpha[k] = al;
#pragma unroll
            for (int j = 0; j < P; j++) {
                q[j].x = fma(-al, y[j].x, q[j].x);
                q[j].y = fma(-al, y[j].y, q[j].y);
            }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < P; j++) {
            q[j].x = q[j].x * lb_gamma;
            q[j].y = q[j].y * lb_gamma;
        }
        for (int k = lb_active - 1; k >= 0; k--) {
            const int sl = slot(k);
            double2 s[P], y[P];
            ld(V_S + sl, s);
            ld(V_Y + sl, y);
            const double beta = rho[sl] * wdot<P>(y, q);
            const double co = alpha[k] - beta;
#pragma unroll
            for (int j = 0; j < P; j++) {
                q[j].x = fma(co, s[j].x, q[j].x);
                q[j].y = fma(co, s[j].y, q[j].y);
            }
        }
    }

    // fpr = u - uhalf, norm
    __device__ __forceinline__ void compute_fpr(const double2 (&u)[P], const double2 (&uh)[P], double2 (&fpr)[P]) {
        double e[P];
#pragma unroll
        for (int j = 0; j < P; j++) {
            double d0 = u[j].x - uh[j].x, d1 = u[j].y - uh[j].y;
            fpr[j] = make_double2(d0, d1);
            e[j] = fma(d1, d1, d0 * d0);
        }
        norm_fpr = sqrt(hsum<P>(e));
    }

    __device__ void panoc_init(double2 (&u)[P]) {
        lb_reset();
        tau = 1.0;
        iteration = 0;
        double2 g[P], gh[P], hv[P], gs[P], uh[P];
        double pen;
        cost = eval<1>(u, g, pen);
        double e[P];
#pragma unroll
        for (int j = 0; j < P; j++) {
            hv[j].x = act[j] ? fmax(DELTA_LIPSCHITZ, EPSILON_LIPSCHITZ * u[j].x) : 0.0;
            hv[j].y = act[j] ? fmax(DELTA_LIPSCHITZ, EPSILON_LIPSCHITZ * u[j].y) : 0.0;
            e[j] = fma(hv[j].y, hv[j].y, hv[j].x * hv[j].x);
        }
        const double norm_h = sqrt(hsum<P>(e));
#pragma unroll
        for (int j = 0; j < P; j++) {
            u[j].x = u[j].x + hv[j].x;
            u[j].y = u[j].y + hv[j].y;
        }
        eval<1>(u, gh, pen);
        lip = sqrt(wdiff2<P>(gh, g)) / norm_h;
        set_gamma(GAMMA_L_COEFF / fmax(lip, MIN_L_ESTIMATE));
        sigma = (1.0 - GAMMA_L_COEFF) / (4.0 * gamma);
        st(V_GRAD, g);
        grad_step_half(u, g, gs, uh);
    }

    // one PANOC iteration; returns false when the exit condition holds
    __device__ bool panoc_step(double2 (&u)[P]) {
        double2 g[P], uh[P], fpr[P];
        double pen;
        ld(V_GRAD, g);
        ld(V_UHALF, uh);
        compute_fpr(u, uh, fpr);
        if (norm_fpr < cfg.tolerance) {
            double e[P];
#pragma unroll
            for (int j = 0; j < P; j++) {
                double p0 = iteration ? g[j].x : 0.0, p1 = iteration ? g[j].y : 0.0;
                double r0 = fma(fpr[j].x, inv_gamma, g[j].x) - p0;
                double r1 = fma(fpr[j].y, inv_gamma, g[j].y) - p1;
                e[j] = fma(r1, r1, r0 * r0);
            }
            if (sqrt(hsum<P>(e)) < akkt_tol) return false;
        }
        // update_lipschitz_constant()
        double2 dummy[P];
        double cost_half = eval<0>(uh, dummy, pen);
        if (iteration == 0) cost = eval<0>(u, dummy, pen);  // u was perturbed by the Lipschitz estimate
        else n_cost++;  // OpEn re-evaluates psi(u); the value is bit-identical to the cached one
        int it = 0;
        for (;;) {
            const double ip = wdot<P>(g, fpr);
            const double rhs = cost + LIPSCHITZ_UPDATE_EPSILON * fabs(cost) - ip +
                               (GAMMA_L_COEFF * 0.5 * inv_gamma) * (norm_fpr * norm_fpr);
            if (!(cost_half > rhs && it < MAX_LIPSCHITZ_UPDATE_ITERATIONS && lip < MAX_LIPSCHITZ_CONSTANT)) break;
            lb_reset();
            lip *= 2.0;
            set_gamma(gamma / 2.0);
            double2 gs[P];
            grad_step_half(u, g, gs, uh);
            cost_half = eval<0>(uh, dummy, pen);
            compute_fpr(u, uh, fpr);
            it++;
        }
        sigma = (1.0 - GAMMA_L_COEFF) / (4.0 * gamma);
        // lbfgs_direction()
        lb_update(fpr, u);
        if (iteration == 0) {  // update_no_linesearch()
#pragma unroll
            for (int j = 0; j < P; j++) u[j] = uh[j];
            cost = eval<1>(u, g, pen);
            st(V_GRAD, g);
            double2 gs[P];
            grad_step_half(u, g, gs, uh);
        } else {  // linesearch() on the forward-backward envelope
            double2 dir[P], gs[P];
#pragma unroll
            for (int j = 0; j < P; j++) dir[j] = fpr[j];
            lb_apply(dir);
            ld(V_GSTEP, gs);
            const double dist2 = wdiff2<P>(gs, uh);
            const double fbe = cost - (0.5 * gamma) * wdot<P>(g, g) + (0.5 * dist2) * inv_gamma;
            const double rhs_ls = fbe - sigma * (norm_fpr * norm_fpr);
            tau = 1.0;
            int nls = 0;
            double2 up[P];
            for (;;) {
                const double om = 1.0 - tau;
#pragma unroll
                for (int j = 0; j < P; j++) {
                    up[j].x = fma(-tau, dir[j].x, fma(-om, fpr[j].x, u[j].x));
                    up[j].y = fma(-tau, dir[j].y, fma(-om, fpr[j].y, u[j].y));
                }
                cost = eval<1>(up, g, pen);
                grad_step_half(up, g, gs, uh);
                const double d2 = wdiff2<P>(gs, uh);
                const double lhs = cost - (0.5 * gamma) * wdot<P>(g, g) + (0.5 * d2) * inv_gamma;
                if (!(lhs > rhs_ls && nls < MAX_LINESEARCH_ITERATIONS)) break;
                tau /= 2.0;
                nls++;
            }
            st(V_GRAD, g);
#pragma unroll
            for (int j = 0; j < P; j++) u[j] = up[j];
        }
        iteration++;
        return true;
    }

    __device__ int panoc_solve(double2 (&u)[P], int& iters) {
        panoc_init(u);
        int num_iter = 0;
        bool cont = true;
        bool flag = panoc_step(u);
        while (flag && cont) {
            num_iter++;
            cont = num_iter < cfg.max_inner_iterations;
            flag = panoc_step(u);
        }
        iters = num_iter;
        bool fin = true;
#pragma unroll
        for (int j = 0; j < P; j++) fin = fin && isfinite(u[j].x) && isfinite(u[j].y);
        if (!__all_sync(FULL, fin)) return NMPC_NOT_FINITE;
        ld(V_UHALF, u);
        return cont ? NMPC_CONVERGED : NMPC_NOT_CONVERGED_ITERATIONS;
    }

    // ALM / penalty outer loop; u in/out (lane-distributed), y in/out in yl
    __device__ int solve(double2 (&u)[P], nmpc_stats& st_out) {
        const int N = cfg.N_hor;
        const int nf2 = cfg.Nobs + cfg.Ndynobs;
        pn = make_pen(cfg.initial_penalty);
        akkt_tol = cfg.initial_tolerance;
        lb_head = 0;
        n_cost = n_grad = 0;
        norm_fpr = 0.0;
        int alm_iter = 0, inner_total = 0, num_outer = 0, status = NMPC_CONVERGED, done = 0;
        double f2n = 0.0, f2np = 0.0, dyn = 0.0, dynp = 0.0;
        const double inv_ts = smem[wb + L.hdr + H_INVTS];
        for (int outer = 0; outer < cfg.max_outer_iterations; outer++) {
            num_outer++;
#pragma unroll
            for (int j = 0; j < P; j++) {
                yl[j].x = clampd(yl[j].x, -Y_SET_BOUND, Y_SET_BOUND);
                yl[j].y = clampd(yl[j].y, -Y_SET_BOUND, Y_SET_BOUND);
            }
            int iters = 0;
            const int inner = panoc_solve(u, iters);
            inner_total += iters;
            if (inner == NMPC_NOT_FINITE) {
                status = NMPC_NOT_FINITE;
                done = 2;
                break;
            }
            status = inner;
            // multipliers: y+ = y + c*(F1 - Proj_C(F1 + y/c)); infeasibilities
            double2 dummy[P], yp[P];
            double pen;
            eval_psi<P, 2>(cfg, L, wb, lane, u, pn, yl, dummy, pen, nullptr);
            double e[P];
#pragma unroll
            for (int j = 0; j < P; j++) {
                const double v = u[j].x, w = u[j].y;
                double vp = __shfl_up_sync(FULL, v, 1), wp_ = __shfl_up_sync(FULL, w, 1);
                if (j > 0) {
                    double v31 = __shfl_sync(FULL, u[j > 0 ? j - 1 : 0].x, 31), w31 = __shfl_sync(FULL, u[j > 0 ? j - 1 : 0].y, 31);
                    if (lane == 0) {
                        vp = v31;
                        wp_ = w31;
                    }
                } else if (lane == 0) {
                    vp = smem[wb + L.hdr + H_VINIT];
                    wp_ = smem[wb + L.hdr + H_WINIT];
                }
                const double wa = (v - vp) * inv_ts, ww = (w - wp_) * inv_ts;
                double za = wa + yl[j].x / pn.c, zw = ww + yl[j].y / pn.c;
                za = clampd(za, cfg.lin_acc_min, cfg.lin_acc_max);
                zw = clampd(zw, -cfg.ang_acc_max, cfg.ang_acc_max);
                yp[j].x = act[j] ? fma(pn.c, wa - za, yl[j].x) : 0.0;
                yp[j].y = act[j] ? fma(pn.c, ww - zw, yl[j].y) : 0.0;
                double d0 = yp[j].x - yl[j].x, d1 = yp[j].y - yl[j].y;
                e[j] = act[j] ? fma(d1, d1, d0 * d0) : 0.0;
            }
            dynp = sqrt(hsum<P>(e));
            f2np = sqrt(pen);
            const bool crit1 = alm_iter > 0 && dynp <= pn.c * cfg.delta_tolerance + DBL_EPS;
            const bool crit2 = (nf2 == 0) || f2np <= cfg.delta_tolerance + DBL_EPS;
            const bool crit3 = akkt_tol <= cfg.tolerance + DBL_EPS;
            if (crit1 && crit2 && crit3) {
                done = 1;
                break;
            }
            bool stall;
            if (alm_iter == 0) stall = true;
            else {
                const bool ca = dynp <= cfg.sufficient_decrease_coeff * dyn + DBL_EPS;
                const bool cp = f2np <= cfg.sufficient_decrease_coeff * f2n + DBL_EPS;
                stall = (nf2 > 0) ? (ca && cp) : ca;
            }
            if (!stall) pn = make_pen(pn.c * cfg.penalty_update_factor);
            akkt_tol = fmax(akkt_tol * cfg.inner_tolerance_update, cfg.tolerance);
            alm_iter++;
            dyn = dynp;
            f2n = f2np;
#pragma unroll
            for (int j = 0; j < P; j++) yl[j] = yp[j];
        }
        (void)N;
        if (done != 2 && num_outer == cfg.max_outer_iterations) status = NMPC_NOT_CONVERGED_ITERATIONS;
        st_out.exit_status = status;
        st_out.outer_iterations = num_outer;
        st_out.inner_iterations = inner_total;
        st_out.last_norm_fpr = norm_fpr;
        st_out.delta_y_norm_over_c = dynp / pn.c;
        st_out.f2_norm = f2np;
        st_out.penalty = pn.c;
        if (status == NMPC_NOT_FINITE) st_out.cost = CUDART_NAN;
        else {
            double2 dummy[P];
            double pen;
            const Pen keep = pn;
            pn = make_pen(0.0);
            st_out.cost = eval_psi<P, 0>(cfg, L, wb, lane, u, pn, yl, dummy, pen, nullptr);
            pn = keep;
        }
        st_out.n_cost_evals = n_cost;
        st_out.n_grad_evals = n_grad;
        st_out.reserved = 0;
        return status;
    }
};

// ---------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(512, 1) nmpc_solve_kernel(const __grid_constant__ KArgs a) {
    const nmpc_config& cfg = a.cfg;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Lay L = make_layout(cfg.N_hor, cfg.Nobs, cfg.Ndynobs);
    const int wb = warp * L.total;
    const int N = cfg.N_hor;
    for (;;) {
        int b = 0;
        if (lane == 0) b = (int)atomicAdd(a.counter, 1u);
        b = __shfl_sync(FULL, b, 0);
        if (b >= a.B) break;
        __syncwarp();
        stage_problem(cfg, L, wb, lane, a.P + (size_t)b * a.np);
        Solver<P> S(cfg, L, wb, lane);
        double2 u[P];
        const double* U0 = a.U + (size_t)b * 2 * N;
        const double* Y0 = a.Y ? a.Y + (size_t)b * 2 * N : nullptr;
#pragma unroll
        for (int j = 0; j < P; j++) {
            const int t = lane + 32 * j;
            u[j] = (t < N) ? *reinterpret_cast<const double2*>(U0 + 2 * t) : make_double2(0.0, 0.0);
            S.yl[j] = (t < N && Y0) ? make_double2(Y0[t], Y0[N + t]) : make_double2(0.0, 0.0);
        }
        nmpc_stats st;
        const int status = S.solve(u, st);
#pragma unroll
        for (int j = 0; j < P; j++) {
            const int t = lane + 32 * j;
            if (t < N) {
                *reinterpret_cast<double2*>(a.U + (size_t)b * 2 * N + 2 * t) = u[j];
                if (a.Y) {
                    a.Y[(size_t)b * 2 * N + t] = S.yl[j].x;
                    a.Y[(size_t)b * 2 * N + N + t] = S.yl[j].y;
                }
            }
        }
        if (lane == 0) {
            if (a.status) a.status[b] = status;
            if (a.stats) a.stats[b] = st;
        }
    }
}

// parity hook: psi, grad, F1, F2 for B (p, u, c, y) tuples
template <int P>
__global__ void __launch_bounds__(512, 1) nmpc_eval_kernel(const __grid_constant__ KArgs a) {
    const nmpc_config& cfg = a.cfg;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Lay L = make_layout(cfg.N_hor, cfg.Nobs, cfg.Ndynobs);
    const int wb = warp * L.total;
    const int N = cfg.N_hor, nf2 = cfg.Nobs + cfg.Ndynobs;
    const int wpb = blockDim.x >> 5;
    for (int b = blockIdx.x * wpb + warp; b < a.B; b += gridDim.x * wpb) {
        __syncwarp();
        stage_problem(cfg, L, wb, lane, a.P + (size_t)b * a.np);
        double2 u[P], yl[P], g[P];
        const double* U0 = a.U + (size_t)b * 2 * N;
        const double* Y0 = a.Y ? a.Y + (size_t)b * 2 * N : nullptr;
#pragma unroll
        for (int j = 0; j < P; j++) {
            const int t = lane + 32 * j;
            u[j] = (t < N) ? make_double2(U0[2 * t], U0[2 * t + 1]) : make_double2(0.0, 0.0);
            yl[j] = (t < N && Y0) ? make_double2(Y0[t], Y0[N + t]) : make_double2(0.0, 0.0);
        }
        const Pen pn = make_pen(a.cvec[b]);
        double pen;
        const double psi = eval_psi<P, 1>(cfg, L, wb, lane, u, pn, yl, g, pen, a.F2 ? a.F2 + (size_t)b * nf2 : nullptr);
        if (lane == 0 && a.psi) a.psi[b] = psi;
        const double inv_ts = smem[wb + L.hdr + H_INVTS];
#pragma unroll
        for (int j = 0; j < P; j++) {
            const int t = lane + 32 * j;
            const double v = u[j].x, w = u[j].y;
            double vp = __shfl_up_sync(FULL, v, 1), wp_ = __shfl_up_sync(FULL, w, 1);
            if (j > 0) {
                double v31 = __shfl_sync(FULL, u[j > 0 ? j - 1 : 0].x, 31), w31 = __shfl_sync(FULL, u[j > 0 ? j - 1 : 0].y, 31);
                if (lane == 0) {
                    vp = v31;
                    wp_ = w31;
                }
            } else if (lane == 0) {
                vp = smem[wb + L.hdr + H_VINIT];
                wp_ = smem[wb + L.hdr + H_WINIT];
            }
            if (t < N) {
                if (a.grad) {
                    a.grad[(size_t)b * 2 * N + 2 * t] = g[j].x;
                    a.grad[(size_t)b * 2 * N + 2 * t + 1] = g[j].y;
                }
                if (a.F1) {
                    a.F1[(size_t)b * 2 * N + t] = (v - vp) * inv_ts;
                    a.F1[(size_t)b * 2 * N + N + t] = (w - wp_) * inv_ts;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// host side: the C ABI (include/nmpc_b200.h)
struct nmpc_handle {
    nmpc_config cfg;
    int device, sm_count, np, P, warps_per_cta;
    size_t smem_bytes;
    cudaStream_t stream;
    unsigned int* counter;
    // scratch for the host-buffer entry points
    double *dP, *dU, *dY;
    int32_t* dstatus;
    nmpc_stats* dstats;
    int cap;
    // nmpc_call state (what OpEn's TCP server keeps between requests)
    double *call_u, *call_y;
    int64_t launches;
    double last_ms;
    cudaEvent_t ev0, ev1;
    char err[512];
};

static int set_err(nmpc_handle* h, int code, const char* fmt, const char* detail) {
    if (h) snprintf(h->err, sizeof(h->err), fmt, detail ? detail : "");
    return code;
}
#define CUDA_TRY(h, call)                                                                  \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) return set_err(h, NMPC_ERR_CUDA, #call ": %s", cudaGetErrorString(e_)); \
    } while (0)

extern "C" {

void nmpc_default_config(nmpc_config* cfg) {
    memset(cfg, 0, sizeof(*cfg));
    cfg->N_hor = 20; cfg->Nobs = 10; cfg->Ndynobs = 3;  // configs/default.yaml:7,38,39
    cfg->lbfgs_memory = 10; cfg->max_inner_iterations = 500; cfg->max_outer_iterations = 10;
    cfg->ts = 0.2;                                                            // :18
    cfg->lin_vel_min = -0.5; cfg->lin_vel_max = 1.5; cfg->ang_vel_max = 0.5;  // :8-9,12
    cfg->lin_acc_min = -1.0; cfg->lin_acc_max = 1.0; cfg->ang_acc_max = 3.0;  // :10-11,13
    cfg->tolerance = 1e-4; cfg->initial_tolerance = 1e-4; cfg->delta_tolerance = 1e-4;
    cfg->inner_tolerance_update = 0.1; cfg->penalty_update_factor = 5.0; cfg->initial_penalty = 1.0;
    cfg->sufficient_decrease_coeff = 0.1;
}

int32_t nmpc_param_len(const nmpc_config* cfg) {
    return NMPC_NZ + cfg->N_hor + 3 * cfg->Nobs + 5 * cfg->Ndynobs * cfg->N_hor + 3 * cfg->N_hor;
}

int32_t nmpc_abi_version(void) { return NMPC_ABI_VERSION; }

const char* nmpc_exit_status_name(int32_t s) {
    switch (s) {
        case NMPC_CONVERGED: return "Converged";
        case NMPC_NOT_CONVERGED_ITERATIONS: return "NotConvergedIterations";
        case NMPC_NOT_CONVERGED_OUT_OF_TIME: return "NotConvergedOutOfTime";
        case NMPC_NOT_FINITE: return "NotFiniteComputation";
        default: return "Unknown";
    }
}

const char* nmpc_last_error(nmpc_handle* h) { return h ? h->err : "null handle"; }

static const void* solve_kernel_for(int P) {
    switch (P) {
        case 1: return (const void*)nmpc_solve_kernel<1>;
        case 2: return (const void*)nmpc_solve_kernel<2>;
        default: return (const void*)nmpc_solve_kernel<3>;
    }
}
static const void* eval_kernel_for(int P) {
    switch (P) {
        case 1: return (const void*)nmpc_eval_kernel<1>;
        case 2: return (const void*)nmpc_eval_kernel<2>;
        default: return (const void*)nmpc_eval_kernel<3>;
    }
}

int nmpc_create(const nmpc_config* cfg, int device, nmpc_handle** out) {
    if (!cfg || !out) return NMPC_ERR_INVALID;
    *out = nullptr;
    if (cfg->N_hor < 2 || cfg->N_hor > NMPC_MAX_HORIZON || cfg->Nobs < 0 || cfg->Ndynobs < 0 || cfg->lbfgs_memory < 1 ||
        cfg->lbfgs_memory > NMPC_LBFGS_MAX || !(cfg->ts > 0.0) || cfg->max_inner_iterations < 1 ||
        cfg->max_outer_iterations < 1)
        return NMPC_ERR_INVALID;
    nmpc_handle* h = new (std::nothrow) nmpc_handle();
    if (!h) return NMPC_ERR_NOMEM;
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg;
    h->device = device;
    h->np = nmpc_param_len(cfg);
    h->P = (cfg->N_hor + 31) / 32;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        delete h;
        return NMPC_ERR_CUDA;
    }
    cudaError_t e = cudaSetDevice(device);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        delete h;
        return NMPC_ERR_CUDA;
    }
    h->sm_count = prop.multiProcessorCount;
    const Lay L = make_layout(cfg->N_hor, cfg->Nobs, cfg->Ndynobs);
    const size_t per_warp = (size_t)L.total * sizeof(double);
    const size_t max_smem = prop.sharedMemPerBlockOptin;
    int w = (int)(max_smem / per_warp);
    if (w < 1) {
        delete h;
        return NMPC_ERR_INVALID;  // problem too large for one warp's arena
    }
    if (w > 16) w = 16;
    h->warps_per_cta = w;
    h->smem_bytes = per_warp * w;
    e = cudaFuncSetAttribute(solve_kernel_for(h->P), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(eval_kernel_for(h->P), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&h->counter, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&h->call_u, 2 * cfg->N_hor * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&h->call_y, 2 * cfg->N_hor * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(h->call_u, 0, 2 * cfg->N_hor * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(h->call_y, 0, 2 * cfg->N_hor * sizeof(double));
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
    if (e != cudaSuccess) {
        nmpc_destroy(h);
        return NMPC_ERR_CUDA;
    }
    *out = h;
    return NMPC_OK;
}

int nmpc_destroy(nmpc_handle* h) {
    if (!h) return NMPC_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->counter);
    cudaFree(h->call_u);
    cudaFree(h->call_y);
    cudaFree(h->dP);
    cudaFree(h->dU);
    cudaFree(h->dY);
    cudaFree(h->dstatus);
    cudaFree(h->dstats);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return NMPC_OK;
}

int nmpc_ping(nmpc_handle* h) {
    if (!h) return NMPC_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return NMPC_OK;
}

static int launch_solve(nmpc_handle* h, int32_t B, const double* dP, double* dU, double* dY, int32_t* dstatus,
                        nmpc_stats* dstats, cudaStream_t s) {
    KArgs a;
    memset(&a, 0, sizeof(a));
    a.cfg = h->cfg;
    a.B = B;
    a.np = h->np;
    a.P = dP;
    a.U = dU;
    a.Y = dY;
    a.status = dstatus;
    a.stats = dstats;
    a.counter = h->counter;
    CUDA_TRY(h, cudaMemsetAsync(h->counter, 0, sizeof(unsigned int), s));
    const int warps_needed = B;
    int grid = h->sm_count;
    const int ctas_needed = (warps_needed + h->warps_per_cta - 1) / h->warps_per_cta;
    if (grid > ctas_needed) grid = ctas_needed;
    if (grid < 1) grid = 1;
    void* args[] = {&a};
    CUDA_TRY(h, cudaLaunchKernel(solve_kernel_for(h->P), dim3(grid), dim3(32 * h->warps_per_cta), args, h->smem_bytes, s));
    h->launches++;
    return NMPC_OK;
}

int nmpc_solve_batch_device(nmpc_handle* h, int32_t B, const double* dP, double* dU, double* dY, int32_t* dstatus,
                            nmpc_stats* dstats, void* stream) {
    if (!h || B < 0 || !dP || !dU) return set_err(h, NMPC_ERR_INVALID, "nmpc_solve_batch_device: bad argument%s", "");
    if (B == 0) return NMPC_OK;
    CUDA_TRY(h, cudaSetDevice(h->device));
    return launch_solve(h, B, dP, dU, dY, dstatus, dstats, stream ? (cudaStream_t)stream : h->stream);
}

static int ensure_scratch(nmpc_handle* h, int B) {
    if (B <= h->cap) return NMPC_OK;
    cudaFree(h->dP); cudaFree(h->dU); cudaFree(h->dY); cudaFree(h->dstatus); cudaFree(h->dstats);
    h->dP = h->dU = h->dY = nullptr; h->dstatus = nullptr; h->dstats = nullptr; h->cap = 0;
    const size_t n2 = 2 * (size_t)h->cfg.N_hor;
    CUDA_TRY(h, cudaMalloc(&h->dP, (size_t)B * h->np * sizeof(double)));
    CUDA_TRY(h, cudaMalloc(&h->dU, (size_t)B * n2 * sizeof(double)));
    CUDA_TRY(h, cudaMalloc(&h->dY, (size_t)B * n2 * sizeof(double)));
    CUDA_TRY(h, cudaMalloc(&h->dstatus, (size_t)B * sizeof(int32_t)));
    CUDA_TRY(h, cudaMalloc(&h->dstats, (size_t)B * sizeof(nmpc_stats)));
    h->cap = B;
    return NMPC_OK;
}

int nmpc_solve_batch(nmpc_handle* h, int32_t B, const double* P, double* U, double* Y, int32_t* status,
                     nmpc_stats* stats) {
    if (!h || B < 0 || !P || !U) return set_err(h, NMPC_ERR_INVALID, "nmpc_solve_batch: bad argument%s", "");
    if (B == 0) return NMPC_OK;
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = ensure_scratch(h, B);
    if (rc) return rc;
    const size_t n2 = 2 * (size_t)h->cfg.N_hor;
    cudaStream_t s = h->stream;
    CUDA_TRY(h, cudaMemcpyAsync(h->dP, P, (size_t)B * h->np * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaMemcpyAsync(h->dU, U, (size_t)B * n2 * sizeof(double), cudaMemcpyHostToDevice, s));
    if (Y) CUDA_TRY(h, cudaMemcpyAsync(h->dY, Y, (size_t)B * n2 * sizeof(double), cudaMemcpyHostToDevice, s));
    else CUDA_TRY(h, cudaMemsetAsync(h->dY, 0, (size_t)B * n2 * sizeof(double), s));
    CUDA_TRY(h, cudaEventRecord(h->ev0, s));
    rc = launch_solve(h, B, h->dP, h->dU, h->dY, h->dstatus, h->dstats, s);
    if (rc) return rc;
    CUDA_TRY(h, cudaEventRecord(h->ev1, s));
    CUDA_TRY(h, cudaMemcpyAsync(U, h->dU, (size_t)B * n2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (Y) CUDA_TRY(h, cudaMemcpyAsync(Y, h->dY, (size_t)B * n2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (status) CUDA_TRY(h, cudaMemcpyAsync(status, h->dstatus, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (stats) CUDA_TRY(h, cudaMemcpyAsync(stats, h->dstats, (size_t)B * sizeof(nmpc_stats), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    float ms = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms = ms;
    return NMPC_OK;
}

int nmpc_call(nmpc_handle* h, const double* p, double* u_out, int32_t* exit_status, nmpc_stats* stats_out) {
    if (!h || !p || !u_out) return set_err(h, NMPC_ERR_INVALID, "nmpc_call: bad argument%s", "");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = ensure_scratch(h, 1);
    if (rc) return rc;
    const size_t n2 = 2 * (size_t)h->cfg.N_hor;
    cudaStream_t s = h->stream;
    CUDA_TRY(h, cudaMemcpyAsync(h->dP, p, (size_t)h->np * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaEventRecord(h->ev0, s));
    // the persisted (u, y) are solved in place: warm start from the previous reply, un-shifted
    rc = launch_solve(h, 1, h->dP, h->call_u, h->call_y, h->dstatus, h->dstats, s);
    if (rc) return rc;
    CUDA_TRY(h, cudaEventRecord(h->ev1, s));
    int32_t st = 0;
    nmpc_stats stt;
    CUDA_TRY(h, cudaMemcpyAsync(u_out, h->call_u, n2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaMemcpyAsync(&st, h->dstatus, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaMemcpyAsync(&stt, h->dstats, sizeof(nmpc_stats), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    float ms = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms = ms;
    if (exit_status) *exit_status = st;
    if (stats_out) *stats_out = stt;
    return NMPC_OK;
}

int nmpc_reset_warm_start(nmpc_handle* h) {
    if (!h) return NMPC_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n2 = 2 * (size_t)h->cfg.N_hor;
    CUDA_TRY(h, cudaMemsetAsync(h->call_u, 0, n2 * sizeof(double), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->call_y, 0, n2 * sizeof(double), h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return NMPC_OK;
}

int nmpc_eval_batch(nmpc_handle* h, int32_t B, const double* P, const double* U, const double* c, const double* Y,
                    double* psi, double* grad, double* F1, double* F2) {
    if (!h || B < 0 || !P || !U || !c) return set_err(h, NMPC_ERR_INVALID, "nmpc_eval_batch: bad argument%s", "");
    if (B == 0) return NMPC_OK;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n2 = 2 * (size_t)h->cfg.N_hor, nf2 = (size_t)h->cfg.Nobs + h->cfg.Ndynobs;
    cudaStream_t s = h->stream;
    double *dP = 0, *dU = 0, *dY = 0, *dc = 0, *dpsi = 0, *dgrad = 0, *dF1 = 0, *dF2 = 0;
    int rc = NMPC_OK;
    cudaError_t e = cudaSuccess;
#define TRY_(call) if (e == cudaSuccess) e = (call)
    TRY_(cudaMalloc(&dP, (size_t)B * h->np * sizeof(double)));
    TRY_(cudaMalloc(&dU, (size_t)B * n2 * sizeof(double)));
    TRY_(cudaMalloc(&dY, (size_t)B * n2 * sizeof(double)));
    TRY_(cudaMalloc(&dc, (size_t)B * sizeof(double)));
    TRY_(cudaMalloc(&dpsi, (size_t)B * sizeof(double)));
    TRY_(cudaMalloc(&dgrad, (size_t)B * n2 * sizeof(double)));
    TRY_(cudaMalloc(&dF1, (size_t)B * n2 * sizeof(double)));
    TRY_(cudaMalloc(&dF2, (size_t)(B * nf2 + 1) * sizeof(double)));
    TRY_(cudaMemcpyAsync(dP, P, (size_t)B * h->np * sizeof(double), cudaMemcpyHostToDevice, s));
    TRY_(cudaMemcpyAsync(dU, U, (size_t)B * n2 * sizeof(double), cudaMemcpyHostToDevice, s));
    if (Y) TRY_(cudaMemcpyAsync(dY, Y, (size_t)B * n2 * sizeof(double), cudaMemcpyHostToDevice, s));
    else TRY_(cudaMemsetAsync(dY, 0, (size_t)B * n2 * sizeof(double), s));
    TRY_(cudaMemcpyAsync(dc, c, (size_t)B * sizeof(double), cudaMemcpyHostToDevice, s));
    if (e == cudaSuccess) {
        KArgs a;
        memset(&a, 0, sizeof(a));
        a.cfg = h->cfg; a.B = B; a.np = h->np; a.P = dP; a.U = dU; a.Y = dY; a.cvec = dc;
        a.psi = dpsi; a.grad = dgrad; a.F1 = dF1; a.F2 = dF2;
        int grid = (B + h->warps_per_cta - 1) / h->warps_per_cta;
        if (grid > 8 * h->sm_count) grid = 8 * h->sm_count;
        void* args[] = {&a};
        e = cudaLaunchKernel(eval_kernel_for(h->P), dim3(grid), dim3(32 * h->warps_per_cta), args, h->smem_bytes, s);
        h->launches++;
    }
    if (psi) TRY_(cudaMemcpyAsync(psi, dpsi, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (grad) TRY_(cudaMemcpyAsync(grad, dgrad, (size_t)B * n2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (F1) TRY_(cudaMemcpyAsync(F1, dF1, (size_t)B * n2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (F2 && nf2) TRY_(cudaMemcpyAsync(F2, dF2, (size_t)B * nf2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    TRY_(cudaStreamSynchronize(s));
#undef TRY_
    if (e != cudaSuccess) rc = set_err(h, NMPC_ERR_CUDA, "nmpc_eval_batch: %s", cudaGetErrorString(e));
    cudaFree(dP); cudaFree(dU); cudaFree(dY); cudaFree(dc); cudaFree(dpsi); cudaFree(dgrad); cudaFree(dF1); cudaFree(dF2);
    return rc;
}

int64_t nmpc_launch_count(nmpc_handle* h) { return h ? h->launches : 0; }
double nmpc_last_kernel_ms(nmpc_handle* h) { return h ? h->last_ms : 0.0; }

}  // extern "C"
