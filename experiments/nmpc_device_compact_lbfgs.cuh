// nmpc_device.cuh — device code of the batched NMPC solver (one warp per problem).
//
// What it computes is the problem of MpcModule.build() (src/mpc/mpc_generator.py:66-193) solved
// the way the reference's OpEn solver does (PANOC + L-BFGS inside an ALM/penalty loop); the
// control flow mirrors oracle/nmpc_oracle.c step by step and the arithmetic follows the contract
// in DESIGN.md §4 (explicit fma, own sincos, warp-ordered reductions), so results are bit-identical
// to the oracle.
//
// Organisation:
//   * lane l owns horizon steps t = l + 32*j (P = ceil(N/32) register passes);
//   * rollout and adjoint sweep are Kogge-Stone scans over lanes; reductions are xor-butterflies;
//   * the per-problem constants (segments, circles, ellipses, weights) and the PANOC / L-BFGS
//     vectors live in the warp's shared-memory arena, addressed with explicit 32-bit shared
//     addresses (ld.shared / st.shared) so no generic-address arithmetic is left in the loops;
//   * the solver is a phase machine with ONE evaluation site: every psi / grad psi / F2
//     evaluation of PANOC, the line search, the Lipschitz backtracking and the ALM update goes
//     through the same code, which keeps the kernel small enough for the instruction caches.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/nmpc_b200.h"

#define FULL 0xffffffffu
#ifndef NMPC_OBS_QUICK
#define NMPC_OBS_QUICK 0
#endif
#ifndef NMPC_CTE_STAGE
#define NMPC_CTE_STAGE 0
#endif
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
enum { V_GRAD = 0, V_UHALF, V_FPR, V_DIR, V_GSTEP, V_OLDS, V_OLDG, V_S, V_Y = V_S + MEMP1, V_END = V_Y + MEMP1 };
enum { H_X0 = 0, H_Y0, H_TH0, H_VINIT, H_WINIT, H_XREF, H_YREF, H_THREF, H_Q, H_QV, H_QTH, H_RV, H_RW, H_QN, H_QTHN,
       H_QCTE, H_AP, H_WP, H_INVTS, H_COUNT = 20 };
#define SEG_STRIDE 6   // s1x s1y | dx dy | inv pad
#define CIRC_STRIDE 4  // cx cy | r2 (original slot index as int in the 4th double)
#define ELL_STRIDE 6   // ex ey | cosA sinA | 1/rx^2 1/ry^2

struct Lay {
    int n2, seg, circ, ell, rho, alpha, cco, gram, res, hdr, vref, total;
};
__host__ __device__ inline int even_up(int x) { return (x + 1) & ~1; }
__host__ __device__ inline Lay make_layout(int N, int Nobs, int Nd) {
    Lay L;
    L.n2 = 2 * N;
    int o = V_END * 2 * N;
    L.seg = o; o += SEG_STRIDE * N;
    L.circ = o; o += CIRC_STRIDE * Nobs;
    L.ell = o; o += ELL_STRIDE * Nd * N;
    L.rho = o; o += 12;
    L.alpha = o; o += 12;
    L.cco = o; o += 12;
    L.gram = o; o += 2 * MEMP1 * MEMP1;  // SY then YY, indexed by physical ring slot
    L.res = o; o += 64;                  // results of the batched dot products of one PANOC iteration
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
    long long* dbg;  // NMPC_PROFILE builds only: 8 cycle counters per problem
};

// ---------------------------------------------------------------------------------
// explicit shared-memory access (32-bit shared addresses)
__device__ __forceinline__ double lds1(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double2 lds2(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ int ldsi(uint32_t a) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts1(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts2(uint32_t a, double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void stsi(uint32_t a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// Rectangle::project of OpEn is comparison-based: a NaN stays a NaN (and ends the solve as NotFinite)
__device__ __forceinline__ double clampd(double x, double lo, double hi) { return (x < lo) ? lo : ((x > hi) ? hi : x); }
// min/max as compare-selects (same forms as the oracle): NaN -> the constant, zero results are +0
// (written as setp/selp PTX: the C ternaries get canonicalised to max.f64/min.f64, which sm_100
//  expands into a ~12-instruction DSETP.MAX/FSEL/SEL/NaN-fix-up sequence each)
__device__ __forceinline__ double sel_clamp01(double t) {
    double r;
    asm("{\n\t.reg .pred p;\n\t"
        "setp.gt.f64 p, %1, 0d0000000000000000;\n\tselp.f64 %0, %1, 0d0000000000000000, p;\n\t"
        "setp.lt.f64 p, %0, 0d3FF0000000000000;\n\tselp.f64 %0, %0, 0d3FF0000000000000, p;\n\t}"
        : "=d"(r)
        : "d"(t));
    return r;
}
// if (d2 < best) { best = d2; bi = idx; }  — strict '<': the first minimal segment keeps the gradient
__device__ __forceinline__ void take_if_less(double d2, int idx, double& best, int& bi) {
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f64 p, %2, %0;\n\tselp.f64 %0, %2, %0, p;\n\tselp.s32 %1, %3, %1, p;\n\t}"
        : "+d"(best), "+r"(bi)
        : "d"(d2), "r"(idx));
}
__device__ __forceinline__ double sel_excess(double z, double lo, double hi) { return (z > hi) ? z - hi : ((z < lo) ? z - lo : 0.0); }
// l += y on the lanes where `on` holds, as one predicated DADD (no select pair)
__device__ __forceinline__ void add_if(double& l, double y, bool on) {
    asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p add.f64 %0, %0, %1;\n\t}" : "+d"(l) : "d"(y), "r"((int)on));
}

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
// two sums at once (interleaved shuffles)
template <int P>
__device__ __forceinline__ void hsum2(const double (&e)[P], const double (&f)[P], double& se, double& sf) {
    double a = e[0], b = f[0];
#pragma unroll
    for (int j = 1; j < P; j++) {
        a = a + e[j];
        b = b + f[j];
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
        double ya = __shfl_xor_sync(FULL, a, off), yb = __shfl_xor_sync(FULL, b, off);
        a = a + ya;
        b = b + yb;
    }
    se = a;
    sf = b;
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
            add_if(l, y, lane >= off);
        }
        double lm1 = __shfl_up_sync(FULL, l, 1);
        double g = (j == 0) ? l : carry + l;
        excl[j] = (lane == 0) ? carry : ((j == 0) ? lm1 : carry + lm1);
        incl[j] = g;
        if (j + 1 < P) carry = __shfl_sync(FULL, g, 31);
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
            add_if(la, ya, lane >= off);
            add_if(lb, yb, lane >= off);
        }
        double ma = __shfl_up_sync(FULL, la, 1), mb = __shfl_up_sync(FULL, lb, 1);
        double ga = (j == 0) ? la : ca + la, gb = (j == 0) ? lb : cb + lb;
        ea[j] = (lane == 0) ? ca : ((j == 0) ? ma : ca + ma);
        eb[j] = (lane == 0) ? cb : ((j == 0) ? mb : cb + mb);
        ia[j] = ga;
        ib[j] = gb;
        if (j + 1 < P) {
            ca = __shfl_sync(FULL, ga, 31);
            cb = __shfl_sync(FULL, gb, 31);
        }
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
            add_if(l, y, lane + off < 32);
        }
        double g = (j == P - 1) ? l : carry + l;
        suf[j] = g;
        if (j > 0) carry = __shfl_sync(FULL, g, 0);
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
            add_if(la, ya, lane + off < 32);
            add_if(lb, yb, lane + off < 32);
        }
        double ga = (j == P - 1) ? la : ca + la, gb = (j == P - 1) ? lb : cb + lb;
        sa[j] = ga;
        sb[j] = gb;
        if (j > 0) {
            ca = __shfl_sync(FULL, ga, 0);
            cb = __shfl_sync(FULL, gb, 0);
        }
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


// Transposed warp reduction of 32 per-lane values: on return lane l holds, in v[0], the sum over the 32
// lanes of value l, accumulated in the same order as butterfly() (xor 16, 8, 4, 2, 1 — bit-identical to it).
// 31 exchanges for 32 inner products instead of 160; each stage issues all its shuffles back to back.
__device__ __forceinline__ void treduce32(double (&v)[32], const int lane) {
#pragma unroll
    for (int D = 16; D >= 1; D >>= 1) {
        const bool hi = (lane & D) != 0;
        double recv[16];
#pragma unroll
        for (int k = 0; k < D; k++) {
            const double send = hi ? v[k] : v[k + D];
            recv[k] = __shfl_xor_sync(FULL, send, D);
        }
#pragma unroll
        for (int k = 0; k < D; k++) {
            const double keep = hi ? v[k + D] : v[k];
            v[k] = keep + recv[k];
        }
    }
}
// per-lane partial of a 2N-vector inner product (what hsum() would butterfly)
template <int P>
__device__ __forceinline__ double pdot(const double2 (&a)[P], const double2 (&b)[P]) {
    double e = fma(a[0].y, b[0].y, a[0].x * b[0].x);
#pragma unroll
    for (int j = 1; j < P; j++) e = e + fma(a[j].y, b[j].y, a[j].x * b[j].x);
    return e;
}

// ---------------------------------------------------------------------------------
enum { MODE_COST = 0, MODE_GRAD = 1, MODE_F2 = 2 };
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

// One warp's view of its problem: arena addresses + lane mapping.
template <int P>
struct Warp {
    const nmpc_config& cfg;
    uint32_t sb;         // shared byte address of the arena
    uint32_t la[P];      // sb + 16*t : this lane's element inside vector 0
    uint32_t vstride;    // bytes per vector (2N doubles)
    uint32_t a_seg, a_circ, a_ell, a_rho, a_alpha, a_cco, a_gram, a_res, a_hdr, a_vref;
    int lane, n_circ;    // n_circ: circles with r != 0 (zero-padded slots are skipped: they add exact zeros)
    bool act[P];
    int tix[P];

    __device__ __forceinline__ Warp(const nmpc_config& c, const Lay& L, int warp, int lane_) : cfg(c), lane(lane_) {
        sb = (uint32_t)__cvta_generic_to_shared(smem) + (uint32_t)(warp * L.total) * 8u;
        vstride = (uint32_t)L.n2 * 8u;
        a_seg = sb + L.seg * 8u; a_circ = sb + L.circ * 8u; a_ell = sb + L.ell * 8u; a_rho = sb + L.rho * 8u;
        a_alpha = sb + L.alpha * 8u; a_hdr = sb + L.hdr * 8u; a_vref = sb + L.vref * 8u;
        a_cco = sb + L.cco * 8u; a_gram = sb + L.gram * 8u; a_res = sb + L.res * 8u;
        n_circ = 0;
#pragma unroll
        for (int j = 0; j < P; j++) {
            tix[j] = lane + 32 * j;
            act[j] = tix[j] < cfg.N_hor;
            la[j] = sb + 16u * tix[j];
        }
    }
    __device__ __forceinline__ double hdr(int i) const { return lds1(a_hdr + 8u * i); }
    __device__ __forceinline__ void ld(int k, double2 (&r)[P]) const {
#pragma unroll
        for (int j = 0; j < P; j++) r[j] = act[j] ? lds2(la[j] + k * vstride) : make_double2(0.0, 0.0);
    }
    __device__ __forceinline__ void st(int k, const double2 (&r)[P]) const {
#pragma unroll
        for (int j = 0; j < P; j++)
            if (act[j]) sts2(la[j] + k * vstride, r[j]);
    }

    // unpack the parameter row (layout: include/nmpc_b200.h) into the arena
    __device__ void stage(const double* __restrict__ p) {
        const int N = cfg.N_hor, Nobs = cfg.Nobs, Nd = cfg.Ndynobs;
        __syncwarp();
        if (lane < 8) sts1(a_hdr + 8u * lane, p[lane]);
        if (lane >= 8 && lane < 18) sts1(a_hdr + 8u * lane, p[lane + 2]);
        if (lane == 18) sts1(a_hdr + 8u * H_INVTS, 1.0 / cfg.ts);
        for (int t = lane; t < N; t += 32) sts1(a_vref + 8u * t, p[NMPC_NZ + t]);
        const double* pc = p + NMPC_NZ + N;
        int nreal = 0;
        for (int k0 = 0; k0 < Nobs; k0 += 32) {  // order-preserving compaction of the non-padded circles
            const int k = k0 + lane;
            double cx = 0.0, cy = 0.0, r = 0.0;
            if (k < Nobs) {
                cx = pc[3 * k];
                cy = pc[3 * k + 1];
                r = pc[3 * k + 2];
            }
            const bool real = (k < Nobs) && (r != 0.0);
            const unsigned m = __ballot_sync(FULL, real);
            if (real) {
                const int pos = nreal + __popc(m & ((1u << lane) - 1u));
                const uint32_t a = a_circ + 32u * pos;
                sts2(a, make_double2(cx, cy));
                sts1(a + 16u, r * r);
                stsi(a + 24u, k);
            }
            nreal += __popc(m);
        }
        n_circ = nreal;
        const double* pe = pc + 3 * Nobs;
        const int ne = Nd * N;
        for (int i = lane; i < ne; i += 32) {
            const double* e = pe + 5 * i;  // obstacle-major then time: offset k*5N + 5t = 5*(k*N + t)
            double sa, ca;
            nm_sincos(e[4], sa, ca);
            const uint32_t a = a_ell + 48u * i;
            sts2(a, make_double2(e[0], e[1]));
            sts2(a + 16u, make_double2(ca, sa));
            sts2(a + 32u, make_double2(1.0 / (e[2] * e[2]), 1.0 / (e[3] * e[3])));
        }
        const double* pr = pe + 5 * ne;
        for (int i = lane; i < N; i += 32) {
            if (i >= 1) {
                double ax = pr[3 * (i - 1)], ay = pr[3 * (i - 1) + 1];
                double dx = pr[3 * i] - ax, dy = pr[3 * i + 1] - ay;
                const uint32_t a = a_seg + 48u * i;
                sts2(a, make_double2(ax, ay));
                sts2(a + 16u, make_double2(dx, dy));
                sts1(a + 32u, 1.0 / (fma(dx, dx, dy * dy) + 1e-16));
            }
        }
        __syncwarp();
    }

    // previous step's control for lane-distributed (v, w): lane-1, pass carry, or (v_init, w_init)
    __device__ __forceinline__ void prev_controls(const double2 (&uv)[P], int j, double& vp, double& wp) const {
        vp = __shfl_up_sync(FULL, uv[j].x, 1);
        wp = __shfl_up_sync(FULL, uv[j].y, 1);
        if (j > 0) {
            double v31 = __shfl_sync(FULL, uv[j > 0 ? j - 1 : 0].x, 31), w31 = __shfl_sync(FULL, uv[j > 0 ? j - 1 : 0].y, 31);
            if (lane == 0) {
                vp = v31;
                wp = w31;
            }
        } else if (lane == 0) {
            vp = hdr(H_VINIT);
            wp = hdr(H_WINIT);
        }
    }

    // psi / grad psi / F2 for the staged problem (mode is warp-uniform)
    __device__ double eval(const int mode, const double2 (&uv)[P], const Pen pn, const double2 (&yl)[P],
                           double2 (&gout)[P], double& pen_out, double* __restrict__ F2g) {
        const bool GRAD = (mode == MODE_GRAD);
        const int N = cfg.N_hor;
        const double ts = cfg.ts;
        double tw[P], inclT[P], exclT[P];
#pragma unroll
        for (int j = 0; j < P; j++) tw[j] = act[j] ? ts * uv[j].y : 0.0;
        prefix_scan<P>(tw, inclT, exclT, lane);
        double sn[P], cs[P], thpre[P], TH[P], a[P], b[P];
        const double th0 = hdr(H_TH0);
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
            const double x0 = hdr(H_X0), y0 = hdr(H_Y0);
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
        const double qcte = hdr(H_QCTE);

        if (mode != MODE_F2) {
            // cross-track error: each lane scans the N-1 segments for its own predicted point
            double best[P];
            int bi[P];
#pragma unroll
            for (int j = 0; j < P; j++) {
                best[j] = CUDART_INF;
                bi[j] = 1;
            }
            constexpr int UNR = (P == 1) ? 4 : 2;
            uint32_t as = a_seg + 48u;
            int i = 1;
            // UNR segments per trip, written stage by stage: the SM issues in order, so independent
            // chains only overlap if they are interleaved in the instruction stream
            for (; i + UNR <= N; i += UNR, as += 48u * UNR) {
                double2 s1[UNR], d[UNR];
                double inv[UNR];
#pragma unroll
                for (int q = 0; q < UNR; q++) {
                    s1[q] = lds2(as + 48u * q);
                    d[q] = lds2(as + 48u * q + 16u);
                    inv[q] = lds1(as + 48u * q + 32u);
                }
#pragma unroll
                for (int j = 0; j < P; j++) {
#if NMPC_CTE_STAGE
                    double px[UNR], py[UNR], tt[UNR], ex[UNR], ey[UNR], d2[UNR];
#pragma unroll
                    for (int q = 0; q < UNR; q++) {
                        px[q] = X[j] - s1[q].x;
                        py[q] = Y[j] - s1[q].y;
                    }
#pragma unroll
                    for (int q = 0; q < UNR; q++) tt[q] = py[q] * d[q].y;
#pragma unroll
                    for (int q = 0; q < UNR; q++) tt[q] = fma(px[q], d[q].x, tt[q]);
#pragma unroll
                    for (int q = 0; q < UNR; q++) tt[q] = tt[q] * inv[q];
#pragma unroll
                    for (int q = 0; q < UNR; q++) tt[q] = sel_clamp01(tt[q]);
#pragma unroll
                    for (int q = 0; q < UNR; q++) {
                        ex[q] = fma(tt[q], d[q].x, -px[q]);
                        ey[q] = fma(tt[q], d[q].y, -py[q]);
                    }
#pragma unroll
                    for (int q = 0; q < UNR; q++) d2[q] = ey[q] * ey[q];
#pragma unroll
                    for (int q = 0; q < UNR; q++) d2[q] = fma(ex[q], ex[q], d2[q]);
#pragma unroll
#else
                    double d2[UNR];
#pragma unroll
                    for (int q = 0; q < UNR; q++) {
                        double px = X[j] - s1[q].x, py = Y[j] - s1[q].y;
                        double that = fma(px, d[q].x, py * d[q].y) * inv[q];
                        double tst = sel_clamp01(that);
                        double ex = fma(tst, d[q].x, -px), ey = fma(tst, d[q].y, -py);
                        d2[q] = fma(ex, ex, ey * ey);
                    }
#pragma unroll
#endif
                    for (int q = 0; q < UNR; q++) take_if_less(d2[q], i + q, best[j], bi[j]);
                }
            }
            for (; i < N; i++, as += 48u) {
                const double2 s1 = lds2(as), d = lds2(as + 16u);
                const double inv = lds1(as + 32u);
#pragma unroll
                for (int j = 0; j < P; j++) {
                    double px = X[j] - s1.x, py = Y[j] - s1.y;
                    double that = fma(px, d.x, py * d.y) * inv;
                    double tst = sel_clamp01(that);
                    double ex = fma(tst, d.x, -px), ey = fma(tst, d.y, -py);
                    double d2 = fma(ex, ex, ey * ey);
                    take_if_less(d2, i, best[j], bi[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < P; j++) {
                mind2[j] = best[j];
                if (GRAD) {  // redo the arg-min segment (same operations, same bits) for the gradient
                    const uint32_t ab = a_seg + 48u * bi[j];
                    const double2 s1 = lds2(ab), d = lds2(ab + 16u);
                    const double inv = lds1(ab + 32u);
                    double px = X[j] - s1.x, py = Y[j] - s1.y;
                    double that = fma(px, d.x, py * d.y) * inv;
                    double tst = sel_clamp01(that);
                    double ex = fma(tst, d.x, -px), ey = fma(tst, d.y, -py);
                    double ed = (that >= 0.0 && that <= 1.0) ? fma(ex, d.x, ey * d.y) * inv : 0.0;
                    double k2 = 2.0 * qcte;
                    gX[j] = k2 * fma(ed, d.x, -ex);
                    gY[j] = k2 * fma(ed, d.y, -ey);
                }
            }
        }

        // obstacle penalty F2: circles (non-padded ones), then this lane's time slice of each ellipse.
        // Obstacles are tested in chunks: all inside-tests and votes of a chunk are issued back to back and
        // ONE branch decides whether any of them needs the (rare) ordered time-sum + gradient path.
        double pen = 0.0;
        auto ordered_sum = [&](const double(&h)[P], const unsigned(&m)[P]) {
            double g = 0.0;  // sum over active steps in ascending t (skipped terms are exact zeros)
#pragma unroll
            for (int j = 0; j < P; j++) {
                unsigned mm = m[j];
                while (mm) {
                    int src = __ffs(mm) - 1;
                    g = g + __shfl_sync(FULL, h[j], src);
                    mm &= mm - 1;
                }
            }
            return g;
        };
        {
            constexpr int CH = (P == 1) ? 4 : 2;
            uint32_t ac = a_circ;
            for (int k0 = 0; k0 < n_circ; k0 += CH, ac += 32u * CH) {
                double h[CH][P], dx[CH][P], dy[CH][P];
                unsigned m[CH][P];
                unsigned any = 0;
#pragma unroll
                for (int q = 0; q < CH; q++) {
                    const bool valid = (k0 + q) < n_circ;  // warp-uniform; slots past n_circ hold stale data
                    const double2 cxy = lds2(ac + 32u * q);
                    const double r2 = lds1(ac + 32u * q + 16u);
#pragma unroll
                    for (int j = 0; j < P; j++) {
                        dx[q][j] = X[j] - cxy.x;
                        dy[q][j] = Y[j] - cxy.y;
                        const double hh = fma(-dy[q][j], dy[q][j], fma(-dx[q][j], dx[q][j], r2));
                        h[q][j] = valid ? hh : -1.0;
                    }
                }
#pragma unroll
                for (int q = 0; q < CH; q++)
#pragma unroll
                    for (int j = 0; j < P; j++) {
                        m[q][j] = __ballot_sync(FULL, act[j] && h[q][j] > 0.0);
                        any |= m[q][j];
                    }
                if (any) {
#pragma unroll
                    for (int q = 0; q < CH; q++) {
                        unsigned anyq = 0;
#pragma unroll
                        for (int j = 0; j < P; j++) anyq |= m[q][j];
                        if (anyq) {
                            const double g = ordered_sum(h[q], m[q]);
                            if (F2g && lane == 0) F2g[ldsi(ac + 32u * q + 24u)] = g;
                            pen = fma(g, g, pen);
                            if (GRAD && g > 0.0) {
                                const double cg = pn.c * g;
#pragma unroll
                                for (int j = 0; j < P; j++)
                                    if (act[j] && h[q][j] > 0.0) {
                                        gX[j] = fma(cg, -2.0 * dx[q][j], gX[j]);
                                        gY[j] = fma(cg, -2.0 * dy[q][j], gY[j]);
                                    }
                            }
                        }
                    }
                }
            }
            constexpr int CE = (P == 1) ? 3 : 1;
            const int Nd = cfg.Ndynobs;
            for (int k0 = 0; k0 < Nd; k0 += CE) {
                double h[CE][P], ta[CE][P], tb[CE][P], eca[CE][P], esa[CE][P];
                unsigned m[CE][P];
                unsigned any = 0;
#pragma unroll
                for (int q = 0; q < CE; q++) {
                    const bool valid = (k0 + q) < Nd;
                    const int k = valid ? k0 + q : k0;
#pragma unroll
                    for (int j = 0; j < P; j++) {
                        const uint32_t ae = a_ell + 48u * (k * N + (act[j] ? tix[j] : 0));
                        const double2 exy = lds2(ae), csa = lds2(ae + 16u), ir = lds2(ae + 32u);
                        const double dx = X[j] - exy.x, dy = Y[j] - exy.y;
                        eca[q][j] = csa.x;
                        esa[q][j] = csa.y;
                        const double ea = fma(dx, csa.x, dy * csa.y);
                        const double eb = fma(dx, csa.y, -(dy * csa.x));
                        const double hh = fma(-(eb * eb), ir.y, fma(-(ea * ea), ir.x, 1.0));
                        h[q][j] = valid ? hh : -1.0;
                        ta[q][j] = ea * ir.x;
                        tb[q][j] = eb * ir.y;
                    }
                }
#pragma unroll
                for (int q = 0; q < CE; q++)
#pragma unroll
                    for (int j = 0; j < P; j++) {
                        m[q][j] = __ballot_sync(FULL, act[j] && h[q][j] > 0.0);
                        any |= m[q][j];
                    }
                if (any) {
#pragma unroll
                    for (int q = 0; q < CE; q++) {
                        unsigned anyq = 0;
#pragma unroll
                        for (int j = 0; j < P; j++) anyq |= m[q][j];
                        if (anyq) {
                            const double g = ordered_sum(h[q], m[q]);
                            if (F2g && lane == 0) F2g[cfg.Nobs + k0 + q] = g;
                            pen = fma(g, g, pen);
                            if (GRAD && g > 0.0) {
                                const double cg = pn.c * g;
#pragma unroll
                                for (int j = 0; j < P; j++)
                                    if (act[j] && h[q][j] > 0.0) {
                                        double hX = -2.0 * fma(ta[q][j], eca[q][j], tb[q][j] * esa[q][j]);
                                        double hY = -2.0 * fma(ta[q][j], esa[q][j], -(tb[q][j] * eca[q][j]));
                                        gX[j] = fma(cg, hX, gX[j]);
                                        gY[j] = fma(cg, hY, gY[j]);
                                    }
                            }
                        }
                    }
                }
            }
        }
        pen_out = pen;
        if (mode == MODE_F2) return 0.0;

        // stage cost, acceleration cost, ALM term
        const double inv_ts = hdr(H_INVTS);
        const double xref = hdr(H_XREF), yref = hdr(H_YREF), thref = hdr(H_THREF);
        const double w_rv = hdr(H_RV), w_rw = hdr(H_RW), w_qv = hdr(H_QV), w_q = hdr(H_Q), w_qth = hdr(H_QTH);
        const double w_ap = hdr(H_AP), w_wp = hdr(H_WP), w_qN = hdr(H_QN), w_qthN = hdr(H_QTHN);
        double cl[P], Aa[P], Aw[P], vref[P];
#pragma unroll
        for (int j = 0; j < P; j++) {
            const double v = uv[j].x, w = uv[j].y;
            double vp, wp_;
            prev_controls(uv, j, vp, wp_);
            double c0 = w_rv * (v * v);
            c0 = fma(w_rw, w * w, c0);
            vref[j] = act[j] ? lds1(a_vref + 8u * tix[j]) : 0.0;
            double dv = v - vref[j];
            c0 = fma(w_qv, dv * dv, c0);
            double ex = xpre[j] - xref, ey = ypre[j] - yref, et = thpre[j] - thref;
            c0 = fma(w_q, fma(ex, ex, ey * ey), c0);
            c0 = fma(w_qth, et * et, c0);
            c0 = fma(qcte, mind2[j], c0);
            double acc = (v - vp) * inv_ts, aac = (w - wp_) * inv_ts;
            c0 = fma(w_ap, acc * acc, c0);
            c0 = fma(w_wp, aac * aac, c0);
            double za = fma(yl[j].x, pn.inv_c, acc), zw = fma(yl[j].y, pn.inv_c, aac);
            double da = sel_excess(za, cfg.lin_acc_min, cfg.lin_acc_max);
            double dw = sel_excess(zw, -cfg.ang_acc_max, cfg.ang_acc_max);
            c0 = fma(pn.hc, fma(da, da, dw * dw), c0);
            cl[j] = act[j] ? c0 : 0.0;
            Aa[j] = act[j] ? fma(pn.c, da, (2.0 * w_ap) * acc) * inv_ts : 0.0;
            Aw[j] = act[j] ? fma(pn.c, dw, (2.0 * w_wp) * aac) * inv_ts : 0.0;
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
        const double term = fma(w_qN, fma(eXN, eXN, eYN * eYN), w_qthN * (eTN * eTN));
        const double psi = fma(pn.hc, pen, hsum<P>(cl) + term);
        if (!GRAD) return psi;

        // backward sweep
        double mth[P];
#pragma unroll
        for (int j = 0; j < P; j++) {
            const bool last = !(tix[j] + 1 < N);
            const double qq = last ? w_qN : w_q, qt = last ? w_qthN : w_qth;
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
            if (j + 1 < P) {
                double n0 = __shfl_sync(FULL, nn[(j + 1 < P) ? j + 1 : j], 0);
                if (lane == 31) nx = n0;
            } else if (lane == 31) nx = 0.0;
            rr[j] = act[j] ? mth[j] + nx : 0.0;
        }
        suffix_scan<P>(rr, TT, lane);
#pragma unroll
        for (int j = 0; j < P; j++) {
            const double v = uv[j].x, w = uv[j].y;
            double An = __shfl_down_sync(FULL, Aa[j], 1), Wn = __shfl_down_sync(FULL, Aw[j], 1);
            if (j + 1 < P) {
                double A0 = __shfl_sync(FULL, Aa[(j + 1 < P) ? j + 1 : j], 0), W0 = __shfl_sync(FULL, Aw[(j + 1 < P) ? j + 1 : j], 0);
                if (lane == 31) {
                    An = A0;
                    Wn = W0;
                }
            } else if (lane == 31) {
                An = 0.0;
                Wn = 0.0;
            }
            double lv = fma(2.0 * w_rv, v, (2.0 * w_qv) * (v - vref[j])) + (Aa[j] - An);
            double lw = (2.0 * w_rw) * w + (Aw[j] - Wn);
            double gv = fma(ts, fma(cs[j], LX[j], sn[j] * LY[j]), lv);
            double gw = fma(ts, TT[j], lw);
            gout[j] = act[j] ? make_double2(gv, gw) : make_double2(0.0, 0.0);
        }
        return psi;
    }
};

// ---------------------------------------------------------------------------------
// The solver: ALM/PM outer loop around PANOC as a phase machine with one evaluation site.
// Phases that end in an evaluation set (x, mode) and fall through to it; the others `continue`.
enum Phase {
    PH_OUTER_BEGIN, PH_INIT_A, PH_INIT_B, PH_STEP_BEGIN, PH_LIP, PH_COST_U, PH_LIP_LOOP, PH_LIP_RETRY, PH_IT0, PH_LS,
    PH_STEP_DONE, PH_SOLVE_END, PH_F2, PH_FINAL, PH_EXIT
};

template <int P>
__device__ int solve_problem(Warp<P>& W, double2 (&u)[P], double2 (&yl)[P], nmpc_stats& st_out, long long* prof_out = nullptr) {
#ifdef NMPC_PROFILE
    long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long tstart = clock64();
#endif
    const nmpc_config& cfg = W.cfg;
    const int lane = W.lane;
    const int mem = cfg.lbfgs_memory, mem1 = cfg.lbfgs_memory + 1;
    const int nf2 = cfg.Nobs + cfg.Ndynobs;
    // warp-uniform state
    double gamma = 0.0, inv_gamma = 0.0, sigma = 0.0, lip = 0.0, cost = 0.0, norm_fpr = 0.0, tau = 1.0;
    double akkt_tol = cfg.initial_tolerance, cost_half = 0.0, norm_h = 0.0, rhs_ls = 0.0, lb_gamma = 1.0;
    Pen pn = make_pen(cfg.initial_penalty);
    Pen pn_eval = pn;
    int iteration = 0, n_cost = 0, n_grad = 0, lb_active = 0, lb_first = 1, lb_head = 0;
    int alm_iter = 0, inner_total = 0, num_outer = 0, status = NMPC_CONVERGED, inner_status = NMPC_CONVERGED;
    int num_iter = 0, it_lip = 0, nls = 0;
    bool cont = true;
    double f2n = 0.0, f2np = 0.0, dyn = 0.0, dynp = 0.0;
    const double inv_ts = W.hdr(H_INVTS);

    double2 x[P], g[P];  // evaluation point / gradient out
    double pen = 0.0;
    int mode = MODE_GRAD;
    int phase = PH_OUTER_BEGIN;

    auto set_gamma = [&](double gm) {
        gamma = gm;
        inv_gamma = 1.0 / gm;
    };
    // gradient_step() + half_step(): gstep = p - gamma*grad ; uhalf = Proj_U(gstep); both stored
    auto grad_step_half = [&](const double2(&p)[P], const double2(&gr)[P], double2(&gs)[P], double2(&uh)[P]) {
#pragma unroll
        for (int j = 0; j < P; j++) {
            gs[j].x = fma(-gamma, gr[j].x, p[j].x);
            gs[j].y = fma(-gamma, gr[j].y, p[j].y);
            uh[j].x = W.act[j] ? clampd(gs[j].x, cfg.lin_vel_min, cfg.lin_vel_max) : 0.0;
            uh[j].y = W.act[j] ? clampd(gs[j].y, -cfg.ang_vel_max, cfg.ang_vel_max) : 0.0;
        }
        W.st(V_GSTEP, gs);
        W.st(V_UHALF, uh);
    };
    auto compute_fpr = [&](const double2(&uh)[P], double2(&fpr)[P]) {
        double e[P];
#pragma unroll
        for (int j = 0; j < P; j++) {
            double d0 = u[j].x - uh[j].x, d1 = u[j].y - uh[j].y;
            fpr[j] = make_double2(d0, d1);
            e[j] = fma(d1, d1, d0 * d0);
        }
        norm_fpr = sqrt(hsum<P>(e));
    };
    auto slot = [&](int i) {
        int s = lb_head + i;
        return (s >= mem1) ? s - mem1 : s;
    };
    auto RES = [&](int i) { return lds1(W.a_res + 8u * i); };
    // RES index of the c-th inner product of stored pair i (see batch_dots)
    auto RIX = [&](int i, int c) { return (i < 4) ? 9 + 5 * i + c : 32 + 5 * (i - 4) + c; };
    // All inner products of one PANOC iteration as ONE batched reduction (results -> RES(i)):
    //   0 fpr.fpr  1 grad.fpr  2 grad.grad  3 |gstep-uhalf|^2  4 s.y  5 s.s  6 y.y  7 s.fpr  8 y.fpr
    //   RIX(i, c) for stored pair i:  c = 0 s_i.fpr  1 y_i.fpr  2 s.y_i  3 s_i.y  4 y.y_i   (s = u-old_state, y = fpr-old_g)
    // Values are bit-identical to wdot()/wdiff2() of the same vectors (same lane partials, same butterfly order).
    auto batch_dots = [&](const double2(&fpr)[P], const double2(&uh)[P]) {
#ifdef NMPC_PROFILE
        const long long tb0 = clock64();
#endif
        const int kact = lb_active;
        const bool pair = !lb_first;
        double2 sv[P], yv[P];
        double v[32];
        {
            double2 gr[P], gs[P];
            W.ld(V_GRAD, gr);
            W.ld(V_GSTEP, gs);
            v[0] = pdot<P>(fpr, fpr);
            v[1] = pdot<P>(gr, fpr);
            v[2] = pdot<P>(gr, gr);
            double e = 0.0;
#pragma unroll
            for (int j = 0; j < P; j++) {
                double d0 = gs[j].x - uh[j].x, d1 = gs[j].y - uh[j].y;
                double ej = fma(d1, d1, d0 * d0);
                e = (j == 0) ? ej : e + ej;
            }
            v[3] = e;
        }
        if (pair) {
            double2 os[P], og[P];
            W.ld(V_OLDS, os);
            W.ld(V_OLDG, og);
#pragma unroll
            for (int j = 0; j < P; j++) {
                sv[j] = make_double2(u[j].x - os[j].x, u[j].y - os[j].y);
                yv[j] = make_double2(fpr[j].x - og[j].x, fpr[j].y - og[j].y);
            }
        } else {
#pragma unroll
            for (int j = 0; j < P; j++) sv[j] = yv[j] = make_double2(0.0, 0.0);
        }
        v[4] = pdot<P>(sv, yv);
        v[5] = pdot<P>(sv, sv);
        v[6] = pdot<P>(yv, yv);
        v[7] = pdot<P>(sv, fpr);
        v[8] = pdot<P>(yv, fpr);
        auto pair_dots = [&](int i, double& p0, double& p1, double& p2, double& p3, double& p4) {
            p0 = p1 = p2 = p3 = p4 = 0.0;
            if (i < kact) {
                double2 si[P], yi[P];
                const int sl = slot(i);
                W.ld(V_S + sl, si);
                W.ld(V_Y + sl, yi);
                p0 = pdot<P>(si, fpr);
                p1 = pdot<P>(yi, fpr);
                p2 = pdot<P>(sv, yi);
                p3 = pdot<P>(si, yv);
                p4 = pdot<P>(yv, yi);
            }
        };
#pragma unroll
        for (int i = 0; i < 4; i++) pair_dots(i, v[9 + 5 * i], v[10 + 5 * i], v[11 + 5 * i], v[12 + 5 * i], v[13 + 5 * i]);
        v[29] = v[30] = v[31] = 0.0;
        treduce32(v, lane);
        sts1(W.a_res + 8u * lane, v[0]);
        if (kact > 4) {
#pragma unroll
            for (int i = 0; i < 6; i++) pair_dots(4 + i, v[5 * i], v[5 * i + 1], v[5 * i + 2], v[5 * i + 3], v[5 * i + 4]);
            v[30] = v[31] = 0.0;
            treduce32(v, lane);
            sts1(W.a_res + 8u * (32 + lane), v[0]);
        }
        __syncwarp();
#ifdef NMPC_PROFILE
        prof[7] += clock64() - tb0;
#endif
    };

    for (;;) {
        // ------------------------------------------------------------------ pre: pick (x, mode)
        switch (phase) {
            case PH_OUTER_BEGIN: {
                num_outer++;
#pragma unroll
                for (int j = 0; j < P; j++) {  // project_on_set_y
                    yl[j].x = clampd(yl[j].x, -Y_SET_BOUND, Y_SET_BOUND);
                    yl[j].y = clampd(yl[j].y, -Y_SET_BOUND, Y_SET_BOUND);
                }
                // panoc init
                lb_active = 0;
                lb_first = 1;
                tau = 1.0;
                iteration = 0;
#pragma unroll
                for (int j = 0; j < P; j++) x[j] = u[j];
                mode = MODE_GRAD;
                phase = PH_INIT_A;
                break;
            }
            case PH_STEP_BEGIN: {
                double2 uh[P], fpr[P];
                W.ld(V_UHALF, uh);
#pragma unroll
                for (int j = 0; j < P; j++) fpr[j] = make_double2(u[j].x - uh[j].x, u[j].y - uh[j].y);
                W.st(V_FPR, fpr);
                batch_dots(fpr, uh);  // every inner product this iteration needs, in one batched reduction
                norm_fpr = sqrt(RES(0));
                bool exit_now = false;
                if (norm_fpr < cfg.tolerance) {
                    double2 gr[P];
                    W.ld(V_GRAD, gr);
                    double e[P];
#pragma unroll
                    for (int j = 0; j < P; j++) {
                        double p0 = iteration ? gr[j].x : 0.0, p1 = iteration ? gr[j].y : 0.0;
                        double r0 = fma(fpr[j].x, inv_gamma, gr[j].x) - p0;
                        double r1 = fma(fpr[j].y, inv_gamma, gr[j].y) - p1;
                        e[j] = fma(r1, r1, r0 * r0);
                    }
                    exit_now = sqrt(hsum<P>(e)) < akkt_tol;
                }
                if (exit_now) {
                    phase = PH_SOLVE_END;
                    continue;
                }
                it_lip = 0;
#pragma unroll
                for (int j = 0; j < P; j++) x[j] = uh[j];
                mode = MODE_COST;
                phase = PH_LIP;
                break;
            }
            case PH_LIP_LOOP: {
                const double ip = RES(1);
                const double rhs = cost + LIPSCHITZ_UPDATE_EPSILON * fabs(cost) - ip +
                                   (GAMMA_L_COEFF * 0.5 * inv_gamma) * (norm_fpr * norm_fpr);
                if (cost_half > rhs && it_lip < MAX_LIPSCHITZ_UPDATE_ITERATIONS && lip < MAX_LIPSCHITZ_CONSTANT) {
                    lb_active = 0;
                    lb_first = 1;
                    lip *= 2.0;
                    set_gamma(gamma / 2.0);
                    double2 gr[P], gs[P], uh[P];
                    W.ld(V_GRAD, gr);
                    grad_step_half(u, gr, gs, uh);
#pragma unroll
                    for (int j = 0; j < P; j++) x[j] = uh[j];
                    mode = MODE_COST;
                    phase = PH_LIP_RETRY;
                    break;
                }
                sigma = (1.0 - GAMMA_L_COEFF) / (4.0 * gamma);
                double2 fpr[P];
                W.ld(V_FPR, fpr);
                // lbfgs_direction(): update_hessian(g = fpr, state = u) with the batched ys, ss, yy
                const int k_old = lb_active;
                bool have_new = false;
                if (lb_first) {
                    lb_first = 0;
                    W.st(V_OLDS, u);
                    W.st(V_OLDG, fpr);
                } else {
                    const double ys = RES(4), ss = RES(5), yy = RES(6);
                    const double rho_new = 1.0 / ys;
                    bool accept = !(ss <= DBL_EPS || ys <= SY_EPSILON);
                    if (accept) {
                        const double lhs = ys / ss, rhsb = CBFGS_EPSILON * sqrt(RES(0));
                        accept = (lhs > rhsb && isfinite(lhs) && isfinite(rhsb));
                    }
                    if (accept) {
                        double2 os[P], og[P], sv[P], yv[P];
                        W.ld(V_OLDS, os);
                        W.ld(V_OLDG, og);
#pragma unroll
                        for (int j = 0; j < P; j++) {
                            sv[j] = make_double2(u[j].x - os[j].x, u[j].y - os[j].y);
                            yv[j] = make_double2(fpr[j].x - og[j].x, fpr[j].y - og[j].y);
                        }
                        const int tmp = slot(mem);
                        W.st(V_S + tmp, sv);
                        W.st(V_Y + tmp, yv);
                        W.st(V_OLDS, u);
                        W.st(V_OLDG, fpr);
                        // Gram rows/columns of the new pair: lane i copies the three entries against old pair i
                        if (lane < k_old) {
                            const int pi = slot(lane);
                            const double sy_new_i = RES(RIX(lane, 2)), sy_i_new = RES(RIX(lane, 3));
                            const double yy_i = RES(RIX(lane, 4));
                            sts1(W.a_gram + 8u * (tmp * MEMP1 + pi), sy_new_i);
                            sts1(W.a_gram + 8u * (pi * MEMP1 + tmp), sy_i_new);
                            sts1(W.a_gram + 8u * (MEMP1 * MEMP1 + tmp * MEMP1 + pi), yy_i);
                            sts1(W.a_gram + 8u * (MEMP1 * MEMP1 + pi * MEMP1 + tmp), yy_i);
                        }
                        if (lane == 0) {
                            sts1(W.a_gram + 8u * (tmp * MEMP1 + tmp), ys);
                            sts1(W.a_gram + 8u * (MEMP1 * MEMP1 + tmp * MEMP1 + tmp), yy);
                            sts1(W.a_rho + 8u * tmp, rho_new);
                        }
                        lb_head = (lb_head + mem >= mem1) ? lb_head + mem - mem1 : lb_head + mem;
                        lb_gamma = (1.0 / rho_new) / yy;
                        lb_active = (lb_active + 1 < mem) ? lb_active + 1 : mem;
                        have_new = true;
                        __syncwarp();
                    }
                }
                if (iteration == 0) {  // update_no_linesearch(): u <- uhalf
                    W.ld(V_UHALF, u);
#pragma unroll
                    for (int j = 0; j < P; j++) x[j] = u[j];
                    mode = MODE_GRAD;
                    phase = PH_IT0;
                    break;
                }
                // direction = H * fpr: two-loop recursion in compact form (scalar recursion on the Gram entries)
#ifdef NMPC_PROFILE
                const long long tl0 = clock64();
#endif
                double2 q[P];
                const int k = lb_active;
                if (k > 0) {
                    // Lane i owns stored pair i (new logical order: 0 = the pair just accepted, if any).
                    //   forward solve :  alpha_j = rho_j acc_j ;  acc_i -= alpha_j s_i.y_j (i > j) ;  t_i -= alpha_j y_i.y_j
                    //   backward solve:  c_l = alpha_l - rho_l t_l ;  t_i += c_l s_l.y_i (i < l)        (t_i scaled by gamma between)
                    // one shuffle + one fma per step instead of a 40-element reduction per step.
                    const bool mine = lane < k;
                    const int li = mine ? lane : 0;
                    const int pi = slot(li);
                    const int ri = have_new ? (li == 0 ? 7 : RIX(li - 1, 0)) : RIX(li, 0);  // RES index of s_i.fpr (y_i.fpr follows)
                    double acc = RES(ri), t = RES(ri + 1);
                    const double rho_i = lds1(W.a_rho + 8u * pi);
                    const uint32_t row_sy = W.a_gram + 8u * (pi * MEMP1);                    // s_i.y_*
                    const uint32_t row_yy = W.a_gram + 8u * (MEMP1 * MEMP1 + pi * MEMP1);    // y_i.y_*
                    const uint32_t col_sy = W.a_gram + 8u * pi;                              // s_*.y_i
                    double al_i = 0.0, c_i = 0.0;
                    for (int j = 0; j < k; j++) {
                        const int pj = slot(j);
                        const double alj = __shfl_sync(FULL, rho_i * acc, j);
                        if (lane == j) al_i = alj;
                        const double syij = lds1(row_sy + 8u * pj), yyij = lds1(row_yy + 8u * pj);
                        if (lane > j) acc = fma(-alj, syij, acc);
                        t = fma(-alj, yyij, t);
                    }
                    t = lb_gamma * t;
                    for (int l = k - 1; l >= 0; l--) {
                        const int pl = slot(l);
                        const double cl = __shfl_sync(FULL, al_i - rho_i * t, l);
                        if (lane == l) c_i = cl;
                        const double syli = lds1(col_sy + 8u * (pl * MEMP1));
                        if (lane < l) t = fma(cl, syli, t);
                    }
                    if (mine) {
                        sts1(W.a_alpha + 8u * lane, al_i);
                        sts1(W.a_cco + 8u * lane, c_i);
                    }
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < P; j++) q[j] = make_double2(lb_gamma * fpr[j].x, lb_gamma * fpr[j].y);
                    for (int jj = 0; jj < k; jj++) {
                        const double co = -(lb_gamma * lds1(W.a_alpha + 8u * jj));
                        double2 yv[P];
                        W.ld(V_Y + slot(jj), yv);
#pragma unroll
                        for (int j = 0; j < P; j++) {
                            q[j].x = fma(co, yv[j].x, q[j].x);
                            q[j].y = fma(co, yv[j].y, q[j].y);
                        }
                    }
                    for (int l = k - 1; l >= 0; l--) {
                        const double co = lds1(W.a_cco + 8u * l);
                        double2 sv[P];
                        W.ld(V_S + slot(l), sv);
#pragma unroll
                        for (int j = 0; j < P; j++) {
                            q[j].x = fma(co, sv[j].x, q[j].x);
                            q[j].y = fma(co, sv[j].y, q[j].y);
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < P; j++) q[j] = fpr[j];
                }
                W.st(V_DIR, q);
#ifdef NMPC_PROFILE
                prof[4] += clock64() - tl0;
                prof[5]++;
#endif
                // linesearch(): right-hand side on the forward-backward envelope (gg, dist2 from the batch)
                {
                    const double dist2 = RES(3), gg = RES(2);
                    const double fbe = cost - (0.5 * gamma) * gg + (0.5 * dist2) * inv_gamma;
                    rhs_ls = fbe - sigma * (norm_fpr * norm_fpr);
                }
                tau = 1.0;
                nls = 0;
#pragma unroll
                for (int j = 0; j < P; j++) {  // tau = 1: u - 0*fpr - 1*dir
                    x[j].x = fma(-tau, q[j].x, fma(-0.0, fpr[j].x, u[j].x));
                    x[j].y = fma(-tau, q[j].y, fma(-0.0, fpr[j].y, u[j].y));
                }
                mode = MODE_GRAD;
                phase = PH_LS;
                break;
            }
            case PH_STEP_DONE: {
                if (!cont) {
                    phase = PH_SOLVE_END;
                    continue;
                }
                num_iter++;
                cont = num_iter < cfg.max_inner_iterations;
                phase = PH_STEP_BEGIN;
                continue;
            }
            case PH_SOLVE_END: {
                inner_total += num_iter;
                bool fin = true;
#pragma unroll
                for (int j = 0; j < P; j++) fin = fin && isfinite(u[j].x) && isfinite(u[j].y);
                if (!__all_sync(FULL, fin)) {
                    status = NMPC_NOT_FINITE;
                    phase = PH_EXIT;
                    continue;
                }
                W.ld(V_UHALF, u);
                inner_status = cont ? NMPC_CONVERGED : NMPC_NOT_CONVERGED_ITERATIONS;
                status = inner_status;
#pragma unroll
                for (int j = 0; j < P; j++) x[j] = u[j];
                mode = MODE_F2;
                phase = PH_F2;
                break;
            }
            case PH_EXIT: {
#ifdef NMPC_PROFILE
                prof[6] = clock64() - tstart;
                if (prof_out && lane == 0)
                    for (int i = 0; i < 8; i++) prof_out[i] = prof[i];
#endif
                st_out.exit_status = status;
                st_out.outer_iterations = num_outer;
                st_out.inner_iterations = inner_total;
                st_out.last_norm_fpr = norm_fpr;
                st_out.delta_y_norm_over_c = dynp / pn.c;
                st_out.f2_norm = f2np;
                st_out.penalty = pn.c;
                if (status == NMPC_NOT_FINITE) st_out.cost = CUDART_NAN;
                st_out.n_cost_evals = n_cost;
                st_out.n_grad_evals = n_grad;
                st_out.reserved = 0;
                return status;
            }
            default:
                break;  // phases entered with (x, mode) already set
        }

        // ------------------------------------------------------------------ the one evaluation site
        pn_eval = (phase == PH_FINAL) ? make_pen(0.0) : pn;
#ifdef NMPC_PROFILE
        const long long tp0 = clock64();
#endif
        const double psi = W.eval(mode, x, pn_eval, yl, g, pen, nullptr);
#ifdef NMPC_PROFILE
        {
            const long long dt = clock64() - tp0;
            if (mode == MODE_GRAD) { prof[0] += dt; prof[1]++; } else { prof[2] += dt; prof[3]++; }
        }
#endif
        if (mode == MODE_GRAD) n_grad++;
        if (mode == MODE_COST && phase != PH_FINAL) n_cost++;

        // ------------------------------------------------------------------ post
        switch (phase) {
            case PH_INIT_A: {  // cost/gradient at u; then perturb u by h (estimate_loc_lip leaves it perturbed)
                cost = psi;
                W.st(V_GRAD, g);
                double e[P];
#pragma unroll
                for (int j = 0; j < P; j++) {
                    const double ex_ = EPSILON_LIPSCHITZ * u[j].x, ey_ = EPSILON_LIPSCHITZ * u[j].y;
                    double hx = W.act[j] ? ((ex_ > DELTA_LIPSCHITZ) ? ex_ : DELTA_LIPSCHITZ) : 0.0;
                    double hy = W.act[j] ? ((ey_ > DELTA_LIPSCHITZ) ? ey_ : DELTA_LIPSCHITZ) : 0.0;
                    e[j] = fma(hy, hy, hx * hx);
                    u[j].x = u[j].x + hx;
                    u[j].y = u[j].y + hy;
                    x[j] = u[j];
                }
                norm_h = sqrt(hsum<P>(e));
                mode = MODE_GRAD;
                phase = PH_INIT_B;
                break;
            }
            case PH_INIT_B: {
                double2 gr[P], gs[P], uh[P];
                W.ld(V_GRAD, gr);
                lip = sqrt(wdiff2<P>(g, gr)) / norm_h;
                set_gamma(GAMMA_L_COEFF / fmax(lip, MIN_L_ESTIMATE));
                sigma = (1.0 - GAMMA_L_COEFF) / (4.0 * gamma);
                grad_step_half(u, gr, gs, uh);
                num_iter = 0;
                cont = true;
                phase = PH_STEP_BEGIN;
                break;
            }
            case PH_LIP: {  // psi(uhalf); OpEn then re-evaluates psi(u): needed only when u was perturbed (iteration 0)
                cost_half = psi;
                if (iteration == 0) {
#pragma unroll
                    for (int j = 0; j < P; j++) x[j] = u[j];
                    mode = MODE_COST;
                    phase = PH_COST_U;
                } else {
                    n_cost++;  // the re-evaluation OpEn performs; its value is bit-identical to the cached cost
                    phase = PH_LIP_LOOP;
                }
                break;
            }
            case PH_COST_U: {
                cost = psi;
                phase = PH_LIP_LOOP;
                break;
            }
            case PH_LIP_RETRY: {
                cost_half = psi;
                double2 uh[P], fpr[P];
                W.ld(V_UHALF, uh);
#pragma unroll
                for (int j = 0; j < P; j++) fpr[j] = make_double2(u[j].x - uh[j].x, u[j].y - uh[j].y);
                W.st(V_FPR, fpr);
                batch_dots(fpr, uh);  // gamma changed: every batched inner product is recomputed
                norm_fpr = sqrt(RES(0));
                it_lip++;
                phase = PH_LIP_LOOP;
                break;
            }
            case PH_IT0: {
                cost = psi;
                W.st(V_GRAD, g);
                double2 gs[P], uh[P];
                grad_step_half(u, g, gs, uh);
                iteration++;
                phase = PH_STEP_DONE;
                break;
            }
            case PH_LS: {
                cost = psi;
                double2 gs[P], uh[P];
                grad_step_half(x, g, gs, uh);
                double d2, gg;
                {
                    double e[P], f[P];
#pragma unroll
                    for (int j = 0; j < P; j++) {
                        double d0 = gs[j].x - uh[j].x, d1 = gs[j].y - uh[j].y;
                        e[j] = fma(d1, d1, d0 * d0);
                        f[j] = fma(g[j].y, g[j].y, g[j].x * g[j].x);
                    }
                    hsum2<P>(e, f, d2, gg);
                }
                const double lhs = cost - (0.5 * gamma) * gg + (0.5 * d2) * inv_gamma;
                if (lhs > rhs_ls && nls < MAX_LINESEARCH_ITERATIONS) {
                    tau /= 2.0;
                    nls++;
                    const double om = 1.0 - tau;
                    double2 fpr[P], dir[P];
                    W.ld(V_FPR, fpr);
                    W.ld(V_DIR, dir);
#pragma unroll
                    for (int j = 0; j < P; j++) {
                        x[j].x = fma(-tau, dir[j].x, fma(-om, fpr[j].x, u[j].x));
                        x[j].y = fma(-tau, dir[j].y, fma(-om, fpr[j].y, u[j].y));
                    }
                    mode = MODE_GRAD;
                    phase = PH_LS;
                } else {
                    W.st(V_GRAD, g);
#pragma unroll
                    for (int j = 0; j < P; j++) u[j] = x[j];
                    iteration++;
                    phase = PH_STEP_DONE;
                }
                break;
            }
            case PH_F2: {  // multipliers y+ = y + c*(F1 - Proj_C(F1 + y/c)); infeasibilities; outer-loop logic
                double2 yp[P];
                double e[P];
#pragma unroll
                for (int j = 0; j < P; j++) {
                    double vp, wp_;
                    W.prev_controls(u, j, vp, wp_);
                    const double wa = (u[j].x - vp) * inv_ts, ww = (u[j].y - wp_) * inv_ts;
                    double za = wa + yl[j].x / pn.c, zw = ww + yl[j].y / pn.c;
                    za = clampd(za, cfg.lin_acc_min, cfg.lin_acc_max);
                    zw = clampd(zw, -cfg.ang_acc_max, cfg.ang_acc_max);
                    yp[j].x = W.act[j] ? fma(pn.c, wa - za, yl[j].x) : 0.0;
                    yp[j].y = W.act[j] ? fma(pn.c, ww - zw, yl[j].y) : 0.0;
                    double d0 = yp[j].x - yl[j].x, d1 = yp[j].y - yl[j].y;
                    e[j] = W.act[j] ? fma(d1, d1, d0 * d0) : 0.0;
                }
                dynp = sqrt(hsum<P>(e));
                f2np = sqrt(pen);
                const bool crit1 = alm_iter > 0 && dynp <= pn.c * cfg.delta_tolerance + DBL_EPS;
                const bool crit2 = (nf2 == 0) || f2np <= cfg.delta_tolerance + DBL_EPS;
                const bool crit3 = akkt_tol <= cfg.tolerance + DBL_EPS;
                bool finished = crit1 && crit2 && crit3;
                if (!finished) {
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
                    if (num_outer >= cfg.max_outer_iterations) {
                        status = NMPC_NOT_CONVERGED_ITERATIONS;
                        finished = true;
                    }
                } else if (num_outer == cfg.max_outer_iterations) {
                    status = NMPC_NOT_CONVERGED_ITERATIONS;
                }
                if (finished) {
#pragma unroll
                    for (int j = 0; j < P; j++) x[j] = u[j];
                    mode = MODE_COST;
                    phase = PH_FINAL;
                } else {
                    phase = PH_OUTER_BEGIN;
                }
                break;
            }
            case PH_FINAL: {
                st_out.cost = psi;
                phase = PH_EXIT;
                break;
            }
            default:
                break;
        }
    }
}
