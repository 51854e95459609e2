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
#define PROBE_BUCKETS 1024
// Code size matters as much as instruction count here: the per-iteration hot loop of the solver is about the
// size of the SM's 32 KB L1.5 instruction cache, and a loop that no longer fits misses on every line (measured:
// a lone warp's two-loop recursion slows from 5.7k to 7.9k cycles when the evaluation code grows by 15 %).
// NMPC_UNROLL_LOOPS=1 lets ptxas unroll the latency-bound loops again (tools/variants.py).
#ifndef NMPC_UNROLL_LOOPS
#define NMPC_UNROLL_LOOPS 0
#endif
#if NMPC_UNROLL_LOOPS
#define NMPC_NOUNROLL
#else
#define NMPC_NOUNROLL _Pragma("unroll 1")
#endif
// the five exchange stages of every warp reduction / scan: unrolled (1) or a real loop (0, smaller code)
#ifndef NMPC_STAGE_UNROLL
#define NMPC_STAGE_UNROLL 1
#endif
#if NMPC_STAGE_UNROLL
#define NMPC_STAGES _Pragma("unroll")
#else
#define NMPC_STAGES _Pragma("unroll 1")
#endif
#ifndef NMPC_ICLAMP
#define NMPC_ICLAMP 0  // measured neutral on B200 (round 1); the compare-select form is the oracle's
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
#ifndef NMPC_HELP_COST
#define NMPC_HELP_COST 1  // also hand psi(uhalf) (Lipschitz test) to a helper and run the two-loop recursion meanwhile
#endif
#ifndef NMPC_HELP_PART_MAX
#define NMPC_HELP_PART_MAX 0  // helpers work from SM sub-partitions with at most this many owners
#endif
#ifndef NMPC_HELP_COST_MAXLIVE
#define NMPC_HELP_COST_MAXLIVE 12
#endif
#ifndef NMPC_HELP_EXTRA
#define NMPC_HELP_EXTRA 9  // trials offered beyond what the previous search needed (9 = always all NMPC_HELP_R)
#endif
#ifndef NMPC_HELP_SLEEP
#define NMPC_HELP_SLEEP 100  // ns between two polls of an idle helper
#endif
// Speculative line search on idle warps (experiments/README.md has the measurements): once the problem queue is
// empty, warps without a problem evaluate the next line-search trials of the warps that still have one.  Bit-exact
// (who evaluates a trial never changes its bits); on the final round-1 kernel +5 % on the B=4096 batch (the step
// lasts as long as its hardest problem) and +1 % on a saturated batch.  -DNMPC_HELP_R=0 compiles it out.
#ifndef NMPC_HELP_R
#define NMPC_HELP_R 4  // line-search trials a problem may have in flight on idle warps of its CTA (0 = feature off)
#endif
enum { V_GRAD = 0, V_UHALF, V_FPR, V_DIR, V_GSTEP, V_OLDS, V_OLDG, V_S, V_Y = V_S + MEMP1,
       V_U = V_Y + MEMP1, V_YL, V_JG, V_END = V_JG + NMPC_HELP_R };  // V_U, V_YL, V_JG*: what a helper warp reads / writes
enum { H_X0 = 0, H_Y0, H_TH0, H_VINIT, H_WINIT, H_XREF, H_YREF, H_THREF, H_Q, H_QV, H_QTH, H_RV, H_RW, H_QN, H_QTHN,
       H_QCTE, H_AP, H_WP, H_INVTS,
       // warp-uniform solver state that is touched once per outer iteration (kept out of the registers)
       H_F2N, H_DYN, H_F2NP, H_DYNP, H_NORMH, H_LIP, H_AKKT, H_NCIRC /* int */, H_COUNT = 28 };
// job record of one speculative line-search trial (bytes): state, trial, seq (ints) | gamma | c | psi | lhs
#define JOB_BYTES 64u
enum { JOB_EMPTY = 0, JOB_POSTED = 1, JOB_TAKEN = 2, JOB_DONE = 3 };
#define SEG_STRIDE 6   // s1x s1y | dx dy | inv pad
#define SEG_PAD 3      // copies of the last segment behind the table: the cross-track loop needs no remainder trips
#define CIRC_STRIDE 4  // cx cy | r2 (original slot index as int in the 4th double)
#define ELL_STRIDE 6   // ex ey | cosA sinA | 1/rx^2 1/ry^2

struct Lay {
    int n2, seg, circ, ell, rho, alpha, hdr, vref, job, total;
};
__host__ __device__ inline int even_up(int x) { return (x + 1) & ~1; }
__host__ __device__ inline Lay make_layout(int N, int Nobs, int Nd) {
    Lay L;
    L.n2 = 2 * N;
    int o = V_END * 2 * N;
    L.seg = o; o += SEG_STRIDE * (N + SEG_PAD + 1);
    L.circ = o; o += CIRC_STRIDE * Nobs;
    L.ell = o; o += ELL_STRIDE * Nd * N;
    L.rho = o; o += 12;
    L.alpha = o; o += 12;
    L.hdr = o; o += H_COUNT;
    L.vref = o; o += even_up(N);
    L.job = o; o += (NMPC_HELP_R + 1) * (int)(JOB_BYTES / 8);  // line-search trials + one cost-evaluation record
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
    const int32_t* skip;  // nullable: rows with skip[b] != 0 are left untouched (fleet: robots that have terminated)
    const int32_t* order; // nullable: permutation of 0..B-1, the order in which problems are handed out
    int32_t* probe_bucket; // probe kernel: sort bucket of every problem
    int32_t* probe_hist;   // probe kernel: bucket histogram (PROBE_BUCKETS ints)
    // eval kernel only
    const double* cvec;
    double *psi, *grad, *F1, *F2;
    long long* dbg;  // NMPC_PROFILE builds only: 16 cycle counters per problem
};
#ifdef NMPC_PROFILE
#define PROF_BEGIN() long long plast_ = clock64()
#define PROF_MARK(i) do { const long long t_ = clock64(); pt[i] += t_ - plast_; plast_ = t_; } while (0)
#else
#define PROF_BEGIN() do { } while (0)
#define PROF_MARK(i) do { } while (0)
#endif

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
// predicated forms (one instruction each, no branch): active lanes only
__device__ __forceinline__ double2 lds2_if(uint32_t a, bool on) {
    double2 v;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\tmov.f64 %1, 0d0000000000000000;\n\t"
                 "@p ld.shared.v2.f64 {%0, %1}, [%2];\n\t}"
                 : "=d"(v.x), "=d"(v.y)
                 : "r"(a), "r"((int)on));
    return v;
}
__device__ __forceinline__ void sts2_if(uint32_t a, double2 v, bool on) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\t@p st.shared.v2.f64 [%0], {%1, %2};\n\t}" ::"r"(a), "d"(v.x), "d"(v.y),
                 "r"((int)on)
                 : "memory");
}
__device__ __forceinline__ void stsi(uint32_t a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// volatile / atomic access to the job words shared between warps of one CTA
__device__ __forceinline__ int ldv_shared(uint32_t a) {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void stv_shared(uint32_t a, int v) { asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ int cas_shared(uint32_t a, int cmp, int val) {
    int old;
    asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(a), "r"(cmp), "r"(val) : "memory");
    return old;
}
__device__ __forceinline__ int add_shared(uint32_t a, int val) {
    int old;
    asm volatile("atom.shared.add.s32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(val) : "memory");
    return old;
}

// Rectangle::project of OpEn is comparison-based: a NaN stays a NaN (and ends the solve as NotFinite)
__device__ __forceinline__ double clampd(double x, double lo, double hi) { return (x < lo) ? lo : ((x > hi) ? hi : x); }
// min/max as compare-selects (same forms as the oracle): NaN -> the constant, zero results are +0
// (written as setp/selp PTX: the C ternaries get canonicalised to max.f64/min.f64, which sm_100
//  expands into a ~12-instruction DSETP.MAX/FSEL/SEL/NaN-fix-up sequence each)
__device__ __forceinline__ double sel_clamp01(double t) {
#if NMPC_ICLAMP
    // Same result as the two compare-selects for every non-NaN t, computed on the integer pipe from the
    // high word (sign and exponent order doubles like signed integers for t >= 0): hi' = min(max(hi, 0), hi(1.0)),
    // lo' = lo only while 0 <= hi < hi(1.0).  Two FP64-pipe compares and four selects become four ALU ops.
    // (NaN: the selects give 0, this gives 0 or 1 by the sign bit; either way the distance stays NaN because
    //  a NaN projection parameter comes from a NaN point, which is already in ex/ey.)
    double r;
    asm("{\n\t.reg .b32 lo, hi, h2;\n\t.reg .pred p;\n\t"
        "mov.b64 {lo, hi}, %1;\n\t"
        "max.s32 h2, hi, 0;\n\t"
        "min.s32 h2, h2, 0x3FF00000;\n\t"
        "setp.lt.u32 p, hi, 0x3FF00000;\n\t"
        "selp.b32 lo, lo, 0, p;\n\t"
        "mov.b64 %0, {lo, h2};\n\t}"
        : "=d"(r)
        : "d"(t));
    return r;
#else
    double r;
    asm("{\n\t.reg .pred p;\n\t"
        "setp.gt.f64 p, %1, 0d0000000000000000;\n\tselp.f64 %0, %1, 0d0000000000000000, p;\n\t"
        "setp.lt.f64 p, %0, 0d3FF0000000000000;\n\tselp.f64 %0, %0, 0d3FF0000000000000, p;\n\t}"
        : "=d"(r)
        : "d"(t));
    return r;
#endif
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
    // out-of-range / non-finite arguments give NaN (as in the oracle), without a branch: the reduction runs on 0
    const bool bad = !(fabs(x) < 1.0e8);
    x = bad ? 0.0 : x;
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
    s = bad ? CUDART_NAN : s;
    c = bad ? CUDART_NAN : c;
}

// ---------------------------------------------------------------------------------
// warp-ordered reductions
__device__ __forceinline__ double butterfly(double a) {
    NMPC_STAGES
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
    NMPC_STAGES
    for (int off = 16; off; off >>= 1) {
        double ya = __shfl_xor_sync(FULL, a, off), yb = __shfl_xor_sync(FULL, b, off);
        a = a + ya;
        b = b + yb;
    }
    se = a;
    sf = b;
}
// four sums at once
template <int P>
__device__ __forceinline__ void hsum4(const double (&e0)[P], const double (&e1)[P], const double (&e2)[P],
                                      const double (&e3)[P], double& s0, double& s1, double& s2, double& s3) {
    double a = e0[0], b = e1[0], c = e2[0], d = e3[0];
#pragma unroll
    for (int j = 1; j < P; j++) {
        a = a + e0[j];
        b = b + e1[j];
        c = c + e2[j];
        d = d + e3[j];
    }
    NMPC_STAGES
    for (int off = 16; off; off >>= 1) {
        double ya = __shfl_xor_sync(FULL, a, off), yb = __shfl_xor_sync(FULL, b, off);
        double yc = __shfl_xor_sync(FULL, c, off), yd = __shfl_xor_sync(FULL, d, off);
        a = a + ya;
        b = b + yb;
        c = c + yc;
        d = d + yd;
    }
    s0 = a;
    s1 = b;
    s2 = c;
    s3 = d;
}
template <int P>
__device__ __forceinline__ void prefix_scan(const double (&x)[P], double (&incl)[P], double (&excl)[P], int lane) {
    double carry = 0.0;
#pragma unroll
    for (int j = 0; j < P; j++) {
        double l = x[j];
        NMPC_STAGES
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
        NMPC_STAGES
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
        NMPC_STAGES
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
        NMPC_STAGES
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
// suffix_scan2 with an independent xor-butterfly sum riding along in the same instruction stream: the stages of
// the two reductions interleave, so the cost sum of a gradient evaluation costs no extra latency.
// Each of the three results is bit-identical to suffix_scan2 / hsum.
template <int P>
__device__ __forceinline__ void suffix_scan2_hsum(const double (&xa)[P], const double (&xb)[P], double (&sa)[P],
                                                  double (&sb)[P], const double (&e)[P], double& esum, int lane) {
    double acc = e[0];
#pragma unroll
    for (int j = 1; j < P; j++) acc = acc + e[j];
    double ca = 0.0, cb = 0.0;
#pragma unroll
    for (int j = P - 1; j >= 0; j--) {
        double la = xa[j], lb = xb[j];
        NMPC_STAGES
        for (int st = 0; st < 5; st++) {
            const int off = 1 << st;
            double ya = __shfl_down_sync(FULL, la, off);
            double yb = __shfl_down_sync(FULL, lb, off);
            double yc = 0.0;
            if (j == P - 1) yc = __shfl_xor_sync(FULL, acc, 16 >> st);
            add_if(la, ya, lane + off < 32);
            add_if(lb, yb, lane + off < 32);
            if (j == P - 1) acc = acc + yc;
        }
        double ga = (j == P - 1) ? la : ca + la, gb = (j == P - 1) ? lb : cb + lb;
        sa[j] = ga;
        sb[j] = gb;
        if (j > 0) {
            ca = __shfl_sync(FULL, ga, 0);
            cb = __shfl_sync(FULL, gb, 0);
        }
    }
    esum = acc;
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
// NF > 0: the horizon is a compile-time constant (NF == cfg.N_hor, checked by the host): the cross-track
// loop is fully unrolled with a tree arg-min, so a lone warp (the tail of a small batch) gets ILP.
template <int P, int NF = 0>
struct Warp {
    const nmpc_config& cfg;
    uint32_t sb;         // shared byte address of the arena
    uint32_t la[P];      // sb + 16*t : this lane's element inside vector 0
    uint32_t vstride;    // bytes per vector (2N doubles)
    uint32_t a_seg, a_circ, a_ell, a_rho, a_alpha, a_hdr, a_vref, a_job;
    uint32_t sb0, arena_bytes;  // arena of warp 0 / bytes per arena (retarget)
    int lane, n_circ;    // n_circ: circles with r != 0 (zero-padded slots are skipped: they add exact zeros)
    bool act[P];
    int tix[P];
#ifdef NMPC_PROFILE
    long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // cycles per eval section (tools/prof_cycles.py)
#endif

    __device__ __forceinline__ Warp(const nmpc_config& c, const Lay& L, int warp, int lane_) : cfg(c), lane(lane_) {
        sb0 = (uint32_t)__cvta_generic_to_shared(smem);
        arena_bytes = (uint32_t)L.total * 8u;
        sb = sb0 + (uint32_t)warp * arena_bytes;
        vstride = (uint32_t)L.n2 * 8u;
        a_seg = sb + L.seg * 8u; a_circ = sb + L.circ * 8u; a_ell = sb + L.ell * 8u; a_rho = sb + L.rho * 8u;
        a_alpha = sb + L.alpha * 8u; a_hdr = sb + L.hdr * 8u; a_vref = sb + L.vref * 8u; a_job = sb + L.job * 8u;
        n_circ = 0;
#pragma unroll
        for (int j = 0; j < P; j++) {
            tix[j] = lane + 32 * j;
            act[j] = tix[j] < (NF ? NF : cfg.N_hor);
            la[j] = sb + 16u * tix[j];
        }
    }
    // point this view at the arena of warp `o` of the CTA (a helper warp evaluates another warp's problem there)
    __device__ __forceinline__ void retarget(int o) {
        const uint32_t nsb = sb0 + (uint32_t)o * arena_bytes, d = nsb - sb;
        sb = nsb;
        a_seg += d; a_circ += d; a_ell += d; a_rho += d; a_alpha += d; a_hdr += d; a_vref += d; a_job += d;
#pragma unroll
        for (int j = 0; j < P; j++) la[j] += d;
        n_circ = ldsi(a_hdr + 8u * H_NCIRC);
    }
    __device__ __forceinline__ double hdr(int i) const { return lds1(a_hdr + 8u * i); }
    __device__ __forceinline__ void ld(int k, double2 (&r)[P]) const {
#pragma unroll
        for (int j = 0; j < P; j++) r[j] = lds2_if(la[j] + k * vstride, act[j]);
    }
    __device__ __forceinline__ void st(int k, const double2 (&r)[P]) const {
#pragma unroll
        for (int j = 0; j < P; j++) sts2_if(la[j] + k * vstride, r[j], act[j]);
    }

    // unpack the parameter row (layout: include/nmpc_b200.h) into the arena
    __device__ void stage(const double* __restrict__ p) {
        const int N = NF ? NF : cfg.N_hor, Nobs = cfg.Nobs, Nd = cfg.Ndynobs;
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
        if (lane == 0) stsi(a_hdr + 8u * H_NCIRC, nreal);
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
        for (int ii = lane; ii < N + SEG_PAD; ii += 32) {
            if (ii >= 1) {
                // slots N .. N+SEG_PAD-1 repeat segment N-1: equal distances never win the strict '<' arg-min
                const int i = (ii < N) ? ii : N - 1;
                double ax = pr[3 * (i - 1)], ay = pr[3 * (i - 1) + 1];
                double dx = pr[3 * i] - ax, dy = pr[3 * i + 1] - ay;
                const uint32_t a = a_seg + 48u * ii;
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
        const int N = NF ? NF : cfg.N_hor;
        const double ts = cfg.ts;
        PROF_BEGIN();
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
        PROF_MARK(0);
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
        PROF_MARK(1);

        if (mode != MODE_F2) {
            // cross-track error: each lane scans the N-1 segments for its own predicted point
            double best[P];
            int bi[P];
#pragma unroll
            for (int j = 0; j < P; j++) {
                best[j] = CUDART_INF;
                bi[j] = 1;
            }
            if constexpr (NF > 0 && P == 1) {
                // all NF-1 segments are independent; arg-min by a tree whose left operand holds the lower
                // indices and wins ties (= the first minimal segment, as the serial strict-'<' scan)
                double dv[NF - 1];
                int iv[NF - 1];
#pragma unroll
                for (int i = 1; i < NF; i++) {
                    const uint32_t as = a_seg + 48u * i;
                    const double2 s1 = lds2(as), d = lds2(as + 16u);
                    const double inv = lds1(as + 32u);
                    double px = X[0] - s1.x, py = Y[0] - s1.y;
                    double that = fma(px, d.x, py * d.y) * inv;
                    double tst = sel_clamp01(that);
                    double ex = fma(tst, d.x, -px), ey = fma(tst, d.y, -py);
                    dv[i - 1] = fma(ex, ex, ey * ey);
                    iv[i - 1] = i;
                }
#pragma unroll
                for (int st = 1; st < NF - 1; st *= 2)
#pragma unroll
                    for (int k = 0; k + st < NF - 1; k += 2 * st) take_if_less(dv[k + st], iv[k + st], dv[k], iv[k]);
                take_if_less(dv[0], iv[0], best[0], bi[0]);
            } else {
            constexpr int UNR = (P == 1) ? 4 : 2;
            uint32_t as = a_seg + 48u;
            int i = 1;
            // UNR segments per trip, written stage by stage: the SM issues in order, so independent
            // chains only overlap if they are interleaved in the instruction stream
            NMPC_NOUNROLL
            for (; i < N; i += UNR, as += 48u * UNR) {
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
                    int iq[UNR];
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
                    for (int q = 0; q < UNR; q++) iq[q] = i + q;
                    // arg-min of the trip as a tree (the left operand holds the lower indices and wins ties, like the
                    // serial strict-'<' scan), then ONE merge into the running minimum: the chain between trips is short
#pragma unroll
                    for (int st = 1; st < UNR; st *= 2)
#pragma unroll
                        for (int q = 0; q + st < UNR; q += 2 * st) take_if_less(d2[q + st], iq[q + st], d2[q], iq[q]);
                    take_if_less(d2[0], iq[0], best[j], bi[j]);
                }
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

        PROF_MARK(2);
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
        PROF_MARK(3);
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
        if (!GRAD) {
            const double psi0 = fma(pn.hc, pen, hsum<P>(cl) + term);
            PROF_MARK(4);
            return psi0;
        }

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
        double LX[P], LY[P], csum;
        suffix_scan2_hsum<P>(gX, gY, LX, LY, cl, csum, lane);  // position adjoints + the cost sum
        const double psi = fma(pn.hc, pen, csum + term);
        PROF_MARK(4);
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
        PROF_MARK(5);
        return psi;
    }
};

// ---------------------------------------------------------------------------------
// The solver: ALM/PM outer loop around PANOC as a phase machine with one evaluation site.
// Phases that end in an evaluation set (x, mode) and fall through to it; the others `continue`.
enum Phase {
    PH_OUTER_BEGIN, PH_INIT_A, PH_INIT_B, PH_STEP_BEGIN, PH_LIP, PH_COST_U, PH_LIP_LOOP, PH_LIP_RETRY, PH_IT0, PH_LS,
    PH_STEP_DONE, PH_SOLVE_END, PH_F2, PH_FINAL, PH_EXIT, PH_HELP_WAIT, PH_HELP_EVAL
};

// Owner mode (helper == false) solves the staged problem.  Helper mode (entered once the problem queue is empty)
// serves the other warps of the CTA: it polls their job records for posted line-search trials
// x = u - (1-tau) fpr - tau dir, evaluates them in the owner's arena through the same evaluation site and writes
// back psi, the gradient and the trial's envelope value; it returns when no warp of the CTA owns a problem any
// more.  Who evaluates a trial never changes its bits, so results do not depend on timing.
// HC (latency mode, chosen by the host for batches of at most two problems per SM): psi(uhalf) of every
// iteration is also handed to a helper while the owner runs the L-BFGS update and the two-loop recursion: a lone
// problem's iteration drops from 29.5k to 23k cycles, but the extra code costs 3-6 % on full batches, hence two
// instantiations.
template <int P, int NF, bool HC>
__device__ int solve_problem(Warp<P, NF>& W, double2 (&u)[P], double2 (&yl)[P], nmpc_stats& st_out, const bool helper,
                             const uint32_t a_live, const int nwarps, long long* prof_out = nullptr) {
#ifdef NMPC_PROFILE
    long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long tstart = clock64();
#endif
    const nmpc_config& cfg = W.cfg;
    const int lane = W.lane;
    const int mem = cfg.lbfgs_memory, mem1 = cfg.lbfgs_memory + 1;
    const int nf2 = cfg.Nobs + cfg.Ndynobs;
    // warp-uniform state
    double gamma = 0.0, inv_gamma = 0.0, sigma = 0.0, cost = 0.0, norm_fpr = 0.0, tau = 1.0;
    double cost_half = 0.0, rhs_ls = 0.0, lb_gamma = 1.0;
    // rarely touched warp-uniform scalars live in the arena header: every lane stores the same value and
    // reads back its own store, so no synchronisation is involved
    auto sget = [&](int i) { return lds1(W.a_hdr + 8u * i); };
    auto sput = [&](int i, double v) { sts1(W.a_hdr + 8u * i, v); };
    sput(H_AKKT, cfg.initial_tolerance);
    sput(H_F2N, 0.0);
    sput(H_DYN, 0.0);
    sput(H_F2NP, 0.0);
    sput(H_DYNP, 0.0);
    sput(H_LIP, 0.0);
    Pen pn = make_pen(cfg.initial_penalty);
    Pen pn_eval = pn;
    int iteration = 0, n_cost = 0, n_grad = 0, lb_active = 0, lb_first = 1, lb_head = 0;
    int alm_iter = 0, inner_total = 0, num_outer = 0, status = NMPC_CONVERGED, inner_status = NMPC_CONVERGED;
    int num_iter = 0, it_lip = 0, nls = 0;
    bool cont = true, fbe_valid = false;
    double fbe_u = 0.0;
    const double inv_ts = W.hdr(H_INVTS);
    bool cost_pending = false;  // psi(uhalf) of this iteration is being evaluated by a helper warp
    bool spec_done = false;     // L-BFGS update and direction of this iteration are already done (speculatively)
    int ls_hint = 0;  // trials the previous line search needed beyond tau = 1
    int ls_seq = 0;   // owner: line searches started (tags job records); helper: slot r of the job being served

    double2 x[P], g[P];  // evaluation point / gradient out
    double pen = 0.0;
    int mode = MODE_GRAD;
    int phase = helper ? PH_HELP_WAIT : PH_OUTER_BEGIN;

    auto set_gamma = [&](double gm) {  // sigma only changes with gamma: computed here, not once per iteration
        gamma = gm;
        inv_gamma = 1.0 / gm;
        sigma = (1.0 - GAMMA_L_COEFF) / (4.0 * gm);
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

#ifdef NMPC_PROFILE
    long long last_t = clock64();
    int last_slot = -1;  // per-phase cycles outside the evaluations: slots 16+phase (pre), 32+phase (post)
#define PH_ACCOUNT(next_slot)                                                                                  \
    do {                                                                                                       \
        const long long now_ = clock64();                                                                      \
        if (last_slot >= 0 && prof_out && lane == 0)                                                           \
            atomicAdd((unsigned long long*)&prof_out[last_slot], (unsigned long long)(now_ - last_t));        \
        last_t = now_;                                                                                         \
        last_slot = (next_slot);                                                                               \
    } while (0)
#else
#define PH_ACCOUNT(next_slot) do { } while (0)
#endif
    for (;;) {
        PH_ACCOUNT(16 + phase);
        // ------------------------------------------------------------------ pre: pick (x, mode)
        switch (phase) {
            case PH_OUTER_BEGIN: {
                num_outer++;
#pragma unroll
                for (int j = 0; j < P; j++) {  // project_on_set_y
                    yl[j].x = clampd(yl[j].x, -Y_SET_BOUND, Y_SET_BOUND);
                    yl[j].y = clampd(yl[j].y, -Y_SET_BOUND, Y_SET_BOUND);
                }
                if (NMPC_HELP_R > 0) W.st(V_YL, yl);  // helper warps read the multipliers from the arena
                // panoc init
                lb_active = 0;
                lb_first = 1;
                fbe_valid = false;
                tau = 1.0;
                iteration = 0;
#pragma unroll
                for (int j = 0; j < P; j++) x[j] = u[j];
                mode = MODE_GRAD;
                phase = PH_INIT_A;
                break;
            }
            case PH_STEP_BEGIN: {
                double2 gr[P], uh[P], fpr[P];
                W.ld(V_GRAD, gr);
                W.ld(V_UHALF, uh);
                compute_fpr(uh, fpr);
                bool exit_now = false;
                if (norm_fpr < cfg.tolerance) {
                    double e[P];
#pragma unroll
                    for (int j = 0; j < P; j++) {
                        double p0 = iteration ? gr[j].x : 0.0, p1 = iteration ? gr[j].y : 0.0;
                        double r0 = fma(fpr[j].x, inv_gamma, gr[j].x) - p0;
                        double r1 = fma(fpr[j].y, inv_gamma, gr[j].y) - p1;
                        e[j] = fma(r1, r1, r0 * r0);
                    }
                    exit_now = sqrt(hsum<P>(e)) < sget(H_AKKT);
                }
                if (exit_now) {
                    phase = PH_SOLVE_END;
                    continue;
                }
                W.st(V_FPR, fpr);
                it_lip = 0;
#if NMPC_HELP_R > 0
                if constexpr (HC) {
                // With an idle sub-partition in the CTA, psi(uhalf) (only needed for the Lipschitz test) goes to a
                // helper warp while this warp already updates the L-BFGS memory and runs the two-loop recursion.
                // If the test then fails (rare), the update is discarded exactly as the reference discards its memory.
                // Only in the deep tail (few owners left in the CTA): otherwise the cost jobs take helper time from the
                // line-search trials of the other owners, which are worth more.
                if (__builtin_expect(iteration > 0 && ldv_shared(a_live) <= NMPC_HELP_COST_MAXLIVE, 0) &&
                    __any_sync(FULL, lane < 4 && lane < nwarps && ldv_shared(a_live + 4u + 4u * lane) <= NMPC_HELP_PART_MAX)) {
                    const uint32_t aj = W.a_job + JOB_BYTES * NMPC_HELP_R;
                    int posted = 0;
                    __threadfence_block();
                    __syncwarp();
                    if (lane == 0) {
                        const int stt = ldv_shared(aj);
                        if (stt == JOB_EMPTY || stt == JOB_DONE) {
                            stsi(aj + 4u, 0);  // trial 0 = cost evaluation at uhalf
                            stsi(aj + 8u, ls_seq);
                            sts1(aj + 16u, gamma);
                            sts1(aj + 24u, pn.c);
                            __threadfence_block();
                            stv_shared(aj, JOB_POSTED);
                            posted = 1;
                        }
                    }
                    if (__shfl_sync(FULL, posted, 0)) {
                        cost_pending = true;
                        phase = PH_LIP_LOOP;
                        continue;
                    }
                }
                }
#endif
#pragma unroll
                for (int j = 0; j < P; j++) x[j] = uh[j];
                mode = MODE_COST;
                phase = PH_LIP;
                break;
            }
            case PH_LIP_LOOP: {
                double2 gr[P], fpr[P], s[P], y[P];
                W.ld(V_GRAD, gr);
                W.ld(V_FPR, fpr);
                // <grad, fpr> for the Lipschitz test and, when a previous (state, fpr) pair exists, the three
                // inner products of the L-BFGS update (s.y, s.s, y.y) in ONE interleaved butterfly
                double ip, ys = 0.0, ss = 0.0, yy = 0.0;
                if (lb_first) {
                    ip = wdot<P>(gr, fpr);
                } else {
                    double2 os[P], og[P];
                    W.ld(V_OLDS, os);
                    W.ld(V_OLDG, og);
                    double e0[P], e1[P], e2[P], e3[P];
#pragma unroll
                    for (int j = 0; j < P; j++) {
                        s[j] = make_double2(u[j].x - os[j].x, u[j].y - os[j].y);
                        y[j] = make_double2(fpr[j].x - og[j].x, fpr[j].y - og[j].y);
                        e0[j] = fma(gr[j].y, fpr[j].y, gr[j].x * fpr[j].x);
                        e1[j] = fma(s[j].y, y[j].y, s[j].x * y[j].x);
                        e2[j] = fma(s[j].y, s[j].y, s[j].x * s[j].x);
                        e3[j] = fma(y[j].y, y[j].y, y[j].x * y[j].x);
                    }
                    hsum4<P>(e0, e1, e2, e3, ip, ys, ss, yy);
                }
                // Lipschitz test of PANOC's step size; on failure: halve gamma, drop the L-BFGS memory, re-evaluate
                auto lip_test_fails = [&]() -> bool {
                    const double rhs = cost + LIPSCHITZ_UPDATE_EPSILON * fabs(cost) - ip +
                                       (GAMMA_L_COEFF * 0.5 * inv_gamma) * (norm_fpr * norm_fpr);
                    if (!(cost_half > rhs && it_lip < MAX_LIPSCHITZ_UPDATE_ITERATIONS && sget(H_LIP) < MAX_LIPSCHITZ_CONSTANT))
                        return false;
                    lb_active = 0;
                    lb_first = 1;
                    fbe_valid = false;
                    spec_done = false;
                    sput(H_LIP, sget(H_LIP) * 2.0);
                    set_gamma(gamma / 2.0);
                    double2 gs[P], uh[P];
                    grad_step_half(u, gr, gs, uh);
#pragma unroll
                    for (int j = 0; j < P; j++) x[j] = uh[j];
                    mode = MODE_COST;
                    phase = PH_LIP_RETRY;
                    return true;
                };
                if (!(HC && __builtin_expect(cost_pending, 0)) && lip_test_fails()) break;
                double2 q[P];
                if (!(HC && __builtin_expect(spec_done, 0))) {
                // lbfgs_direction(): update_hessian(g = fpr, state = u)
                if (lb_first) {
                    lb_first = 0;
                    W.st(V_OLDS, u);
                    W.st(V_OLDG, fpr);
                } else {
                    const int tmp = slot(mem);
                    W.st(V_S + tmp, s);
                    W.st(V_Y + tmp, y);
                    const double rho_new = 1.0 / ys;
                    bool accept = !(ss <= DBL_EPS || ys <= SY_EPSILON);
                    if (accept) {
                        // sqrt(<fpr, fpr>) is norm_fpr: same vector, same expression, same reduction order
                        const double lhs = ys / ss, rhsb = CBFGS_EPSILON * norm_fpr;
                        accept = (lhs > rhsb && isfinite(lhs) && isfinite(rhsb));
                    }
                    if (accept) {
                        W.st(V_OLDS, u);
                        W.st(V_OLDG, fpr);
                        if (lane == 0) sts1(W.a_rho + 8u * tmp, rho_new);
                        lb_head = (lb_head + mem >= mem1) ? lb_head + mem - mem1 : lb_head + mem;
                        lb_gamma = (1.0 / rho_new) / yy;
                        lb_active = (lb_active + 1 < mem) ? lb_active + 1 : mem;
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
                // direction = H * fpr (two-loop recursion)
#ifdef NMPC_PROFILE
                const long long tl0 = clock64();
#endif
#pragma unroll
                for (int j = 0; j < P; j++) q[j] = fpr[j];
                if (lb_active > 0) {
                    // each step's (s, y) pair is fetched while the previous step's butterfly is in flight
                    double2 sv[P], yv[P], sn[P], yn[P];
                    int sl = slot(0);
                    W.ld(V_S + sl, sv);
                    W.ld(V_Y + sl, yv);
                    double rho = lds1(W.a_rho + 8u * sl);
                    NMPC_NOUNROLL
                    for (int k = 0; k < lb_active; k++) {
                        const int sl_n = slot((k + 1 < lb_active) ? k + 1 : k);
                        W.ld(V_S + sl_n, sn);
                        W.ld(V_Y + sl_n, yn);
                        const double rho_n = lds1(W.a_rho + 8u * sl_n);
                        const double al = rho * wdot<P>(sv, q);
                        if (lane == 0) sts1(W.a_alpha + 8u * k, al);
#pragma unroll
                        for (int j = 0; j < P; j++) {
                            q[j].x = fma(-al, yv[j].x, q[j].x);
                            q[j].y = fma(-al, yv[j].y, q[j].y);
                            sv[j] = sn[j];
                            yv[j] = yn[j];
                        }
                        rho = rho_n;
                    }
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < P; j++) {
                        q[j].x = q[j].x * lb_gamma;
                        q[j].y = q[j].y * lb_gamma;
                    }
                    // (sv, yv, rho) now hold the newest pair (k = lb_active-1): the backward loop starts there
                    NMPC_NOUNROLL
                    for (int k = lb_active - 1; k >= 0; k--) {
                        const int sl_n = slot((k > 0) ? k - 1 : 0);
                        W.ld(V_S + sl_n, sn);
                        W.ld(V_Y + sl_n, yn);
                        const double rho_n = lds1(W.a_rho + 8u * sl_n);
                        const double alk = lds1(W.a_alpha + 8u * k);
                        const double beta = rho * wdot<P>(yv, q);
                        const double co = alk - beta;
#pragma unroll
                        for (int j = 0; j < P; j++) {
                            q[j].x = fma(co, sv[j].x, q[j].x);
                            q[j].y = fma(co, sv[j].y, q[j].y);
                            sv[j] = sn[j];
                            yv[j] = yn[j];
                        }
                        rho = rho_n;
                    }
                }
                W.st(V_DIR, q);
#ifdef NMPC_PROFILE
                prof[4] += clock64() - tl0;
                prof[5]++;
#endif
                } else {  // update and direction were done before psi(uhalf) was known
                    W.ld(V_DIR, q);
                    spec_done = false;
                }
#if NMPC_HELP_R > 0
                if constexpr (HC) if (__builtin_expect(cost_pending, 0)) {
                    cost_pending = false;
                    const uint32_t aj = W.a_job + JOB_BYTES * NMPC_HELP_R;
                    int got = 0;
                    if (lane == 0) {
                        int stt = ldv_shared(aj);
                        if (stt == JOB_POSTED && cas_shared(aj, JOB_POSTED, JOB_EMPTY) == JOB_POSTED) stt = JOB_EMPTY;
                        if (stt != JOB_EMPTY) {
                            while (ldv_shared(aj) != JOB_DONE) __nanosleep(20);
                            got = 1;
                        }
                    }
                    got = __shfl_sync(FULL, got, 0);
                    if (!got) {  // no helper picked it up: evaluate here and come back (update / direction are kept)
                        spec_done = true;
                        W.ld(V_UHALF, x);
                        mode = MODE_COST;
                        phase = PH_LIP;
                        break;
                    }
                    __threadfence_block();
                    cost_half = lds1(aj + 32u);
                    n_cost += 2;  // the evaluation and OpEn's re-evaluation of psi(u) (see PH_LIP)
                    __syncwarp();
                    if (lane == 0) stv_shared(aj, JOB_EMPTY);
                    if (lip_test_fails()) break;
                }
#endif
                // linesearch(): right-hand side on the forward-backward envelope
                if (fbe_valid) {
                    // the envelope at u is the accepted trial's left-hand side of the previous line search
                    // (same cost, gradient, gamma and stored gstep/uhalf: bit-identical), unless gamma changed
                    rhs_ls = fbe_u - sigma * (norm_fpr * norm_fpr);
                } else {
                    double2 gs[P], uh[P];
                    W.ld(V_GSTEP, gs);
                    W.ld(V_UHALF, uh);
                    double dist2, gg;
                    {
                        double e[P], f[P];
#pragma unroll
                        for (int j = 0; j < P; j++) {
                            double d0 = gs[j].x - uh[j].x, d1 = gs[j].y - uh[j].y;
                            e[j] = fma(d1, d1, d0 * d0);
                            f[j] = fma(gr[j].y, gr[j].y, gr[j].x * gr[j].x);
                        }
                        hsum2<P>(e, f, dist2, gg);
                    }
                    const double fbe = cost - (0.5 * gamma) * gg + (0.5 * dist2) * inv_gamma;
                    rhs_ls = fbe - sigma * (norm_fpr * norm_fpr);
                }
                tau = 1.0;
                nls = 0;
#if NMPC_HELP_R > 0
                ls_seq++;
                // a_live[0]: warps of the CTA in owner mode; a_live[1..4]: the same per SM sub-partition (warp % 4).
                // Helpers only work from a sub-partition without owners, so they never take issue slots or FP64
                // pipe cycles from a warp that is solving a problem.
                const bool idle_part = __any_sync(FULL, lane < 4 && lane < nwarps && ldv_shared(a_live + 4u + 4u * lane) <= NMPC_HELP_PART_MAX);
                if (idle_part) {
                    // offer the next trials (tau = 1/2, 1/4, ...) while this warp evaluates tau = 1.
                    // A record still held by a late helper of an earlier search is skipped.
                    W.st(V_U, u);
                    __threadfence_block();
                    __syncwarp();
                    // NMPC_HELP_EXTRA < NMPC_HELP_R would limit the offer to what the previous search needed plus
                    // EXTRA trials (less speculative work); measured worse than always offering all of them
                    if (lane < NMPC_HELP_R && lane < ls_hint + NMPC_HELP_EXTRA) {
                        const uint32_t aj = W.a_job + JOB_BYTES * lane;
                        const int stt = ldv_shared(aj);
                        if (stt == JOB_EMPTY || stt == JOB_DONE) {
                            stsi(aj + 4u, lane + 1);
                            stsi(aj + 8u, ls_seq);
                            sts1(aj + 16u, gamma);
                            sts1(aj + 24u, pn.c);
                            __threadfence_block();
                            stv_shared(aj, JOB_POSTED);
                        }
                    }
                    __syncwarp();
                }
#endif
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
#if NMPC_HELP_R > 0
            case PH_HELP_WAIT: {
                // poll the job records of every warp of the CTA: lane l looks at record (r, o) = (l / nwarps, l % nwarps),
                // so the lowest set bit of the vote is the most urgent trial (smallest r) on offer
                const int njobs = nwarps * (NMPC_HELP_R + 1);
                int j = -1;
                const uint32_t a_part = a_live + 4u + 4u * ((threadIdx.x >> 5) & 3u);  // this warp's sub-partition
                for (;;) {
                    if (ldv_shared(a_live) <= 0) return 0;
                    if (ldv_shared(a_part) > NMPC_HELP_PART_MAX) {  // owners share this sub-partition: stay out of their way
                        __nanosleep(2000);
                        continue;
                    }
                    unsigned m = 0;
                    int base = 0;
                    for (; base < njobs && !m; base += 32) {
                        const int l = base + lane;
                        bool posted = false;
                        if (l < njobs) {
                            const int o = l % nwarps, r = l / nwarps;
                            posted = ldv_shared(W.sb0 + (uint32_t)o * W.arena_bytes + (W.a_job - W.sb) + JOB_BYTES * r) == JOB_POSTED;
                        }
                        m = __ballot_sync(FULL, posted);
                    }
                    if (m) {
                        const int l = base - 32 + __ffs(m) - 1;
                        const int o = l % nwarps, r = l / nwarps;
                        int ok = 0;
                        if (lane == 0)
                            ok = cas_shared(W.sb0 + (uint32_t)o * W.arena_bytes + (W.a_job - W.sb) + JOB_BYTES * r, JOB_POSTED,
                                            JOB_TAKEN) == JOB_POSTED;
                        ok = __shfl_sync(FULL, ok, 0);
                        if (ok) {
                            j = l;
                            break;
                        }
                        continue;
                    }
                    __nanosleep(NMPC_HELP_SLEEP);
                }
                __threadfence_block();
                const int o = j % nwarps, r = j / nwarps;
                W.retarget(o);
                const uint32_t aj = W.a_job + JOB_BYTES * r;
                const int trial = ldsi(aj + 4u);
                set_gamma(lds1(aj + 16u));
                pn = make_pen(lds1(aj + 24u));
                ls_seq = r;
                W.ld(V_YL, yl);
                if (trial == 0) {  // psi(uhalf) for the owner's Lipschitz test
                    W.ld(V_UHALF, x);
                    mode = MODE_COST;
                    phase = PH_HELP_EVAL;
                    break;
                }
                tau = 1.0;
                for (int k = 0; k < trial; k++) tau /= 2.0;
                const double om = 1.0 - tau;
                double2 uo[P], fpr[P], dir[P];
                W.ld(V_U, uo);
                W.ld(V_FPR, fpr);
                W.ld(V_DIR, dir);
#pragma unroll
                for (int jj = 0; jj < P; jj++) {
                    x[jj].x = fma(-tau, dir[jj].x, fma(-om, fpr[jj].x, uo[jj].x));
                    x[jj].y = fma(-tau, dir[jj].y, fma(-om, fpr[jj].y, uo[jj].y));
                }
                mode = MODE_GRAD;
                phase = PH_HELP_EVAL;
                break;
            }
#endif
            case PH_EXIT: {
#ifdef NMPC_PROFILE
                prof[6] = clock64() - tstart;
                if (prof_out && lane == 0)
                    for (int i = 0; i < 8; i++) {
                        prof_out[i] = prof[i];
                        prof_out[8 + i] = W.pt[i];
                    }
#endif
                st_out.exit_status = status;
                st_out.outer_iterations = num_outer;
                st_out.inner_iterations = inner_total;
                st_out.last_norm_fpr = norm_fpr;
                st_out.delta_y_norm_over_c = sget(H_DYNP) / pn.c;
                st_out.f2_norm = sget(H_F2NP);
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
        PH_ACCOUNT(-1);
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
        PH_ACCOUNT(32 + phase);

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
                sput(H_NORMH, sqrt(hsum<P>(e)));
                mode = MODE_GRAD;
                phase = PH_INIT_B;
                break;
            }
            case PH_INIT_B: {
                double2 gr[P], gs[P], uh[P];
                W.ld(V_GRAD, gr);
                const double lip = sqrt(wdiff2<P>(g, gr)) / sget(H_NORMH);
                sput(H_LIP, lip);
                set_gamma(GAMMA_L_COEFF / fmax(lip, MIN_L_ESTIMATE));
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
                compute_fpr(uh, fpr);
                W.st(V_FPR, fpr);
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
                double lhs = cost - (0.5 * gamma) * gg + (0.5 * d2) * inv_gamma;
                bool evaluate = false;
                while (lhs > rhs_ls && nls < MAX_LINESEARCH_ITERATIONS) {
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
#if NMPC_HELP_R > 0
                    // was this trial offered to a helper?  DONE: take its result; TAKEN: wait for it;
                    // POSTED: nobody picked it up -> withdraw it and evaluate here.
                    const int r = (nls - 1) % NMPC_HELP_R;
                    const uint32_t aj = W.a_job + JOB_BYTES * r;
                    int got = 0;
                    if (lane == 0) {
                        int stt = ldv_shared(aj);
                        if (stt != JOB_EMPTY && ldsi(aj + 4u) == nls && ldsi(aj + 8u) == ls_seq) {
                            if (stt == JOB_POSTED && cas_shared(aj, JOB_POSTED, JOB_EMPTY) == JOB_POSTED) stt = JOB_EMPTY;
                            if (stt != JOB_EMPTY) {
                                while (ldv_shared(aj) != JOB_DONE) __nanosleep(20);
                                got = 1;
                            }
                        }
                    }
                    got = __shfl_sync(FULL, got, 0);
                    if (got) {
                        __threadfence_block();
                        cost = lds1(aj + 32u);
                        lhs = lds1(aj + 40u);
                        W.ld(V_JG + r, g);
                        n_grad++;
                        __syncwarp();
                        // the record is free again: offer the trial NMPC_HELP_R steps ahead
                        if (lane == 0) {
                            if (nls + NMPC_HELP_R <= MAX_LINESEARCH_ITERATIONS && lhs > rhs_ls &&
                                nls + NMPC_HELP_R <= ls_hint + NMPC_HELP_EXTRA + 1) {
                                stsi(aj + 4u, nls + NMPC_HELP_R);
                                __threadfence_block();
                                stv_shared(aj, JOB_POSTED);
                            } else {
                                stv_shared(aj, JOB_EMPTY);
                            }
                        }
                        if (!(lhs > rhs_ls && nls < MAX_LINESEARCH_ITERATIONS)) grad_step_half(x, g, gs, uh);  // accepted
                        continue;
                    }
#endif
                    evaluate = true;
                    break;
                }
                if (evaluate) {
                    mode = MODE_GRAD;
                    phase = PH_LS;
                } else {
#if NMPC_HELP_R > 0
                    if (lane < NMPC_HELP_R) {  // withdraw what is still on offer
                        const uint32_t aj = W.a_job + JOB_BYTES * lane;
                        if (ldv_shared(aj) == JOB_POSTED) cas_shared(aj, JOB_POSTED, JOB_EMPTY);
                    }
#endif
                    W.st(V_GRAD, g);
#pragma unroll
                    for (int j = 0; j < P; j++) u[j] = x[j];
                    ls_hint = nls;
                    fbe_u = lhs;
                    fbe_valid = true;
                    iteration++;
                    phase = PH_STEP_DONE;
                }
                break;
            }
#if NMPC_HELP_R > 0
            case PH_HELP_EVAL: {  // helper: envelope value of the trial, results into the owner's arena
                if (mode == MODE_COST) {
                    const uint32_t ajc = W.a_job + JOB_BYTES * ls_seq;
                    if (lane == 0) sts1(ajc + 32u, psi);
                    __threadfence_block();
                    __syncwarp();
                    if (lane == 0) stv_shared(ajc, JOB_DONE);
                    phase = PH_HELP_WAIT;
                    break;
                }
                double e[P], f[P], d2, gg;
#pragma unroll
                for (int j = 0; j < P; j++) {
                    const double gsx = fma(-gamma, g[j].x, x[j].x), gsy = fma(-gamma, g[j].y, x[j].y);
                    const double uhx = W.act[j] ? clampd(gsx, cfg.lin_vel_min, cfg.lin_vel_max) : 0.0;
                    const double uhy = W.act[j] ? clampd(gsy, -cfg.ang_vel_max, cfg.ang_vel_max) : 0.0;
                    const double d0 = gsx - uhx, d1 = gsy - uhy;
                    e[j] = fma(d1, d1, d0 * d0);
                    f[j] = fma(g[j].y, g[j].y, g[j].x * g[j].x);
                }
                hsum2<P>(e, f, d2, gg);
                const double lhs = psi - (0.5 * gamma) * gg + (0.5 * d2) * inv_gamma;
                const uint32_t aj = W.a_job + JOB_BYTES * ls_seq;
                W.st(V_JG + ls_seq, g);
                if (lane == 0) {
                    sts1(aj + 32u, psi);
                    sts1(aj + 40u, lhs);
                }
                __threadfence_block();
                __syncwarp();
                if (lane == 0) stv_shared(aj, JOB_DONE);
                phase = PH_HELP_WAIT;
                break;
            }
#endif
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
                const double dynp = sqrt(hsum<P>(e)), f2np = sqrt(pen);
                const double akkt_tol = sget(H_AKKT);
                sput(H_DYNP, dynp);
                sput(H_F2NP, f2np);
                const bool crit1 = alm_iter > 0 && dynp <= pn.c * cfg.delta_tolerance + DBL_EPS;
                const bool crit2 = (nf2 == 0) || f2np <= cfg.delta_tolerance + DBL_EPS;
                const bool crit3 = akkt_tol <= cfg.tolerance + DBL_EPS;
                bool finished = crit1 && crit2 && crit3;
                if (!finished) {
                    bool stall;
                    if (alm_iter == 0) stall = true;
                    else {
                        const bool ca = dynp <= cfg.sufficient_decrease_coeff * sget(H_DYN) + DBL_EPS;
                        const bool cp = f2np <= cfg.sufficient_decrease_coeff * sget(H_F2N) + DBL_EPS;
                        stall = (nf2 > 0) ? (ca && cp) : ca;
                    }
                    if (!stall) pn = make_pen(pn.c * cfg.penalty_update_factor);
                    sput(H_AKKT, fmax(akkt_tol * cfg.inner_tolerance_update, cfg.tolerance));
                    alm_iter++;
                    sput(H_DYN, dynp);
                    sput(H_F2N, f2np);
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
