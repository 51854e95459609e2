// nmpc_device.cuh — device code of the batched NMPC solver (one warp per problem, G-lane evaluation groups).
//
// What it computes is the problem of MpcModule.build() (src/mpc/mpc_generator.py:66-193) solved
// the way the reference's OpEn solver does (PANOC + L-BFGS inside an ALM/penalty loop); the
// results are those of oracle/nmpc_oracle.c bit for bit (same statements, same arithmetic contract,
// explicit fma, own sincos, group-ordered reductions).
//
// Organisation (round 2):
//   * a warp owns one problem.  Its 32 lanes form NG = 32 / G evaluation groups of G lanes (G = 8 for
//     N <= 24, else 16); inside a group lane i owns the S consecutive horizon steps S*i .. S*i+S-1, so
//     every per-step quantity is S independent dependency chains per lane (the latency-bound FP64 code
//     gets its instruction-level parallelism from there) and every reduction / scan over the horizon is
//     S-1 serial adds plus log2(G) shuffle stages;
//   * the PANOC / L-BFGS vector algebra is replicated in every group (same loads, same results), the
//     EVALUATIONS are not: one call of eval() computes psi and grad psi at NG different points at once,
//     one per group.  PANOC's iteration needs psi(u_half) (Lipschitz test) and then psi, grad psi at the
//     line-search trials tau = 1, 1/2, 1/4, ... one after the other; here the L-BFGS update and the
//     two-loop recursion run first (they do not need psi(u_half)) and ONE call evaluates u_half and the
//     first NG-1 trials together.  76 % of the iterations of the BASELINE config-2 batch end with that
//     single call (3.4 sequential evaluations per iteration in round 1).  The speculation is exact: a
//     trial's value does not depend on who evaluates it or when, the Lipschitz test that fails (rare)
//     discards the L-BFGS update exactly like the reference discards its memory, and the evaluation
//     counters report what the reference's serial loop would have evaluated;
//   * the per-problem constants and the PANOC / L-BFGS vectors live in the warp's shared-memory arena
//     (vector element (s, i) at 16*(G*s + i): a group's lanes read 16*G contiguous bytes, conflict-free),
//     addressed with explicit 32-bit shared addresses;
//   * the solver is a phase machine with ONE evaluation site, which keeps the hot loop near the size of
//     the instruction cache.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/nmpc_b200.h"

#define FULL 0xffffffffu
#define PROBE_BUCKETS 1024
#define MEMP1 (NMPC_LBFGS_MAX + 1)
#ifndef NMPC_SEG_UNR
#define NMPC_SEG_UNR 1  // reference segments per half-trip of the cross-track loop (x S steps per lane in flight); 2 is
                        // as fast per segment but 110 instructions longer: -3 % at saturation (instruction cache)
#endif

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
// horizon layout: G lanes per evaluation group, S consecutive steps per lane (same rule as nmpc_oracle_layout)
__host__ __device__ inline void nmpc_layout_for(int N, int& G, int& S) {
    if (N <= 16) { G = 8; S = 2; }
    else if (N <= 24) { G = 8; S = 3; }
    else if (N <= 32) { G = 16; S = 2; }
    else if (N <= 48) { G = 16; S = 3; }
    else if (N <= 64) { G = 16; S = 4; }
    else { G = 16; S = 6; }
}

// per-warp shared-memory arena (offsets in doubles; every block is 16-byte aligned)
enum { V_GRAD = 0, V_UHALF, V_FPR, V_DIR, V_GSTEP, V_OLDS, V_OLDG, V_U, V_YL, V_T0, V_S, V_Y = V_S + MEMP1,
       V_END = V_Y + MEMP1 };
enum { H_X0 = 0, H_Y0, H_TH0, H_VINIT, H_WINIT, H_XREF, H_YREF, H_THREF, H_Q, H_QV, H_QTH, H_RV, H_RW, H_QN, H_QTHN,
       H_QCTE, H_AP, H_WP, H_INVTS,
       // warp-uniform solver state that is touched once per outer iteration (kept out of the registers)
       H_F2N, H_DYN, H_F2NP, H_DYNP, H_NORMH, H_LIP, H_AKKT, H_NCIRC /* int */,
       // PANOC scalars: every lane stores the same value and reads back its own store.  Keeping them (and the
       // counters below) here instead of in registers is what lets the evaluation have the register file
       H_GAMMA = 28, H_INVG, H_SIGMA, H_COST, H_NFPR, H_RHSLS, H_LBG, H_FBEU, H_IP, H_PENC, H_PINV, H_TBEG,
       H_INTS = 40,
       H_HRES = 48,  // helper mode: (psi, left-hand side) of the trial each group evaluated, 2 doubles x NG
       H_COUNT = 56 };
enum { I_NCOST = 0, I_NGRAD, I_ALM, I_INNER, I_NOUTER, I_STATUS, I_ISTATUS, I_ITLIP, I_NUMIT, I_FLAGS };
#define SEG_STRIDE 6   // s1x s1y | dx dy | inv pad
#define CIRC_STRIDE 4  // cx cy | r2 (original slot index as int in the 4th double)
#define ELL_STRIDE 6   // ex ey | cosA sinA | 1/rx^2 1/ry^2

struct Lay {
    int vlen, seg, circ, ell, ebd, rho, alpha, syd, hdr, vref, total;
};
__host__ __device__ inline int even_up(int x) { return (x + 1) & ~1; }
__host__ __device__ inline Lay make_layout(int N, int Nobs, int Nd) {
    int G, S;
    nmpc_layout_for(N, G, S);
    Lay L;
    L.vlen = 2 * G * S;  // doubles per vector: (v, w) for G*S step slots, slots >= N stay zero
    int o = V_END * L.vlen;
    L.seg = o; o += SEG_STRIDE * (N + 3 * NMPC_SEG_UNR);  // + copies of the last segment: no remainder trips, prefetch overrun
    L.circ = o; o += CIRC_STRIDE * (Nobs + 4);
    L.ell = o; o += ELL_STRIDE * Nd * N;
    L.ebd = o; o += 4 * Nd;  // per dynamic obstacle: centre and squared radius of a disc around all its poses
    L.rho = o; o += 12;
    L.alpha = o; o += 12;
    L.syd = o; o += 12;  // Gram entries <s_p, y_(pair accepted after p)> by slot (paired two-loop recursion)
    L.hdr = o; o += H_COUNT;
    L.vref = o; o += even_up(G * S);
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
    long long* dbg;  // NMPC_PROFILE builds only: cycle counters per problem
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
__device__ __forceinline__ void sts2_if(uint32_t a, double2 v, bool on) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\t@p st.shared.v2.f64 [%0], {%1, %2};\n\t}" ::"r"(a), "d"(v.x), "d"(v.y),
                 "r"((int)on)
                 : "memory");
}
// IEEE division / square root out of line: the hot loop holds one copy of the ~30-instruction sequences
#ifndef NMPC_OOL_DIV
#define NMPC_OOL_DIV 0  // measured: -2.4 % at saturation (the call costs more than the instruction-cache lines it saves)
#endif
#if NMPC_OOL_DIV
__device__ __noinline__ double nm_div(double a, double b) { return a / b; }
__device__ __noinline__ double nm_sqrt(double a) { return sqrt(a); }
#else
__device__ __forceinline__ double nm_div(double a, double b) { return a / b; }
__device__ __forceinline__ double nm_sqrt(double a) { return sqrt(a); }
#endif
// one lane stores (predicated instruction, no divergent branch)
__device__ __forceinline__ void sts1_if(uint32_t a, double v, bool on) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p st.shared.f64 [%0], %1;\n\t}" ::"r"(a), "d"(v), "r"((int)on) : "memory");
}
__device__ __forceinline__ void stsi(uint32_t a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// ---------------------------------------------------------------------------------
// Helper warps.  A warp that has run out of problems attaches itself to a warp of its CTA that is still solving (the
// "owner") and, every PANOC iteration, evaluates the owner's SECOND batch of line-search trials while the owner's own
// call evaluates psi(u_half) and the first batch (solve_problem: PH_HELP / take_from_helper).  When the owner's call
// accepts nothing (a fifth to half of the iterations) the next batch is already there.  Same operations on the same
// inputs: the results do not depend on who evaluates a trial.  One mailbox per warp slot; flags move with
// release / acquire at CTA scope.  Every wait is a loop that ALL lanes of the warp execute (one warp-wide load
// returns one value, so the loop is warp-uniform): a polling loop run by lane 0 alone leaves the warp split in two
// for everything that follows, at a third of the speed.
struct Mailbox {
    unsigned seq;     // owner: bumped after the trial inputs (V_U, V_FPR, V_DIR, step size) of an iteration are in place
    unsigned done;    // owner's box: the seq whose results are in the helper's arena; helper's own box: the last seq it took
    unsigned state;   // MB_RUNNING while the warp owns problems, MB_DONE once it has retired
    unsigned helper;  // owner's box: 0, or 1 + the slot of the attached helper; helper's own box: the slot of its owner
};
struct HelpShared {
    Mailbox mb[32];
    unsigned arena_bytes, hdr_off, smem_base, pad;  // so that cold paths need no registers for them
};
__shared__ HelpShared g_help;
enum { MB_RUNNING = 1, MB_DONE = 2 };
enum { MB_SEQ = 0, MB_DONEQ = 4, MB_STATE = 8, MB_HELPER = 12 };  // byte offsets inside a mailbox
enum { HELP_NONE = 0, HELP_OWNER = 1, HELP_HELPER = 2 };           // solve_problem's `help` argument
__device__ __forceinline__ uint32_t help_base() { return (uint32_t)__cvta_generic_to_shared(&g_help); }
__device__ __forceinline__ uint32_t mbox_of(int slot) { return help_base() + 16u * (uint32_t)slot; }
__device__ __forceinline__ uint32_t help_u32(int field) { return ((const volatile unsigned*)&g_help.arena_bytes)[field]; }
__device__ __forceinline__ unsigned ld_acquire(uint32_t a) {
    unsigned v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(uint32_t a, unsigned v) { asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned cas_acq_rel(uint32_t a, unsigned cmp, unsigned val) {
    unsigned old;
    asm volatile("atom.acq_rel.cta.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(a), "r"(cmp), "r"(val) : "memory");
    return old;
}
// a short pause between two polls of a flag (nanosleep's granularity is far too coarse here)
__device__ __forceinline__ void spin_wait(int cycles) {
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {
    }
}

// Rectangle::project of OpEn is comparison-based: a NaN stays a NaN (and ends the solve as NotFinite)
__device__ __forceinline__ double clampd(double x, double lo, double hi) { return (x < lo) ? lo : ((x > hi) ? hi : x); }
// min/max as compare-selects (same forms as the oracle): NaN -> the constant, zero results are +0
// (written as setp/selp PTX: the C ternaries get canonicalised to max.f64/min.f64, which sm_100
//  expands into a ~12-instruction DSETP.MAX/FSEL/SEL/NaN-fix-up sequence each)
#ifndef NMPC_ICLAMP
#define NMPC_ICLAMP 1  // round 2: -4 % cycles on the cross-track loop now that it is FP64-pipe bound
#endif
__device__ __forceinline__ double sel_clamp01(double t) {
#if NMPC_ICLAMP
    // Same result as the two compare-selects for every non-NaN t, computed on the integer pipe from the
    // high word (sign and exponent order doubles like signed integers for t >= 0): hi' = min(max(hi, 0), hi(1.0)),
    // lo' = lo only while 0 <= hi < hi(1.0).  Two FP64-pipe compares and four selects become four ALU ops.
    // (NaN: the selects give 0, this gives 0 or 1 by the sign bit; either way the distance stays NaN because
    //  a NaN projection parameter comes from a NaN point, which is already in ex/ey.)
    double r0;
    asm("{\n\t.reg .b32 lo, hi, h2;\n\t.reg .pred p;\n\t"
        "mov.b64 {lo, hi}, %1;\n\t"
        "max.s32 h2, hi, 0;\n\t"
        "min.s32 h2, h2, 0x3FF00000;\n\t"
        "setp.lt.u32 p, hi, 0x3FF00000;\n\t"
        "selp.b32 lo, lo, 0, p;\n\t"
        "mov.b64 %0, {lo, h2};\n\t}"
        : "=d"(r0)
        : "d"(t));
    return r0;
#endif
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
// group-ordered reductions: lane partials are serial sums over the lane's S steps (the callers
// form them); these combine the G partials of a group.  Shuffle distances < G never leave an aligned group.
template <int G>
__device__ __forceinline__ double gsum(double a) {
#pragma unroll
    for (int off = G / 2; off; off >>= 1) a = a + __shfl_xor_sync(FULL, a, off);
    return a;
}
template <int G>
__device__ __forceinline__ void gsum2(double& a, double& b) {
#pragma unroll
    for (int off = G / 2; off; off >>= 1) {
        const double ya = __shfl_xor_sync(FULL, a, off), yb = __shfl_xor_sync(FULL, b, off);
        a = a + ya;
        b = b + yb;
    }
}
template <int G>
__device__ __forceinline__ void gsum4(double& a, double& b, double& c, double& d) {
#pragma unroll
    for (int off = G / 2; off; off >>= 1) {
        const double ya = __shfl_xor_sync(FULL, a, off), yb = __shfl_xor_sync(FULL, b, off);
        const double yc = __shfl_xor_sync(FULL, c, off), yd = __shfl_xor_sync(FULL, d, off);
        a = a + ya;
        b = b + yb;
        c = c + yc;
        d = d + yd;
    }
}
template <int G>
__device__ __forceinline__ void gsum5(double& a, double& b, double& c, double& d, double& e) {
#pragma unroll
    for (int off = G / 2; off; off >>= 1) {
        const double ya = __shfl_xor_sync(FULL, a, off), yb = __shfl_xor_sync(FULL, b, off);
        const double yc = __shfl_xor_sync(FULL, c, off), yd = __shfl_xor_sync(FULL, d, off);
        const double ye = __shfl_xor_sync(FULL, e, off);
        a = a + ya;
        b = b + yb;
        c = c + yc;
        d = d + yd;
        e = e + ye;
    }
}
// Kogge-Stone inclusive scan of the lane totals over the group; returns the EXCLUSIVE prefix of this lane
// (the scanned total of lane gl-1; 0.0 for the first lane)
template <int G>
__device__ __forceinline__ double gscan_up_excl(double T, int gl) {
#pragma unroll
    for (int off = 1; off < G; off <<= 1) {
        const double y = __shfl_up_sync(FULL, T, off, G);
        add_if(T, y, gl >= off);
    }
    const double E = __shfl_up_sync(FULL, T, 1, G);
    return gl == 0 ? 0.0 : E;
}
template <int G>
__device__ __forceinline__ void gscan_up_excl2(double Ta, double Tb, int gl, double& Ea, double& Eb) {
#pragma unroll
    for (int off = 1; off < G; off <<= 1) {
        const double ya = __shfl_up_sync(FULL, Ta, off, G), yb = __shfl_up_sync(FULL, Tb, off, G);
        add_if(Ta, ya, gl >= off);
        add_if(Tb, yb, gl >= off);
    }
    Ea = __shfl_up_sync(FULL, Ta, 1, G);
    Eb = __shfl_up_sync(FULL, Tb, 1, G);
    Ea = gl == 0 ? 0.0 : Ea;
    Eb = gl == 0 ? 0.0 : Eb;
}
// suffix direction: exclusive suffix = scanned total of lane gl+1 (0.0 for the last lane)
template <int G>
__device__ __forceinline__ double gscan_down_excl(double T, int gl) {
#pragma unroll
    for (int off = 1; off < G; off <<= 1) {
        const double y = __shfl_down_sync(FULL, T, off, G);
        add_if(T, y, gl + off < G);
    }
    const double E = __shfl_down_sync(FULL, T, 1, G);
    return gl == G - 1 ? 0.0 : E;
}
// two suffix scans with an independent group sum riding along in the same instruction stream
template <int G>
__device__ __forceinline__ void gscan_down_excl2_sum(double Ta, double Tb, int gl, double& Ea, double& Eb, double& acc) {
    int st = G / 2;
#pragma unroll
    for (int off = 1; off < G; off <<= 1, st >>= 1) {
        const double ya = __shfl_down_sync(FULL, Ta, off, G), yb = __shfl_down_sync(FULL, Tb, off, G);
        const double yc = __shfl_xor_sync(FULL, acc, st);
        add_if(Ta, ya, gl + off < G);
        add_if(Tb, yb, gl + off < G);
        acc = acc + yc;
    }
    Ea = __shfl_down_sync(FULL, Ta, 1, G);
    Eb = __shfl_down_sync(FULL, Tb, 1, G);
    Ea = gl == G - 1 ? 0.0 : Ea;
    Eb = gl == G - 1 ? 0.0 : Eb;
}

// ---------------------------------------------------------------------------------
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
template <int G, int S>
struct Warp {
    static constexpr int NG = 32 / G;  // evaluation groups
    const nmpc_config& cfg;
    uint32_t sb;       // shared byte address of the arena
    uint32_t la;       // sb + 16*gl : this lane's element of slot row 0 inside vector 0
    uint32_t vstride;  // bytes per vector
    uint32_t a_hdr;
    int o_seg, o_circ, o_ell, o_ebd, o_rho, o_alpha, o_syd, o_vref;  // byte offsets inside the arena (the same for every warp)
    int lane, grp, gl, N;
    __device__ __forceinline__ uint32_t a_seg() const { return sb + o_seg; }
    __device__ __forceinline__ uint32_t a_circ() const { return sb + o_circ; }
    __device__ __forceinline__ uint32_t a_ell() const { return sb + o_ell; }
    __device__ __forceinline__ uint32_t a_ebd() const { return sb + o_ebd; }
    __device__ __forceinline__ uint32_t a_rho() const { return sb + o_rho; }
    __device__ __forceinline__ uint32_t a_alpha() const { return sb + o_alpha; }
    __device__ __forceinline__ uint32_t a_syd() const { return sb + o_syd; }
    __device__ __forceinline__ uint32_t a_vref() const { return sb + o_vref; }
    __device__ __forceinline__ int tix(int s) const { return S * gl + s; }
    __device__ __forceinline__ bool act(int s) const { return S * gl + s < N; }
    // n_circ: circles with r != 0 (zero-padded slots are skipped: they add exact zeros)
    __device__ __forceinline__ int n_circ() const { return ldsi(a_hdr + 8u * H_NCIRC); }
#ifdef NMPC_PROFILE
    long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // cycles per eval section (tools/prof_cycles.py)
#endif

    __device__ __forceinline__ Warp(const nmpc_config& c, const Lay& L, int warp, int lane_) : cfg(c), lane(lane_) {
        grp = lane / G;
        gl = lane % G;
        N = cfg.N_hor;
        sb = (uint32_t)__cvta_generic_to_shared(smem) + (uint32_t)warp * (uint32_t)L.total * 8u;
        la = sb + 16u * gl;
        vstride = (uint32_t)L.vlen * 8u;
        o_seg = L.seg * 8; o_circ = L.circ * 8; o_ell = L.ell * 8; o_ebd = L.ebd * 8; o_rho = L.rho * 8;
        o_alpha = L.alpha * 8; o_syd = L.syd * 8; o_vref = L.vref * 8;
        a_hdr = sb + L.hdr * 8u;
    }
    // point this view at the arena of warp slot `slot` (a helper warp looks at its owner's arena)
    __device__ __forceinline__ void rebase(int slot) {
        sb = help_u32(2) + (uint32_t)slot * help_u32(0);
        la = sb + 16u * gl;
        a_hdr = sb + help_u32(1);
    }
    __device__ __forceinline__ double hdr(int i) const { return lds1(a_hdr + 8u * i); }
    // vector k of the arena: every lane reads its S (v, w) pairs; all groups see the same vector
    __device__ __forceinline__ void ld(int k, double2 (&r)[S]) const {
#pragma unroll
        for (int s = 0; s < S; s++) r[s] = lds2(la + k * vstride + 16u * G * s);
    }
    // stores come from ONE group (by default group 0; `from` = the group whose registers hold the vector)
    // (the callers place the warp barriers: one before a store to a vector that other groups may still be reading in
    //  the same phase, one after the last store before anybody loads; phases are separated by barriers anyway)
    __device__ __forceinline__ void st(int k, const double2 (&r)[S], int from = 0) const {
        const bool mine = grp == from;
#pragma unroll
        for (int s = 0; s < S; s++) sts2_if(la + k * vstride + 16u * G * s, r[s], mine);  // predicated: no branch
    }

    // unpack the parameter row (layout: include/nmpc_b200.h) into the arena
    __device__ void stage(const double* __restrict__ p) {
        const int Nobs = cfg.Nobs, Nd = cfg.Ndynobs;
        __syncwarp();
        if (lane < 8) sts1(a_hdr + 8u * lane, p[lane]);
        if (lane >= 8 && lane < 18) sts1(a_hdr + 8u * lane, p[lane + 2]);
        if (lane == 18) sts1(a_hdr + 8u * H_INVTS, 1.0 / cfg.ts);
        for (int t = lane; t < G * S; t += 32) sts1(a_vref() + 8u * t, t < N ? p[NMPC_NZ + t] : 0.0);
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
                const uint32_t a = a_circ() + 32u * pos;
                sts2(a, make_double2(cx, cy));
                sts1(a + 16u, r * r);
                stsi(a + 24u, k);
            }
            nreal += __popc(m);
        }
        if (lane < 4) {  // the circle loop reads four at a time: dummies that can never be entered (r^2 = -1)
            const uint32_t a = a_circ() + 32u * (nreal + lane);
            sts2(a, make_double2(0.0, 0.0));
            sts1(a + 16u, -1.0);
            stsi(a + 24u, 0);
        }
        if (lane == 0) stsi(a_hdr + 8u * H_NCIRC, nreal);
        const double* pe = pc + 3 * Nobs;
        const int ne = Nd * N;
        for (int i = lane; i < ne; i += 32) {
            const double* e = pe + 5 * i;  // obstacle-major then time: offset k*5N + 5t = 5*(k*N + t)
            double sa, ca;
            nm_sincos(e[4], sa, ca);
            const uint32_t a = a_ell() + 48u * i;
            sts2(a, make_double2(e[0], e[1]));
            sts2(a + 16u, make_double2(ca, sa));
            sts2(a + 32u, make_double2(1.0 / (e[2] * e[2]), 1.0 / (e[3] * e[3])));
        }
        // A disc that contains every pose of dynamic obstacle k over the horizon (plus a relative margin far above
        // rounding): a predicted point outside it is outside the ellipse at every step, so eval() can skip the
        // obstacle without changing a bit (it would add exact zeros).  NaN / inf parameters give a NaN radius: the
        // test `inside the disc` is then false for every point, and so is the ellipse test itself.
        for (int k = lane; k < Nd; k += 32) {
            const double* e = pe + 5 * (size_t)k * N;
            double xlo = e[0], xhi = e[0], ylo = e[1], yhi = e[1];
            for (int t = 1; t < N; t++) {
                xlo = fmin(xlo, e[5 * t]); xhi = fmax(xhi, e[5 * t]);
                ylo = fmin(ylo, e[5 * t + 1]); yhi = fmax(yhi, e[5 * t + 1]);
            }
            const double bx = 0.5 * (xlo + xhi), by = 0.5 * (ylo + yhi);
            double R = 0.0;
            bool bad = false;
            for (int t = 0; t < N; t++) {
                const double ddx = e[5 * t] - bx, ddy = e[5 * t + 1] - by;
                const double r = sqrt(ddx * ddx + ddy * ddy) + fmax(fabs(e[5 * t + 2]), fabs(e[5 * t + 3]));
                bad = bad || !(r == r) || !(e[5 * t + 4] == e[5 * t + 4]);
                R = fmax(R, r);
            }
            R = R * (1.0 + 1e-6) + 1e-9;
            const uint32_t a = a_ebd() + 32u * k;
            sts2(a, make_double2(bx, by));
            sts1(a + 16u, bad ? CUDART_NAN : R * R);
        }
        const double* pr = pe + 5 * ne;
        for (int ii = lane; ii < N + 3 * NMPC_SEG_UNR; ii += 32) {
            if (ii >= 1) {
                // slots N .. repeat segment N-1: equal distances never win the strict '<' arg-min
                const int i = (ii < N) ? ii : N - 1;
                double ax = pr[3 * (i - 1)], ay = pr[3 * (i - 1) + 1];
                double dx = pr[3 * i] - ax, dy = pr[3 * i + 1] - ay;
                const uint32_t a = a_seg() + 48u * ii;
                sts2(a, make_double2(ax, ay));
                sts2(a + 16u, make_double2(dx, dy));
                sts1(a + 32u, 1.0 / (fma(dx, dx, dy * dy) + 1e-16));
            }
        }
        __syncwarp();
    }

    // control of the step before this lane's step s: in-lane, the previous lane's last step, or (v_init, w_init)
    __device__ __forceinline__ void prev_controls(const double2 (&uv)[S], double& vp0, double& wp0) const {
        vp0 = __shfl_up_sync(FULL, uv[S - 1].x, 1, G);
        wp0 = __shfl_up_sync(FULL, uv[S - 1].y, 1, G);
        if (gl == 0) {
            vp0 = hdr(H_VINIT);
            wp0 = hdr(H_WINIT);
        }
    }

    // psi, grad psi and |F2|^2 of the staged problem at this GROUP's point uv (every group evaluates its own
    // point with its own penalty parameter; the multipliers come from the arena vector V_YL)
    // The penalty parameter c comes from the arena header (H_PENC, H_PINV = 1 / max(c, 1)); zero_c: this lane's
    // group evaluates with c = 0 (f(u) alone).
    __device__ double eval(const double2 (&uv)[S], const bool zero_c, double2 (&gout)[S], double& pen_out,
                           double* __restrict__ F2g) {
        const double ts = cfg.ts;
        PROF_BEGIN();
        // ---- rollout (src/mpc/mpc_generator.py:88-90): heading by a scan of ts*w, position by a joint scan
        double cth[S];
#pragma unroll
        for (int s = 0; s < S; s++) {
            const double v = ts * uv[s].y;  // padded steps hold (0, 0): no mask needed
            cth[s] = (s == 0) ? v : cth[s - 1] + v;
        }
        const double Eth = gscan_up_excl<G>(cth[S - 1], gl);
        const double th0 = hdr(H_TH0);
        const double thp0 = th0 + Eth;  // heading before this lane's first step
        double TH[S], sn[S], cs[S];
#pragma unroll
        for (int s = 0; s < S; s++) TH[s] = th0 + (Eth + cth[s]);
#pragma unroll
        for (int s = 0; s < S; s++) nm_sincos(s == 0 ? thp0 : TH[s > 0 ? s - 1 : 0], sn[s], cs[s]);
        PROF_MARK(0);
        double X[S], Y[S], xp0, yp0;
        {
            double ca[S], cb[S];
#pragma unroll
            for (int s = 0; s < S; s++) {
                const double va = ts * (uv[s].x * cs[s]);  // padded steps: v = 0
                const double vb = ts * (uv[s].x * sn[s]);
                ca[s] = (s == 0) ? va : ca[s - 1] + va;
                cb[s] = (s == 0) ? vb : cb[s - 1] + vb;
            }
            double Ea, Eb;
            gscan_up_excl2<G>(ca[S - 1], cb[S - 1], gl, Ea, Eb);
            const double x0 = hdr(H_X0), y0 = hdr(H_Y0);
            xp0 = x0 + Ea;
            yp0 = y0 + Eb;
#pragma unroll
            for (int s = 0; s < S; s++) {
                X[s] = x0 + (Ea + ca[s]);
                Y[s] = y0 + (Eb + cb[s]);
            }
        }
        PROF_MARK(1);

        // ---- cross-track error (:122-144): every lane scans the N-1 segments for its S predicted points
        double gX[S], gY[S], mind2[S];
        const double qcte = hdr(H_QCTE);
        {
            double best[S];
            int bi[S];
#pragma unroll
            for (int s = 0; s < S; s++) {
                best[s] = CUDART_INF;
                bi[s] = 1;
            }
            constexpr int U = NMPC_SEG_UNR;
            struct Seg {
                double2 s1[U], d[U];
                double inv[U];
            };
            auto ldseg = [&](uint32_t a, Seg& g) {
#pragma unroll
                for (int q = 0; q < U; q++) {
                    g.s1[q] = lds2(a + 48u * q);
                    g.d[q] = lds2(a + 48u * q + 16u);
                    g.inv[q] = lds1(a + 48u * q + 32u);
                }
            };
            auto body = [&](int i, const Seg& g) {
                double d2[U][S];
#pragma unroll
                for (int q = 0; q < U; q++)
#pragma unroll
                    for (int s = 0; s < S; s++) {
                        const double px = X[s] - g.s1[q].x, py = Y[s] - g.s1[q].y;
                        const double that = fma(px, g.d[q].x, py * g.d[q].y) * g.inv[q];
                        const double tst = sel_clamp01(that);
                        const double ex = fma(tst, g.d[q].x, -px), ey = fma(tst, g.d[q].y, -py);
                        d2[q][s] = fma(ex, ex, ey * ey);
                    }
#pragma unroll
                for (int q = 0; q < U; q++)
#pragma unroll
                    for (int s = 0; s < S; s++) take_if_less(d2[q][s], i + q, best[s], bi[s]);
            };
            // two half-trips per trip, each working on segments fetched one half-trip earlier (no moves, the loads
            // of the next half-trip are in flight behind the arithmetic of this one)
            Seg ga, gb;
            uint32_t as = a_seg() + 48u;
            ldseg(as, ga);
#pragma unroll 1
            for (int i = 1; i < N; i += 2 * U, as += 96u * U) {
                ldseg(as + 48u * U, gb);
                body(i, ga);
                ldseg(as + 96u * U, ga);
                body(i + U, gb);
            }
#pragma unroll
            for (int s = 0; s < S; s++) {  // redo the arg-min segment (same operations, same bits) for the gradient
                mind2[s] = best[s];
                const uint32_t ab = a_seg() + 48u * bi[s];
                const double2 s1 = lds2(ab), d = lds2(ab + 16u);
                const double inv = lds1(ab + 32u);
                const double px = X[s] - s1.x, py = Y[s] - s1.y;
                const double that = fma(px, d.x, py * d.y) * inv;
                const double tst = sel_clamp01(that);
                const double ex = fma(tst, d.x, -px), ey = fma(tst, d.y, -py);
                const double ed = (that >= 0.0 && that <= 1.0) ? fma(ex, d.x, ey * d.y) * inv : 0.0;
                const double k2 = 2.0 * qcte;
                gX[s] = k2 * fma(ed, d.x, -ex);
                gY[s] = k2 * fma(ed, d.y, -ey);
            }
        }
        PROF_MARK(2);

        // ---- obstacle penalty F2 (:106-119): F2_k = sum over the horizon of max(0, inside_k(t)).
        // All inside-tests of a chunk are issued back to back and ONE warp-wide vote decides whether any group
        // needs the group sums and gradient terms (skipping adds exact zeros only).
        double pen = 0.0;
        {
            // pass 1: the inside-tests of all circles back to back (no votes, no branches: the chains overlap); every
            // lane notes the circles one of its points is inside of, one warp-wide OR gives the circles that need
            // pass 2 (group sum, penalty, gradient) — rarely any.  32 circles per round.
            const int ncirc = n_circ();
#pragma unroll 1
            for (int kb = 0; kb < ncirc; kb += 32) {
                const int kn = min(32, ncirc - kb);
                unsigned mine = 0;
                uint32_t ac = a_circ() + 32u * kb;
#pragma unroll 1
                for (int k = 0; k < kn; k += 4, ac += 128u) {
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const double2 cxy = lds2(ac + 32u * q);
                        const double r2 = lds1(ac + 32u * q + 16u);  // the slot behind the last circle holds r^2 = -1
                        // (bitwise | and &: a short-circuit || would make every test wait for the one before)
                        unsigned in = 0u;
#pragma unroll
                        for (int s = 0; s < S; s++) {
                            const double dx = X[s] - cxy.x, dy = Y[s] - cxy.y;
                            in |= (unsigned)(fma(-dy, dy, fma(-dx, dx, r2)) > 0.0) & (unsigned)act(s);
                        }
                        mine |= in << (k + q);
                    }
                }
                unsigned todo = __reduce_or_sync(FULL, mine);
                while (todo) {
                    const int k = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const uint32_t a1 = a_circ() + 32u * (kb + k);
                    const double2 cxy = lds2(a1);
                    const double r2 = lds1(a1 + 16u);
                    double hp[S], dx[S], dy[S];
#pragma unroll
                    for (int s = 0; s < S; s++) {
                        dx[s] = X[s] - cxy.x;
                        dy[s] = Y[s] - cxy.y;
                        const double hh = fma(-dy[s], dy[s], fma(-dx[s], dx[s], r2));
                        hp[s] = (act(s) && hh > 0.0) ? hh : 0.0;
                    }
                    double g = hp[0];
#pragma unroll
                    for (int s = 1; s < S; s++) g = g + hp[s];
                    g = gsum<G>(g);
                    if (F2g && lane == 0 && g > 0.0) F2g[ldsi(a1 + 24u)] = g;
                    pen = fma(g, g, pen);
                    const double cg = (zero_c ? 0.0 : hdr(H_PENC)) * g;
#pragma unroll
                    for (int s = 0; s < S; s++)
                        if (hp[s] > 0.0) {
                            gX[s] = fma(cg, -2.0 * dx[s], gX[s]);
                            gY[s] = fma(cg, -2.0 * dy[s], gY[s]);
                        }
                }
            }
            const int Nd = cfg.Ndynobs;
            unsigned near = 0u;  // dynamic obstacles (32 per round) whose bounding disc holds one of this lane's points
#pragma unroll 1
            for (int kb = 0; kb < Nd; kb += 32) {
            const int kn = min(32, Nd - kb);
            near = 0u;
#pragma unroll 1
            for (int k = 0; k < kn; k++) {
                const double2 bxy = lds2(a_ebd() + 32u * (kb + k));
                const double R2 = lds1(a_ebd() + 32u * (kb + k) + 16u);
                unsigned in = 0u;
#pragma unroll
                for (int s = 0; s < S; s++) {
                    const double dx = X[s] - bxy.x, dy = Y[s] - bxy.y;
                    in |= (unsigned)(fma(dx, dx, dy * dy) < R2) & (unsigned)act(s);
                }
                near |= in << k;
            }
            unsigned todo_e = __reduce_or_sync(FULL, near);  // nobody near: the obstacle adds exact zeros
            while (todo_e) {
                const int k = kb + __ffs(todo_e) - 1;
                todo_e &= todo_e - 1;
                double hp[S], ta[S], tb[S], eca[S], esa[S];
                bool any = false;
#pragma unroll
                for (int s = 0; s < S; s++) {
                    const uint32_t ae = a_ell() + 48u * (k * N + (act(s) ? tix(s) : 0));
                    const double2 exy = lds2(ae), csa = lds2(ae + 16u), ir = lds2(ae + 32u);
                    const double dx = X[s] - exy.x, dy = Y[s] - exy.y;
                    eca[s] = csa.x;
                    esa[s] = csa.y;
                    const double ea = fma(dx, csa.x, dy * csa.y);
                    const double eb = fma(dx, csa.y, -(dy * csa.x));
                    const double hh = fma(-(eb * eb), ir.y, fma(-(ea * ea), ir.x, 1.0));
                    const bool in = act(s) & (hh > 0.0);
                    hp[s] = in ? hh : 0.0;
                    any = any | in;
                    ta[s] = ea * ir.x;
                    tb[s] = eb * ir.y;
                }
                if (__any_sync(FULL, any)) {
                    double g = hp[0];
#pragma unroll
                    for (int s = 1; s < S; s++) g = g + hp[s];
                    g = gsum<G>(g);
                    if (F2g && lane == 0 && g > 0.0) F2g[cfg.Nobs + k] = g;
                    pen = fma(g, g, pen);
                    const double cg = (zero_c ? 0.0 : hdr(H_PENC)) * g;
#pragma unroll
                    for (int s = 0; s < S; s++)
                        if (hp[s] > 0.0) {
                            const double hX = -2.0 * fma(ta[s], eca[s], tb[s] * esa[s]);
                            const double hY = -2.0 * fma(ta[s], esa[s], -(tb[s] * eca[s]));
                            gX[s] = fma(cg, hX, gX[s]);
                            gY[s] = fma(cg, hY, gY[s]);
                        }
                }
            }
            }
        }
        pen_out = pen;
        PROF_MARK(3);
        Pen pn;
        pn.c = zero_c ? 0.0 : hdr(H_PENC);
        pn.hc = 0.5 * pn.c;
        pn.inv_c = zero_c ? 1.0 : hdr(H_PINV);

        // ---- stage cost (:84-86), acceleration cost and ALM term (:157-171)
        const double inv_ts = hdr(H_INVTS);
        const double xref = hdr(H_XREF), yref = hdr(H_YREF), thref = hdr(H_THREF);
        const double w_rv = hdr(H_RV), w_rw = hdr(H_RW), w_qv = hdr(H_QV), w_q = hdr(H_Q), w_qth = hdr(H_QTH);
        const double w_ap = hdr(H_AP), w_wp = hdr(H_WP), w_qN = hdr(H_QN), w_qthN = hdr(H_QTHN);
        double cl[S], Aa[S], Aw[S], vref[S];
        {
            double vp0, wp0;
            prev_controls(uv, vp0, wp0);
#pragma unroll
            for (int s = 0; s < S; s++) {
                const double v = uv[s].x, w = uv[s].y;
                const double vp = (s == 0) ? vp0 : uv[s > 0 ? s - 1 : 0].x, wp_ = (s == 0) ? wp0 : uv[s > 0 ? s - 1 : 0].y;
                const double2 yl = lds2(la + V_YL * vstride + 16u * G * s);
                double c0 = w_rv * (v * v);
                c0 = fma(w_rw, w * w, c0);
                vref[s] = lds1(a_vref() + 8u * tix(s));
                const double dv = v - vref[s];
                c0 = fma(w_qv, dv * dv, c0);
                const double ex = ((s == 0) ? xp0 : X[s > 0 ? s - 1 : 0]) - xref, ey = ((s == 0) ? yp0 : Y[s > 0 ? s - 1 : 0]) - yref;
                const double et = ((s == 0) ? thp0 : TH[s > 0 ? s - 1 : 0]) - thref;
                c0 = fma(w_q, fma(ex, ex, ey * ey), c0);
                c0 = fma(w_qth, et * et, c0);
                c0 = fma(qcte, mind2[s], c0);
                const double acc = (v - vp) * inv_ts, aac = (w - wp_) * inv_ts;
                c0 = fma(w_ap, acc * acc, c0);
                c0 = fma(w_wp, aac * aac, c0);
                const double za = fma(yl.x, pn.inv_c, acc), zw = fma(yl.y, pn.inv_c, aac);
                const double da = sel_excess(za, cfg.lin_acc_min, cfg.lin_acc_max);
                const double dw = sel_excess(zw, -cfg.ang_acc_max, cfg.ang_acc_max);
                c0 = fma(pn.hc, fma(da, da, dw * dw), c0);
                cl[s] = act(s) ? c0 : 0.0;
                Aa[s] = act(s) ? fma(pn.c, da, (2.0 * w_ap) * acc) * inv_ts : 0.0;
                Aw[s] = act(s) ? fma(pn.c, dw, (2.0 * w_wp) * aac) * inv_ts : 0.0;
            }
        }
        // terminal cost at t = N-1 (:148)
        const int glN = (N - 1) / S, sN = (N - 1) - glN * S;
        double XN = X[0], YN = Y[0], TN = TH[0];
#pragma unroll
        for (int s = 1; s < S; s++)
            if (s == sN) {
                XN = X[s];
                YN = Y[s];
                TN = TH[s];
            }
        XN = __shfl_sync(FULL, XN, glN, G);
        YN = __shfl_sync(FULL, YN, glN, G);
        TN = __shfl_sync(FULL, TN, glN, G);
        const double eXN = XN - xref, eYN = YN - yref, eTN = TN - thref;
        const double term = fma(w_qN, fma(eXN, eXN, eYN * eYN), w_qthN * (eTN * eTN));

        // ---- backward sweep: position adjoints are suffix sums; the cost sum rides in the same shuffle stages
        double mth[S];
#pragma unroll
        for (int s = 0; s < S; s++) {
            const bool last = !(tix(s) + 1 < N);
            const double qq = last ? w_qN : w_q, qt = last ? w_qthN : w_qth;
            gX[s] = act(s) ? fma(2.0 * qq, X[s] - xref, gX[s]) : 0.0;
            gY[s] = act(s) ? fma(2.0 * qq, Y[s] - yref, gY[s]) : 0.0;
            mth[s] = act(s) ? (2.0 * qt) * (TH[s] - thref) : 0.0;
        }
        double LX[S], LY[S], csum = cl[0];
#pragma unroll
        for (int s = 1; s < S; s++) csum = csum + cl[s];
        {
            double dXs[S], dYs[S];
#pragma unroll
            for (int s = S - 1; s >= 0; s--) {
                dXs[s] = (s == S - 1) ? gX[s] : dXs[s < S - 1 ? s + 1 : s] + gX[s];
                dYs[s] = (s == S - 1) ? gY[s] : dYs[s < S - 1 ? s + 1 : s] + gY[s];
            }
            double EX, EY;
            gscan_down_excl2_sum<G>(dXs[0], dYs[0], gl, EX, EY, csum);
#pragma unroll
            for (int s = 0; s < S; s++) {
                LX[s] = EX + dXs[s];
                LY[s] = EY + dYs[s];
            }
        }
        const double psi = fma(pn.hc, pen, csum + term);
        PROF_MARK(4);
        double nn[S], TT[S];
#pragma unroll
        for (int s = 0; s < S; s++) nn[s] = act(s) ? (ts * uv[s].x) * fma(cs[s], LY[s], -(sn[s] * LX[s])) : 0.0;
        {
            double nnx = __shfl_down_sync(FULL, nn[0], 1, G);  // the next lane's first step
            if (gl == G - 1) nnx = 0.0;
            double dT[S];
#pragma unroll
            for (int s = S - 1; s >= 0; s--) {
                const double nx = (s == S - 1) ? nnx : nn[s < S - 1 ? s + 1 : s];
                const double rr = act(s) ? mth[s] + nx : 0.0;
                dT[s] = (s == S - 1) ? rr : dT[s < S - 1 ? s + 1 : s] + rr;
            }
            const double ET = gscan_down_excl<G>(dT[0], gl);
#pragma unroll
            for (int s = 0; s < S; s++) TT[s] = ET + dT[s];
        }
        {
            double Anx = __shfl_down_sync(FULL, Aa[0], 1, G), Wnx = __shfl_down_sync(FULL, Aw[0], 1, G);
            if (gl == G - 1) {
                Anx = 0.0;
                Wnx = 0.0;
            }
#pragma unroll
            for (int s = 0; s < S; s++) {
                const double v = uv[s].x, w = uv[s].y;
                const double An = (s == S - 1) ? Anx : Aa[s < S - 1 ? s + 1 : s], Wn = (s == S - 1) ? Wnx : Aw[s < S - 1 ? s + 1 : s];
                const double lv = fma(2.0 * w_rv, v, (2.0 * w_qv) * (v - vref[s])) + (Aa[s] - An);
                const double lw = (2.0 * w_rw) * w + (Aw[s] - Wn);
                const double gv = fma(ts, fma(cs[s], LX[s], sn[s] * LY[s]), lv);
                const double gw = fma(ts, TT[s], lw);
                gout[s] = act(s) ? make_double2(gv, gw) : make_double2(0.0, 0.0);
            }
        }
        PROF_MARK(5);
        return psi;
    }
};

// ---------------------------------------------------------------------------------
// The solver: ALM/PM outer loop around PANOC as a phase machine with one evaluation site.
// Phases that end in an evaluation set x (per group) and fall through to it; the others `continue`.
enum Phase { PH_OUTER_BEGIN, PH_INIT, PH_STEP_BEGIN, PH_A, PH_RETRY, PH_LS, PH_STEP_DONE, PH_SOLVE_END, PH_F2, PH_EXIT, PH_HELP };

__device__ __forceinline__ unsigned long long nm_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Owner side of the helper scheme, out of line (it runs in a fraction of the iterations and must not cost the solver's
// hot loop registers or instruction-cache lines): wait for the helper's results of this iteration's post, pick the first
// of its NG trials that passes the line search (trial exponents NG-1 .. 2NG-2), copy that trial's (x, grad, gradient
// step, half step) from the helper's arena over V_U, V_GRAD, V_GSTEP, V_UHALF and its (psi, lhs) into the header.
// Returns the accepted trial's index 0 .. NG-1, -1 if none passes, -2 if the helper did not answer (never seen; then
// the caller evaluates the batch itself).
template <int G, int S>
__device__ __noinline__ int take_from_helper(uint32_t la, uint32_t vstride, uint32_t a_hdr, int lane) {
    constexpr int NG = 32 / G;
    const uint32_t mbs = mbox_of((int)(threadIdx.x >> 5));
    const unsigned want = ld_acquire(mbs + MB_SEQ);  // this warp's own last post
    int polls = 0;
    while (ld_acquire(mbs + MB_DONEQ) != want) {  // every lane polls (warp-uniform)
        spin_wait(32);
        if (++polls > (1 << 22)) return -2;
    }
    __syncwarp();
    const int hw = (int)ld_acquire(mbs + MB_HELPER) - 1;
    const uint32_t hsb = help_u32(2) + (uint32_t)hw * help_u32(0);  // the helper's arena
    const uint32_t hhdr = hsb + help_u32(1) + 8u * H_HRES;
    const double rhs = lds1(a_hdr + 8u * H_RHSLS);
    int k = -1;
#pragma unroll
    for (int j = NG - 1; j >= 0; j--) {
        const double lj = lds1(hhdr + 16u * j + 8u);
        if (!(lj > rhs) || (NG - 1 + j) >= MAX_LINESEARCH_ITERATIONS) k = j;  // the first one wins
    }
    if (k < 0) return -1;
    const int gl = lane % G;
    const bool g0 = lane < G;  // group 0 stores (like Warp::st)
    const uint32_t hla = hsb + 16u * gl + (uint32_t)(4 * k) * vstride;
    const int dst[4] = {V_U, V_GRAD, V_GSTEP, V_UHALF};
#pragma unroll
    for (int v = 0; v < 4; v++) {
        double2 t[S];
#pragma unroll
        for (int s = 0; s < S; s++) t[s] = lds2(hla + (uint32_t)v * vstride + 16u * G * s);
#pragma unroll
        for (int s = 0; s < S; s++) sts2_if(la + (uint32_t)dst[v] * vstride + 16u * G * s, t[s], g0);
    }
    sts1(a_hdr + 8u * H_COST, lds1(hhdr + 16u * k));
    sts1(a_hdr + 8u * H_FBEU, lds1(hhdr + 16u * k + 8u));
    __syncwarp();
    return k;
}

// Solves the staged problem.  The decision vector lives in the arena (V_U: start point in, solution out) and the
// multipliers in V_YL; the warp-uniform solver state lives in the arena header (H_GAMMA ..., I_*), so that across an
// evaluation only a handful of registers stay live.
template <int G, int S>
// help: HELP_NONE, HELP_OWNER (the kernel has mailboxes: post the trial inputs once a helper is attached) or HELP_HELPER
// (W views the OWNER's arena; this warp's own mailbox holds the owner's slot and the last seq taken; the call returns
// when the owner retires).
__device__ int solve_problem(Warp<G, S>& W, nmpc_stats& st_out, long long* prof_out = nullptr, int help = HELP_NONE) {
    constexpr int NG = 32 / G;
#ifdef NMPC_PROFILE
    long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long tstart = clock64();
#endif
    const nmpc_config& cfg = W.cfg;
    const int lane = W.lane, grp = W.grp;
    const int mem = cfg.lbfgs_memory, mem1 = cfg.lbfgs_memory + 1;
    // warp-uniform state in the arena header: every lane stores the same value and reads back its own store
    auto sget = [&](int i) { return lds1(W.a_hdr + 8u * i); };
    auto sput = [&](int i, double v) { sts1(W.a_hdr + 8u * i, v); };
    auto iget = [&](int i) { return ldsi(W.a_hdr + 8u * H_INTS + 4u * i); };
    auto iput = [&](int i, int v) { stsi(W.a_hdr + 8u * H_INTS + 4u * i, v); };
    if (help != HELP_HELPER) {
    sput(H_AKKT, cfg.initial_tolerance);
    sput(H_F2N, 0.0);
    sput(H_DYN, 0.0);
    sput(H_F2NP, 0.0);
    sput(H_DYNP, 0.0);
    sput(H_LIP, 0.0);
    sput(H_NFPR, 0.0);
    sput(H_PENC, cfg.initial_penalty);
    sput(H_PINV, 1.0 / fmax(cfg.initial_penalty, 1.0));
    iput(I_NCOST, 0); iput(I_NGRAD, 0); iput(I_ALM, 0); iput(I_INNER, 0); iput(I_NOUTER, 0);
    iput(I_STATUS, NMPC_CONVERGED); iput(I_ISTATUS, NMPC_CONVERGED); iput(I_ITLIP, 0); iput(I_NUMIT, 0);
    }
    // the few values that stay in registers
    int iteration = 0, lb_active = 0, lb_head = 0, e0 = 0;
    // flags: 1 lb_first, 2 gfirst (group 0 of the current call evaluates u_half), 4 cont, 8 fbe_valid, 16 timed_out
    //        32 posted (a helper evaluates the second batch of trials of this iteration), 64 the kernel has mailboxes
    enum { F_LBFIRST = 1, F_GFIRST = 2, F_CONT = 4, F_FBE = 8, F_TIMEOUT = 16, F_POSTED = 32, F_MBOX = 64 };
    int flags = F_LBFIRST | F_CONT | (help == HELP_OWNER ? F_MBOX : 0);
    if (help != HELP_HELPER && cfg.max_duration_micros > 0) sput(H_TBEG, __longlong_as_double((long long)nm_globaltimer()));
    auto out_of_time = [&]() -> bool {
        if (cfg.max_duration_micros <= 0) return false;
        const unsigned long long t0 = (unsigned long long)__double_as_longlong(sget(H_TBEG));
        return nm_globaltimer() - t0 > (unsigned long long)cfg.max_duration_micros * 1000ull;
    };

    double2 x[S], g[S];  // this group's evaluation point / gradient out
    double pen = 0.0;
    bool zero_c = false;
    int phase = (help == HELP_HELPER) ? PH_HELP : PH_OUTER_BEGIN;

    auto set_gamma = [&](double gm) {  // sigma only changes with gamma: computed here, not once per iteration
        sput(H_GAMMA, gm);
        sput(H_INVG, 1.0 / gm);
        sput(H_SIGMA, (1.0 - GAMMA_L_COEFF) / (4.0 * gm));
    };
    // gradient_step() + half_step(): gs = p - gamma*grad ; uh = Proj_U(gs)
    auto grad_step_half = [&](const double2(&p)[S], const double2(&gr)[S], double2(&gs)[S], double2(&uh)[S]) {
        const double gamma = sget(H_GAMMA);
#pragma unroll
        for (int s = 0; s < S; s++) {
            gs[s].x = fma(-gamma, gr[s].x, p[s].x);
            gs[s].y = fma(-gamma, gr[s].y, p[s].y);
            uh[s].x = W.act(s) ? clampd(gs[s].x, cfg.lin_vel_min, cfg.lin_vel_max) : 0.0;
            uh[s].y = W.act(s) ? clampd(gs[s].y, -cfg.ang_vel_max, cfg.ang_vel_max) : 0.0;
        }
    };
    // per-lane partials (serial over the lane's steps), then the group sum
    auto dot = [&](const double2(&a)[S], const double2(&b)[S]) {
        double e = fma(a[0].y, b[0].y, a[0].x * b[0].x);
#pragma unroll
        for (int s = 1; s < S; s++) e = e + fma(a[s].y, b[s].y, a[s].x * b[s].x);
        return e;
    };
    auto diff2 = [&](const double2(&a)[S], const double2(&b)[S]) {
        double e = 0.0;
#pragma unroll
        for (int s = 0; s < S; s++) {
            const double d0 = a[s].x - b[s].x, d1 = a[s].y - b[s].y;
            const double t = fma(d1, d1, d0 * d0);
            e = (s == 0) ? t : e + t;
        }
        return e;
    };
    // fpr = u - u_half and its norm (returned; also kept in the header)
    auto compute_fpr = [&](const double2(&u)[S], const double2(&uh)[S], double2(&fpr)[S]) -> double {
#pragma unroll
        for (int s = 0; s < S; s++) fpr[s] = make_double2(u[s].x - uh[s].x, u[s].y - uh[s].y);
        const double nf = nm_sqrt(gsum<G>(dot(fpr, fpr)));
        sput(H_NFPR, nf);
        return nf;
    };
    // Lipschitz test of PANOC's step size (update_lipschitz_constant): true = the step size must be halved
    auto lip_test_fails = [&](double cost_half) -> bool {
        const double cost = sget(H_COST), nf = sget(H_NFPR);
        const double rhs = cost + LIPSCHITZ_UPDATE_EPSILON * fabs(cost) - sget(H_IP) + (GAMMA_L_COEFF * 0.5 * sget(H_INVG)) * (nf * nf);
        return cost_half > rhs && iget(I_ITLIP) < MAX_LIPSCHITZ_UPDATE_ITERATIONS && sget(H_LIP) < MAX_LIPSCHITZ_CONSTANT;
    };
    // halve gamma, drop the L-BFGS memory, recompute the half step; every group then evaluates psi(u_half)
    auto lip_halve = [&]() {
        lb_active = 0;
        flags = (flags | F_LBFIRST) & ~(F_FBE | F_POSTED);
        sput(H_LIP, sget(H_LIP) * 2.0);
        set_gamma(sget(H_GAMMA) / 2.0);
        double2 u[S], gr[S], gs[S], uh[S];
        W.ld(V_U, u);
        W.ld(V_GRAD, gr);
        grad_step_half(u, gr, gs, uh);
        __syncwarp();
        W.st(V_GSTEP, gs);
        W.st(V_UHALF, uh);
        __syncwarp();
#pragma unroll
        for (int s = 0; s < S; s++) x[s] = uh[s];
        zero_c = false;
        phase = PH_RETRY;
    };
    // line-search trial points of a call whose groups hold the exponents e0 + grp - gfirst (tau = 2^-e, at most 2^-10)
    auto form_trials = [&]() {
        const int gfirst = (flags & F_GFIRST) ? 1 : 0;
        int e = e0 + grp - gfirst;
        e = e > MAX_LINESEARCH_ITERATIONS ? MAX_LINESEARCH_ITERATIONS : e;
        // tau = 2^-e exactly (the reference halves tau e times)
        const double tau = __longlong_as_double((long long)(1023 - (e < 0 ? 0 : e)) << 52);
        const double om = 1.0 - tau;
        double2 u[S], fpr[S], dir[S];
        W.ld(V_U, u);
        W.ld(V_FPR, fpr);
        W.ld(V_DIR, dir);
#pragma unroll
        for (int s = 0; s < S; s++) {
            x[s].x = fma(-tau, dir[s].x, fma(-om, fpr[s].x, u[s].x));
            x[s].y = fma(-tau, dir[s].y, fma(-om, fpr[s].y, u[s].y));
        }
        if (grp < gfirst) W.ld(V_UHALF, x);  // group 0 of a first call evaluates psi(u_half)
        zero_c = false;
    };
    // right-hand side of the line search from the forward-backward envelope at u (compute_rhs_ls)
    auto rhs_from_envelope = [&]() {
        double2 gs[S], uh[S], gr[S];
        W.ld(V_GSTEP, gs);
        W.ld(V_UHALF, uh);
        W.ld(V_GRAD, gr);
        double dist2 = diff2(gs, uh), gg = dot(gr, gr);
        gsum2<G>(dist2, gg);
        const double nf = sget(H_NFPR);
        const double fbe = sget(H_COST) - (0.5 * sget(H_GAMMA)) * gg + (0.5 * dist2) * sget(H_INVG);
        sput(H_RHSLS, fbe - sget(H_SIGMA) * (nf * nf));
    };
    // update_no_linesearch() of iteration 0: u <- u_half with the cost and gradient group 0 has just evaluated there
    auto first_iteration_update = [&](double cost_half) {
        __syncwarp();
        W.st(V_GRAD, g);  // group 0 evaluated u_half
        __syncwarp();
        double2 u[S], gr[S], gs[S], uh[S];
        W.ld(V_UHALF, u);
        W.ld(V_GRAD, gr);
        sput(H_COST, cost_half);
        grad_step_half(u, gr, gs, uh);
        __syncwarp();
        W.st(V_U, u);
        W.st(V_GSTEP, gs);
        W.st(V_UHALF, uh);
        __syncwarp();
        iput(I_NGRAD, iget(I_NGRAD) + 1);
        iteration++;
        phase = PH_STEP_DONE;
    };

    // PANOCOptimizer::solve's loop condition after a step (inlined where a step ends: no trip through the phase switch)
    auto step_done = [&]() {
        if (!(flags & F_CONT)) {
            phase = PH_SOLVE_END;
            return;
        }
        const int num_iter = iget(I_NUMIT) + 1;
        iput(I_NUMIT, num_iter);
        if (!(num_iter < cfg.max_inner_iterations)) flags &= ~F_CONT;
        if (out_of_time()) {  // PANOCOptimizer::solve: time ran out
            flags |= F_TIMEOUT;
            phase = PH_SOLVE_END;
            return;
        }
        phase = PH_STEP_BEGIN;
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
        __syncwarp();  // phases exchange vectors between groups through the arena
        // ------------------------------------------------------------------ pre: pick this group's x
        switch (phase) {
            case PH_OUTER_BEGIN: {
                if (out_of_time()) {  // AlmOptimizer::solve: no time left
                    iput(I_STATUS, NMPC_NOT_CONVERGED_OUT_OF_TIME);
                    phase = PH_EXIT;
                    continue;
                }
                iput(I_NOUTER, iget(I_NOUTER) + 1);
                {
                    double2 yl[S];
                    W.ld(V_YL, yl);
#pragma unroll
                    for (int s = 0; s < S; s++) {  // project_on_set_y
                        yl[s].x = clampd(yl[s].x, -Y_SET_BOUND, Y_SET_BOUND);
                        yl[s].y = clampd(yl[s].y, -Y_SET_BOUND, Y_SET_BOUND);
                    }
                    __syncwarp();
                    W.st(V_YL, yl);
                }
                // panoc init: cost and gradient at u (group 0) and the gradient at u + h (the other groups) in one
                // call; estimate_loc_lip leaves u perturbed by h
                lb_active = 0;
                flags = (flags | F_LBFIRST) & ~F_FBE;
                iteration = 0;
                {
                    double2 u[S];
                    W.ld(V_U, u);
                    double e = 0.0;
#pragma unroll
                    for (int s = 0; s < S; s++) {
                        const double ex_ = EPSILON_LIPSCHITZ * u[s].x, ey_ = EPSILON_LIPSCHITZ * u[s].y;
                        const double hx = W.act(s) ? ((ex_ > DELTA_LIPSCHITZ) ? ex_ : DELTA_LIPSCHITZ) : 0.0;
                        const double hy = W.act(s) ? ((ey_ > DELTA_LIPSCHITZ) ? ey_ : DELTA_LIPSCHITZ) : 0.0;
                        const double t = fma(hy, hy, hx * hx);
                        e = (s == 0) ? t : e + t;
                        const double2 up = make_double2(u[s].x + hx, u[s].y + hy);
                        x[s] = (grp == 0) ? u[s] : up;
                        u[s] = up;
                    }
                    sput(H_NORMH, sqrt(gsum<G>(e)));
                    __syncwarp();
                    W.st(V_U, u);
                }
                zero_c = false;
                __syncwarp();
                phase = PH_INIT;
                break;
            }
            case PH_STEP_BEGIN: {
                flags &= ~F_POSTED;
                double2 u[S], gr[S], uh[S], fpr[S];
                W.ld(V_U, u);
                W.ld(V_GRAD, gr);
                W.ld(V_UHALF, uh);
                const double norm_fpr = compute_fpr(u, uh, fpr);
                bool exit_now = false;
                if (__builtin_expect(norm_fpr < cfg.tolerance, 0)) {
                    const double inv_gamma = sget(H_INVG);
                    double e = 0.0;
#pragma unroll
                    for (int s = 0; s < S; s++) {
                        const double p0 = iteration ? gr[s].x : 0.0, p1 = iteration ? gr[s].y : 0.0;
                        const double r0 = fma(fpr[s].x, inv_gamma, gr[s].x) - p0;
                        const double r1 = fma(fpr[s].y, inv_gamma, gr[s].y) - p1;
                        const double t = fma(r1, r1, r0 * r0);
                        e = (s == 0) ? t : e + t;
                    }
                    exit_now = sqrt(gsum<G>(e)) < sget(H_AKKT);
                }
                if (exit_now) {
                    phase = PH_SOLVE_END;
                    continue;
                }
                __syncwarp();
                W.st(V_FPR, fpr);
                iput(I_ITLIP, 0);
                // <grad, fpr> for the Lipschitz test and, when a previous (state, fpr) pair exists, the three
                // inner products of the L-BFGS update (s.y, s.s, y.y) in ONE interleaved group sum.
                // The update and the direction are computed BEFORE psi(u_half) is known: if the Lipschitz test then
                // fails (rare), lip_halve() drops the memory exactly like the reference does before its update.
                double2 q[S];
                // rho and the initial scaling of a pair accepted in THIS iteration go to the recursion in registers;
                // their stores (and the divisions behind them) stay off the recursion's critical path
                bool fresh = false;
                double rho_fresh = 0.0, lbg_fresh = 0.0;
                if (flags & F_LBFIRST) {
                    sput(H_IP, gsum<G>(dot(gr, fpr)));
                    flags &= ~F_LBFIRST;
                    W.st(V_OLDS, u);
                    W.st(V_OLDG, fpr);
                } else {
                    double2 os[S], og[S], sv[S], yv[S];
                    W.ld(V_OLDS, os);
                    W.ld(V_OLDG, og);
#pragma unroll
                    for (int s = 0; s < S; s++) {
                        sv[s] = make_double2(u[s].x - os[s].x, u[s].y - os[s].y);
                        yv[s] = make_double2(fpr[s].x - og[s].x, fpr[s].y - og[s].y);
                    }
                    // ... and <s_newest, y_new>, the Gram entry the paired two-loop recursion needs (below)
                    double2 sp[S];
                    W.ld(V_S + lb_head, sp);
                    double ys = dot(sv, yv), ss = dot(sv, sv), yy = dot(yv, yv), ip = dot(gr, fpr), sy1 = dot(sp, yv);
                    gsum5<G>(ip, ys, ss, yy, sy1);
                    sput(H_IP, ip);
                    // lbfgs update_hessian(g = fpr, state = u)
                    int tmp = lb_head + mem;
                    tmp = (tmp >= mem1) ? tmp - mem1 : tmp;
                    W.st(V_S + tmp, sv);
                    W.st(V_Y + tmp, yv);
                    const double rho_new = nm_div(1.0, ys);
                    bool accept = !(ss <= DBL_EPS || ys <= SY_EPSILON);
                    if (accept) {
                        // sqrt(<fpr, fpr>) is norm_fpr: same vector, same expression, same reduction order
                        const double lhs = nm_div(ys, ss), rhsb = CBFGS_EPSILON * norm_fpr;
                        accept = (lhs > rhsb && isfinite(lhs) && isfinite(rhsb));
                    }
                    if (accept) {
                        __syncwarp();  // the other groups have read the old (state, fpr) pair above
                        W.st(V_OLDS, u);
                        W.st(V_OLDG, fpr);
                        sts1_if(W.a_syd() + 8u * lb_head, sy1, lane == 0 && lb_active > 0);
                        lb_head = tmp;  // rotate_right(1): the staging slot becomes slot 0
                        fresh = true;
                        rho_fresh = rho_new;
                        lbg_fresh = nm_div(nm_div(1.0, rho_new), yy);
                        lb_active = (lb_active + 1 < mem) ? lb_active + 1 : mem;
                    }
                }
                __syncwarp();
                if (__builtin_expect(iteration == 0, 0)) {
                    if (fresh) {  // (cannot happen at iteration 0: the first update only remembers the pair)
                        sts1_if(W.a_rho() + 8u * lb_head, rho_fresh, lane == 0);
                        sput(H_LBG, lbg_fresh);
                    }
                    // group 0: psi and grad psi at u_half (Lipschitz test, then update_no_linesearch);
                    // the other groups: psi(u) — u was perturbed by the Lipschitz estimate, its cost is stale
#pragma unroll
                    for (int s = 0; s < S; s++) x[s] = (grp == 0) ? uh[s] : u[s];
                    zero_c = false;
                    phase = PH_A;
                    break;
                }
                // direction = H * fpr (two-loop recursion)
#ifdef NMPC_PROFILE
                const long long tl0 = clock64();
#endif
#pragma unroll
                for (int s = 0; s < S; s++) q[s] = fpr[s];
                if (lb_active > 0) {
                    // The recursion alpha_k = rho_k <s_k, q>; q -= alpha_k y_k (k = 0 newest) taken TWO pairs at a time:
                    //   alpha_{k+1} = rho_{k+1} ( <s_{k+1}, q> - alpha_k <s_{k+1}, y_k> )
                    // so both inner products of a trip use the same q (ONE interleaved group sum instead of two
                    // dependent ones) and <s_{k+1}, y_k> is the stored Gram entry syd[slot k+1].  Same for the
                    // backward loop with <y_{k-1}, s_k> = syd[slot k].  (oracle: lb_apply, contract build)
                    const uint32_t a_s0 = W.la + V_S * W.vstride, a_y0 = W.la + V_Y * W.vstride;
                    auto ldv = [&](uint32_t base, int sl, double2(&r)[S]) {
#pragma unroll
                        for (int s = 0; s < S; s++) r[s] = lds2(base + sl * W.vstride + 16u * G * s);
                    };
                    auto nxt = [&](int sl) { return (sl + 1 >= mem1) ? 0 : sl + 1; };
                    auto prv = [&](int sl) { return (sl == 0) ? mem1 - 1 : sl - 1; };
                    double2 sa[S], ya[S], sb[S], yb[S];
                    int sl = lb_head;  // physical slot of pair k; k + 1 (older) is the next slot of the ring
                    ldv(a_s0, sl, sa);
                    ldv(a_y0, sl, ya);
                    double rhoa = fresh ? rho_fresh : lds1(W.a_rho() + 8u * sl);
                    const double rho_newest = rhoa;
                    int k = 0;
#pragma unroll 1
                    for (; k + 1 < lb_active; k += 2) {
                        const int sl1 = nxt(sl);
                        ldv(a_s0, sl1, sb);
                        ldv(a_y0, sl1, yb);
                        const double rhob = lds1(W.a_rho() + 8u * sl1), g1 = lds1(W.a_syd() + 8u * sl1);
                        double pa = dot(sa, q), pb = dot(sb, q);
                        gsum2<G>(pa, pb);
                        const double al0 = rhoa * pa;
                        const double al1 = rhob * fma(-al0, g1, pb);
                        sts1_if(W.a_alpha() + 8u * k, al0, lane == 0);
                        sts1_if(W.a_alpha() + 8u * k + 8u, al1, lane == 0);
#pragma unroll
                        for (int s = 0; s < S; s++) {
                            q[s].x = fma(-al1, yb[s].x, fma(-al0, ya[s].x, q[s].x));
                            q[s].y = fma(-al1, yb[s].y, fma(-al0, ya[s].y, q[s].y));
                        }
                        sl = nxt(sl1);
                        ldv(a_s0, sl, sa);  // next trip's first pair (past the end on the last trip: a valid slot, unused)
                        ldv(a_y0, sl, ya);
                        rhoa = lds1(W.a_rho() + 8u * sl);
                    }
                    if (k < lb_active) {  // odd count: the oldest pair on its own
                        const double al = rhoa * gsum<G>(dot(sa, q));
                        sts1_if(W.a_alpha() + 8u * k, al, lane == 0);
#pragma unroll
                        for (int s = 0; s < S; s++) {
                            q[s].x = fma(-al, ya[s].x, q[s].x);
                            q[s].y = fma(-al, ya[s].y, q[s].y);
                        }
                    }
                    __syncwarp();
                    const double lb_gamma = fresh ? lbg_fresh : sget(H_LBG);
#pragma unroll
                    for (int s = 0; s < S; s++) {
                        q[s].x = q[s].x * lb_gamma;
                        q[s].y = q[s].y * lb_gamma;
                    }
                    // backward: pairs (k, k-1) from the oldest, k = lb_active-1
                    k = lb_active - 1;
                    sl = lb_head + k;
                    sl = (sl >= mem1) ? sl - mem1 : sl;
                    ldv(a_s0, sl, sa);
                    ldv(a_y0, sl, ya);
                    rhoa = (sl == lb_head) ? rho_newest : lds1(W.a_rho() + 8u * sl);
#pragma unroll 1
                    for (; k >= 1; k -= 2) {
                        const int sl1 = prv(sl);
                        ldv(a_s0, sl1, sb);
                        ldv(a_y0, sl1, yb);
                        const double rhob = (sl1 == lb_head) ? rho_newest : lds1(W.a_rho() + 8u * sl1), g1 = lds1(W.a_syd() + 8u * sl);
                        const double alk = lds1(W.a_alpha() + 8u * k), alk1 = lds1(W.a_alpha() + 8u * k - 8u);
                        double qa = dot(ya, q), qb = dot(yb, q);
                        gsum2<G>(qa, qb);
                        const double c0 = alk - rhoa * qa;
                        const double c1 = alk1 - rhob * fma(c0, g1, qb);
#pragma unroll
                        for (int s = 0; s < S; s++) {
                            q[s].x = fma(c1, sb[s].x, fma(c0, sa[s].x, q[s].x));
                            q[s].y = fma(c1, sb[s].y, fma(c0, sa[s].y, q[s].y));
                        }
                        sl = prv(sl1);
                        ldv(a_s0, sl, sa);
                        ldv(a_y0, sl, ya);
                        rhoa = (sl == lb_head) ? rho_newest : lds1(W.a_rho() + 8u * sl);
                    }
                    if (k == 0) {  // odd count: the newest pair on its own
                        const double alk = lds1(W.a_alpha());
                        const double co = alk - rhoa * gsum<G>(dot(ya, q));
#pragma unroll
                        for (int s = 0; s < S; s++) {
                            q[s].x = fma(co, sa[s].x, q[s].x);
                            q[s].y = fma(co, sa[s].y, q[s].y);
                        }
                    }
                }
                W.st(V_DIR, q);
                if (fresh) {
                    sts1_if(W.a_rho() + 8u * lb_head, rho_fresh, lane == 0);
                    sput(H_LBG, lbg_fresh);
                }
                __syncwarp();
#ifdef NMPC_PROFILE
                prof[4] += clock64() - tl0;
                prof[5]++;
#endif
                if (flags & F_MBOX) {  // a helper warp is attached: it evaluates the NG trials after this call's (V_FPR, V_DIR, V_U
                    const uint32_t mbs = mbox_of((int)(threadIdx.x >> 5));  // are in place: the barrier above)
                    if (ldsi(mbs + MB_HELPER) != 0) {
                        const unsigned nseq = (unsigned)ldsi(mbs + MB_SEQ) + 2u;
                        if (lane == 0) st_release(mbs + MB_SEQ, nseq);
                        flags |= F_POSTED;
                    }
                }
                // ONE call: group 0 evaluates psi(u_half), groups 1.. the trials tau = 1, 1/2, ...
                e0 = 0;
                flags |= F_GFIRST;
                form_trials();
                phase = PH_A;
                break;
            }
            case PH_STEP_DONE: {
                step_done();
                continue;
            }
            case PH_SOLVE_END: {
                iput(I_INNER, iget(I_INNER) + iget(I_NUMIT));
                double2 u[S];
                W.ld(V_U, u);
                bool fin = true;
#pragma unroll
                for (int s = 0; s < S; s++) fin = fin && isfinite(u[s].x) && isfinite(u[s].y);
                if (!__all_sync(FULL, fin)) {
                    iput(I_STATUS, NMPC_NOT_FINITE);
                    phase = PH_EXIT;
                    continue;
                }
                W.ld(V_UHALF, u);
                __syncwarp();
                W.st(V_U, u);
                const int inner_status = (flags & F_TIMEOUT) ? NMPC_NOT_CONVERGED_OUT_OF_TIME
                                                             : ((flags & F_CONT) ? NMPC_CONVERGED : NMPC_NOT_CONVERGED_ITERATIONS);
                iput(I_ISTATUS, inner_status);
                iput(I_STATUS, inner_status);
                // F2(u) for the outer loop (group 0) and f(u) = psi with c = 0 (group 1), should this be the end
#pragma unroll
                for (int s = 0; s < S; s++) x[s] = u[s];
                zero_c = (grp == 1);
                phase = PH_F2;
                break;
            }
            case PH_EXIT: {
#ifdef NMPC_PROFILE
                prof[6] = clock64() - tstart;
                if (prof_out && lane == 0)
                    for (int i = 0; i < 8; i++) {
                        prof_out[i] = prof[i];
                        prof_out[8 + i] = W.pt[i];
                    }
#endif
                const int status = iget(I_STATUS);
                const double c = sget(H_PENC);
                st_out.exit_status = status;
                st_out.outer_iterations = iget(I_NOUTER);
                st_out.inner_iterations = iget(I_INNER);
                st_out.last_norm_fpr = sget(H_NFPR);
                st_out.delta_y_norm_over_c = sget(H_DYNP) / c;
                st_out.f2_norm = sget(H_F2NP);
                st_out.penalty = c;
                if (status == NMPC_NOT_FINITE) st_out.cost = CUDART_NAN;
                st_out.n_cost_evals = iget(I_NCOST);
                st_out.n_grad_evals = iget(I_NGRAD);
                st_out.reserved = 0;
                return status;
            }
            case PH_HELP: {  // helper mode: wait for the owner's next post (or its retirement), then take the second batch
                const uint32_t mbm = mbox_of((int)(threadIdx.x >> 5));
                const uint32_t mbo = mbox_of(ldsi(mbm + MB_HELPER));
                const unsigned last = (unsigned)ldsi(mbm + MB_DONEQ);
                unsigned sq;
                for (;;) {  // every lane polls (warp-uniform)
                    if (ld_acquire(mbo + MB_STATE) == MB_DONE) return 0;
                    sq = ld_acquire(mbo + MB_SEQ);
                    if (sq != last) break;
                    spin_wait(64);
                }
                __syncwarp();
                stsi(mbm + MB_DONEQ, (int)sq);
                e0 = NG - 1;  // the owner's call holds psi(u_half) and the trials 0 .. NG-2
                flags &= ~F_GFIRST;
                form_trials();
                break;
            }
            default:
                break;
        }

        // ------------------------------------------------------------------ the one evaluation site
        PH_ACCOUNT(-1);
#ifdef NMPC_PROFILE
        const long long tp0 = clock64();
#endif
        const double psi = W.eval(x, zero_c, g, pen, nullptr);
#ifdef NMPC_PROFILE
        prof[0] += clock64() - tp0;
        prof[1]++;
#endif
        PH_ACCOUNT(32 + phase);
        __syncwarp();

        // ------------------------------------------------------------------ post
        switch (phase) {
            case PH_INIT: {  // group 0: cost / gradient at u; group 1: gradient at u + h -> local Lipschitz estimate
                sput(H_COST, __shfl_sync(FULL, psi, 0));
                W.st(V_GRAD, g, 0);
                W.st(V_T0, g, 1);
                __syncwarp();
                double2 u[S], gr[S], gh[S], gs[S], uh[S];
                W.ld(V_U, u);
                W.ld(V_GRAD, gr);
                W.ld(V_T0, gh);
                const double lip = sqrt(gsum<G>(diff2(gh, gr))) / sget(H_NORMH);
                sput(H_LIP, lip);
                set_gamma(GAMMA_L_COEFF / fmax(lip, MIN_L_ESTIMATE));
                grad_step_half(u, gr, gs, uh);
                W.st(V_GSTEP, gs);
                W.st(V_UHALF, uh);
                __syncwarp();
                iput(I_NGRAD, iget(I_NGRAD) + 2);
                iput(I_NUMIT, 0);
                flags = (flags | F_CONT) & ~F_TIMEOUT;
                phase = PH_STEP_BEGIN;
                break;
            }
            case PH_A: {  // first call of an iteration: group 0 holds psi(u_half), the other groups trials (or psi(u))
                const double cost_half = __shfl_sync(FULL, psi, 0);
                if (iteration == 0) sput(H_COST, __shfl_sync(FULL, psi, G));
                iput(I_NCOST, iget(I_NCOST) + 2);  // psi(u_half) and OpEn's re-evaluation of psi(u) (bit-identical to the cached cost)
                if (__builtin_expect(lip_test_fails(cost_half), 0)) {
                    lip_halve();
                    break;
                }
                if (__builtin_expect(iteration == 0, 0)) {
                    first_iteration_update(cost_half);
                    break;
                }
                if (flags & F_FBE) {
                    // the envelope at u is the accepted trial's left-hand side of the previous line search
                    // (same cost, gradient, gamma and stored gstep/uhalf: bit-identical), unless gamma changed
                    const double nf = sget(H_NFPR);
                    sput(H_RHSLS, sget(H_FBEU) - sget(H_SIGMA) * (nf * nf));
                } else {
                    rhs_from_envelope();
                }
            }  // fall through: check the trials this call evaluated
            case PH_LS: {
                double2 gs[S], uh[S];
                grad_step_half(x, g, gs, uh);
                double d2 = diff2(gs, uh), gg = dot(g, g);
                gsum2<G>(d2, gg);
                const double lhs = psi - (0.5 * sget(H_GAMMA)) * gg + (0.5 * d2) * sget(H_INVG);
                const int gfirst = (flags & F_GFIRST) ? 1 : 0;
                int e = e0 + grp - gfirst;  // this group's trial
                // linesearch(): the first trial with lhs <= rhs is kept; the trial after MAX halvings is kept anyway
                const bool ok = grp >= gfirst && (!(lhs > sget(H_RHSLS)) || e >= MAX_LINESEARCH_ITERATIONS);
                const unsigned m = __ballot_sync(FULL, ok);
                if (m) {
                    const int src = __ffs(m) - 1;  // first lane of the accepting group
                    const int k = src / G;
                    e = e0 + k - gfirst;
                    iput(I_NGRAD, iget(I_NGRAD) + (e > MAX_LINESEARCH_ITERATIONS ? MAX_LINESEARCH_ITERATIONS : e) + 1);
                    __syncwarp();
                    W.st(V_U, x, k);
                    W.st(V_GRAD, g, k);
                    W.st(V_GSTEP, gs, k);
                    W.st(V_UHALF, uh, k);
                    sput(H_COST, __shfl_sync(FULL, psi, src));
                    sput(H_FBEU, __shfl_sync(FULL, lhs, src));
                    flags |= F_FBE;
                    __syncwarp();
                    iteration++;
                    step_done();
                } else {
                    int k = -2;
                    if ((flags & F_POSTED) && gfirst) {  // the helper warp has evaluated the next NG trials meanwhile
                        flags &= ~F_POSTED;
                        k = take_from_helper<G, S>(W.la, W.vstride, W.a_hdr, lane);
                    }
                    if (k >= 0) {
                        e = NG - 1 + k;
                        iput(I_NGRAD, iget(I_NGRAD) + (e > MAX_LINESEARCH_ITERATIONS ? MAX_LINESEARCH_ITERATIONS : e) + 1);
                        flags |= F_FBE;
                        iteration++;
                        step_done();
                    } else {
                        e0 += ((k == -1) ? 2 * NG : NG) - gfirst;
                        flags &= ~F_GFIRST;
                        form_trials();
                        phase = PH_LS;
                    }
                }
                break;
            }
            case PH_HELP: {  // helper mode: this warp evaluated the owner's trials NG-1 .. 2NG-2; the results go to its OWN arena
                double2 gs[S], uh[S];
                grad_step_half(x, g, gs, uh);
                double d2 = diff2(gs, uh), gg = dot(g, g);
                gsum2<G>(d2, gg);
                const double lhs = psi - (0.5 * sget(H_GAMMA)) * gg + (0.5 * d2) * sget(H_INVG);
                const int me = (int)(threadIdx.x >> 5);
                const uint32_t osb = help_u32(2) + (uint32_t)me * help_u32(0);
                const uint32_t mla = osb + 16u * W.gl + (uint32_t)(4 * grp) * W.vstride;
#pragma unroll
                for (int s2 = 0; s2 < S; s2++) {
                    sts2(mla + 16u * G * s2, x[s2]);
                    sts2(mla + W.vstride + 16u * G * s2, g[s2]);
                    sts2(mla + 2u * W.vstride + 16u * G * s2, gs[s2]);
                    sts2(mla + 3u * W.vstride + 16u * G * s2, uh[s2]);
                }
                if (W.gl == 0) sts2(osb + help_u32(1) + 8u * H_HRES + 16u * grp, make_double2(psi, lhs));
                __syncwarp();
                const uint32_t mbm = mbox_of(me);
                if (lane == 0) st_release(mbox_of(ldsi(mbm + MB_HELPER)) + MB_DONEQ, (unsigned)ldsi(mbm + MB_DONEQ));
                break;  // phase stays PH_HELP
            }
            case PH_RETRY: {  // psi(u_half) after a halving of gamma (every group evaluated the same point)
                const double cost_half = psi;
                iput(I_NCOST, iget(I_NCOST) + 1);
                double2 u[S], uh[S], fpr[S], gr[S];
                W.ld(V_U, u);
                W.ld(V_UHALF, uh);
                W.ld(V_GRAD, gr);
                compute_fpr(u, uh, fpr);
                __syncwarp();
                W.st(V_FPR, fpr);
                sput(H_IP, gsum<G>(dot(gr, fpr)));
                iput(I_ITLIP, iget(I_ITLIP) + 1);
                if (lip_test_fails(cost_half)) {
                    lip_halve();
                    break;
                }
                // lbfgs update_hessian on the emptied buffer: remember (u, fpr)
                flags &= ~F_LBFIRST;
                W.st(V_OLDS, u);
                W.st(V_OLDG, fpr);
                if (iteration == 0) {
                    first_iteration_update(cost_half);
                    break;
                }
                W.st(V_DIR, fpr);  // empty memory: direction = fpr
                __syncwarp();
                rhs_from_envelope();
                e0 = 0;
                flags &= ~F_GFIRST;
                form_trials();
                phase = PH_LS;
                break;
            }
            case PH_F2: {  // multipliers y+ = y + c*(F1 - Proj_C(F1 + y/c)); infeasibilities; outer-loop logic
                const double f_cost = __shfl_sync(FULL, psi, G);  // group 1 evaluated with c = 0
                const int nf2 = cfg.Nobs + cfg.Ndynobs;
                const double c = sget(H_PENC), inv_ts = W.hdr(H_INVTS);
                double2 u[S], yl[S], yp[S];
                W.ld(V_U, u);
                W.ld(V_YL, yl);
                double e = 0.0;
                {
                    double vp0, wp0;
                    W.prev_controls(u, vp0, wp0);
#pragma unroll
                    for (int s = 0; s < S; s++) {
                        const double vp = (s == 0) ? vp0 : u[s > 0 ? s - 1 : 0].x, wp_ = (s == 0) ? wp0 : u[s > 0 ? s - 1 : 0].y;
                        const double wa = (u[s].x - vp) * inv_ts, ww = (u[s].y - wp_) * inv_ts;
                        double za = wa + yl[s].x / c, zw = ww + yl[s].y / c;
                        za = clampd(za, cfg.lin_acc_min, cfg.lin_acc_max);
                        zw = clampd(zw, -cfg.ang_acc_max, cfg.ang_acc_max);
                        yp[s].x = W.act(s) ? fma(c, wa - za, yl[s].x) : 0.0;
                        yp[s].y = W.act(s) ? fma(c, ww - zw, yl[s].y) : 0.0;
                        const double d0 = yp[s].x - yl[s].x, d1 = yp[s].y - yl[s].y;
                        const double t = W.act(s) ? fma(d1, d1, d0 * d0) : 0.0;
                        e = (s == 0) ? t : e + t;
                    }
                }
                const double dynp = sqrt(gsum<G>(e)), f2np = sqrt(__shfl_sync(FULL, pen, 0));
                const double akkt_tol = sget(H_AKKT);
                sput(H_DYNP, dynp);
                sput(H_F2NP, f2np);
                const int alm_iter = iget(I_ALM), num_outer = iget(I_NOUTER);
                const bool crit1 = alm_iter > 0 && dynp <= c * cfg.delta_tolerance + DBL_EPS;
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
                    if (!stall) {
                        const double cn = c * cfg.penalty_update_factor;
                        sput(H_PENC, cn);
                        sput(H_PINV, 1.0 / fmax(cn, 1.0));
                    }
                    sput(H_AKKT, fmax(akkt_tol * cfg.inner_tolerance_update, cfg.tolerance));
                    iput(I_ALM, alm_iter + 1);
                    sput(H_DYN, dynp);
                    sput(H_F2N, f2np);
                    __syncwarp();
                    W.st(V_YL, yp);
                    __syncwarp();
                    if (num_outer >= cfg.max_outer_iterations) {
                        iput(I_STATUS, NMPC_NOT_CONVERGED_ITERATIONS);
                        finished = true;
                    }
                } else if (num_outer == cfg.max_outer_iterations) {
                    iput(I_STATUS, NMPC_NOT_CONVERGED_ITERATIONS);
                }
                st_out.cost = f_cost;
                phase = finished ? PH_EXIT : PH_OUTER_BEGIN;
                break;
            }
            default:
                break;
        }
    }
}
