// nmpc_kernels.cu — B200 (sm_100a) batched NMPC solver: kernels + the C ABI of include/nmpc_b200.h.
//
// Replaces what sits behind `mng.call(parameters)` in the reference
// (src/mpc/mpc_generator.py:206): the OpEn-generated solver for the problem that
// MpcModule.build() defines (src/mpc/mpc_generator.py:66-193).  The device code is in
// nmpc_device.cuh: one persistent warp per NMPC instance, problems pulled from an atomic queue.
// Compile with --fmad=false: the arithmetic contract places every fma explicitly.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <vector>

#include "nmpc_device.cuh"
#include "nmpc_fleet.cuh"

// Warps (problems in flight) per SM.  One CTA per SM; the cap is what the register file allows for each
// instantiation (32 * cap threads per CTA bound the registers per thread through __launch_bounds__).
#ifndef NMPC_WARPS
#define NMPC_WARPS 12  // three warps per SM sub-partition (168 registers per thread); 8 (255 registers) is 10 % slower at saturation
#endif
__host__ __device__ constexpr int warps_cap(int G, int S) {
    return (G == 8 || S <= 2) ? NMPC_WARPS : (S == 3 ? 8 : (S == 4 ? 6 : 4));
}

// load (u0, y0) of problem b into the arena (V_U, V_YL); u also stays in registers (every group holds the vector)
template <int G, int S>
__device__ __forceinline__ void load_start(const KArgs& a, Warp<G, S>& W, int b, double2 (&u)[S]) {
    const int N = W.N;
    const double* U0 = a.U + (size_t)b * 2 * N;
    const double* Y0 = a.Y ? a.Y + (size_t)b * 2 * N : nullptr;
    double2 yl[S];
#pragma unroll
    for (int s = 0; s < S; s++) {
        const int t = W.tix(s);
        u[s] = (t < N) ? *reinterpret_cast<const double2*>(U0 + 2 * t) : make_double2(0.0, 0.0);
        yl[s] = (t < N && Y0) ? make_double2(Y0[t], Y0[N + t]) : make_double2(0.0, 0.0);
    }
    W.st(V_U, u);
    W.st(V_YL, yl);
    __syncwarp();
}

#ifndef NMPC_HELP_SHARE
#define NMPC_HELP_SHARE 0  // solving warps a helper tolerates on its own scheduler (1: measured neutral to worse)
#endif
#ifndef NMPC_HELP_MAX_WAVES
#define NMPC_HELP_MAX_WAVES 2
#endif
// HELP: the instantiation with helper warps (launch_solve picks it for batches whose tail matters; a batch that keeps every
// warp busy for most of the launch runs the plain one: its solver loop is a few percent faster without the mailbox code)
template <int G, int S, bool HELP>
__global__ void __launch_bounds__(32 * warps_cap(G, S), 1) nmpc_solve_kernel(const __grid_constant__ KArgs a) {
    const nmpc_config& cfg = a.cfg;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const Lay L = make_layout(cfg.N_hor, cfg.Nobs, cfg.Ndynobs);
    const int N = cfg.N_hor;
    Warp<G, S> W(cfg, L, warp, lane);
    if (HELP) {
    if (lane == 0) {
        g_help.mb[warp].seq = 0;
        g_help.mb[warp].done = 0;
        g_help.mb[warp].state = MB_RUNNING;
        g_help.mb[warp].helper = 0;
    }
    if (threadIdx.x == 0) {
        g_help.arena_bytes = (unsigned)L.total * 8u;
        g_help.hdr_off = (unsigned)L.hdr * 8u;
        g_help.smem_base = (unsigned)__cvta_generic_to_shared(smem);
    }
    __syncthreads();
    }
    // First wave: warp w of CTA c takes problem c + gridDim.x * w, so a batch smaller than the machine is spread
    // over the SMs instead of filling a few CTAs; afterwards the problems come from the atomic queue.
    // Out of problems: the warp retires and attaches itself as the helper of a warp of this CTA that is still solving and
    // has none yet — the one that has run the most iterations so far (the long solves are the ones that bound the launch).
    // It then runs the solver's helper mode through the SAME call site (the kernel holds one copy of the solver).
    bool first = true, retired = false;
    for (;;) {
        int b = 0;
        if (!retired) {
            if (first) {
                b = (int)(blockIdx.x + gridDim.x * warp);
                first = false;
            } else {
                if (lane == 0) b = (int)(gridDim.x * nwarps + atomicAdd(a.counter, 1u));
                b = __shfl_sync(FULL, b, 0);
            }
            if (b < a.B && a.order) b = a.order[b];
            if (b >= a.B) {  // queue empty
                retired = true;
                if (HELP) {
                    __syncwarp();
                    if (lane == 0) st_release(mbox_of(warp) + MB_STATE, MB_DONE);
                }
            } else if (a.skip && a.skip[b]) {
                continue;
            }
        }
        if (retired) {
            if (!HELP) break;
            // Lane w looks at warp slot w; every step of the search is warp-uniform.  A helper must not take issue slots from
            // a warp that is still solving: it attaches only once no such warp is left on its own sub-partition (slots
            // w, w+4, w+8 share one scheduler), and waits — asleep — until then.
            int ow = -1;
            unsigned seq0 = 0;
            for (;;) {
                const uint32_t mbw = mbox_of(lane < nwarps ? lane : 0);
                const bool running = lane < nwarps && lane != warp && ld_acquire(mbw + MB_STATE) == MB_RUNNING;
                const bool cand = running && ld_acquire(mbw + MB_HELPER) == 0u;
                const unsigned rm = __ballot_sync(FULL, running), cm = __ballot_sync(FULL, cand);
                if (!cm) break;  // nobody left to help
                if (__popc(rm & (0x11111111u << (warp & 3))) > NMPC_HELP_SHARE) {  // a solving warp shares this warp's scheduler: not yet
                    __nanosleep(4000);
                    continue;
                }
                int key = -1;
                if (cand) {
                    const uint32_t oh = g_help.smem_base + (uint32_t)lane * g_help.arena_bytes + g_help.hdr_off + 8u * H_INTS;
                    key = ldsi(oh + 4u * I_INNER) + ldsi(oh + 4u * I_NUMIT);
                    key = key < 0 ? 0 : (key > 0x00ffffff ? 0x00ffffff : key);  // (a warp staging its problem: stale counters)
                    key = (key << 5) | (31 - lane);
                }
                const int best = __reduce_max_sync(FULL, key);
                const int w = 31 - (best & 31);
                const uint32_t mbo = mbox_of(w);
                seq0 = ld_acquire(mbo + MB_SEQ);  // before attaching: the owner posts only once it sees a helper
                unsigned old = 0;
                if (lane == 0) old = cas_acq_rel(mbo + MB_HELPER, 0u, (unsigned)warp + 1u);
                old = __shfl_sync(FULL, old, 0);
                if (old == 0u) {
                    ow = w;
                    break;
                }
            }
            if (ow < 0) break;
            if (lane == 0) {  // this warp's own mailbox now describes its job: the owner's slot and the last seq taken
                g_help.mb[warp].helper = (unsigned)ow;
                g_help.mb[warp].done = seq0;
            }
            __syncwarp();
            W.rebase(ow);
        } else {
            W.stage(a.P + (size_t)b * a.np);
            double2 u0[S];
            load_start<G, S>(a, W, b, u0);
        }
        nmpc_stats st;
        st.cost = 0.0;
        const int help = !HELP ? HELP_NONE : (retired ? HELP_HELPER : HELP_OWNER);
#ifdef NMPC_PROFILE
        const int status = solve_problem<G, S>(W, st, (a.dbg && !retired) ? a.dbg + (size_t)b * 48 : nullptr, help);
#else
        const int status = solve_problem<G, S>(W, st, nullptr, help);
#endif
        __syncwarp();
        if (retired) {  // the owner has retired: look for another one
            W.rebase(warp);
            continue;
        }
        if (W.grp == 0) {
            double2 u[S], yl[S];
            W.ld(V_U, u);
            W.ld(V_YL, yl);
#pragma unroll
            for (int s = 0; s < S; s++) {
                const int t = W.tix(s);
                if (t < N) {
                    *reinterpret_cast<double2*>(a.U + (size_t)b * 2 * N + 2 * t) = u[s];
                    if (a.Y) {
                        a.Y[(size_t)b * 2 * N + t] = yl[s].x;
                        a.Y[(size_t)b * 2 * N + N + t] = yl[s].y;
                    }
                }
            }
        }
        if (lane == 0) {
            if (a.status) a.status[b] = status;
            if (a.stats) a.stats[b] = st;
        }
    }
}

// Probe for the longest-first schedule: |grad psi(u0)|^2 of every problem — ONE evaluation, a few thousandths of an
// average solve — ranks the problems by the work they will need (Spearman 0.82 with the inner-iteration count on the
// BASELINE config-2 batch).  The kernel writes a sort bucket per problem (exponent and three
// mantissa bits of the squared norm, descending) and the bucket histogram; two tiny kernels turn that into the order
// in which the solve kernel hands the problems out.  Scheduling only: results do not depend on it.
template <int G, int S>
__global__ void __launch_bounds__(32 * warps_cap(G, S), 1) nmpc_probe_kernel(const __grid_constant__ KArgs a) {
    const nmpc_config& cfg = a.cfg;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Lay L = make_layout(cfg.N_hor, cfg.Nobs, cfg.Ndynobs);
    const int wpb = blockDim.x >> 5;
    Warp<G, S> W(cfg, L, warp, lane);
    for (int b = blockIdx.x * wpb + warp; b < a.B; b += gridDim.x * wpb) {
        int bucket = PROBE_BUCKETS - 1;  // skipped rows go last
        if (!(a.skip && a.skip[b])) {
            W.stage(a.P + (size_t)b * a.np);
            double2 u[S], g[S];
            load_start<G, S>(a, W, b, u);
            sts1(W.a_hdr + 8u * H_PENC, cfg.initial_penalty);
            sts1(W.a_hdr + 8u * H_PINV, 1.0 / fmax(cfg.initial_penalty, 1.0));
            double pen;
            W.eval(u, false, g, pen, nullptr);
            double e = fma(g[0].y, g[0].y, g[0].x * g[0].x);
#pragma unroll
            for (int s = 1; s < S; s++) e = e + fma(g[s].y, g[s].y, g[s].x * g[s].x);
            const double k = gsum<G>(e);
            // exponent + 3 mantissa bits, window 2^-64 .. 2^64; NaN / inf count as hardest
            const int hi = (int)((unsigned long long)__double_as_longlong(k) >> 49) & 0x7fff;
            int idx = hi - ((1023 - 64) << 3);
            idx = idx < 0 ? 0 : (idx > PROBE_BUCKETS - 2 ? PROBE_BUCKETS - 2 : idx);
            bucket = PROBE_BUCKETS - 2 - idx;
        }
        if (lane == 0) {
            a.probe_bucket[b] = bucket;
            atomicAdd(&a.probe_hist[bucket], 1);
        }
        __syncwarp();
    }
}
__global__ void __launch_bounds__(PROBE_BUCKETS) order_scan_kernel(int32_t* hist) {
    __shared__ int sh[PROBE_BUCKETS];
    const int t = threadIdx.x;
    const int own = hist[t];
    sh[t] = own;
    __syncthreads();
    for (int off = 1; off < PROBE_BUCKETS; off <<= 1) {
        const int v = (t >= off) ? sh[t - off] : 0;
        __syncthreads();
        sh[t] += v;
        __syncthreads();
    }
    hist[t] = sh[t] - own;  // exclusive prefix: first slot of the bucket
}
__global__ void __launch_bounds__(256) order_scatter_kernel(const int32_t* bucket, int32_t* hist, int32_t* order, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) order[atomicAdd(&hist[bucket[b]], 1)] = b;
}

// parity hook: psi, grad, F1, F2 for B (p, u, c, y) tuples
template <int G, int S>
__global__ void __launch_bounds__(32 * warps_cap(G, S), 1) nmpc_eval_kernel(const __grid_constant__ KArgs a) {
    const nmpc_config& cfg = a.cfg;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Lay L = make_layout(cfg.N_hor, cfg.Nobs, cfg.Ndynobs);
    const int N = cfg.N_hor, nf2 = cfg.Nobs + cfg.Ndynobs;
    const int wpb = blockDim.x >> 5;
    Warp<G, S> W(cfg, L, warp, lane);
    for (int b = blockIdx.x * wpb + warp; b < a.B; b += gridDim.x * wpb) {
        W.stage(a.P + (size_t)b * a.np);
        double2 u[S], g[S];
        load_start<G, S>(a, W, b, u);
        double* F2g = a.F2 ? a.F2 + (size_t)b * nf2 : nullptr;
        if (F2g)
            for (int k = lane; k < nf2; k += 32) F2g[k] = 0.0;  // slots nobody is inside of report exactly 0
        __syncwarp();
        sts1(W.a_hdr + 8u * H_PENC, a.cvec[b]);
        sts1(W.a_hdr + 8u * H_PINV, 1.0 / fmax(a.cvec[b], 1.0));
        double pen;
        const double psi = W.eval(u, false, g, pen, F2g);
        if (lane == 0 && a.psi) a.psi[b] = psi;
        const double inv_ts = W.hdr(H_INVTS);
        double vp0, wp0;
        W.prev_controls(u, vp0, wp0);
        if (W.grp == 0) {
#pragma unroll
            for (int s = 0; s < S; s++) {
                const int t = W.tix(s);
                const double vp = (s == 0) ? vp0 : u[s > 0 ? s - 1 : 0].x, wp_ = (s == 0) ? wp0 : u[s > 0 ? s - 1 : 0].y;
                if (t < N) {
                    if (a.grad) {
                        a.grad[(size_t)b * 2 * N + 2 * t] = g[s].x;
                        a.grad[(size_t)b * 2 * N + 2 * t + 1] = g[s].y;
                    }
                    if (a.F1) {
                        a.F1[(size_t)b * 2 * N + t] = (u[s].x - vp) * inv_ts;
                        a.F1[(size_t)b * 2 * N + N + t] = (u[s].y - wp_) * inv_ts;
                    }
                }
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// host side: the C ABI (include/nmpc_b200.h)
// Work-queue state of one launch.  A handle keeps a small ring of these so that launches on different streams
// (double buffering, or a device-pointer call followed by a host-buffer call) never share a queue counter or the
// probe / order scratch: a context is reused only after the launch that last used it has finished (event wait on
// the new launch's stream).
#define NMPC_LAUNCH_CTXS 4
struct launch_ctx {
    unsigned int* counter;
    int32_t *pbucket, *porder, *phist;  // longest-first schedule of large batches (probe + counting sort)
    int pcap;
    cudaEvent_t done;
    bool used;
};
struct nmpc_handle {
    nmpc_config cfg;
    int device, sm_count, np, G, S, warps_per_cta;
    int help_max_waves;  // batches of up to this many waves of warp slots run the kernel with helper warps (0: never)
    size_t smem_bytes;
    cudaStream_t stream;
    launch_ctx ctx[NMPC_LAUNCH_CTXS];
    int next_ctx;
    // scratch for the host-buffer entry points
    double *dP, *dU, *dY;
    int32_t* dstatus;
    nmpc_stats* dstats;
    int cap;
    // scratch of nmpc_eval_batch
    double* ebuf;
    size_t ebuf_len;
    // nmpc_call state (what OpEn's TCP server keeps between requests)
    double *call_u, *call_y;
    int64_t launches;
    double last_ms;
    cudaEvent_t ev0, ev1;
    char err[512];
};

#ifdef NMPC_PROFILE
static long long* g_dbg = nullptr;  // profiling builds only (scratch tooling, never shipped)
#endif

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

#ifndef NMPC_AUTO_ORDER
#define NMPC_AUTO_ORDER 1  // batches larger than the warp slots: probe + longest-first order before the solve
#endif
#ifndef NMPC_FLEET_ORDER
#define NMPC_FLEET_ORDER 1  // fleets: longest-first order from the previous step's iteration counts
#endif
#ifndef NMPC_ZEROCOPY
#define NMPC_ZEROCOPY 1  // nmpc_solve_batch on page-locked host buffers: no staging copies
#endif
// one instantiation per horizon layout (nmpc_layout_for)
#define NMPC_FOR_LAYOUT(KERNEL, G, S)                                   \
    ((G) == 8 ? ((S) == 2 ? (const void*)KERNEL<8, 2> : (const void*)KERNEL<8, 3>)  \
              : ((S) == 2 ? (const void*)KERNEL<16, 2>                               \
                          : ((S) == 3 ? (const void*)KERNEL<16, 3> : ((S) == 4 ? (const void*)KERNEL<16, 4> : (const void*)KERNEL<16, 6>))))
#define NMPC_FOR_LAYOUT_H(KERNEL, G, S, H)                                        \
    ((G) == 8 ? ((S) == 2 ? (const void*)KERNEL<8, 2, H> : (const void*)KERNEL<8, 3, H>)  \
              : ((S) == 2 ? (const void*)KERNEL<16, 2, H>                                  \
                          : ((S) == 3 ? (const void*)KERNEL<16, 3, H> : ((S) == 4 ? (const void*)KERNEL<16, 4, H> : (const void*)KERNEL<16, 6, H>))))
static const void* solve_kernel_for(const nmpc_handle* h, bool help) {
    return help ? NMPC_FOR_LAYOUT_H(nmpc_solve_kernel, h->G, h->S, true) : NMPC_FOR_LAYOUT_H(nmpc_solve_kernel, h->G, h->S, false);
}
static const void* probe_kernel_for(const nmpc_handle* h) { return NMPC_FOR_LAYOUT(nmpc_probe_kernel, h->G, h->S); }
static const void* eval_kernel_for(const nmpc_handle* h) { return NMPC_FOR_LAYOUT(nmpc_eval_kernel, h->G, h->S); }

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
    nmpc_layout_for(cfg->N_hor, h->G, h->S);
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
    const size_t max_smem = prop.sharedMemPerBlockOptin - 1024;  // the kernels keep some static shared memory as well (mailboxes)
    int w = (int)(max_smem / per_warp);
    if (w < 1) {
        delete h;
        return NMPC_ERR_INVALID;  // problem too large for one warp's arena
    }
    if (w > warps_cap(h->G, h->S)) w = warps_cap(h->G, h->S);
    h->warps_per_cta = w;
    h->help_max_waves = NMPC_HELP_MAX_WAVES;
    if (const char* ev = getenv("NMPC_B200_HELP_MAX_WAVES")) h->help_max_waves = atoi(ev);  // tuning / A-B runs
    h->smem_bytes = per_warp * w;
    e = cudaFuncSetAttribute(solve_kernel_for(h, false), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(solve_kernel_for(h, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(eval_kernel_for(h), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(probe_kernel_for(h), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    for (int i = 0; i < NMPC_LAUNCH_CTXS && e == cudaSuccess; i++) {
        e = cudaMalloc(&h->ctx[i].counter, sizeof(unsigned int));
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ctx[i].done, cudaEventDisableTiming);
    }
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
    cudaDeviceSynchronize();  // launches on caller streams may still use the queue contexts
    for (int i = 0; i < NMPC_LAUNCH_CTXS; i++) {
        cudaFree(h->ctx[i].counter);
        cudaFree(h->ctx[i].pbucket); cudaFree(h->ctx[i].porder); cudaFree(h->ctx[i].phist);
        if (h->ctx[i].done) cudaEventDestroy(h->ctx[i].done);
    }
    cudaFree(h->ebuf);
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

#ifdef NMPC_DEBUG_ORDER
static const int32_t* g_dbg_order = nullptr;
static int32_t g_dbg_order_n = 0;
#endif
static int launch_solve(nmpc_handle* h, int32_t B, const double* dP, double* dU, double* dY, int32_t* dstatus,
                        nmpc_stats* dstats, cudaStream_t s, const int32_t* dskip = nullptr, const int32_t* dorder = nullptr) {
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
    a.skip = dskip;
    a.order = dorder;
    // this launch's own queue state; if the context was used before, wait (on the device) for that launch
    launch_ctx& c = h->ctx[h->next_ctx];
    h->next_ctx = (h->next_ctx + 1) % NMPC_LAUNCH_CTXS;
    if (c.used) CUDA_TRY(h, cudaStreamWaitEvent(s, c.done, 0));
    c.used = true;
    a.counter = c.counter;
#ifdef NMPC_DEBUG_ORDER
    if (!dorder && g_dbg_order && g_dbg_order_n == B) dorder = g_dbg_order, a.order = dorder;
#endif
    if (NMPC_AUTO_ORDER && !dorder && B > h->sm_count * h->warps_per_cta) {
        // more problems than warp slots: rank them with one gradient evaluation each and start the long ones first
        if (B > c.pcap) {
            // the context's previous launch has at most been waited for on stream s, not on the host: drain it
            // before its scratch is freed (only when a batch grows)
            CUDA_TRY(h, cudaEventSynchronize(c.done));
            cudaFree(c.pbucket); cudaFree(c.porder); cudaFree(c.phist);
            c.pbucket = c.porder = c.phist = nullptr; c.pcap = 0;
            CUDA_TRY(h, cudaMalloc(&c.pbucket, (size_t)B * sizeof(int32_t)));
            CUDA_TRY(h, cudaMalloc(&c.porder, (size_t)B * sizeof(int32_t)));
            CUDA_TRY(h, cudaMalloc(&c.phist, PROBE_BUCKETS * sizeof(int32_t)));
            c.pcap = B;
        }
        a.probe_bucket = c.pbucket;
        a.probe_hist = c.phist;
        CUDA_TRY(h, cudaMemsetAsync(c.phist, 0, PROBE_BUCKETS * sizeof(int32_t), s));
        int pgrid = (B + h->warps_per_cta - 1) / h->warps_per_cta;
        if (pgrid > h->sm_count) pgrid = h->sm_count;
        void* pargs[] = {&a};
        CUDA_TRY(h, cudaLaunchKernel(probe_kernel_for(h), dim3(pgrid), dim3(32 * h->warps_per_cta), pargs, h->smem_bytes, s));
        order_scan_kernel<<<1, PROBE_BUCKETS, 0, s>>>(c.phist);
        order_scatter_kernel<<<(B + 255) / 256, 256, 0, s>>>(c.pbucket, c.phist, c.porder, B);
        h->launches += 3;
        a.order = c.porder;
    }
#ifdef NMPC_PROFILE
    a.dbg = g_dbg;
#endif
    CUDA_TRY(h, cudaMemsetAsync(c.counter, 0, sizeof(unsigned int), s));
    int grid = h->sm_count;  // one persistent CTA per SM; fewer only if there are fewer problems than SMs
    if (grid > B) grid = B;
    if (grid < 1) grid = 1;
    void* args[] = {&a};
    // helper warps pay when a good part of the launch is a tail of long solves on a partly idle machine: small batches
    // (down to the single problem of nmpc_call) and batches of a few waves; a batch that keeps every warp busy does not
    const long long slots = (long long)h->sm_count * h->warps_per_cta;
    const bool help = h->help_max_waves > 0 && (long long)B <= (long long)h->help_max_waves * slots;
    CUDA_TRY(h, cudaLaunchKernel(solve_kernel_for(h, help), dim3(grid), dim3(32 * h->warps_per_cta), args, h->smem_bytes, s));
    CUDA_TRY(h, cudaEventRecord(c.done, s));
    h->launches++;
    return NMPC_OK;
}

int nmpc_solve_batch_device(nmpc_handle* h, int32_t B, const double* dP, double* dU, double* dY, int32_t* dstatus,
                            nmpc_stats* dstats, void* stream) {
    if (!h || B < 0 || !dP || !dU) return set_err(h, NMPC_ERR_INVALID, "nmpc_solve_batch_device: bad argument%s", "");
    if (B == 0) return NMPC_OK;
    CUDA_TRY(h, cudaSetDevice(h->device));
    return launch_solve(h, B, dP, dU, dY, dstatus, dstats, (cudaStream_t)stream);  // NULL = the default stream
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

// device-visible alias of a page-locked host buffer (cudaHostAlloc / cudaHostRegister / torch pin_memory), or NULL
static void* pinned_alias(const void* p) {
    if (!p) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return (at.type == cudaMemoryTypeHost) ? at.devicePointer : nullptr;
}

int nmpc_solve_batch(nmpc_handle* h, int32_t B, const double* P, double* U, double* Y, int32_t* status,
                     nmpc_stats* stats) {
    if (!h || B < 0 || !P || !U) return set_err(h, NMPC_ERR_INVALID, "nmpc_solve_batch: bad argument%s", "");
    if (B == 0) return NMPC_OK;
    CUDA_TRY(h, cudaSetDevice(h->device));
    {
        // Page-locked caller buffers are read and written by the kernel in place: every row is touched once per
        // solve (3.4 KB in, 0.7 KB out against ~10 ms of arithmetic), so the PCIe latency hides behind the other
        // warps and no staging copy sits in front of the kernel.
        void* aP = pinned_alias(P);
        void* aU = pinned_alias(U);
        void* aY = Y ? pinned_alias(Y) : nullptr;
        void* aS = status ? pinned_alias(status) : nullptr;
        void* aT = stats ? pinned_alias(stats) : nullptr;
        if (NMPC_ZEROCOPY && aP && aU && (!Y || aY) && (!status || aS) && (!stats || aT)) {
            cudaStream_t s = h->stream;
            CUDA_TRY(h, cudaEventRecord(h->ev0, s));
            // Y == NULL: the kernel starts from zero multipliers and drops the multiplier state
            int rc0 = launch_solve(h, B, (const double*)aP, (double*)aU, (double*)aY, (int32_t*)aS, (nmpc_stats*)aT, s);
            if (rc0) return rc0;
            CUDA_TRY(h, cudaEventRecord(h->ev1, s));
            CUDA_TRY(h, cudaStreamSynchronize(s));
            float ms0 = 0.f;
            CUDA_TRY(h, cudaEventElapsedTime(&ms0, h->ev0, h->ev1));
            h->last_ms = ms0;
            return NMPC_OK;
        }
    }
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
    const size_t n2 = 2 * (size_t)h->cfg.N_hor, nf2 = (size_t)h->cfg.Nobs + h->cfg.Ndynobs, np = (size_t)h->np;
    cudaStream_t s = h->stream;
    // one grow-only scratch block of the handle: P | U | Y | c | psi | grad | F1 | F2
    const size_t need = (size_t)B * (np + 4 * n2 + 2 + nf2) + 1;
    if (need > h->ebuf_len) {
        CUDA_TRY(h, cudaStreamSynchronize(s));
        cudaFree(h->ebuf);
        h->ebuf = nullptr;
        h->ebuf_len = 0;
        CUDA_TRY(h, cudaMalloc(&h->ebuf, need * sizeof(double)));
        h->ebuf_len = need;
    }
    double* dP = h->ebuf;
    double* dU = dP + (size_t)B * np;
    double* dY = dU + (size_t)B * n2;
    double* dc = dY + (size_t)B * n2;
    double* dpsi = dc + B;
    double* dgrad = dpsi + B;
    double* dF1 = dgrad + (size_t)B * n2;
    double* dF2 = dF1 + (size_t)B * n2;
    CUDA_TRY(h, cudaMemcpyAsync(dP, P, (size_t)B * np * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaMemcpyAsync(dU, U, (size_t)B * n2 * sizeof(double), cudaMemcpyHostToDevice, s));
    if (Y) CUDA_TRY(h, cudaMemcpyAsync(dY, Y, (size_t)B * n2 * sizeof(double), cudaMemcpyHostToDevice, s));
    else CUDA_TRY(h, cudaMemsetAsync(dY, 0, (size_t)B * n2 * sizeof(double), s));
    CUDA_TRY(h, cudaMemcpyAsync(dc, c, (size_t)B * sizeof(double), cudaMemcpyHostToDevice, s));
    KArgs a;
    memset(&a, 0, sizeof(a));
    a.cfg = h->cfg; a.B = B; a.np = h->np; a.P = dP; a.U = dU; a.Y = dY; a.cvec = dc;
    a.psi = dpsi; a.grad = dgrad; a.F1 = dF1; a.F2 = dF2;
    int grid = (B + h->warps_per_cta - 1) / h->warps_per_cta;
    if (grid > 8 * h->sm_count) grid = 8 * h->sm_count;
    void* args[] = {&a};
    CUDA_TRY(h, cudaLaunchKernel(eval_kernel_for(h), dim3(grid), dim3(32 * h->warps_per_cta), args, h->smem_bytes, s));
    h->launches++;
    if (psi) CUDA_TRY(h, cudaMemcpyAsync(psi, dpsi, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (grad) CUDA_TRY(h, cudaMemcpyAsync(grad, dgrad, (size_t)B * n2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (F1) CUDA_TRY(h, cudaMemcpyAsync(F1, dF1, (size_t)B * n2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (F2 && nf2) CUDA_TRY(h, cudaMemcpyAsync(F2, dF2, (size_t)B * nf2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    return NMPC_OK;
}


// ---------------------------------------------------------------------------------
// fleet stepping (include/nmpc_b200.h "Fleet stepping"; device code in nmpc_fleet.cuh)
}  // extern "C"
struct nmpc_fleet {
    nmpc_handle* h;
    nmpc_fleet_config fc;
    FleetArgs a;  // device pointers
    nmpc_stats* dstats;
    double *dU, *dY;
    int32_t* dstatus;
    int32_t *dorder, *dhist;  // longest-first order of the next solve (fleet_order_* kernels)
    bool have_order;
    bool loaded;  // plans complete (references uploaded or sampled)
    bool staged;  // nmpc_fleet_load has run
    int64_t steps_done;  // receding-horizon steps enqueued since nmpc_fleet_load (bounds the schedule rows in use)
    // grow-only scratch of nmpc_fleet_sample_refs
    int32_t* dn_nodes;
    double* dnodes;
    size_t nodes_cap;
};

template <typename T>
static cudaError_t dalloc(T** p, size_t n) { return cudaMalloc((void**)p, (n ? n : 1) * sizeof(T)); }
extern "C" {

int nmpc_fleet_create(nmpc_handle* h, const nmpc_fleet_config* fc, nmpc_fleet** out) {
    if (!h || !fc || !out) return NMPC_ERR_INVALID;
    *out = nullptr;
    if (fc->n_robots < 1 || fc->max_ref < 1 || fc->max_vert < 0 || fc->n_brake < 1 || fc->n_sched < 0 || fc->log_steps < 0 ||
        !(fc->base_speed > 0.0) || fc->num_steps_taken < 0 || fc->num_steps_taken > h->cfg.N_hor || fc->n_dyn < 0 ||
        fc->n_dyn > h->cfg.Ndynobs)
        return set_err(h, NMPC_ERR_INVALID, "nmpc_fleet_create: bad fleet config%s", "");
    CUDA_TRY(h, cudaSetDevice(h->device));
    nmpc_fleet* f = new (std::nothrow) nmpc_fleet();
    if (!f) return NMPC_ERR_NOMEM;
    memset(f, 0, sizeof(*f));
    f->h = h;
    f->fc = *fc;
    FleetArgs& a = f->a;
    a.cfg = h->cfg;
    a.fc = *fc;
    a.np = h->np;
    const size_t B = fc->n_robots, N = h->cfg.N_hor, Nd = h->cfg.Ndynobs;
    cudaError_t e = cudaSuccess;
#define TRY_(call) if (e == cudaSuccess) e = (call)
    TRY_(dalloc((int32_t**)&a.n_ref, B));
    TRY_(dalloc((double**)&a.ref, B * fc->max_ref * 3));
    TRY_(dalloc((int32_t**)&a.n_vert, B));
    TRY_(dalloc((double**)&a.vert, B * fc->max_vert * 2));
    TRY_(dalloc((double**)&a.goal, B * 3));
    TRY_(dalloc((double**)&a.brake_vel, (size_t)fc->n_brake));
    TRY_(dalloc((double**)&a.brake_dist, (size_t)fc->n_brake));
    TRY_(dalloc((double**)&a.sched_init, N * Nd * 5));
    TRY_(dalloc((double**)&a.sched, (size_t)fc->n_sched * Nd * 5));
    TRY_(dalloc(&a.state, B * 3));
    TRY_(dalloc(&a.last_u, B * 2));
    TRY_(dalloc(&a.t, B));
    TRY_(dalloc(&a.idx, B));
    TRY_(dalloc(&a.done, B));
    TRY_(dalloc(&a.P, B * h->np));
    TRY_(dalloc(&f->dU, B * 2 * N));
    TRY_(dalloc(&f->dY, B * 2 * N));
    TRY_(dalloc(&f->dstatus, B));
    TRY_(dalloc(&f->dstats, B));
    TRY_(dalloc(&f->dorder, B));
    TRY_(dalloc(&f->dhist, (size_t)256));
    if (fc->log_steps > 0) {
        TRY_(dalloc(&a.log, B * fc->log_steps * 5));
        TRY_(dalloc(&a.n_logged, B));
    }
#undef TRY_
    a.U = f->dU;
    a.status = f->dstatus;
    if (e != cudaSuccess) {
        set_err(h, NMPC_ERR_CUDA, "nmpc_fleet_create: %s", cudaGetErrorString(e));
        nmpc_fleet_destroy(f);
        return NMPC_ERR_CUDA;
    }
    *out = f;
    return NMPC_OK;
}

int nmpc_fleet_destroy(nmpc_fleet* f) {
    if (!f) return NMPC_OK;
    cudaSetDevice(f->h->device);
    cudaStreamSynchronize(f->h->stream);
    FleetArgs& a = f->a;
    cudaFree((void*)a.n_ref); cudaFree((void*)a.ref); cudaFree((void*)a.n_vert); cudaFree((void*)a.vert);
    cudaFree((void*)a.goal); cudaFree((void*)a.brake_vel); cudaFree((void*)a.brake_dist);
    cudaFree((void*)a.sched_init); cudaFree((void*)a.sched);
    cudaFree(a.state); cudaFree(a.last_u); cudaFree(a.t); cudaFree(a.idx); cudaFree(a.done); cudaFree(a.P);
    cudaFree(a.log); cudaFree(a.n_logged);
    cudaFree(f->dU); cudaFree(f->dY); cudaFree(f->dstatus); cudaFree(f->dstats);
    cudaFree(f->dorder); cudaFree(f->dhist);
    cudaFree(f->dn_nodes); cudaFree(f->dnodes);
    delete f;
    return NMPC_OK;
}

int nmpc_fleet_load(nmpc_fleet* f, const int32_t* n_ref, const double* ref, const int32_t* n_vert, const double* vert,
                    const double* start, const double* goal, const double* brake_vel, const double* brake_dist,
                    const double* sched_init, const double* sched) {
    if (!f) return NMPC_ERR_INVALID;
    nmpc_handle* h = f->h;
    const nmpc_fleet_config& fc = f->fc;
    const bool have_ref = n_ref && ref;
    if ((!n_ref != !ref) || !n_vert || (!vert && fc.max_vert > 0) || !start || !goal || !brake_vel || !brake_dist ||
        (fc.n_sched > 0 && (!sched_init || !sched)))
        return set_err(h, NMPC_ERR_INVALID, "nmpc_fleet_load: missing array%s", "");
    const size_t B = fc.n_robots, N = h->cfg.N_hor, Nd = h->cfg.Ndynobs;
    for (size_t b = 0; b < B; b++)
        if ((have_ref && (n_ref[b] < 1 || n_ref[b] > fc.max_ref)) || n_vert[b] < 0 || n_vert[b] > fc.max_vert)
            return set_err(h, NMPC_ERR_INVALID, "nmpc_fleet_load: n_ref / n_vert out of range%s", "");
    CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    FleetArgs& a = f->a;
#define UP_(dst, src, n) CUDA_TRY(h, cudaMemcpyAsync((void*)(dst), (src), (n), cudaMemcpyHostToDevice, s))
    if (have_ref) {
        UP_(a.n_ref, n_ref, B * sizeof(int32_t));
        UP_(a.ref, ref, B * fc.max_ref * 3 * sizeof(double));
    }
    UP_(a.n_vert, n_vert, B * sizeof(int32_t));
    if (fc.max_vert > 0) UP_(a.vert, vert, B * fc.max_vert * 2 * sizeof(double));
    UP_(a.goal, goal, B * 3 * sizeof(double));
    UP_(a.brake_vel, brake_vel, fc.n_brake * sizeof(double));
    UP_(a.brake_dist, brake_dist, fc.n_brake * sizeof(double));
    if (fc.n_sched > 0) {
        UP_(a.sched_init, sched_init, N * Nd * 5 * sizeof(double));
        UP_(a.sched, sched, (size_t)fc.n_sched * Nd * 5 * sizeof(double));
    }
    UP_(a.state, start, B * 3 * sizeof(double));
#undef UP_
    CUDA_TRY(h, cudaMemsetAsync(a.last_u, 0, B * 2 * sizeof(double), s));
    CUDA_TRY(h, cudaMemsetAsync(a.t, 0, B * sizeof(int32_t), s));
    CUDA_TRY(h, cudaMemsetAsync(a.idx, 0, B * sizeof(int32_t), s));
    CUDA_TRY(h, cudaMemsetAsync(a.done, 0, B * sizeof(int32_t), s));
    CUDA_TRY(h, cudaMemsetAsync(a.P, 0, B * h->np * sizeof(double), s));
    CUDA_TRY(h, cudaMemsetAsync(f->dU, 0, B * 2 * N * sizeof(double), s));
    CUDA_TRY(h, cudaMemsetAsync(f->dY, 0, B * 2 * N * sizeof(double), s));
    CUDA_TRY(h, cudaMemsetAsync(f->dstatus, 0, B * sizeof(int32_t), s));
    if (a.log) CUDA_TRY(h, cudaMemsetAsync(a.n_logged, 0, B * sizeof(int32_t), s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    f->loaded = have_ref;
    f->staged = true;
    f->have_order = false;
    f->steps_done = 0;
    return NMPC_OK;
}

int nmpc_fleet_sample_refs(nmpc_fleet* f, const int32_t* n_nodes, const double* nodes, int32_t max_nodes, double v,
                           double* ref_out, int32_t* n_ref_out) {
    if (!f) return NMPC_ERR_INVALID;
    nmpc_handle* h = f->h;
    if (!f->staged) return set_err(h, NMPC_ERR_INVALID, "nmpc_fleet_sample_refs: call nmpc_fleet_load first%s", "");
    if (!n_nodes || !nodes || max_nodes < 1 || !(v > 0.0)) return set_err(h, NMPC_ERR_INVALID, "nmpc_fleet_sample_refs: bad argument%s", "");
    const nmpc_fleet_config& fc = f->fc;
    const size_t B = fc.n_robots;
    for (size_t b = 0; b < B; b++)
        if (n_nodes[b] < 1 || n_nodes[b] > max_nodes) return set_err(h, NMPC_ERR_INVALID, "nmpc_fleet_sample_refs: n_nodes out of range%s", "");
    CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    if (B * (size_t)max_nodes > f->nodes_cap) {
        CUDA_TRY(h, cudaStreamSynchronize(s));
        cudaFree(f->dn_nodes); cudaFree(f->dnodes);
        f->dn_nodes = nullptr; f->dnodes = nullptr; f->nodes_cap = 0;
        CUDA_TRY(h, cudaMalloc(&f->dn_nodes, B * sizeof(int32_t)));
        CUDA_TRY(h, cudaMalloc(&f->dnodes, B * max_nodes * 2 * sizeof(double)));
        f->nodes_cap = B * (size_t)max_nodes;
    }
    int32_t* dn = f->dn_nodes;
    double* dnodes = f->dnodes;
    cudaError_t e = cudaMemcpyAsync(dn, n_nodes, B * sizeof(int32_t), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dnodes, nodes, B * max_nodes * 2 * sizeof(double), cudaMemcpyHostToDevice, s);
    std::vector<int32_t> got(B);
    if (e == cudaSuccess) {
        SampleArgs sa;
        sa.B = (int)B; sa.max_nodes = max_nodes; sa.max_ref = fc.max_ref; sa.v = v; sa.ts = h->cfg.ts;
        sa.n_nodes = dn; sa.nodes = dnodes; sa.start = f->a.state;
        sa.ref = (double*)f->a.ref; sa.n_ref = (int32_t*)f->a.n_ref;
        fleet_sample_refs_kernel<<<((int)B + 127) / 128, 128, 0, s>>>(sa);
        h->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(got.data(), f->a.n_ref, B * sizeof(int32_t), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && ref_out) e = cudaMemcpyAsync(ref_out, f->a.ref, B * fc.max_ref * 3 * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return set_err(h, NMPC_ERR_CUDA, "nmpc_fleet_sample_refs: %s", cudaGetErrorString(e));
    for (size_t b = 0; b < B; b++)
        if (got[b] < 1 || got[b] > fc.max_ref) {
            f->loaded = false;
            return set_err(h, NMPC_ERR_INVALID, "nmpc_fleet_sample_refs: a reference needs more than max_ref samples%s", "");
        }
    if (n_ref_out) memcpy(n_ref_out, got.data(), B * sizeof(int32_t));
    f->loaded = true;
    return NMPC_OK;
}

int nmpc_fleet_step(nmpc_fleet* f, int32_t n_steps) {
    if (!f || n_steps < 0) return NMPC_ERR_INVALID;
    nmpc_handle* h = f->h;
    if (!f->loaded) return set_err(h, NMPC_ERR_INVALID, "nmpc_fleet_step: nmpc_fleet_load has not been called%s", "");
    CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const int B = f->fc.n_robots;
    // step t reads schedule rows t .. t + N - 1 (src/path_generator.py:306-316): refuse to run past the end of
    // the uploaded schedule instead of freezing the obstacles at their last pose
    const int64_t spp = f->fc.num_steps_taken > 0 ? f->fc.num_steps_taken : 1;   // plant steps per solve
    // with unused obstacle slots the reference's rotation also reads obstacle 0's entries (index + t - Lp < t + N)
    if (f->fc.n_sched > 0 && (f->steps_done + n_steps - 1) * spp + h->cfg.N_hor > f->fc.n_sched)
        return set_err(h, NMPC_ERR_INVALID, "nmpc_fleet_step: the dynamic-obstacle schedule is too short for this many steps "
                       "(n_sched rows must cover steps + N_hor - 1)%s", "");
    CUDA_TRY(h, cudaEventRecord(h->ev0, s));
    for (int k = 0; k < n_steps; k++) {
        fleet_assemble_kernel<<<(B * 32 + 255) / 256, 256, 0, s>>>(f->a);
        h->launches++;
        // fleets larger than the machine's warp slots: hand the robots out longest-first (by the previous step)
        const bool use_order = NMPC_FLEET_ORDER && f->have_order && B > h->sm_count * h->warps_per_cta;
        int rc = launch_solve(h, B, f->a.P, f->dU, f->dY, f->dstatus, f->dstats, s, f->a.done, use_order ? f->dorder : nullptr);
        if (rc) {
            cudaStreamSynchronize(s);  // what was enqueued so far finishes before the caller sees the error
            return rc;
        }
        f->steps_done++;
        fleet_advance_kernel<<<(B + 255) / 256, 256, 0, s>>>(f->a);
        h->launches++;
        if (NMPC_FLEET_ORDER && B > h->sm_count * h->warps_per_cta) {
            OrderArgs oa;
            oa.B = B; oa.stats = f->dstats; oa.done = f->a.done; oa.hist = f->dhist; oa.order = f->dorder;
            CUDA_TRY(h, cudaMemsetAsync(f->dhist, 0, 256 * sizeof(int32_t), s));
            fleet_order_hist_kernel<<<(B + 255) / 256, 256, 0, s>>>(oa);
            fleet_order_scan_kernel<<<1, 256, 0, s>>>(oa);
            fleet_order_scatter_kernel<<<(B + 255) / 256, 256, 0, s>>>(oa);
            h->launches += 3;
            f->have_order = true;
        }
    }
    CUDA_TRY(h, cudaEventRecord(h->ev1, s));
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(s));
    float ms = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms = ms;
    return NMPC_OK;
}

#define DOWN_(dst, src, n) if (dst) CUDA_TRY(h, cudaMemcpyAsync((dst), (src), (n), cudaMemcpyDeviceToHost, s))
int nmpc_fleet_state(nmpc_fleet* f, double* state, double* last_u, int32_t* t, int32_t* idx, int32_t* done, int32_t* status) {
    if (!f) return NMPC_ERR_INVALID;
    nmpc_handle* h = f->h;
    CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const size_t B = f->fc.n_robots;
    DOWN_(state, f->a.state, B * 3 * sizeof(double));
    DOWN_(last_u, f->a.last_u, B * 2 * sizeof(double));
    DOWN_(t, f->a.t, B * sizeof(int32_t));
    DOWN_(idx, f->a.idx, B * sizeof(int32_t));
    DOWN_(done, f->a.done, B * sizeof(int32_t));
    DOWN_(status, f->dstatus, B * sizeof(int32_t));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    return NMPC_OK;
}

int nmpc_fleet_last(nmpc_fleet* f, double* P, double* U, double* Y) {
    if (!f) return NMPC_ERR_INVALID;
    nmpc_handle* h = f->h;
    CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const size_t B = f->fc.n_robots, n2 = 2 * (size_t)h->cfg.N_hor;
    DOWN_(P, f->a.P, B * h->np * sizeof(double));
    DOWN_(U, f->dU, B * n2 * sizeof(double));
    DOWN_(Y, f->dY, B * n2 * sizeof(double));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    return NMPC_OK;
}

int nmpc_fleet_log(nmpc_fleet* f, double* log, int32_t* n_logged) {
    if (!f) return NMPC_ERR_INVALID;
    nmpc_handle* h = f->h;
    if (!f->a.log) return set_err(h, NMPC_ERR_INVALID, "nmpc_fleet_log: the fleet was created with log_steps = 0%s", "");
    CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const size_t B = f->fc.n_robots;
    DOWN_(log, f->a.log, B * f->fc.log_steps * 5 * sizeof(double));
    DOWN_(n_logged, f->a.n_logged, B * sizeof(int32_t));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    return NMPC_OK;
}
#undef DOWN_

#ifdef NMPC_PROFILE
void nmpc_debug_set_buffer(void* p) { g_dbg = (long long*)p; }
#endif
#ifdef NMPC_DEBUG_ORDER
// experiments only (tools/variants.py builds): hand-out order for the next launches of any handle (device pointer, B entries)
void nmpc_debug_set_order(const int32_t* dorder, int32_t n) { g_dbg_order = dorder; g_dbg_order_n = n; }
#endif

int64_t nmpc_launch_count(nmpc_handle* h) { return h ? h->launches : 0; }
double nmpc_last_kernel_ms(nmpc_handle* h) { return h ? h->last_ms : 0.0; }

}  // extern "C"
