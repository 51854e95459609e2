// nmpc_fleet.cuh — device side of the fleet stepping API (include/nmpc_b200.h, "Fleet stepping").
//
// The reference assembles one robot's parameter vector per receding-horizon step with Python list
// slicing (src/path_generator.py:293-382) and integrates the plant in MpcModule.run
// (src/mpc/mpc_generator.py:223-235).  Here both run on the device for B robots: one warp per robot
// assembles its row of P straight into HBM (coalesced stores, lanes stride the row), the solve kernel
// picks the rows up from there, and one thread per robot applies the first control and advances the plant.
// The arithmetic that decides anything (arg-min distances, brake-ramp filter, termination test, Euler
// step) is written operation for operation like the reference's Python floats; the plant's sin/cos is the
// solver's own nm_sincos (<= 2 ulp from libm), see tests/test_fleet.py for what that means for parity.
#pragma once
#include "nmpc_device.cuh"

struct FleetArgs {
    nmpc_config cfg;
    nmpc_fleet_config fc;
    int np;
    // plans (read only)
    const int32_t* n_ref;
    const double* ref;      // [B, max_ref, 3]
    const int32_t* n_vert;
    const double* vert;     // [B, max_vert, 2]
    const double* goal;     // [B, 3]
    const double* brake_vel;
    const double* brake_dist;
    const double* sched_init;  // [N, Nd, 5]
    const double* sched;       // [n_sched, Nd, 5]
    // per-robot run state
    double* state;    // [B, 3]
    double* last_u;   // [B, 2]
    int32_t* t;       // steps taken
    int32_t* idx;     // reference index
    int32_t* done;    // 0 live, 1 terminal (goal reached and stopped), 2 solver failure (NotFinite)
    // solver I/O
    double* P;        // [B, np]
    const double* U;  // [B, 2N]
    const int32_t* status;
    // log
    double* log;      // [B, log_steps, 5]
    int32_t* n_logged;
};

// (distance^2, index) arg-min over the warp; ties go to the lower index (np.argmin returns the first minimum)
__device__ __forceinline__ void warp_argmin(double& d, int& i) {
#pragma unroll
    for (int off = 16; off; off >>= 1) {
        const double od = __shfl_xor_sync(FULL, d, off);
        const int oi = __shfl_xor_sync(FULL, i, off);
        if (od < d || (od == d && oi < i)) {
            d = od;
            i = oi;
        }
    }
}

// squared 2-norm the way the reference measures closeness (np.linalg.norm -> sqrt(x.x), src/visibility/visibility.py:111-124);
// sqrt is monotone, so the arg-min over d^2 is the arg-min over d except for distances that differ in the last ulp.
__device__ __forceinline__ double dist2(double ax, double ay, double bx, double by) {
    const double dx = ax - bx, dy = ay - by;
    return dx * dx + dy * dy;
}

// One warp per robot: src/path_generator.py:293-382 for the robot's current state.
__global__ void __launch_bounds__(256) fleet_assemble_kernel(const __grid_constant__ FleetArgs f) {
    const int lane = threadIdx.x & 31;
    const int b = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (b >= f.fc.n_robots) return;
    if (f.done[b]) return;
    const int N = f.cfg.N_hor, Nobs = f.cfg.Nobs, Nd = f.cfg.Ndynobs;
    double* __restrict__ p = f.P + (size_t)b * f.np;
    const double x = f.state[3 * b], y = f.state[3 * b + 1], th = f.state[3 * b + 2];
    const double lv = f.last_u[2 * b], lw = f.last_u[2 * b + 1];
    const double* __restrict__ goal = f.goal + 3 * (size_t)b;
    const double gx = goal[0], gy = goal[1];
    const int t = f.t[b];

    // ---- reference index: closest sample inside [idx - steps, idx + 5*steps) (src/path_generator.py:319-324)
    const int n = f.n_ref[b];
    const int steps = max(1, f.fc.num_steps_taken);
    const double* __restrict__ R = f.ref + (size_t)b * f.fc.max_ref * 3;
    int idx = f.idx[b];
    {
        const int lb = max(0, idx - steps), ub = min(n, idx + 5 * steps);
        double d = CUDART_INF;
        int i = 0x7fffffff;
        for (int k = lane; lb + k < ub; k += 32) {  // ascending k per lane + strict '<': the first minimum wins
            const double dk = dist2(x, y, R[3 * (lb + k)], R[3 * (lb + k) + 1]);
            if (dk < d) {
                d = dk;
                i = k;
            }
        }
        warp_argmin(d, i);
        idx = lb + i;
        if (lane == 0) f.idx[b] = idx;
    }

    // ---- header: state, last input (twice), horizon-end reference, weights (src/path_generator.py:378-379)
    if (lane < 3) p[lane] = (lane == 0) ? x : ((lane == 1) ? y : th);
    if (lane == 3 || lane == 8) p[lane] = lv;
    if (lane == 4 || lane == 9) p[lane] = lw;
    if (lane >= 5 && lane < 8) {
        const int c = lane - 5;
        p[lane] = (idx + N >= n) ? goal[c] : R[3 * (idx + N) + c];  // x_finish (:336-344)
    }
    if (lane >= 10 && lane < 20) p[lane] = f.fc.weights[lane - 10];

    // ---- vel_ref with the brake ramp (src/path_generator.py:352-367)
    {
        double* __restrict__ vr = p + NMPC_NZ;
        const double base = f.fc.base_speed;
        const int nbk = f.fc.n_brake;
        if ((double)(idx + N) >= (double)n - f.brake_dist[0] / base) {
            const int nb = min(n - idx - 1, N);
            if (nb == 0) {
                // only brake entries whose stopping distance fits the remaining distance, order kept
                const double ddx = x - gx, ddy = y - gy;
                const double d = sqrt(ddx * ddx + ddy * ddy);
                int out = 0;
                for (int k0 = 0; k0 < nbk; k0 += 32) {
                    const int k = k0 + lane;
                    const bool keep = (k < nbk) && (f.brake_dist[k] <= d);
                    const unsigned m = __ballot_sync(FULL, keep);
                    const int pos = out + __popc(m & ((1u << lane) - 1u));
                    if (keep && pos < N) vr[pos] = f.brake_vel[k];
                    out += __popc(m);
                }
                for (int j = min(out, N) + lane; j < N; j += 32) vr[j] = 0.0;
            } else {
                const int nramp = min(nbk, N - nb);
                for (int j = lane; j < N; j += 32)
                    vr[j] = (j < nb) ? base : ((j < nb + nramp) ? f.brake_vel[j - nb] : 0.0);
            }
        } else {
            for (int j = lane; j < N; j += 32) vr[j] = base;
        }
    }

    // ---- static circles: the corner vertices closest to the robot (src/path_generator.py:299-304,
    //      find_closest_vertices with its slice quirk vert[idx:Nobs], src/visibility/visibility.py:141-148)
    {
        double* __restrict__ pc = p + NMPC_NZ + N;
        const int nv = f.n_vert[b];
        const double* __restrict__ V = f.vert + (size_t)b * f.fc.max_vert * 2;
        int first = 0, cnt = nv;
        if (nv > Nobs) {
            double d = CUDART_INF;
            int i = 0x7fffffff;
            for (int k = lane; k < nv; k += 32) {
                const double dk = dist2(x, y, V[2 * k], V[2 * k + 1]);
                if (dk < d) {
                    d = dk;
                    i = k;
                }
            }
            warp_argmin(d, i);
            first = i;
            cnt = max(0, min(nv, Nobs) - first);
        }
        for (int k = lane; k < Nobs; k += 32) {
            const bool on = k < cnt;
            pc[3 * k] = on ? V[2 * (first + k)] : 0.0;
            pc[3 * k + 1] = on ? V[2 * (first + k) + 1] : 0.0;
            pc[3 * k + 2] = on ? f.fc.circle_radius : 0.0;
        }
    }

    // ---- dynamic ellipses (src/path_generator.py:306-316).  The reference keeps ONE flat list for all obstacle
    //      slots, rotates it left by `steps` entries per iteration and overwrites the tail of every REAL obstacle's
    //      block with its newest poses.  In closed form, with t = plant steps taken so far:
    //        real obstacle k < n_dyn : slot j holds schedule entry m = t + j of obstacle k (the t=0 fill for m < N,
    //                                  the appended entries after);
    //        unused slots k >= n_dyn : phantom unit discs (0,0,1,1,0) (:274-280) — except that the rotation feeds the
    //                                  entries leaving the front of obstacle 0's block into the END of the flat list:
    //                                  position q = (k - n_dyn) N + j of the unused region holds entry q + t - Lp of
    //                                  obstacle 0 once that is >= 0 (Lp = (Nd - n_dyn) N).  Maps without dynamic
    //                                  obstacles rotate identical phantoms: nothing changes.
    {
        double* __restrict__ pe = p + NMPC_NZ + N + 3 * Nobs;
        const int ne = Nd * N;
        const int n_dyn = (f.fc.n_sched == 0) ? 0 : ((f.fc.n_dyn > 0) ? min(f.fc.n_dyn, Nd) : Nd);
        const int Lp = (Nd - n_dyn) * N;
        for (int i = lane; i < ne; i += 32) {
            const int k = i / N, j = i - k * N;
            double* __restrict__ e = pe + 5 * (size_t)i;  // obstacle-major, then time
            int m = -1, col = 0;
            if (k < n_dyn) {
                m = t + j;
                col = k;
            } else if (n_dyn > 0) {
                m = (k - n_dyn) * N + j + t - Lp;
            }
            if (m < 0) {
                e[0] = 0.0; e[1] = 0.0; e[2] = 1.0; e[3] = 1.0; e[4] = 0.0;
            } else {
                const int mm = min(m, f.fc.n_sched - 1);  // (the host refuses to step past the schedule)
                const double* __restrict__ s = (m < N) ? f.sched_init + 5 * ((size_t)m * Nd + col) : f.sched + 5 * ((size_t)mm * Nd + col);
#pragma unroll
                for (int c = 0; c < 5; c++) e[c] = s[c];
            }
        }
    }

    // ---- reference window, padded with the goal pose past the end of the path (src/path_generator.py:336-350,369-373)
    {
        double* __restrict__ pr = p + NMPC_NZ + N + 3 * Nobs + 5 * Nd * N;
        for (int i = lane; i < 3 * N; i += 32) {
            const int j = i / 3, c = i - 3 * j;
            pr[i] = (idx + j < n) ? R[3 * (idx + j) + c] : goal[c];
        }
    }
}

// One thread per robot: MpcModule.run's apply + plant step (src/mpc/mpc_generator.py:223-235) and the
// termination test of the loop (src/path_generator.py:397).
__global__ void __launch_bounds__(256) fleet_advance_kernel(const __grid_constant__ FleetArgs f) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= f.fc.n_robots) return;
    if (f.done[b]) return;
    if (f.status[b] == NMPC_NOT_FINITE) {  // reference: RuntimeError("MPC Solver error") ends the run
        f.done[b] = 2;
        return;
    }
    const int N2 = 2 * f.cfg.N_hor;
    const int steps = max(1, f.fc.num_steps_taken);  // the first `steps` controls of the reply are applied
    double x = f.state[3 * b], y = f.state[3 * b + 1], th = f.state[3 * b + 2];
    const double ts = f.cfg.ts;
    int t = f.t[b];
    double v = 0.0, w = 0.0;
    for (int i = 0; i < steps; i++, t++) {
        v = f.U[(size_t)b * N2 + 2 * i];
        w = f.U[(size_t)b * N2 + 2 * i + 1];
        double s, c;
        nm_sincos(th, s, c);
        x = x + ts * (v * c);
        y = y + ts * (v * s);
        th = th + ts * w;
        if (f.log && t < f.fc.log_steps) {
            double* l = f.log + ((size_t)b * f.fc.log_steps + t) * 5;
            l[0] = x; l[1] = y; l[2] = th; l[3] = v; l[4] = w;
            f.n_logged[b] = t + 1;
        }
    }
    f.state[3 * b] = x;
    f.state[3 * b + 1] = y;
    f.state[3 * b + 2] = th;
    f.last_u[2 * b] = v;
    f.last_u[2 * b + 1] = w;
    f.t[b] = t;
    const double* goal = f.goal + 3 * (size_t)b;
    // np.allclose(states[-3:-1], end[0:2], atol=0.05, rtol=0) and abs(system_input[-2]) < 0.005
    if (fabs(x - goal[0]) <= f.fc.goal_tol && fabs(y - goal[1]) <= f.fc.goal_tol && fabs(v) < f.fc.stop_tol) f.done[b] = 1;
}

// One thread per robot: rough_ref (src/mpc/mpc_generator.py:17-57) — walk the waypoint list at speed v and emit one
// (x, y, heading) sample per sampling interval.  Same statements in the same order as the reference's loop; hypot and
// atan2 are CUDA's (<= 2 ulp from the reference's libm), so samples agree to the last bits, not bit for bit.
struct SampleArgs {
    int B, max_nodes, max_ref;
    double v, ts;
    const int32_t* n_nodes;
    const double* nodes;   // [B, max_nodes, 2] = path[1:] of the global plan
    const double* start;   // [B, 3] (the fleet's state array at step 0)
    double* ref;           // [B, max_ref, 3]
    int32_t* n_ref;        // samples produced (may exceed max_ref: the caller checks)
};

__global__ void __launch_bounds__(128) fleet_sample_refs_kernel(const __grid_constant__ SampleArgs a) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const double* __restrict__ nd = a.nodes + (size_t)b * a.max_nodes * 2;
    double* __restrict__ out = a.ref + (size_t)b * a.max_ref * 3;
    const int nn = a.n_nodes[b];
    double x = a.start[3 * b], y = a.start[3 * b + 1];
    int i = 0, n = 0;
    double tx = nd[0], ty = nd[1];
    double x_dir = 0.0, y_dir = 0.0;
    bool traveling = nn > 0;
    while (traveling) {
        double t = a.ts;
        while (t > 0.0) {
            const double dist = hypot(tx - x, ty - y);
            if (dist == 0.0) {
                traveling = false;
                break;
            }
            x_dir = (tx - x) / dist;
            y_dir = (ty - y) / dist;
            const double time = dist / a.v;
            if (time > t) {
                x = x + x_dir * a.v * t;
                y = y + y_dir * a.v * t;
                t = 0.0;
            } else {
                x = x + x_dir * a.v * time;
                y = y + y_dir * a.v * time;
                t = t - time;
                i++;
                if (i > nn - 1) {
                    traveling = false;
                    break;
                }
                tx = nd[2 * i];
                ty = nd[2 * i + 1];
            }
        }
        if (n < a.max_ref) {
            out[3 * n] = x;
            out[3 * n + 1] = y;
            out[3 * n + 2] = atan2(y_dir, x_dir);
        }
        n++;
    }
    a.n_ref[b] = n;
}

// Longest-first order for the next step's solve (scheduling only, results do not depend on it): a robot's inner
// iteration count of the previous step predicts the next one's, and starting the long solves first shortens the tail
// of a fleet larger than the machine's warp slots.  Counting sort over 256 buckets of 20 iterations, descending.
struct OrderArgs {
    int B;
    const nmpc_stats* stats;
    const int32_t* done;
    int32_t* hist;    // [256], zeroed by the caller
    int32_t* order;   // [B]
};
__device__ __forceinline__ int order_bucket(const OrderArgs& a, int b) {
    if (a.done[b]) return 255;
    const int it = a.stats[b].inner_iterations;
    const int k = it / 20;
    return 254 - (k > 254 ? 254 : k);
}
__global__ void __launch_bounds__(256) fleet_order_hist_kernel(const __grid_constant__ OrderArgs a) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < a.B) atomicAdd(&a.hist[order_bucket(a, b)], 1);
}
__global__ void __launch_bounds__(256) fleet_order_scan_kernel(const __grid_constant__ OrderArgs a) {
    __shared__ int sh[256];
    const int t = threadIdx.x;
    sh[t] = a.hist[t];
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
        const int v = (t >= off) ? sh[t - off] : 0;
        __syncthreads();
        sh[t] += v;
        __syncthreads();
    }
    a.hist[t] = sh[t] - a.hist[t];  // exclusive prefix: first slot of the bucket
}
__global__ void __launch_bounds__(256) fleet_order_scatter_kernel(const __grid_constant__ OrderArgs a) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < a.B) a.order[atomicAdd(&a.hist[order_bucket(a, b)], 1)] = b;
}
