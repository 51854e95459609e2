/*
 * nmpc_b200.h — C ABI of the B200-native batched NMPC solver.
 *
 * This is the drop-in boundary for the one hot path of
 * wljungbergh/mpc-trajectory-generator: the per-step NMPC solve that the
 * reference delegates to an OpEn-generated solver process,
 *
 *     solution = mng.call(parameters)          src/mpc/mpc_generator.py:206
 *
 * where `mng` is `og.tcp.OptimizerTcpManager(...)` (src/path_generator.py:218-222)
 * and the optimisation problem is the one `MpcModule.build()` defines
 * (src/mpc/mpc_generator.py:66-193).  Everything behind `mng` (JSON over TCP,
 * the generated Rust crate, PANOC + ALM, the CasADi cost/gradient code) is
 * replaced by the functions below; everything in front of it (parameter
 * assembly in src/path_generator.py:290-403, A* seeding, plotting) stays in the
 * reference's Python.
 *
 * Conventions
 *   - plain C types only, caller-owned buffers, no exceptions across the ABI;
 *   - every function returns an int status (NMPC_OK == 0) and records a
 *     message retrievable with nmpc_last_error();
 *   - a handle is bound to one CUDA device and is NOT thread-safe (the
 *     reference drives one request at a time through one TCP manager);
 *   - all floating point is IEEE binary64, like the reference
 *     (Python float -> JSON -> Rust f64 -> CasADi double).
 *
 * Layouts
 *   u  (decision vector, src/mpc/mpc_generator.py:70,83)  : [v0,w0,v1,w1,...], 2*N_hor
 *   p  (parameter vector z0, src/mpc/mpc_generator.py:71-79,93-104; assembled at
 *       src/path_generator.py:378-379), length nmpc_param_len():
 *        [0:3]   x, y, theta            initial state
 *        [3:5]   v, omega               last applied input (vel_init, omega_init)
 *        [5:8]   xref, yref, thetaref   horizon-end reference (x_finish)
 *        [8:10]  (unused by the cost)
 *        [10:20] q, qv, qtheta, rv, rw, qN, qthetaN, qCTE, acc_penalty, omega_acc_penalty
 *        [20:20+N]                      vel_ref[t]
 *        next 3*Nobs                    static circles (x, y, r), zero padded
 *        next 5*Ndynobs*N               ellipses, obstacle-major then time: (x, y, rx, ry, angle)
 *        last 3*N                       reference points (x, y, theta) per step
 *   y  (ALM multipliers for F1 = [lin acc (N); ang acc (N)], src/mpc/mpc_generator.py:157-162): 2*N_hor
 */
#ifndef NMPC_B200_H
#define NMPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NMPC_ABI_VERSION 1
#define NMPC_NZ 20          /* scalar header of the parameter vector (configs/default.yaml:35) */
#define NMPC_LBFGS_MAX 10   /* opengen default lbfgs_memory */
#define NMPC_MAX_HORIZON 96 /* three 32-lane passes */

/* status codes returned by the API itself */
enum {
    NMPC_OK = 0,
    NMPC_ERR_INVALID = 1,  /* bad argument / unsupported size            */
    NMPC_ERR_CUDA = 2,     /* CUDA runtime error (message has the detail) */
    NMPC_ERR_NOMEM = 3
};

/* per-problem solver exit status (the strings are OpEn's `exit_status`, which the
 * reference compares against config.bad_exit_codes, src/path_generator.py:393,
 * configs/default.yaml:49) */
enum {
    NMPC_CONVERGED = 0,              /* "Converged"                     */
    NMPC_NOT_CONVERGED_ITERATIONS = 1, /* "NotConvergedIterations"      */
    NMPC_NOT_CONVERGED_OUT_OF_TIME = 2, /* "NotConvergedOutOfTime" (only with nmpc_config.max_duration_micros > 0) */
    NMPC_NOT_FINITE = 3              /* solver error 2000 in the reference's TCP reply:
                                        is_ok() == False (src/mpc/mpc_generator.py:215-221) */
};

/* Build-time constants of the reference's generated solver
 * (configs/default.yaml:6 "Changing these will require a rebuild") plus the
 * opengen 0.6.4 SolverConfiguration defaults the reference relies on
 * (src/mpc/mpc_generator.py:184-186 sets only tolerance and max duration). */
typedef struct nmpc_config {
    int32_t N_hor;    /* horizon length N (configs/default.yaml:7)  */
    int32_t Nobs;     /* static circle slots (configs/default.yaml:38) */
    int32_t Ndynobs;  /* dynamic ellipse slots (configs/default.yaml:39) */
    int32_t lbfgs_memory;         /* 10 */
    int32_t max_inner_iterations; /* 500 */
    int32_t max_outer_iterations; /* 10 */
    int32_t max_duration_micros;  /* 0 = no time limit (default: results stay deterministic).  > 0: OpEn's
                                     with_max_duration_micros (src/mpc/mpc_generator.py:9,186: 500 000): the solve
                                     ends with NotConvergedOutOfTime once that much device time has passed */
    int32_t reserved1;
    double ts;                    /* configs/default.yaml:18 */
    double lin_vel_min, lin_vel_max, ang_vel_max; /* set U, src/mpc/mpc_generator.py:151-153 */
    double lin_acc_min, lin_acc_max, ang_acc_max; /* set C, src/mpc/mpc_generator.py:164-168 */
    double tolerance;                 /* 1e-4, src/mpc/mpc_generator.py:185 */
    double initial_tolerance;         /* 1e-4 */
    double delta_tolerance;           /* 1e-4 */
    double inner_tolerance_update;    /* 0.1  */
    double penalty_update_factor;     /* 5.0  */
    double initial_penalty;           /* 1.0  */
    double sufficient_decrease_coeff; /* 0.1  */
} nmpc_config;

/* per-problem diagnostics; mirrors the fields of OpEn's TCP reply */
typedef struct nmpc_stats {
    int32_t exit_status;        /* same value as the status array */
    int32_t outer_iterations;   /* num_outer_iterations */
    int32_t inner_iterations;   /* num_inner_iterations (summed over outer iterations) */
    int32_t n_cost_evals;       /* psi evaluations without gradient */
    int32_t n_grad_evals;       /* psi + grad psi evaluations       */
    int32_t reserved;
    double last_norm_fpr;       /* last_problem_norm_fpr */
    double delta_y_norm_over_c; /* f1_infeasibility      */
    double f2_norm;             /* f2_norm               */
    double penalty;             /* final penalty c       */
    double cost;                /* f(u) without penalty terms */
} nmpc_stats;

typedef struct nmpc_handle nmpc_handle;

/* fill `cfg` with configs/default.yaml sizes/bounds and the opengen defaults */
void nmpc_default_config(nmpc_config* cfg);

/* length of the parameter vector for a config:
 * nz + N + 3*Nobs + 5*Ndynobs*N + 3*N  (src/mpc/mpc_generator.py:71) */
int32_t nmpc_param_len(const nmpc_config* cfg);

/* Replaces MpcModule.build() + OptimizerTcpManager(...).start()
 * (src/mpc/mpc_generator.py:66-193, src/path_generator.py:218-220):
 * binds a solver instance for `cfg` to CUDA device `device`. */
int nmpc_create(const nmpc_config* cfg, int device, nmpc_handle** out);

/* Replaces mng.kill() (src/path_generator.py:408,417; src/mpc/mpc_generator.py:220). Idempotent on NULL. */
int nmpc_destroy(nmpc_handle* h);

/* Replaces mng.ping() (src/path_generator.py:222): 0 if the device answers. */
int nmpc_ping(nmpc_handle* h);

/* Replaces mng.call(parameters) for ONE problem with the server-side state the
 * reference relies on (src/mpc/mpc_generator.py:206 passes only `p`): the decision
 * vector and the multipliers persist inside the handle between calls (first call:
 * zeros).  u_out: 2*N, stats_out nullable. */
int nmpc_call(nmpc_handle* h, const double* p, double* u_out, int32_t* exit_status, nmpc_stats* stats_out);

/* forget the persisted warm start of nmpc_call (a freshly started server) */
int nmpc_reset_warm_start(nmpc_handle* h);

/* Batched solve, HOST buffers.  Pageable buffers are staged through device scratch (copies inside the call);
 * page-locked buffers (cudaHostAlloc / cudaHostRegister, e.g. torch pin_memory) are read and written by the kernel
 * in place over PCIe — every row moves exactly once either way.
 *   P  [B, np]  row-major parameters
 *   U  [B, 2N]  in: initial guess, out: solution (the projected half step, always inside U)
 *   Y  [B, 2N]  in: initial multipliers, out: multiplier state a server would keep; nullable (zeros)
 *   status [B], stats [B] nullable */
int nmpc_solve_batch(nmpc_handle* h, int32_t B, const double* P, double* U, double* Y,
                     int32_t* status, nmpc_stats* stats);

/* Scheduling note for all batched entry points: a batch with more problems than the GPU has warp slots (148 SMs x 12)
 * is first ranked by one gradient evaluation per problem (|grad psi(u0)| predicts the iteration count) and handed to
 * the persistent warps longest-first; smaller batches are spread over the SMs.  Results never depend on the order. */

/* Batched solve, DEVICE buffers on the handle's device, asynchronous on `stream`
 * (a cudaStream_t passed as void*; NULL = CUDA's default stream, as for any cudaStream_t).
 * Same arrays as nmpc_solve_batch; the caller orders its own copies on that stream. */
int nmpc_solve_batch_device(nmpc_handle* h, int32_t B, const double* dP, double* dU, double* dY,
                            int32_t* dstatus, nmpc_stats* dstats, void* stream);

/* Parity hook: evaluate the augmented cost for B (p, u, c, y) tuples on the device:
 *   psi[B], grad[B,2N], F1[B,2N], F2[B,Nobs+Ndynobs]   (host buffers, any output nullable)
 *   psi = f + c/2 * ( dist^2_C(F1 + y/max(c,1)) + |F2|^2 ) */
int nmpc_eval_batch(nmpc_handle* h, int32_t B, const double* P, const double* U, const double* c,
                    const double* Y, double* psi, double* grad, double* F1, double* F2);

/* ---------------------------------------------------------------------------------------------
 * Fleet stepping: the receding-horizon loop of PathGenerator.run
 * (src/path_generator.py:290-403) for B robots at once, entirely on the device.  One step =
 *   assemble  the per-step parameter vector of every live robot (src/path_generator.py:293-382:
 *             closest-vertex window, reference window + goal padding, vel_ref brake ramp,
 *             dynamic-obstacle ring; helpers src/visibility/visibility.py:111-124,141-148,199-216),
 *   solve     the batch with the persisted, un-shifted warm start (u, y) of each robot — what the
 *             reference's solver process keeps between mng.call()s (src/mpc/mpc_generator.py:206),
 *   advance   apply the first control, integrate the plant (src/mpc/mpc_generator.py:223-235) and
 *             run the termination test (src/path_generator.py:397).
 * n steps are enqueued back to back on the handle's stream without a host round trip.  Fleets larger than the GPU's
 * warp slots are solved longest-first (by each robot's iteration count in the previous step): scheduling only.
 * num_steps_taken controls are applied per solve (configs/default.yaml:17; smooth_velocity.yaml uses 2); a map may
 * have 0 .. Ndynobs dynamic obstacles (unused slots are the reference's phantom unit discs,
 * src/path_generator.py:274-280, including what its flat-list rotation leaks into them, :312).
 */
typedef struct nmpc_fleet nmpc_fleet;

typedef struct nmpc_fleet_config {
    int32_t n_robots;
    int32_t max_ref;   /* capacity (points) of one robot's sampled reference, rough_ref() output */
    int32_t max_vert;  /* capacity of one robot's corner-vertex list (find_original_vertices)    */
    int32_t n_brake;   /* entries of the brake profile (get_brake_vel_ref)                       */
    int32_t n_sched;   /* rows of the dynamic-obstacle schedule; 0 = map without dynamic obstacles */
    int32_t log_steps; /* per-robot trajectory log capacity in steps (0 = no log)               */
    int32_t num_steps_taken; /* controls applied per solve, 1 .. N_hor (configs/default.yaml:17); 0 means 1 */
    int32_t n_dyn;     /* dynamic obstacles the map really has, 1 .. Ndynobs; 0 means Ndynobs (when n_sched > 0) */
    double base_speed;    /* lin_vel_max * throttle_ratio (src/path_generator.py:352)            */
    double circle_radius; /* vehicle_width/2 + vehicle_margin (src/path_generator.py:301)        */
    double goal_tol;      /* 0.05  (src/path_generator.py:397)                                   */
    double stop_tol;      /* 0.005 (src/path_generator.py:397)                                   */
    double weights[10];   /* z0[10:20], order of src/path_generator.py:226-227                   */
} nmpc_fleet_config;

int nmpc_fleet_create(nmpc_handle* h, const nmpc_fleet_config* fc, nmpc_fleet** out);
int nmpc_fleet_destroy(nmpc_fleet* f);

/* Upload the per-robot plans (HOST buffers) and reset every robot to step 0 (state = start, last input 0,
 * reference index 0, warm start zeros):
 *   n_ref[B], ref[B, max_ref, 3]   (x, y, theta) samples of rough_ref (src/mpc/mpc_generator.py:17-57);
 *                                  both NULL if nmpc_fleet_sample_refs produces them on the device
 *   n_vert[B], vert[B, max_vert, 2] original vertices of the A* corners (src/visibility/visibility.py:126-139)
 *   start[B, 3], goal[B, 3]
 *   brake_vel[n_brake], brake_dist[n_brake]     (src/path_generator.py:439-477)
 *   sched_init[N, Ndynobs, 5], sched[n_sched, Ndynobs, 5]  (x, y, rx, ry, angle) of every dynamic obstacle
 *       (columns >= n_dyn unused): entry m of an obstacle's pose sequence as the reference's ring sees it —
 *       sched_init[m] = entry m of the t=0 fill (np.linspace(0, N*ts, N) times, src/visibility/visibility.py:204),
 *       sched[m], m >= N = the (m-N)-th pose appended by the loop (src/path_generator.py:313-316: iteration c appends
 *       num_steps_taken poses at np.linspace((c*s+N-s)*ts, .. + s*ts, s)); rows 0 .. N-1 of sched are unused.
 *       The fleet refuses to step once a step would read past row n_sched-1.  Both NULL when n_sched == 0. */
int nmpc_fleet_load(nmpc_fleet* f, const int32_t* n_ref, const double* ref, const int32_t* n_vert,
                    const double* vert, const double* start, const double* goal, const double* brake_vel,
                    const double* brake_dist, const double* sched_init, const double* sched);

/* sample every robot's reference on the device instead of uploading it — rough_ref
 * (src/mpc/mpc_generator.py:17-57) walks nodes[b] = path[1:] of the robot's global plan from its start position at
 * speed v (the reference passes throttle_ratio * 1.1 * lin_vel_max, src/path_generator.py:262-264), one sample per ts.
 * Call after nmpc_fleet_load (which may then be given ref = NULL); replaces the fleet's ref / n_ref.
 *   n_nodes[B], nodes[B, max_nodes, 2] HOST; ref_out[B, max_ref, 3], n_ref_out[B] nullable HOST read-back.
 * Fails with NMPC_ERR_INVALID if a robot needs more than max_ref samples. */
int nmpc_fleet_sample_refs(nmpc_fleet* f, const int32_t* n_nodes, const double* nodes, int32_t max_nodes, double v,
                           double* ref_out, int32_t* n_ref_out);

/* run n_steps receding-horizon steps for every robot that has not terminated; synchronous */
int nmpc_fleet_step(nmpc_fleet* f, int32_t n_steps);

/* Read-back (HOST buffers, any pointer nullable):
 *   state[B,3], last_u[B,2], t[B] steps taken, idx[B] reference index, done[B] terminal flag,
 *   status[B] exit status of the last solve */
int nmpc_fleet_state(nmpc_fleet* f, double* state, double* last_u, int32_t* t, int32_t* idx, int32_t* done,
                     int32_t* status);
/* parameter rows assembled for the LAST step, P[B, np]; last solution U[B, 2N] and multipliers Y[B, 2N] */
int nmpc_fleet_last(nmpc_fleet* f, double* P, double* U, double* Y);
/* trajectory log: log[B, log_steps, 5] = (x, y, theta after the step, v, omega applied); n_logged[B] */
int nmpc_fleet_log(nmpc_fleet* f, double* log, int32_t* n_logged);

/* number of kernel launches issued through this handle since creation */
int64_t nmpc_launch_count(nmpc_handle* h);

/* cumulative device time (ms) of the last solve launched through the host-buffer entry points */
double nmpc_last_kernel_ms(nmpc_handle* h);

const char* nmpc_last_error(nmpc_handle* h);
const char* nmpc_exit_status_name(int32_t exit_status);
int32_t nmpc_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NMPC_B200_H */
