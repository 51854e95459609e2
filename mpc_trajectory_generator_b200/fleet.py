"""Fleet stepping: the receding-horizon loop of `PathGenerator.run` (src/path_generator.py:290-403)
for B robots at once, on the device (C ABI: the nmpc_fleet_* functions of include/nmpc_b200.h).

The global plans stay on the host (A* seed path, `rough_ref` sampling, brake profile — host.assembly /
host.planner); `FleetPlan.from_scenarios` packs them into the flat arrays the ABI takes.  Per step the
device assembles every live robot's parameter vector, solves the batch with each robot's persisted
warm start and advances the plants; `NmpcFleet.step(n)` enqueues n such steps without a host round trip.
"""
import ctypes as C
import weakref

import numpy as np

from .solver import NmpcError, _dp, _ip, _ptr, param_len


class FleetConfig(C.Structure):
    """struct nmpc_fleet_config"""
    _fields_ = [("n_robots", C.c_int32), ("max_ref", C.c_int32), ("max_vert", C.c_int32), ("n_brake", C.c_int32),
                ("n_sched", C.c_int32), ("log_steps", C.c_int32), ("num_steps_taken", C.c_int32), ("n_dyn", C.c_int32),
                ("base_speed", C.c_double), ("circle_radius", C.c_double), ("goal_tol", C.c_double),
                ("stop_tol", C.c_double), ("weights", C.c_double * 10)]


class FleetPlan:
    """Flat host arrays describing B robots' global plans (what PathGenerator.run prepares before its loop,
    src/path_generator.py:238-287)."""

    def __init__(self, n_ref, ref, n_vert, vert, start, goal, brake_vel, brake_dist, weights, base_speed,
                 circle_radius, sched_init=None, sched=None, n_nodes=None, nodes=None, ref_speed=None,
                 num_steps_taken=1, n_dyn=0):
        self.n_ref = np.ascontiguousarray(n_ref, dtype=np.int32)
        self.ref = np.ascontiguousarray(ref, dtype=np.float64)
        self.n_vert = np.ascontiguousarray(n_vert, dtype=np.int32)
        self.vert = np.ascontiguousarray(vert, dtype=np.float64)
        self.start = np.ascontiguousarray(start, dtype=np.float64)
        self.goal = np.ascontiguousarray(goal, dtype=np.float64)
        self.brake_vel = np.ascontiguousarray(brake_vel, dtype=np.float64)
        self.brake_dist = np.ascontiguousarray(brake_dist, dtype=np.float64)
        self.weights = [float(w) for w in weights]
        self.base_speed = float(base_speed)
        self.circle_radius = float(circle_radius)
        self.sched_init = None if sched_init is None else np.ascontiguousarray(sched_init, dtype=np.float64)
        self.sched = None if sched is None else np.ascontiguousarray(sched, dtype=np.float64)
        # waypoints of the global plan (path[1:]) and the sampling speed: only needed to sample the references on
        # the device (NmpcFleet(..., sample_refs_on_device=True))
        self.n_nodes = None if n_nodes is None else np.ascontiguousarray(n_nodes, dtype=np.int32)
        self.nodes = None if nodes is None else np.ascontiguousarray(nodes, dtype=np.float64)
        self.ref_speed = None if ref_speed is None else float(ref_speed)
        self.num_steps_taken = int(num_steps_taken)   # controls applied per solve (configs/default.yaml:17)
        self.n_dyn = int(n_dyn)                       # dynamic obstacles the map really has (0: none, or all slots)
        B = self.n_ref.shape[0]
        assert self.ref.shape[0] == B and self.ref.shape[2] == 3 and self.start.shape == (B, 3)
        assert self.goal.shape == (B, 3) and self.n_vert.shape == (B,) and self.vert.shape[0] == B

    @property
    def n_robots(self):
        return int(self.n_ref.shape[0])

    @classmethod
    def from_scenarios(cls, scenarios, max_steps=0):
        """Pack host.assembly.Scenario objects (all on the same map and config).  `max_steps` sizes the
        dynamic-obstacle schedule (steps the fleet may run); ignored on maps without dynamic obstacles."""
        sc0 = scenarios[0]
        cfg = sc0.cfg
        B = len(scenarios)
        N, Nd = cfg.N_hor, cfg.Ndynobs
        max_ref = max(len(s.x_ref) for s in scenarios)
        max_vert = max(1, max(len(s.vert) for s in scenarios))
        n_ref = np.zeros(B, dtype=np.int32)
        n_vert = np.zeros(B, dtype=np.int32)
        ref = np.zeros((B, max_ref, 3))
        vert = np.zeros((B, max_vert, 2))
        start = np.zeros((B, 3))
        goal = np.zeros((B, 3))
        for b, s in enumerate(scenarios):
            n = len(s.x_ref)
            n_ref[b] = n
            ref[b, :n, 0], ref[b, :n, 1], ref[b, :n, 2] = s.x_ref, s.y_ref, s.theta_ref
            # the reference fills the circle slots only on maps that have obstacles (src/path_generator.py:295):
            # corner vertices of an obstacle-free, non-convex boundary are never sent
            n_vert[b] = len(s.vert) if len(s.obstacles) else 0
            if n_vert[b]:
                vert[b, :len(s.vert)] = np.asarray(s.vert, dtype=np.float64)
            start[b] = s.start
            goal[b] = s.end
        max_nodes = max(len(s.path) - 1 for s in scenarios)
        n_nodes = np.array([len(s.path) - 1 for s in scenarios], dtype=np.int32)
        nodes = np.zeros((B, max_nodes, 2))
        for b, s in enumerate(scenarios):
            nodes[b, :n_nodes[b]] = np.asarray(s.path[1:], dtype=np.float64)
        sched_init = sched = None
        steps = int(cfg.num_steps_taken)
        n_dyn = len(sc0.dyn_obs)
        if n_dyn:
            if n_dyn > Nd:
                raise NmpcError("the map has more dynamic obstacles than the solver has slots (Ndynobs)")
            if max_steps <= 0:
                raise NmpcError("this map has dynamic obstacles: from_scenarios needs max_steps (the number of steps the "
                                "fleet may run) to size their schedule")
            # An obstacle's pose sequence as the reference's ring sees it: entries 0 .. N-1 are the t=0 fill
            # (np.linspace(0, N*ts, N), src/visibility/visibility.py:204); loop iteration c >= 1 (plant time t = c*steps)
            # appends `steps` poses at np.linspace((t+N-steps)*ts, .. + steps*ts, steps) (src/path_generator.py:313-316),
            # which become entries N + (c-1)*steps .. N + c*steps - 1.
            init = sc0._dyn_obstacles(0 * cfg.ts, N)
            sched_init = np.zeros((N, Nd, 5))
            for k, dob in enumerate(init):
                sched_init[:, k, :] = np.asarray([[float(v) for v in e] for e in dob])
            n_sched = N + max_steps * steps + steps
            sched = np.zeros((n_sched, Nd, 5))
            for c in range(1, max_steps + 2):
                t = c * steps
                for k, dob in enumerate(sc0._dyn_obstacles((t + N - steps) * cfg.ts, steps)):
                    for i in range(steps):
                        m = N + (c - 1) * steps + i
                        if m < n_sched:
                            sched[m, k, :] = [float(v) for v in dob[i]]
        return cls(n_ref, ref, n_vert, vert, start, goal, sc0.brake_vel, sc0.brake_dist, sc0.weights,
                   cfg.lin_vel_max * cfg.throttle_ratio, cfg.vehicle_width / 2 + cfg.vehicle_margin, sched_init, sched,
                   n_nodes=n_nodes, nodes=nodes, ref_speed=cfg.throttle_ratio * 1.1 * cfg.lin_vel_max,
                   num_steps_taken=steps, n_dyn=n_dyn)


class NmpcFleet:
    """B robots stepping in lock-step on one device, bound to an NmpcSolver (its config, device and stream)."""

    def __init__(self, solver, plan, log_steps=0, goal_tol=0.05, stop_tol=0.005, sample_refs_on_device=False):
        self.sample_refs_on_device = bool(sample_refs_on_device)
        self.solver = solver
        self.plan = plan
        L = solver._lib
        self._lib = L
        vp = C.c_void_p
        L.nmpc_fleet_create.argtypes = [vp, C.POINTER(FleetConfig), C.POINTER(vp)]
        L.nmpc_fleet_destroy.argtypes = [vp]
        L.nmpc_fleet_load.argtypes = [vp, _ip, _dp, _ip, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
        L.nmpc_fleet_sample_refs.argtypes = [vp, _ip, _dp, C.c_int32, C.c_double, _dp, _ip]
        L.nmpc_fleet_step.argtypes = [vp, C.c_int32]
        L.nmpc_fleet_state.argtypes = [vp, _dp, _dp, _ip, _ip, _ip, _ip]
        L.nmpc_fleet_last.argtypes = [vp, _dp, _dp, _dp]
        L.nmpc_fleet_log.argtypes = [vp, _dp, _ip]
        self.B = plan.n_robots
        self.log_steps = int(log_steps)
        fc = FleetConfig(n_robots=self.B, max_ref=plan.ref.shape[1], max_vert=plan.vert.shape[1],
                         n_brake=len(plan.brake_vel), n_sched=0 if plan.sched is None else plan.sched.shape[0],
                         log_steps=self.log_steps, num_steps_taken=plan.num_steps_taken, n_dyn=plan.n_dyn,
                         base_speed=plan.base_speed, circle_radius=plan.circle_radius,
                         goal_tol=goal_tol, stop_tol=stop_tol)
        for i, w in enumerate(plan.weights):
            fc.weights[i] = w
        self.fc = fc
        h = vp()
        rc = L.nmpc_fleet_create(solver._h, C.byref(fc), C.byref(h))
        if rc != 0 or not h:
            raise NmpcError(f"nmpc_fleet_create failed (rc={rc}): {solver._lib.nmpc_last_error(solver._h).decode()}")
        self._f = h
        solver._fleets.append(weakref.ref(self))
        self.reset()

    def _check(self, rc, what):
        if rc != 0:
            raise NmpcError(f"{what} failed (rc={rc}): {self._lib.nmpc_last_error(self.solver._h).decode()}")

    def reset(self):
        """(re)upload the plans and put every robot back at step 0"""
        p = self.plan
        ip = lambda a: a.ctypes.data_as(_ip)  # noqa: E731
        dev = self.sample_refs_on_device
        self._check(self._lib.nmpc_fleet_load(self._f, None if dev else ip(p.n_ref), None if dev else _ptr(p.ref),
                                              ip(p.n_vert), _ptr(p.vert), _ptr(p.start), _ptr(p.goal),
                                              _ptr(p.brake_vel), _ptr(p.brake_dist), _ptr(p.sched_init), _ptr(p.sched)),
                    "nmpc_fleet_load")
        if dev:
            self.sample_refs()

    def sample_refs(self, read_back=False):
        """rough_ref on the device for every robot (plan.nodes / plan.ref_speed); -> (ref, n_ref) if read_back"""
        p = self.plan
        if p.nodes is None or p.ref_speed is None:
            raise NmpcError("the plan carries no waypoints (nodes / n_nodes / ref_speed)")
        ref = np.zeros((self.B, self.fc.max_ref, 3)) if read_back else None
        n = np.zeros(self.B, dtype=np.int32) if read_back else None
        self._check(self._lib.nmpc_fleet_sample_refs(self._f, p.n_nodes.ctypes.data_as(_ip), _ptr(p.nodes),
                                                     p.nodes.shape[1], p.ref_speed, _ptr(ref),
                                                     None if n is None else n.ctypes.data_as(_ip)),
                    "nmpc_fleet_sample_refs")
        return ref, n

    def step(self, n_steps=1):
        self._check(self._lib.nmpc_fleet_step(self._f, int(n_steps)), "nmpc_fleet_step")

    def state(self):
        """-> dict(state[B,3], last_u[B,2], t[B], idx[B], done[B], status[B])"""
        B = self.B
        out = {"state": np.zeros((B, 3)), "last_u": np.zeros((B, 2)), "t": np.zeros(B, dtype=np.int32),
               "idx": np.zeros(B, dtype=np.int32), "done": np.zeros(B, dtype=np.int32),
               "status": np.zeros(B, dtype=np.int32)}
        ip = lambda a: a.ctypes.data_as(_ip)  # noqa: E731
        self._check(self._lib.nmpc_fleet_state(self._f, _ptr(out["state"]), _ptr(out["last_u"]), ip(out["t"]),
                                               ip(out["idx"]), ip(out["done"]), ip(out["status"])), "nmpc_fleet_state")
        return out

    def last(self):
        """-> (P[B,np], U[B,2N], Y[B,2N]) of the most recent step"""
        s = self.solver
        P = np.zeros((self.B, param_len(s.cfg)))
        U = np.zeros((self.B, s.n2))
        Y = np.zeros((self.B, s.n2))
        self._check(self._lib.nmpc_fleet_last(self._f, _ptr(P), _ptr(U), _ptr(Y)), "nmpc_fleet_last")
        return P, U, Y

    def log(self):
        """-> (log[B, log_steps, 5] = x, y, theta, v, omega per step; n_logged[B])"""
        lg = np.zeros((self.B, self.log_steps, 5))
        n = np.zeros(self.B, dtype=np.int32)
        self._check(self._lib.nmpc_fleet_log(self._f, _ptr(lg), n.ctypes.data_as(_ip)), "nmpc_fleet_log")
        return lg, n

    def close(self):
        if getattr(self, "_f", None):
            self._lib.nmpc_fleet_destroy(self._f)
            self._f = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
