"""Host mirror of the reference's per-step parameter assembly and plant step, so that
NMPC inputs for batches of robots can be produced without the reference on the path
(the GPU box has no /root/reference).

Mirrors, quirks included:
  rough_ref             src/mpc/mpc_generator.py:17-57    reference sampler along the A* path
  brake_profile         src/path_generator.py:439-477     get_brake_vel_ref
  Scenario              src/path_generator.py:238-287     prepare / A* / reference / init of the run
  Scenario.parameters   src/path_generator.py:290-382     the 430-float vector of one step
  Scenario.apply        src/mpc/mpc_generator.py:223-235  take steps, integrate the plant
  Scenario.terminal     src/path_generator.py:397         termination test
  closest vertices      src/visibility/visibility.py:111-148
  dynamic obstacles     src/visibility/visibility.py:155-216
tests/test_host_assembly.py replays the recorded run of the unmodified reference
(tests/golden/config1_run.npz) through this module and requires identical vectors.
"""
import itertools
import json
import math
import os

import numpy as np

from . import planner

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "maps.json")


class HostConfig(dict):
    """The reference's config keys (configs/default.yaml) with attribute access."""
    __getattr__ = dict.get

    @classmethod
    def default(cls, **kw):
        c = cls(N_hor=20, lin_vel_min=-0.5, lin_vel_max=1.5, lin_acc_min=-1.0, lin_acc_max=1.0, ang_vel_max=0.5,
                ang_acc_max=3.0, throttle_ratio=1.0, num_steps_taken=1, ts=0.2, vel_red_steps=20,
                lin_vel_penalty=0.0, lin_acc_penalty=10.0, ang_vel_penalty=0.0, ang_acc_penalty=5.0,
                cte_penalty=200.0, q=0.0, qv=10.0, qtheta=0.0, qN=0.0, qthetaN=0.0, nx=3, nz=20, nu=2, nobs=3,
                Nobs=10, Ndynobs=3, ndynobs=5, vehicle_width=0.5, vehicle_margin=0.25)
        c.update(kw)
        return c

    @classmethod
    def smooth_velocity(cls, **kw):
        """Weights/bounds of configs/smooth_velocity.yaml:12-29; that file lacks qv, vel_red_steps and
        Ndynobs (rejected by the reference's own loader), so qv=10, vel_red_steps=20 are supplied here."""
        c = cls.default(ang_vel_max=1.0, ang_acc_max=5.0, throttle_ratio=0.9, lin_vel_penalty=0.0,
                        lin_acc_penalty=8.0, ang_vel_penalty=0.0, ang_acc_penalty=20.0, cte_penalty=20.0, q=1.0,
                        qtheta=0.0, qN=5.0, qthetaN=0.2, qv=10.0)
        c.update(kw)
        return c


def load_maps():
    """The reference's 13 scenario maps (src/visibility/graphs.py), from the committed data fixture."""
    with open(_DATA) as f:
        return {m["complexity"]: m for m in json.load(f)["maps"]}


def rough_ref(pos, node_list, v, ts):
    """Walk the waypoints at speed v, one sample per ts (src/mpc/mpc_generator.py:17-57)."""
    x_ref, y_ref, theta_ref = [], [], []
    x, y = pos
    i = 0
    x_target, y_target = node_list[i]
    x_dir = y_dir = 0.0
    traveling = True
    while traveling:
        t = ts
        while t > 0:
            dist = math.hypot(x_target - x, y_target - y)
            if dist == 0:
                traveling = False
                break
            x_dir = (x_target - x) / dist
            y_dir = (y_target - y) / dist
            time = dist / v
            if time > t:
                x, y = x + x_dir * v * t, y + y_dir * v * t
                t = 0
            else:
                x, y = x + x_dir * v * time, y + y_dir * v * time
                t = t - time
                i += 1
                if i > len(node_list) - 1:
                    traveling = False
                    break
                x_target, y_target = node_list[i]
        x_ref.append(x)
        y_ref.append(y)
        theta_ref.append(math.atan2(y_dir, x_dir))
    return x_ref, y_ref, theta_ref


def brake_profile(cfg):
    """Linear ramp base_speed -> 0 and distance-to-goal table (src/path_generator.py:439-477)."""
    base_speed = cfg.lin_vel_max * cfg.throttle_ratio
    brake_acc = max(cfg.lin_acc_min, -base_speed / (cfg.ts * cfg.vel_red_steps))
    brake_time = -base_speed / brake_acc
    brake_dist = base_speed * brake_time + 0.5 * brake_acc * brake_time ** 2
    steps = math.ceil(brake_time / cfg.ts)
    vel = [base_speed - base_speed / (steps - 1) * i for i in range(steps)]
    dist = [0.0] * len(vel)
    dist[0] = brake_dist
    for i, v in enumerate(vel):
        if i < len(dist) - 1:
            dist[i + 1] = dist[i] - v * cfg.ts
    return vel, dist


def _closest(pos, pts):
    """first index of the closest point (np.argmin over 2-norms, src/visibility/visibility.py:111-124)."""
    d = [np.linalg.norm(np.asarray(pos) - np.asarray(p), ord=2) for p in pts]
    return int(np.argmin(d))


def _move_obstacle(p1, p2, freq, time):
    """generate_obstacle, src/visibility/visibility.py:155-166"""
    p1, p2 = np.array(p1), np.array(p2)
    t = abs(np.sin(freq * np.array(time)))
    return t * p1 + (1 - t) * p2


def _sinus_obstacle(p1, p2, freq, time, ampl=1.5):
    """generate_sinus_obstacle + rotate_and_add, src/visibility/visibility.py:168-196"""
    p1, p2 = np.array(p1), np.array(p2)
    dp = p2 - p1
    angle = np.arctan2(dp[1], dp[0])
    t = abs(np.sin(freq * np.array(time)))
    p3 = t * p1 + (1 - t) * p2
    add = ampl * np.cos(10 * freq * time)

    def rot(origin, point, a):
        ox, oy = origin
        px, py = point
        return np.array([math.cos(a) * (px - ox) - math.sin(a) * (py - oy),
                         math.sin(a) * (px - ox) + math.cos(a) * (py - oy)])
    r = rot(p1, p3, angle)
    r[1] += add
    r = rot(np.array([0, 0]), r, -angle)
    return r + p1


class Scenario:
    """One robot on one map: global plan + the state of its receding-horizon run."""

    def __init__(self, cfg, graph_map, start=None, end=None, sinus_object=False, env=None):
        self.cfg = cfg
        self.map = graph_map
        self.start = list(graph_map["start"] if start is None else start)
        self.end = list(graph_map["end"] if end is None else end)
        self.sinus_object = sinus_object
        self.obstacles = [[tuple(p) for p in o] for o in graph_map["obstacles"]]
        self.boundary = [tuple(p) for p in graph_map["boundary"]]
        self.dyn_obs = graph_map.get("dyn_obs", [])
        self.env = env if env is not None else self.make_env(cfg, graph_map)
        self.ok = self._plan()
        self.reset()

    @staticmethod
    def make_env(cfg, graph_map):
        """inflate obstacles / deflate the boundary by vehicle_width and build the visibility graph
        (src/visibility/visibility.py:49-67,90-105); reusable across robots on the same map."""
        S = 2 ** 31   # pyclipper.scale_to_clipper: the reference offsets on Clipper's integer grid

        def offset(poly, delta):
            grid = [(int(round(x * S)) / S, int(round(y * S)) / S) for x, y in poly]
            out = planner.offset_polygon(grid, int(round(delta * S)) / S)
            return [(int(round(x * S)) / S, int(round(y * S)) / S) for x, y in out]

        w = cfg.vehicle_width
        holes = []
        for o in graph_map["obstacles"]:
            infl = offset(o, w)
            infl.reverse()
            holes.append(infl)
        bound = offset(graph_map["boundary"], -w)
        env = planner.PolygonEnvironment()
        env.store(bound, holes)
        env.prepare()
        return env

    def _plan(self):
        cfg = self.cfg
        path, _ = self.env.find_shortest_path((self.start[0], self.start[1]), (self.end[0], self.end[1]))
        self.path = path
        if len(path) < 2:
            return False
        # find_original_vertices, src/visibility/visibility.py:126-139
        self.vert = []
        if len(path) > 2:
            flat = [p for poly in self.obstacles + [self.boundary] for p in poly]
            for v in path[1:-1]:
                self.vert.append(flat[_closest(v, flat)])
        v = cfg.throttle_ratio * 1.1 * cfg.lin_vel_max
        self.x_ref, self.y_ref, self.theta_ref = rough_ref((self.start[0], self.start[1]), path[1:], v, cfg.ts)
        self.ref_points = list(zip(self.x_ref, self.y_ref))
        self.brake_vel, self.brake_dist = brake_profile(cfg)
        return True

    def reset(self):
        cfg = self.cfg
        self.t = 0
        self.idx = 0
        self.states = list(self.start)
        self.system_input = []
        self.constraints = [0.0] * cfg.Nobs * cfg.nobs
        self.dyn_constraints = [0.0] * cfg.Ndynobs * cfg.ndynobs * cfg.N_hor
        self.dyn_constraints[2::cfg.ndynobs] = [1.0] * cfg.Ndynobs * cfg.N_hor
        self.dyn_constraints[3::cfg.ndynobs] = [1.0] * cfg.Ndynobs * cfg.N_hor
        self.weights = [cfg.q, cfg.qv, cfg.qtheta, cfg.lin_vel_penalty, cfg.ang_vel_penalty, cfg.qN, cfg.qthetaN,
                        cfg.cte_penalty, cfg.lin_acc_penalty, cfg.ang_acc_penalty]   # src/path_generator.py:226-227

    # -- helpers ----------------------------------------------------------------------------
    def _closest_vertices(self, pos):
        """find_closest_vertices(pos, Nobs, 0) incl. its slice quirk (src/visibility/visibility.py:141-148)"""
        n = self.cfg.Nobs
        if n >= len(self.vert):
            return self.vert
        idx = _closest(pos, self.vert)
        return self.vert[max(0, idx):min(len(self.vert), n)]

    def _dyn_obstacles(self, t, horizon):
        """get_dyn_obstacle, src/visibility/visibility.py:199-216"""
        cfg = self.cfg
        if len(self.dyn_obs) == 0:
            return []
        times = np.linspace(t, t + horizon * cfg.ts, horizon)
        out = []
        for i, obs in enumerate(self.dyn_obs):
            p1, p2, freq, rx, ry, angle = obs
            rx = rx + cfg.vehicle_width / 2 + cfg.vehicle_margin
            ry = ry + cfg.vehicle_width / 2 + cfg.vehicle_margin
            if self.sinus_object and i == 2:
                out.append([(*_sinus_obstacle(p1, p2, freq, tt), rx, ry, angle) for tt in times])
            else:
                out.append([(*_move_obstacle(p1, p2, freq, tt), rx, ry, angle) for tt in times])
        return out

    # -- one step -----------------------------------------------------------------------------
    def parameters(self):
        """Assemble the parameter vector of the current step (src/path_generator.py:293-382);
        advances the reference index and the dynamic-obstacle ring like the reference does."""
        cfg = self.cfg
        N, steps = cfg.N_hor, cfg.num_steps_taken
        x_init = self.states[-cfg.nx:]
        if len(self.obstacles):
            origin = self._closest_vertices((x_init[0], x_init[1]))
            r = cfg.vehicle_width / 2 + cfg.vehicle_margin
            cons = list(itertools.chain(*[(x, y, r) for x, y in origin]))
            cons += [0.0] * (cfg.Nobs * cfg.nobs - len(cons))
            self.constraints = cons
        per = N * cfg.ndynobs
        if self.t == 0:
            for i, dob in enumerate(self._dyn_obstacles(self.t * cfg.ts, N)):
                self.dyn_constraints[i * per:(i + 1) * per] = [float(v) for v in itertools.chain(*dob)]
        else:
            k = cfg.ndynobs * steps
            self.dyn_constraints = self.dyn_constraints[k:] + self.dyn_constraints[:k]
            for i, dob in enumerate(self._dyn_obstacles((self.t + N - steps) * cfg.ts, steps)):
                self.dyn_constraints[(i + 1) * per - k:(i + 1) * per] = [float(v) for v in itertools.chain(*dob)]
        lb = max(0, self.idx - 1 * steps)
        ub = min(len(self.ref_points), self.idx + 5 * steps)
        self.idx = _closest((x_init[0], x_init[1]), self.ref_points[lb:ub]) + lb
        idx, n = self.idx, len(self.x_ref)
        end = self.end
        if idx + N >= n:
            x_finish = end
            tmpx = self.x_ref[idx:] + [end[0]] * (N - (n - idx))
            tmpy = self.y_ref[idx:] + [end[1]] * (N - (n - idx))
            tmpt = self.theta_ref[idx:] + [end[2]] * (N - (n - idx))
        else:
            x_finish = [self.x_ref[idx + N], self.y_ref[idx + N], self.theta_ref[idx + N]]
            tmpx, tmpy, tmpt = self.x_ref[idx:idx + N], self.y_ref[idx:idx + N], self.theta_ref[idx:idx + N]
        base_speed = cfg.lin_vel_max * cfg.throttle_ratio
        if (idx + N) >= n - self.brake_dist[0] / base_speed:
            nb = min(n - idx - 1, N)
            vel_ref = [base_speed] * nb
            if nb == 0:
                d = math.sqrt((self.states[-3] - end[0]) ** 2 + (self.states[-2] - end[1]) ** 2)
                vel_ref = [v for v, dist in zip(self.brake_vel, self.brake_dist) if dist <= d]
            else:
                vel_ref += self.brake_vel[:min(len(self.brake_vel), N - nb)]
            vel_ref += [0.0] * (N - len(vel_ref))
        else:
            vel_ref = [base_speed] * N
        refs = [0.0] * (N * cfg.nx)
        refs[0::cfg.nx], refs[1::cfg.nx], refs[2::cfg.nx] = tmpx, tmpy, tmpt
        last_u = self.system_input[-cfg.nu:] if len(self.system_input) else [0.0] * cfg.nu
        p = list(x_init) + last_u + list(x_finish) + last_u + self.weights + vel_ref + self.constraints \
            + self.dyn_constraints + refs
        return np.asarray(p, dtype=np.float64)

    def apply(self, u, sincos=None):
        """take num_steps_taken controls and integrate the plant (src/mpc/mpc_generator.py:223-235);
        returns True when the run is terminal (src/path_generator.py:397).  `sincos(theta) -> (sin, cos)`
        replaces libm's (the reference's math.sin/math.cos) when a caller needs the device's own trig."""
        cfg = self.cfg
        u = [float(x) for x in u]
        self.system_input += u[:cfg.nu * cfg.num_steps_taken]
        for i in range(cfg.num_steps_taken):
            uv, uw = u[i * cfg.nu], u[1 + i * cfg.nu]
            x, y, th = self.states[-3], self.states[-2], self.states[-1]
            sn, cs = (math.sin(th), math.cos(th)) if sincos is None else sincos(th)
            self.states += [x + cfg.ts * (uv * cs), y + cfg.ts * (uv * sn), th + cfg.ts * uw]
        self.t += cfg.num_steps_taken
        return self.terminal()

    def terminal(self):
        return bool(np.allclose(self.states[-3:-1], self.end[0:2], atol=0.05, rtol=0)
                    and abs(self.system_input[-2]) < 0.005)
