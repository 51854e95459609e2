"""Oracle-0: independent NumPy / torch(float64) restatement of the NMPC path.

TEST INFRASTRUCTURE ONLY (see oracle/nmpc_oracle.c header).  PARITY UNPINNED: the
reference holds no solver and no golden vectors; this file exists so that the C
oracle is checked by a second, differently written implementation:

  * `cost_graph` follows the reference's CasADi graph statement by statement
    (src/mpc/mpc_generator.py:70-171) with libm cos/sin, true divisions and serial
    sums, in torch.float64 so that torch.autograd supplies a gradient that shares
    no code with the hand-written adjoint in nmpc_oracle.c / the CUDA kernel
    (CasADi, which provides AD in the reference, is not installable here).
  * `psi` adds opengen's augmented-cost terms (builder `__construct_function_psi`).
  * `panoc_alm_solve` is a plain-Python PANOC + ALM/PM loop (same OpEn routines as
    the C oracle restates) for small cross-checks of iteration logic.
"""
import math

import numpy as np
import torch

NZ = 20


def param_len(N, Nobs, Nd):
    return NZ + N + 3 * Nobs + 5 * Nd * N + 3 * N


def cost_graph(u, z0, N, Nobs, Nd, ts):
    """f(u; z0), F1, F2 as torch scalars/vectors.  Mirrors MpcModule.build()
    (src/mpc/mpc_generator.py:70-171) with nu=2, nx=3, nobs=3, ndynobs=5, nz=20."""
    nu, nx, nobs, ndynobs, nz = 2, 3, 3, 5, NZ
    x, y, theta, vel_init, omega_init = z0[0], z0[1], z0[2], z0[3], z0[4]          # :73
    xref, yref, thetaref = z0[5], z0[6], z0[7]                                      # :74
    q, qv, qtheta, rv, rw, qN, qthetaN, qCTE, acc_penalty, omega_acc_penalty = [z0[10 + i] for i in range(10)]  # :75
    cost = torch.zeros((), dtype=torch.float64)
    nF2 = Nobs + Nd
    obstacle_constraints = torch.zeros(nF2, dtype=torch.float64)
    base = nz + N + Nobs * nobs + Nd * ndynobs * N                                  # :79
    for t in range(N):                                                              # :81
        u_t = u[t * nu:(t + 1) * nu]
        cost = cost + rv * u_t[0] ** 2 + rw * u_t[1] ** 2                           # :84
        cost = cost + qv * (u_t[0] - z0[nz + t]) ** 2                               # :85
        cost = cost + q * ((x - xref) ** 2 + (y - yref) ** 2) + qtheta * (theta - thetaref) ** 2  # :86, :59-64
        x = x + ts * (u_t[0] * torch.cos(theta))                                    # :88
        y = y + ts * (u_t[0] * torch.sin(theta))                                    # :89
        theta = theta + ts * u_t[1]                                                 # :90
        s0 = nz + N
        xs_static = z0[s0:s0 + Nobs * nobs:nobs]                                    # :93
        ys_static = z0[s0 + 1:s0 + Nobs * nobs:nobs]                                # :94
        rs_static = z0[s0 + 2:s0 + Nobs * nobs:nobs]                                # :95
        e0 = nz + N + Nobs * nobs                                                   # :98
        e1 = e0 + Nd * ndynobs * N                                                  # :99
        xs_dynamic = z0[e0 + t * ndynobs:e1:ndynobs * N]                            # :100
        ys_dynamic = z0[e0 + t * ndynobs + 1:e1:ndynobs * N]                        # :101
        x_radius = z0[e0 + t * ndynobs + 2:e1:ndynobs * N]                          # :102
        y_radius = z0[e0 + t * ndynobs + 3:e1:ndynobs * N]                          # :103
        As = z0[e0 + t * ndynobs + 4:e1:ndynobs * N]                                # :104
        xdiff_static = x - xs_static
        ydiff_static = y - ys_static
        xdiff_dynamic = x - xs_dynamic
        ydiff_dynamic = y - ys_dynamic
        inside_circle = rs_static ** 2 - xdiff_static ** 2 - ydiff_static ** 2      # :112
        inside_ellipse = 1 - (xdiff_dynamic * torch.cos(As) + ydiff_dynamic * torch.sin(As)) ** 2 / (x_radius ** 2) \
            - (xdiff_dynamic * torch.sin(As) - ydiff_dynamic * torch.cos(As)) ** 2 / (y_radius) ** 2   # :118
        inside = torch.cat([inside_circle, inside_ellipse])
        obstacle_constraints = obstacle_constraints + torch.clamp(inside, min=0.0)  # :119 fmax(0, .)
        # cross-track error (:122-144)
        px, py = x, y
        dists = []
        s2x, s2y = z0[base], z0[base + 1]
        for i in range(1, N):
            s1x, s1y = s2x, s2y
            s2x, s2y = z0[base + i * nx], z0[base + i * nx + 1]
            dx, dy = s2x - s1x, s2y - s1y
            t_hat = ((px - s1x) * dx + (py - s1y) * dy) / (dx ** 2 + dy ** 2 + 1e-16)  # :135
            t_star = torch.clamp(torch.clamp(t_hat, min=0.0), max=1.0)                 # :137
            vx = s1x + t_star * dx - px
            vy = s1y + t_star * dy - py
            dists.append(vx ** 2 + vy ** 2)                                            # :141
        cost = cost + torch.min(torch.stack(dists)) * qCTE                             # :144
    cost = cost + qN * ((x - xref) ** 2 + (y - yref) ** 2) + qthetaN * (theta - thetaref) ** 2   # :148
    v = u[0::2]
    omega = u[1::2]
    acc = (v - torch.cat([vel_init.reshape(1), v[0:-1]])) / ts                      # :160
    omega_acc = (omega - torch.cat([omega_init.reshape(1), omega[0:-1]])) / ts      # :161
    F1 = torch.cat([acc, omega_acc])                                                # :162
    cost = cost + torch.dot(acc, acc) * acc_penalty                                 # :170
    cost = cost + torch.dot(omega_acc, omega_acc) * omega_acc_penalty               # :171
    return cost, F1, obstacle_constraints


def psi(u, z0, c, y, cfg):
    """Augmented cost of opengen's builder: f + c/2*(dist^2_C(F1 + y/max(c,1)) + |F2|^2)."""
    N = cfg["N_hor"]
    f, F1, F2 = cost_graph(u, z0, N, cfg["Nobs"], cfg["Ndynobs"], cfg["ts"])
    lo = torch.tensor([cfg["lin_acc_min"]] * N + [-cfg["ang_acc_max"]] * N, dtype=torch.float64)   # :164-168
    hi = torch.tensor([cfg["lin_acc_max"]] * N + [cfg["ang_acc_max"]] * N, dtype=torch.float64)
    z = F1 + y / max(c, 1.0)
    d = z - torch.minimum(torch.maximum(z, lo), hi)
    return f + 0.5 * c * (torch.dot(d, d) + torch.dot(F2, F2)), f, F1, F2


def eval_psi(u, p, c, y, cfg, want_grad=True):
    """-> (psi, grad, F1, F2) as numpy, gradient by torch.autograd."""
    ut = torch.tensor(np.asarray(u, dtype=np.float64), requires_grad=want_grad)
    z0 = torch.tensor(np.asarray(p, dtype=np.float64))
    yt = torch.tensor(np.asarray(y, dtype=np.float64))
    val, f, F1, F2 = psi(ut, z0, float(c), yt, cfg)
    g = None
    if want_grad:
        (g,) = torch.autograd.grad(val, ut)
        g = g.numpy()
    return float(val), g, F1.detach().numpy(), F2.detach().numpy()


def cfg_dict(c):
    """ctypes Config (oracle_c.Config / product Config) -> plain dict."""
    return {k: getattr(c, k) for k, _ in c._fields_}


# ---------------------------------------------------------------------------
# plain-Python PANOC + ALM (serial sums, libm): same OpEn routines as nmpc_oracle.c
class _Lbfgs:
    def __init__(self, n, mem):
        self.n, self.mem = n, mem
        self.reset()
        self.gamma = 1.0
        self.s, self.y, self.rho = [], [], []

    def reset(self):
        self.active = 0
        self.first_old = True
        self.s, self.y, self.rho = [], [], []

    def update(self, g, state):
        if self.first_old:
            self.first_old = False
            self.old_state, self.old_g = state.copy(), g.copy()
            return
        s, y = state - self.old_state, g - self.old_g
        ys, ss = float(s @ y), float(s @ s)
        if ss <= np.finfo(float).eps or ys <= 1e-10:
            return
        lhs, rhs = ys / ss, 1e-8 * math.sqrt(float(g @ g))
        if not (lhs > rhs and math.isfinite(lhs) and math.isfinite(rhs)):
            return
        self.old_state, self.old_g = state.copy(), g.copy()
        self.s.insert(0, s); self.y.insert(0, y); self.rho.insert(0, 1.0 / ys)
        self.s, self.y, self.rho = self.s[:self.mem], self.y[:self.mem], self.rho[:self.mem]
        self.gamma = ys / float(y @ y)
        self.active = len(self.s)

    def apply(self, q):
        if self.active == 0:
            return q
        q = q.copy()
        al = []
        for k in range(self.active):
            a = self.rho[k] * float(self.s[k] @ q)
            al.append(a)
            q -= a * self.y[k]
        q *= self.gamma
        for k in reversed(range(self.active)):
            b = self.rho[k] * float(self.y[k] @ q)
            q += (al[k] - b) * self.s[k]
        return q


def panoc_alm_solve(p, cfg, u0=None, y0=None):
    """-> dict(u, y, status, outer, inner).  Restates the same OpEn control flow as
    nmpc_oracle.c (see the comments there) in the simplest possible Python."""
    N = cfg["N_hor"]
    n = 2 * N
    u = np.zeros(n) if u0 is None else np.array(u0, dtype=np.float64)
    y = np.zeros(n) if y0 is None else np.array(y0, dtype=np.float64)
    lo_u = np.array([cfg["lin_vel_min"], -cfg["ang_vel_max"]] * N)
    hi_u = np.array([cfg["lin_vel_max"], cfg["ang_vel_max"]] * N)
    lo_c = np.array([cfg["lin_acc_min"]] * N + [-cfg["ang_acc_max"]] * N)
    hi_c = np.array([cfg["lin_acc_max"]] * N + [cfg["ang_acc_max"]] * N)
    c = cfg["initial_penalty"]
    akkt_tol, tol = cfg["initial_tolerance"], cfg["tolerance"]
    EPS = np.finfo(float).eps

    def fg(uu):
        v, g, _, _ = eval_psi(uu, p, c, y, cfg, True)
        return v, g

    def fval(uu):
        return eval_psi(uu, p, c, y, cfg, False)[0]

    def inner(u):
        lb = _Lbfgs(n, cfg["lbfgs_memory"])
        cost, grad = fg(u)
        h = np.maximum(1e-12, 1e-6 * u)
        u = u + h
        _, gh = fg(u)
        lip = np.linalg.norm(gh - grad) / np.linalg.norm(h)
        gamma = 0.95 / max(lip, 1e-10)
        gstep = u - gamma * grad
        uhalf = np.clip(gstep, lo_u, hi_u)
        it = 0
        st = {"cost": cost, "grad": grad, "gamma": gamma, "lip": lip, "gstep": gstep, "uhalf": uhalf}

        def step(u, it):
            fpr = u - st["uhalf"]
            nf = np.linalg.norm(fpr)
            st["nf"] = nf
            if nf < tol:
                prev = st["grad"] if it else 0.0
                if np.linalg.norm(fpr / st["gamma"] + st["grad"] - prev) < akkt_tol:
                    return u, False
            ch = fval(st["uhalf"])
            st["cost"] = fval(u)
            k = 0
            while True:
                rhs = st["cost"] + 1e-6 * abs(st["cost"]) - float(st["grad"] @ fpr) + 0.95 / (2 * st["gamma"]) * nf ** 2
                if not (ch > rhs and k < 10 and st["lip"] < 1e9):
                    break
                lb.reset()
                st["lip"] *= 2; st["gamma"] /= 2
                st["gstep"] = u - st["gamma"] * st["grad"]
                st["uhalf"] = np.clip(st["gstep"], lo_u, hi_u)
                ch = fval(st["uhalf"])
                fpr = u - st["uhalf"]; nf = np.linalg.norm(fpr); st["nf"] = nf
                k += 1
            sigma = 0.05 / (4 * st["gamma"])
            lb.update(fpr, u)
            if it == 0:
                u = st["uhalf"].copy()
                st["cost"], st["grad"] = fg(u)
                st["gstep"] = u - st["gamma"] * st["grad"]
                st["uhalf"] = np.clip(st["gstep"], lo_u, hi_u)
                return u, True
            d = lb.apply(fpr)
            g = st["gamma"]
            fbe = st["cost"] - 0.5 * g * float(st["grad"] @ st["grad"]) + 0.5 * float(np.sum((st["gstep"] - st["uhalf"]) ** 2)) / g
            rhs_ls = fbe - sigma * nf ** 2
            tau, nls = 1.0, 0
            while True:
                up = u - (1 - tau) * fpr - tau * d
                st["cost"], st["grad"] = fg(up)
                st["gstep"] = up - g * st["grad"]
                st["uhalf"] = np.clip(st["gstep"], lo_u, hi_u)
                lhs = st["cost"] - 0.5 * g * float(st["grad"] @ st["grad"]) + 0.5 * float(np.sum((st["gstep"] - st["uhalf"]) ** 2)) / g
                if not (lhs > rhs_ls and nls < 10):
                    break
                tau /= 2; nls += 1
            return up, True

        num, cont = 0, True
        u, flag = step(u, it); it += int(flag)
        while flag and cont:
            num += 1
            cont = num < cfg["max_inner_iterations"]
            u, flag = step(u, it); it += int(flag)
        status = 0 if cont else 1
        if not np.all(np.isfinite(u)):
            status = 3
        return st["uhalf"].copy(), status, num

    iteration, inner_total, num_outer, status = 0, 0, 0, 0
    dyn = f2n = 0.0
    broke = False
    for _ in range(cfg["max_outer_iterations"]):
        num_outer += 1
        y = np.clip(y, -1e12, 1e12)
        u, status, its = inner(u)
        inner_total += its
        if status == 3:
            return dict(u=u, y=y, status=3, outer=num_outer, inner=inner_total)
        _, _, F1, F2 = eval_psi(u, p, 0.0, y, cfg, False)
        w = F1
        yp = y + c * (w - np.clip(w + y / c, lo_c, hi_c))
        dynp, f2np = np.linalg.norm(yp - y), np.linalg.norm(F2)
        if iteration > 0 and dynp <= c * cfg["delta_tolerance"] + EPS and f2np <= cfg["delta_tolerance"] + EPS \
                and akkt_tol <= tol + EPS:
            broke = True
            break
        stall = iteration == 0 or (dynp <= cfg["sufficient_decrease_coeff"] * dyn + EPS
                                   and f2np <= cfg["sufficient_decrease_coeff"] * f2n + EPS)
        if not stall:
            c *= cfg["penalty_update_factor"]
        akkt_tol = max(akkt_tol * cfg["inner_tolerance_update"], tol)
        iteration += 1
        dyn, f2n, y = dynp, f2np, yp
    if num_outer == cfg["max_outer_iterations"]:
        status = 1
    return dict(u=u, y=y, status=status, outer=num_outer, inner=inner_total, c=c)
