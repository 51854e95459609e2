"""Synthetic workloads for the BASELINE.json configs, built with the host
mirror of the reference's pipeline (host.assembly: A* seed path -> rough reference ->
parameter vector).  Everything is seeded and self-contained (maps come from the committed
data fixture), so the same batches can be regenerated on the GPU box.

  first_step_batch   config 2: B random start/goal pairs on one map, the t=0 problem of
                     each pair (last_u = 0, cold start U0 = 0)
  closed_loop_batch  config 3/4: K robots rolled for T receding-horizon steps with a solver
                     callback; records (p_k, u_{k-1}, y_{k-1}) so that all K*T steps can be
                     replayed as ONE batch with the reference's warm-start semantics
  sweep_batch        config 5: horizon / static-obstacle-count sweep
"""
import math

import numpy as np

from .host import assembly
from .solver import NmpcConfig


def solver_config_for(host_cfg, **overrides):
    """NmpcConfig (sizes, bounds, ts) from a HostConfig."""
    return NmpcConfig.from_reference_config(host_cfg, **overrides)


def _sample_free_points(env, bbox, n, rng, clearance=0.3):
    """uniform points in the planner's free space (deflated boundary minus inflated obstacles),
    kept `clearance` away from every polygon edge so that start poses are not degenerate."""
    (x0, y0), (x1, y1) = bbox
    out = []
    ea, eb = env._ea, env._eb
    ex, ey = (eb - ea)[:, 0], (eb - ea)[:, 1]
    ln2 = ex * ex + ey * ey
    while len(out) < n:
        pts = np.stack([rng.uniform(x0, x1, 4 * n), rng.uniform(y0, y1, 4 * n)], axis=1)
        ok = env._free(pts)
        px, py = pts[:, 0:1], pts[:, 1:2]
        t = np.clip(((px - ea[:, 0]) * ex + (py - ea[:, 1]) * ey) / ln2, 0.0, 1.0)
        d2 = (ea[:, 0] + t * ex - px) ** 2 + (ea[:, 1] + t * ey - py) ** 2
        ok &= d2.min(axis=1) > clearance ** 2
        out.extend(pts[ok].tolist())
    return np.asarray(out[:n])


def random_scenarios(host_cfg, complexity, n, seed, min_dist=3.0):
    """n robots with random start/goal poses on map `complexity` (reachable pairs only)."""
    rng = np.random.default_rng(seed)
    gmap = assembly.load_maps()[complexity]
    env = assembly.Scenario.make_env(host_cfg, gmap)
    b = np.asarray(gmap["boundary"])
    bbox = (b.min(axis=0), b.max(axis=0))
    out = []
    while len(out) < n:
        m = n - len(out)
        S = _sample_free_points(env, bbox, m, rng)
        G = _sample_free_points(env, bbox, m, rng)
        hs = rng.uniform(0.0, 2.0 * math.pi, m)
        hg = rng.uniform(0.0, 2.0 * math.pi, m)
        for i in range(m):
            if np.hypot(*(S[i] - G[i])) < min_dist:
                continue
            sc = assembly.Scenario(host_cfg, gmap, [S[i, 0], S[i, 1], hs[i]], [G[i, 0], G[i, 1], hg[i]], env=env)
            if sc.ok:
                out.append(sc)
    return out


def first_step_batch(host_cfg, complexity=3, B=4096, seed=0):
    """BASELINE config 2.  -> (P[B, np], scenarios)"""
    scs = random_scenarios(host_cfg, complexity, B, seed)
    P = np.stack([s.parameters() for s in scs])
    return P, scs


def closed_loop_batch(host_cfg, solve_fn, complexity=11, robots=256, steps=256, seed=1, sincos=None):
    """BASELINE config 3/4.  Rolls `robots` receding-horizon runs for up to `steps` steps;
    `solve_fn(P, U0, Y0) -> (U, Y, status)` solves one step for all live robots (the reference
    sends only p and the server keeps (u, y): warm start from the previous reply, un-shifted).
    -> dict(P, U0, Y0, U, status, robot, step) with one row per recorded step."""
    scs = random_scenarios(host_cfg, complexity, robots, seed)
    n2 = 2 * host_cfg.N_hor
    Uprev = np.zeros((robots, n2))
    Yprev = np.zeros((robots, n2))
    live = np.ones(robots, dtype=bool)
    rec = {k: [] for k in ("P", "U0", "Y0", "U", "status", "robot", "step")}
    for k in range(steps):
        ids = np.nonzero(live)[0]
        if len(ids) == 0:
            break
        P = np.stack([scs[i].parameters() for i in ids])
        U, Y, st = solve_fn(P, Uprev[ids], Yprev[ids])
        rec["P"].append(P); rec["U0"].append(Uprev[ids].copy()); rec["Y0"].append(Yprev[ids].copy())
        rec["U"].append(U.copy()); rec["status"].append(np.asarray(st).copy())
        rec["robot"].append(ids.copy()); rec["step"].append(np.full(len(ids), k))
        Uprev[ids], Yprev[ids] = U, Y
        for j, i in enumerate(ids):
            if scs[i].apply(U[j], sincos=sincos):
                live[i] = False
    return {k: np.concatenate(v) for k, v in rec.items()}


def closed_loop_batch_device(solver, host_cfg, complexity=11, robots=256, steps=256, seed=1):
    """BASELINE config 3/4 recorded on the device: the fleet API rolls `robots` receding-horizon runs for `steps`
    steps and every live robot's step is recorded as one row (p_k, warm start u_{k-1}, y_{k-1}) -> replay batch.
    -> dict(P, U0, Y0, U, status, robot, step)"""
    from .fleet import FleetPlan, NmpcFleet
    scs = random_scenarios(host_cfg, complexity, robots, seed)
    plan = FleetPlan.from_scenarios(scs, max_steps=steps)
    fleet = NmpcFleet(solver, plan)
    n2 = 2 * host_cfg.N_hor
    Uprev = np.zeros((robots, n2))
    Yprev = np.zeros((robots, n2))
    rec = {k: [] for k in ("P", "U0", "Y0", "U", "status", "robot", "step")}
    done_prev = np.zeros(robots, dtype=np.int32)
    for k in range(steps):
        ids = np.nonzero(done_prev == 0)[0]
        if len(ids) == 0:
            break
        fleet.step(1)
        P, U, Y = fleet.last()
        st = fleet.state()
        rec["P"].append(P[ids]); rec["U0"].append(Uprev[ids].copy()); rec["Y0"].append(Yprev[ids].copy())
        rec["U"].append(U[ids]); rec["status"].append(st["status"][ids]); rec["robot"].append(ids.copy())
        rec["step"].append(np.full(len(ids), k))
        Uprev, Yprev, done_prev = U, Y, st["done"]
    fleet.close()
    return {k: np.concatenate(v) for k, v in rec.items()}


def sweep_batch(N, Nobs, B, seed=2, complexity=11, Ndynobs=3):
    """BASELINE config 5: first-step problems at horizon N with Nobs static-circle slots; the slots
    beyond the A*-corner vertices are filled with the map's own obstacle vertices and then uniform
    random centres in free space (r = vehicle_width/2 + vehicle_margin)."""
    hc = assembly.HostConfig.default(N_hor=N, Nobs=Nobs, Ndynobs=Ndynobs)
    scs = random_scenarios(hc, complexity, B, seed)
    rng = np.random.default_rng(seed + 17)
    gmap = assembly.load_maps()[complexity]
    verts = np.asarray([p for o in gmap["obstacles"] for p in o])
    b = np.asarray(gmap["boundary"])
    r = hc.vehicle_width / 2 + hc.vehicle_margin
    P = np.stack([s.parameters() for s in scs])
    c0 = 20 + N
    for i in range(B):
        used = int(np.count_nonzero(P[i, c0 + 2:c0 + 3 * Nobs:3]))
        extra = Nobs - used
        if extra <= 0:
            continue
        pool = verts[rng.permutation(len(verts))][:extra]
        if len(pool) < extra:
            rnd = np.stack([rng.uniform(b[:, 0].min(), b[:, 0].max(), extra - len(pool)),
                            rng.uniform(b[:, 1].min(), b[:, 1].max(), extra - len(pool))], axis=1)
            pool = np.concatenate([pool, rnd]) if len(pool) else rnd
        # keep extra circles off the robot's initial position so the start is feasible
        d = np.hypot(pool[:, 0] - P[i, 0], pool[:, 1] - P[i, 1])
        pool = pool[d > 2 * r]
        blk = np.zeros((extra, 3))
        blk[:len(pool), 0:2] = pool
        blk[:len(pool), 2] = r
        P[i, c0 + 3 * used:c0 + 3 * Nobs] = blk.ravel()
    return P, hc
