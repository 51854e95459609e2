"""Generates the committed fixtures by IMPORTING the reference (read-only, /root/reference)
in the build container.  Run:  python tests/golden/make_fixtures.py

  mpc_trajectory_generator_b200/data/maps.json
      the 13 scenario maps of src/visibility/graphs.py (boundary, obstacles, default
      start/end pose, dynamic obstacles) — data fixture, reused as the input generator's maps.
  tests/golden/config1_run.npz
      BASELINE config 1: the UNMODIFIED reference PathGenerator.run
      (src/path_generator.py:197-437) on map complexity=1 with configs/default.yaml,
      driven end to end with `og.tcp.OptimizerTcpManager` replaced by a manager backed by
      the CPU oracle (no GPU in this container).  Records every parameter vector the
      reference assembled (src/path_generator.py:378-379), every reply, the resulting
      trajectory and the A* path / obstacle vertices.
  tests/golden/ref_runs.npz
      the reference's own benchmark and demo runs, recorded the same way: maps 1-11 with
      configs/default.yaml (src/gen_runtime_plots.py:21-33) and map 12 with sinus_object=True
      (src/main.py:11-23: moving ellipses, ring update src/path_generator.py:306-316).  Every
      STRIDE-th solver call of each run is kept as a self-contained tuple (p, warm start u/y the
      server held, reply u, multipliers y, status, iteration counts) so that a test can replay all of
      them as ONE batch.
  tests/golden/reference_helpers.npz
      outputs of the reference's pure helpers (rough_ref, get_brake_vel_ref) and of its
      PathPreProcessor (on top of our planner substitute) for maps 1, 3, 11, 12.

The third-party packages the reference imports are absent here; `host.shims` provides
them (planner substitute, inert matplotlib/cv2).  OpEn itself cannot run: parity stays
unpinned at that boundary (see oracle/nmpc_oracle.c).
"""
import json
import math
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, "src"))

from mpc_trajectory_generator_b200.host import shims, opengen_compat  # noqa: E402
from oracle import oracle_c  # noqa: E402


class OracleManager:
    """og.tcp.OptimizerTcpManager duck type backed by the CPU oracle, recording the traffic."""
    log = None

    def __init__(self, path=None):
        self.cfg = oracle_c.default_config(**{k: v for k, v in opengen_compat._ACTIVE["cfg"].as_dict().items()
                                              if not k.startswith("reserved")})
        self.u = np.zeros((1, 2 * self.cfg.N_hor))
        self.y = np.zeros((1, 2 * self.cfg.N_hor))
        OracleManager.log = {"P": [], "U": [], "status": [], "inner": [], "outer": [], "U0": [], "Y0": [], "Y": [],
                             "n_grad": [], "n_cost": []}

    def start(self):
        pass

    def ping(self):
        return {"Pong": 1}

    def kill(self):
        pass

    def call(self, p):
        P = np.asarray(p, dtype=np.float64)[None]
        t0 = time.time()
        U, Y, st, stats = oracle_c.solve_batch(self.cfg, P, self.u, self.y, nthreads=1)
        ms = 1e3 * (time.time() - t0)
        lg = OracleManager.log
        lg["U0"].append(self.u[0].copy()); lg["Y0"].append(self.y[0].copy())
        self.u, self.y = U, Y
        lg["P"].append(P[0]); lg["U"].append(U[0].copy()); lg["Y"].append(Y[0].copy()); lg["status"].append(int(st[0]))
        lg["inner"].append(int(stats["inner_iterations"][0])); lg["outer"].append(int(stats["outer_iterations"][0]))
        lg["n_grad"].append(int(stats["n_grad_evals"][0])); lg["n_cost"].append(int(stats["n_cost_evals"][0]))
        return opengen_compat.SolverResponse(opengen_compat.SolverStatus(U[0], st[0], stats[0], ms), True)


def main():
    from utils.config import Configurator
    config = Configurator(os.path.join(REF, "configs", "default.yaml")).configurate()
    shims.install(reference_config=config)
    import opengen as og
    og.tcp.OptimizerTcpManager = OracleManager
    from visibility.graphs import Graphs
    from path_generator import PathGenerator

    graphs = Graphs()
    maps = []
    for i in range(graphs.max_complexity + 1):
        g = graphs.get_graph(i)
        maps.append({"complexity": i, "boundary": [list(map(float, p)) for p in g.boundary_coordinates],
                     "obstacles": [[list(map(float, p)) for p in o] for o in g.obstacle_list],
                     "start": list(map(float, g.start)), "end": list(map(float, g.end)),
                     "dyn_obs": [[list(map(float, d[0])), list(map(float, d[1]))] + [float(x) for x in d[2:]]
                                 for d in g.dyn_obs_list]})
    data_dir = os.path.join(ROOT, "mpc_trajectory_generator_b200", "data")
    os.makedirs(data_dir, exist_ok=True)
    with open(os.path.join(data_dir, "maps.json"), "w") as f:
        json.dump({"source": "src/visibility/graphs.py (reference), dumped by tests/golden/make_fixtures.py",
                   "maps": maps}, f, indent=1)
    print("maps.json:", len(maps), "maps")

    # --- reference helpers -------------------------------------------------------------
    helpers = {}
    pg = PathGenerator(config, build=False)
    bv, bd = pg.get_brake_vel_ref()
    helpers["brake_velocities"], helpers["brake_distances"] = np.array(bv), np.array(bd)
    for cx in (1, 3, 11, 12):
        g = graphs.get_graph(cx)
        pgi = PathGenerator(config, build=False)
        pgi.ppp.prepare(g)
        path, verts = pgi.ppp.get_initial_guess((g.start[0], g.start[1]), (g.end[0], g.end[1]))
        xr, yr, tr = pgi.mpc_generator.rough_ref((g.start[0], g.start[1]), path[1:])
        helpers[f"map{cx}_path"] = np.array(path)
        helpers[f"map{cx}_vertices"] = np.array(verts).reshape(-1, 2)
        helpers[f"map{cx}_ref"] = np.array([xr, yr, tr]).T
        print(f"map {cx}: A* path {[(round(x, 3), round(y, 3)) for x, y in path]} -> {len(xr)} reference points")
    np.savez_compressed(os.path.join(HERE, "reference_helpers.npz"), **helpers)

    # --- config 1: full receding-horizon run through the unmodified reference ------------
    g = graphs.get_graph(1)
    pg = PathGenerator(config, build=False)
    t0 = time.time()
    xx, xy, uv, uomega, solver_times, overhead = pg.run(g, list(g.start), list(g.end))
    lg = OracleManager.log
    print(f"config 1: {len(lg['P'])} NMPC steps in {time.time() - t0:.1f}s; exit status counts "
          f"{np.bincount(lg['status'], minlength=4)}; final pose ({xx[-1]:.3f}, {xy[-1]:.3f}); "
          f"inner iterations mean {np.mean(lg['inner']):.0f}")
    np.savez_compressed(os.path.join(HERE, "config1_run.npz"),
                        P=np.array(lg["P"]), U=np.array(lg["U"]), status=np.array(lg["status"], dtype=np.int32),
                        inner=np.array(lg["inner"], dtype=np.int32), outer=np.array(lg["outer"], dtype=np.int32),
                        xx=np.array(xx), xy=np.array(xy), uv=np.array(uv), uomega=np.array(uomega),
                        path=np.array(pg.ppp.path), vertices=np.array(pg.ppp.vert).reshape(-1, 2),
                        start=np.array(g.start), end=np.array(g.end))


    record_runs(config, graphs, PathGenerator)


STRIDE = 3   # every STRIDE-th solver call of a run is kept (plus the first and the last)


def record_runs(config, graphs, PathGenerator):
    """The reference's benchmark script (maps 1-11) and demo (map 12, sinus_object=True) through the
    UNMODIFIED PathGenerator.run with the oracle-backed manager."""
    out = {k: [] for k in ("P", "U0", "Y0", "U", "Y", "status", "inner", "outer", "n_grad", "n_cost", "map", "step")}
    summary = []
    ctrl, ctrl_map = [], []     # every applied control (v, omega) of every run: lets a test re-walk the whole run
    for cx in range(1, 13):
        g = graphs.get_graph(cx)
        pg = PathGenerator(config, build=False, sinus_object=(cx == 12))
        t0 = time.time()
        xx, xy, uv, uomega, solver_times, overhead = pg.run(g, list(g.start), list(g.end))
        lg = OracleManager.log
        K = len(lg["P"])
        # (a run that never reaches its goal — map 2 stalls in a local minimum facing away from its path — lasts the
        #  reference's full 2500 steps: beyond step 300 only every 50th call is kept)
        keep = sorted(set(range(0, min(K, 300), STRIDE)) | set(range(300, K, 50)) | {K - 1})
        for k in keep:
            for key in ("P", "U0", "Y0", "U", "Y", "status", "inner", "outer", "n_grad", "n_cost"):
                out[key].append(lg[key][k])
            out["map"].append(cx); out["step"].append(k)
        ctrl.extend(zip(uv, uomega)); ctrl_map.extend([cx] * len(uv))
        P = np.array(lg["P"])
        N, Nobs, Nd = config.N_hor, config.Nobs, config.Ndynobs
        circ = P[:, 20 + N:20 + N + 3 * Nobs].reshape(K, Nobs, 3)
        ell = P[:, 20 + N + 3 * Nobs:20 + N + 3 * Nobs + 5 * Nd * N].reshape(K, Nd, N, 5)
        moving = bool(np.any(ell[:, :, :, 0] != 0.0))
        summary.append((cx, K, len(keep), np.bincount(lg["status"], minlength=4).tolist(), int(np.max((circ[:, :, 2] != 0).sum(axis=1))),
                        len(pg.ppp.vert), moving, float(xx[-1]), float(xy[-1])))
        print(f"map {cx}: {K} steps ({len(keep)} kept) in {time.time() - t0:.0f}s, status counts {summary[-1][3]}, "
              f"<= {summary[-1][4]} circles per step of {summary[-1][5]} corner vertices, moving ellipses: {moving}, "
              f"final pose ({xx[-1]:.2f}, {xy[-1]:.2f}) goal ({g.end[0]}, {g.end[1]})")
    np.savez_compressed(os.path.join(HERE, "ref_runs.npz"),
                        **{k: np.array(v, dtype=(np.int32 if k in ("status", "inner", "outer", "n_grad", "n_cost", "map", "step") else np.float64))
                           for k, v in out.items()},
                        ctrl=np.array(ctrl, dtype=np.float64), ctrl_map=np.array(ctrl_map, dtype=np.int32),
                        summary=np.array([[s[0], s[1], s[2], s[4], s[5], int(s[6])] for s in summary], dtype=np.int32))


if __name__ == "__main__":
    main()
