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
        OracleManager.log = {"P": [], "U": [], "status": [], "inner": [], "outer": []}

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
        self.u, self.y = U, Y
        lg = OracleManager.log
        lg["P"].append(P[0]); lg["U"].append(U[0].copy()); lg["status"].append(int(st[0]))
        lg["inner"].append(int(stats["inner_iterations"][0])); lg["outer"].append(int(stats["outer_iterations"][0]))
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


if __name__ == "__main__":
    main()
