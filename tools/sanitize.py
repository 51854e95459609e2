#!/usr/bin/env python
"""tools/sanitize.py — a small mixed workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize.py
small batches (first wave and queue, obstacle slow paths), N=20 and N=40, plus a few fleet steps."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import mpc_trajectory_generator_b200 as pkg
    import nmpc_problems as problems
    from mpc_trajectory_generator_b200 import workloads
    from mpc_trajectory_generator_b200.fleet import FleetPlan
    from mpc_trajectory_generator_b200.host import assembly
    for N, B in ((20, 5), (20, 40), (40, 6)):
        cfg = pkg.NmpcConfig.default(N_hor=N, max_inner_iterations=40, max_outer_iterations=3)
        P = problems.synth(N, 10, 3, B, seed=N + B, active=True)
        s = pkg.NmpcSolver(cfg, device=0)
        U, Y, st, stats = s.solve_batch(P)
        print(N, B, np.bincount(st, minlength=4).tolist(), int(stats["inner_iterations"].sum()), flush=True)
        s.close()
    hc = assembly.HostConfig.default()
    scs = workloads.random_scenarios(hc, 3, 6, seed=3)
    s = pkg.NmpcSolver(workloads.solver_config_for(hc, max_inner_iterations=30, max_outer_iterations=2), device=0)
    f = pkg.NmpcFleet(s, FleetPlan.from_scenarios(scs), log_steps=4, sample_refs_on_device=True)
    f.step(3)
    print("fleet", f.state()["t"].tolist(), flush=True)
    f.close()
    s.close()


if __name__ == "__main__":
    main()
