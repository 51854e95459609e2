#!/usr/bin/env python
"""tools/bench_fleet.py — closed-loop fleet stepping on the device (GPU box).

    python tools/bench_fleet.py [--robots 16384] [--steps 20] [--complexity 11]

Replicates a few hundred planned robots up to `--robots` (plans are per-robot arrays; duplicates are as good as
distinct robots for timing), runs `--steps` receding-horizon steps in one nmpc_fleet_step call and prints robot-steps/s
plus the per-kernel split when run under `ncu --metrics gpu__time_duration.sum`.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--robots", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--complexity", type=int, default=11)
    ap.add_argument("--distinct", type=int, default=256)
    args = ap.parse_args()
    import mpc_trajectory_generator_b200 as pkg
    from mpc_trajectory_generator_b200 import workloads
    from mpc_trajectory_generator_b200.fleet import FleetPlan
    from mpc_trajectory_generator_b200.host import assembly
    hc = assembly.HostConfig.default()
    scs = workloads.random_scenarios(hc, args.complexity, args.distinct, seed=5)
    plan = FleetPlan.from_scenarios(scs, max_steps=args.steps)
    rep = (args.robots + args.distinct - 1) // args.distinct
    tile = lambda a: np.ascontiguousarray(np.concatenate([a] * rep)[:args.robots])  # noqa: E731
    big = FleetPlan(tile(plan.n_ref), tile(plan.ref), tile(plan.n_vert), tile(plan.vert), tile(plan.start),
                    tile(plan.goal), plan.brake_vel, plan.brake_dist, plan.weights, plan.base_speed,
                    plan.circle_radius, plan.sched_init, plan.sched)
    solver = pkg.NmpcSolver(workloads.solver_config_for(hc), device=0)
    fleet = pkg.NmpcFleet(solver, big)
    fleet.step(1)            # warm-up (first-step cold-start solves are the expensive ones)
    t0 = time.perf_counter()
    fleet.step(args.steps)
    wall = time.perf_counter() - t0
    ms = solver.last_kernel_ms
    st = fleet.state()
    npar = pkg.param_len(solver.cfg)
    print(json.dumps({"robots": args.robots, "steps": args.steps, "device_ms": ms, "wall_ms": 1e3 * wall,
                      "robot_steps_per_s": args.robots * args.steps / (ms * 1e-3),
                      "assembled_bytes_per_step": args.robots * npar * 8,
                      "done": int((st["done"] == 1).sum()), "failed": int((st["done"] == 2).sum()),
                      "status_counts": np.bincount(st["status"], minlength=4).tolist()}))
    fleet.close()
    solver.close()


if __name__ == "__main__":
    main()
