#!/usr/bin/env python
"""tools/stress.py — randomized parity campaign on the GPU box: random sizes, weights, batch sizes (so that the plain and
the latency instantiation, first-wave and queue scheduling, helper warps on/off all occur) against the oracle, bit for bit.
    python tools/stress.py [cases=40] [seed=0]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    import mpc_trajectory_generator_b200 as pkg
    import nmpc_problems as problems
    from oracle import oracle_c
    oracle_c.build()
    rng = np.random.default_rng(seed)
    bad = 0
    t0 = time.time()
    for c in range(cases):
        N = int(rng.choice([5, 10, 20, 20, 20, 31, 32, 33, 40, 64, 65, 80]))
        Nobs = int(rng.choice([0, 1, 7, 10, 10, 33, 50]))
        Nd = int(rng.choice([0, 1, 3, 3]))
        B = int(rng.choice([1, 2, 7, 31, 150, 296, 297, 700, 1900])) if N <= 40 else int(rng.choice([1, 3, 40]))
        w = [problems.DEFAULT_WEIGHTS, problems.SMOOTH_WEIGHTS, problems.MIXED_WEIGHTS, None][int(rng.integers(0, 4))]
        kw = dict(N_hor=N, Nobs=Nobs, Ndynobs=Nd, max_inner_iterations=int(rng.choice([30, 120, 500])),
                  max_outer_iterations=int(rng.choice([2, 5, 10])), lbfgs_memory=int(rng.choice([3, 10])))
        g = pkg.NmpcConfig.default(**kw)
        o = oracle_c.default_config(**kw)
        P = problems.synth(N, Nobs, Nd, B, seed=int(rng.integers(0, 1 << 30)), active=bool(rng.integers(0, 2)), weights=w)
        warm = bool(rng.integers(0, 2))
        U0 = problems.random_controls(N, B, seed=c) * 0.5 if warm else None
        s = pkg.NmpcSolver(g, device=0)
        U, Y, st, stats = s.solve_batch(P, U0)
        s.close()
        Uo, Yo, sto, statso = oracle_c.solve_batch(o, P, U0)
        ok = (np.array_equal(st, sto) and np.array_equal(U, Uo, equal_nan=True) and np.array_equal(Y, Yo, equal_nan=True)
              and all(np.array_equal(stats[k], statso[k]) for k in ("inner_iterations", "n_grad_evals", "n_cost_evals")))
        bad += (not ok)
        print(f"case {c:3d} N={N:2d} Nobs={Nobs:2d} Nd={Nd} B={B:4d} warm={int(warm)} {kw['max_inner_iterations']}x{kw['max_outer_iterations']} "
              f"mem={kw['lbfgs_memory']} flags={np.bincount(st, minlength=4).tolist()} {'ok' if ok else 'MISMATCH'}", flush=True)
    print(f"{cases - bad}/{cases} bit-exact in {time.time() - t0:.0f} s")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
