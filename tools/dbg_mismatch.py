#!/usr/bin/env python
"""tools/dbg_mismatch.py — GPU box: solve a synthetic batch on the GPU and with the oracle, list the problems whose
replies differ and both sides' counters (which path of the phase machine they took)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402


def main():
    N, Nobs, B, seed, active = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in
                                ((1, 20), (2, 10), (3, 900), (4, 1920), (5, 1)))
    import mpc_trajectory_generator_b200 as pkg
    import nmpc_problems as problems
    from oracle import oracle_c
    g = pkg.NmpcConfig.default(N_hor=N, Nobs=Nobs, Ndynobs=3)
    o = oracle_c.default_config(N_hor=N, Nobs=Nobs, Ndynobs=3)
    P = problems.synth(N, Nobs, 3, B, seed=seed, active=bool(active))
    s = pkg.NmpcSolver(g, device=0)
    U, Y, st, stats = s.solve_batch(P)
    Uo, Yo, sto, so = oracle_c.solve_batch(o, P)
    bad = np.nonzero((U != Uo).any(axis=1) | (st != sto))[0]
    print("mismatching problems:", len(bad), bad[:20].tolist())
    for b in bad[:8]:
        print(b, "gpu", int(st[b]), [int(stats[k][b]) for k in ("outer_iterations", "inner_iterations", "n_grad_evals", "n_cost_evals")],
              "oracle", int(sto[b]), [int(so[k][b]) for k in ("outer_iterations", "inner_iterations", "n_grad_evals", "n_cost_evals")],
              "max|dU|", float(np.abs(U[b] - Uo[b]).max()))
        # truncated budgets: where does the first difference appear?
        for cap in (1, 2, 3, 5, 10, 20, 50, 100, 200, 500):
            gc = pkg.NmpcConfig.default(N_hor=N, Nobs=Nobs, Ndynobs=3, max_inner_iterations=cap, max_outer_iterations=1)
            oc = oracle_c.default_config(N_hor=N, Nobs=Nobs, Ndynobs=3, max_inner_iterations=cap, max_outer_iterations=1)
            sc = pkg.NmpcSolver(gc, device=0)
            u1, y1, s1, t1 = sc.solve_batch(P[b:b + 1])
            sc.close()
            u2, y2, s2, t2 = oracle_c.solve_batch(oc, P[b:b + 1])
            same = np.array_equal(u1, u2)
            print("   cap", cap, "same" if same else "DIFF", [int(t1[k][0]) for k in ("inner_iterations", "n_grad_evals", "n_cost_evals")],
                  [int(t2[k][0]) for k in ("inner_iterations", "n_grad_evals", "n_cost_evals")])
            if not same:
                break
    s.close()


if __name__ == "__main__":
    main()
