#!/usr/bin/env python
"""tools/sensitivity.py — how much do the throughput-relevant statistics depend on the details of OpEn that were
restated from recollection?  (CPU only; oracle/ is test infrastructure.)

The oracle reads NMPC_ORACLE_VARIANT (bit mask) once per batch call; each bit flips one recalled behaviour to the
plausible alternative (oracle/nmpc_oracle.c, g_variant).  For each variant this prints, on 1 024 problems of BASELINE
config 2 and on the recorded reference runs: share of solves that end NotConvergedIterations, mean inner / outer
iterations, mean evaluations, and how far the converged replies move from the baseline restatement."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = [(0, "baseline restatement"),
            (1, "AKKT residual uses the previous iterate's gradient"),
            (2, "exhausted line search falls back to the half step"),
            (4, "Lipschitz estimate restores u"),
            (8, "ALM criterion 1 may hold in the first outer iteration"),
            (16, "penalty may grow after the first outer iteration"),
            (31, "all five flipped")]

CHILD = r'''
import sys, json, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
from oracle import oracle_c as oc
from mpc_trajectory_generator_b200 import workloads
from mpc_trajectory_generator_b200.host import assembly
P, _ = workloads.first_step_batch(assembly.HostConfig.default(), complexity=3, B=1024, seed=0)
cfg = oc.default_config()
U, Y, st, stats = oc.solve_batch(cfg, P)
g = np.load(%r + "/tests/golden/ref_runs.npz")
U2, Y2, st2, stats2 = oc.solve_batch(cfg, g["P"], g["U0"], g["Y0"])
np.savez(sys.argv[1], U=U, st=st, it=stats["inner_iterations"], ot=stats["outer_iterations"],
         ev=stats["n_grad_evals"] + stats["n_cost_evals"], U2=U2, st2=st2, it2=stats2["inner_iterations"])
''' % (ROOT, ROOT, ROOT)


def main():
    import numpy as np
    res = {}
    for bits, name in VARIANTS:
        out = f"/tmp/_sens_{bits}.npz"
        subprocess.check_call([sys.executable, "-c", CHILD, out], env=dict(os.environ, NMPC_ORACLE_VARIANT=str(bits)))
        res[bits] = np.load(out)
    base = res[0]
    print("| variant | config 2: not converged | mean inner it | mean outer it | mean evals | converged replies vs baseline (max rel-L2) "
          "| flags equal | reference runs: not converged | mean inner it |")
    print("|---|---|---|---|---|---|---|---|---|")
    for bits, name in VARIANTS:
        r = res[bits]
        both = (r["st"] == 0) & (base["st"] == 0)
        rel = np.linalg.norm(r["U"] - base["U"], axis=1) / np.maximum(np.linalg.norm(base["U"], axis=1), 1e-12)
        print(f"| {name} | {100 * (r['st'] == 1).mean():.1f} % | {r['it'].mean():.0f} | {r['ot'].mean():.2f} | {r['ev'].mean():.0f} | "
              f"{rel[both].max():.1e} | {100 * (r['st'] == base['st']).mean():.1f} % | {100 * (r['st2'] == 1).mean():.1f} % | {r['it2'].mean():.0f} |")


if __name__ == "__main__":
    main()
