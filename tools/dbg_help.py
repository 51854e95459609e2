#!/usr/bin/env python
"""tools/dbg_help.py — small solves with the helper warps active (debug builds print what they find)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import variants
import mpc_trajectory_generator_b200 as pkg
P2, _ = variants.workload(4096, 32768)
s = pkg.NmpcSolver(pkg.NmpcConfig.default(), device=0)
for B in [int(a) for a in sys.argv[1:]] or [1, 7]:
    U, Y, st, stats = s.solve_batch(P2[:B])
    print(B, "ms", round(s.last_kernel_ms, 3), "status", np.bincount(st, minlength=3).tolist(), "iters", int(stats["inner_iterations"].sum()),
          "chk", float(np.nansum(U)), flush=True)
s.close()
