#!/usr/bin/env python
"""tools/sat.py — the saturated launch (synthetic B = 32 768: every warp always owns a problem), twice; for ncu captures."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import variants  # noqa: E402


def main():
    import mpc_trajectory_generator_b200 as pkg
    _, Ps = variants.workload(4096, 32768)
    s = pkg.NmpcSolver(pkg.NmpcConfig.default(), device=0)
    for _ in range(2):
        U, Y, st, stats = s.solve_batch(Ps)
        print(len(Ps), s.last_kernel_ms, flush=True)
    s.close()


if __name__ == "__main__":
    main()
