#!/usr/bin/env python
"""tools/lone.py — solve ONE problem of the config-2 batch alone on the GPU (the regime that bounds the B=4096 step):
    python tools/lone.py [index=1207] [repeats=2]        (run it under ncu for the lone-warp stall profile)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import variants  # noqa: E402


def main():
    idx = int(sys.argv[1]) if len(sys.argv) > 1 else 1207
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    import mpc_trajectory_generator_b200 as pkg
    P2, _ = variants.workload(4096, 32768)
    s = pkg.NmpcSolver(pkg.NmpcConfig.default(), device=0)
    for _ in range(reps):
        U, Y, st, stats = s.solve_batch(P2[idx:idx + 1])
        print(idx, int(st[0]), int(stats["inner_iterations"][0]), s.last_kernel_ms, flush=True)
    s.close()


if __name__ == "__main__":
    main()
