#!/usr/bin/env python
"""tools/prof_cycles.py — clock64 section profile of the solve kernel (GPU box).

Builds the sources with -DNMPC_PROFILE (tools/_build/libnmpc_b200_prof.so), solves the K hardest
problems of the config-2 batch ONE AT A TIME (a lone warp on an otherwise idle GPU: the regime that
bounds the B=4096 makespan) and prints cycles per PANOC iteration and per evaluation section.
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import variants  # noqa: E402

PHASES = ["OUTER_BEGIN", "INIT", "STEP_BEGIN", "A", "RETRY", "LS", "STEP_DONE", "SOLVE_END", "F2", "EXIT"]
SECT = ["theta_scan_sincos", "xy_scan", "cte", "obstacles", "cost_butterfly", "adjoint"]


def main():
    extra = [a for a in sys.argv[1:] if a.startswith("NMPC_")]
    name = "prof"
    if "--lib" in sys.argv:
        name = sys.argv[sys.argv.index("--lib") + 1]
    lib = variants.lib_path(name)
    if not os.path.exists(lib) or "--rebuild" in sys.argv:
        from mpc_trajectory_generator_b200 import _build
        os.makedirs(variants.OUT, exist_ok=True)
        _build.build_variant(lib, ["NMPC_PROFILE"] + extra)
    os.environ["NMPC_B200_LIB"] = lib
    import numpy as np
    import torch
    import mpc_trajectory_generator_b200 as pkg
    P2, _ = variants.workload(4096, 32768)
    s = pkg.NmpcSolver(pkg.NmpcConfig.default(), device=0)
    U, Y, st, stats = s.solve_batch(P2)
    order = np.argsort(-stats["inner_iterations"])
    dbg = torch.zeros(48, dtype=torch.int64, device="cuda")
    s._lib.nmpc_debug_set_buffer(C.c_void_p(dbg.data_ptr()))
    for b in list(order[:3]) + [int(order[len(order) // 2])]:
        dbg.zero_()
        _, _, st1, stats1 = s.solve_batch(P2[b:b + 1])
        d = dbg.cpu().numpy()
        it = int(stats1["inner_iterations"][0])
        ng, nc = int(d[1]), 0   # every call of eval() computes psi and grad psi at one point per group
        out = {"problem": int(b), "inner_iterations": it, "cycles_per_iteration": round(d[6] / max(it, 1)),
               "eval_calls": ng, "cycles_per_eval_call": round(d[0] / max(ng, 1)),
               "cost_evals": nc, "cycles_per_cost_eval": round(d[2] / max(nc, 1)),
               "lbfgs_calls": int(d[5]), "cycles_per_lbfgs": round(d[4] / max(int(d[5]), 1)),
               "eval_share": round(float(d[0] + d[2]) / d[6], 3), "lbfgs_share": round(float(d[4]) / d[6], 3),
               "sections_cycles_per_eval": {k: round(d[8 + i] / max(ng + nc, 1)) for i, k in enumerate(SECT)},
               "kernel_ms": s.last_kernel_ms,
               "phase_pre_cycles_per_iteration": {PHASES[i]: round(d[16 + i] / max(it, 1)) for i in range(len(PHASES)) if d[16 + i]},
               "phase_post_cycles_per_iteration": {PHASES[i]: round(d[32 + i] / max(it, 1)) for i in range(len(PHASES)) if d[32 + i]}}
        print(json.dumps(out), flush=True)
    s.close()


if __name__ == "__main__":
    main()
