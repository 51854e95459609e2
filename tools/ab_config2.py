#!/usr/bin/env python
"""tools/ab_config2.py — careful A/B of the config-2 step between library variants / helper thresholds: the variants take
turns, many launches each, medians and quartiles (the step time repeats to about 1.5 % between single launches).
    python tools/ab_config2.py [--reps 12] name ...     names as in tools/variants.py (base, wavesN, <built variant>)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import variants  # noqa: E402


def one(reps):
    import numpy as np
    import torch
    import mpc_trajectory_generator_b200 as pkg
    P2, _ = variants.workload(4096, 32768)
    s = pkg.NmpcSolver(pkg.NmpcConfig.default(), device=0)
    dev = torch.device("cuda", 0)
    B = P2.shape[0]
    dP = torch.from_numpy(P2).to(dev)
    dU = torch.zeros((B, 40), dtype=torch.float64, device=dev)
    dY = torch.zeros_like(dU)
    dst = torch.zeros(B, dtype=torch.int32, device=dev)
    ms = []
    for r in range(reps + 1):
        dU.zero_(); dY.zero_()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        s.solve_batch_device(B, dP.data_ptr(), dU.data_ptr(), dY.data_ptr(), dst.data_ptr(), 0, 0)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    s.close()
    print(json.dumps(ms[1:]))


def main():
    if sys.argv[1] == "_one":
        return one(int(sys.argv[2]))
    reps, names = 12, []
    it = iter(sys.argv[1:])
    for a in it:
        if a == "--reps":
            reps = int(next(it))
        else:
            names.append(a)
    import numpy as np
    variants.workload(4096, 32768)
    res = {n: [] for n in names}
    for rnd in range(3):
        for n in names:
            env = dict(os.environ)
            if n.startswith("waves"):
                env["NMPC_B200_HELP_MAX_WAVES"] = n[5:]
            elif n != "base":
                env["NMPC_B200_LIB"] = variants.lib_path(n)
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "_one", str(reps // 3)], env=env, capture_output=True, text=True)
            res[n] += json.loads(out.stdout.strip().splitlines()[-1])
    for n in names:
        v = np.array(res[n])
        print(f"{n:10s} n={len(v)} median {np.median(v):.2f} ms  q25 {np.percentile(v, 25):.2f}  q75 {np.percentile(v, 75):.2f}  min {v.min():.2f}  "
              f"-> {4096 / np.median(v):.1f} k solves/s", flush=True)


if __name__ == "__main__":
    main()
