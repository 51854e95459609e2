#!/usr/bin/env python
"""tools/variants.py — build and time tuning variants of the solve kernel (same sources, extra -D flags:
NMPC_WARPS, NMPC_SEG_UNR, NMPC_ICLAMP, NMPC_OOL_DIV, NMPC_PROFILE).

    python tools/variants.py build  name=FLAG1,FLAG2 ...     (CPU: nvcc cross-compiles; .so files go to tools/_build/)
    python tools/variants.py run    [--batch 4096] [--big 32768] name ...   (GPU box)

Each variant is loaded in its own process (NMPC_B200_LIB), solves the config-2 batch and a
large synthetic batch, checks bit-exactness against the shipped library's output and prints one line.
Not part of the product path; the shipped library is always mpc_trajectory_generator_b200/libnmpc_b200.so.
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "tools", "_build")


def lib_path(name):
    return os.path.join(OUT, f"libnmpc_b200_{name}.so")


def cmd_build(specs):
    from mpc_trajectory_generator_b200 import _build
    os.makedirs(OUT, exist_ok=True)
    for spec in specs:
        name, _, flags = spec.partition("=")
        defines = [f for f in flags.split(",") if f]
        t = time.time()
        _build.build_variant(lib_path(name), defines)
        print(f"built {name} {defines} in {time.time() - t:.1f}s")


def workload(batch, big):
    import numpy as np
    cache = f"/tmp/nmpc_variants_{batch}_{big}.npz"
    if os.path.exists(cache):
        z = np.load(cache)
        return z["P2"], z["Ps"]
    from mpc_trajectory_generator_b200 import workloads
    from mpc_trajectory_generator_b200.host import assembly
    import nmpc_problems
    hc = assembly.HostConfig.default()
    P2, _ = workloads.first_step_batch(hc, complexity=3, B=batch, seed=0)
    Ps = nmpc_problems.synth(20, 10, 3, big, seed=0, active=False)
    np.savez(cache, P2=P2, Ps=Ps)
    return P2, Ps


def cmd_one(name, batch, big, reps):
    import numpy as np
    import torch
    import mpc_trajectory_generator_b200 as pkg
    P2, Ps = workload(batch, big)
    s = pkg.NmpcSolver(pkg.NmpcConfig.default(), device=0)
    dev = torch.device("cuda", 0)
    res = {"variant": name}
    for tag, P in (("config2", P2), ("synth", Ps)):
        B = P.shape[0]
        dP = torch.from_numpy(P).to(dev)
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
        best = min(ms[1:])
        U = dU.cpu().numpy()
        ref = f"/tmp/nmpc_variants_ref_{tag}_{B}.npy"
        if name == "base" or not os.path.exists(ref):
            np.save(ref, U)
            same = None
        else:
            R = np.load(ref)
            same = bool(np.array_equal(R, U, equal_nan=True))
            if not same:
                bad = np.nonzero(~np.all((R == U) | (np.isnan(R) & np.isnan(U)), axis=1))[0]
                res[tag + "_mismatch_rows"] = [int(len(bad))] + [int(x) for x in bad[:8]]
        res[tag] = {"B": B, "ms": round(best, 3), "solves_per_s": round(B / best * 1e3, 1), "bit_exact_vs_base": same}
    # one problem alone (the hardest of config 2): the latency that bounds the config-2 step
    ms = []
    for r in range(3):
        _, _, st1, stats1 = s.solve_batch(P2[1207:1208])
        ms.append(s.last_kernel_ms)
    res["lone_ms"] = round(min(ms), 3)
    s.close()
    print(json.dumps(res), flush=True)


def cmd_run(args):
    batch, big, reps, names = 4096, 32768, 3, []
    it = iter(args)
    for a in it:
        if a == "--batch":
            batch = int(next(it))
        elif a == "--big":
            big = int(next(it))
        elif a == "--reps":
            reps = int(next(it))
        else:
            names.append(a)
    workload(batch, big)
    for name in names:
        env = dict(os.environ)
        if name.startswith("waves"):   # the shipped library with another helper-warp threshold (wavesN; waves0 = no helpers)
            env["NMPC_B200_HELP_MAX_WAVES"] = name[5:]
        elif name != "base":
            env["NMPC_B200_LIB"] = lib_path(name)
        subprocess.call([sys.executable, os.path.abspath(__file__), "_one", name, str(batch), str(big), str(reps)], env=env)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        cmd_build(sys.argv[2:])
    elif sys.argv[1] == "run":
        cmd_run(sys.argv[2:])
    elif sys.argv[1] == "_one":
        cmd_one(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))
