#!/usr/bin/env python
"""tools/perfect_order.py — what would a perfect longest-first order be worth on config 2?  (NMPC_DEBUG_ORDER build:
tools/variants.py build dord=NMPC_DEBUG_ORDER.)  Solves the batch once, then again handing the problems out by their TRUE
inner-iteration counts (descending), and with a few other orders."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import variants  # noqa: E402
os.environ["NMPC_B200_LIB"] = variants.lib_path("dord")


def main():
    import numpy as np
    import torch
    import mpc_trajectory_generator_b200 as pkg
    P2, _ = variants.workload(4096, 32768)
    s = pkg.NmpcSolver(pkg.NmpcConfig.default(), device=0)
    U, Y, st, stats = s.solve_batch(P2)
    it = stats["inner_iterations"].astype(np.int64)
    dev = torch.device("cuda", 0)
    B = P2.shape[0]
    dP = torch.from_numpy(P2).to(dev)
    dU = torch.zeros((B, 40), dtype=torch.float64, device=dev)
    dY = torch.zeros_like(dU)
    dst = torch.zeros(B, dtype=torch.int32, device=dev)
    s._lib.nmpc_debug_set_order.argtypes = [C.c_void_p, C.c_int32]

    def run(order, reps=7):
        if order is None:
            s._lib.nmpc_debug_set_order(None, 0)
        else:
            d = torch.from_numpy(order.astype(np.int32)).to(dev)
            s._lib.nmpc_debug_set_order(C.c_void_p(d.data_ptr()), B)
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
        return float(np.median(ms[1:])), float(min(ms[1:]))
    desc = np.argsort(-it, kind="stable")
    # snake over the 592 schedulers: ranks 0..591 one per scheduler, 592..1183 reversed, ...
    slots = 148 * 12
    print("probe order (shipped)     ", run(None))
    print("true longest first        ", run(desc))
    rng = np.random.default_rng(0)
    print("random order              ", run(rng.permutation(B)))
    # LPT by scheduler: the first wave's slot j = c + 148 w sits on scheduler (c, w % 4); pair the longest with the shortest
    first = desc[:slots].copy()
    per = 592
    lay = np.empty(slots, dtype=np.int64)
    lay[:per] = first[:per]
    lay[per:2 * per] = first[per:2 * per][::-1]
    lay[2 * per:] = first[2 * per:]
    # slot index j -> (c, w): j = c + 148 w; scheduler of slot = (c, w % 4): ranks r < 592 map to w in 0..3 (r = c + 148 w), so
    # rank r and rank 592 + r share a scheduler: reversing the second block pairs long with short
    order = np.concatenate([lay, desc[slots:]])
    print("true order, snake pairing ", run(order))
    f = os.path.join(ROOT, "tools", "_build", "c2_orders.npz")
    if os.path.exists(f):   # orders made on the CPU from the probe's key (oracle evaluation): exact sort, bucketed, snake pairing
        z = np.load(f)
        for k in z.files:
            print(f"{k:26s}", run(z[k]))
    s.close()


if __name__ == "__main__":
    main()
