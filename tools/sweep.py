#!/usr/bin/env python
"""tools/sweep.py — BASELINE config 5: horizon x static-obstacle-slot sweep on one GPU (GPU box).

    python tools/sweep.py [--distinct 512] [--batch 16384] > profiles/r01_sweep.json

For every (N, Nobs) grid point: `--distinct` first-step problems from workloads.sweep_batch (map 11, extra circles from
the map's vertices and random free-space centres), tiled to `--batch`, solved device-resident (best of 3 launches,
CUDA events).  Prints one JSON object per grid point: solves/s, ms per solve, algorithmic bytes and the HBM fraction as
the task defines it, mean inner iterations and exit flags, warps per SM the arena allowed.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--distinct", type=int, default=512)
    ap.add_argument("--batch", type=int, default=16384)
    ap.add_argument("--horizons", default="10,20,40,80")
    ap.add_argument("--slots", default="10,50,100,200")
    args = ap.parse_args()
    import torch
    import bench
    import mpc_trajectory_generator_b200 as pkg
    from mpc_trajectory_generator_b200 import workloads
    peak, _ = bench.load_peaks()
    dev = torch.device("cuda", 0)
    for N in [int(x) for x in args.horizons.split(",")]:
        for Nobs in [int(x) for x in args.slots.split(",")]:
            P0, hc = workloads.sweep_batch(N, Nobs, B=args.distinct, seed=2)
            rep = (args.batch + args.distinct - 1) // args.distinct
            P = np.ascontiguousarray(np.concatenate([P0] * rep)[:args.batch])
            B = P.shape[0]
            cfg = workloads.solver_config_for(hc)
            try:
                s = pkg.NmpcSolver(cfg, device=0)
            except pkg.NmpcError as e:
                print(json.dumps({"N": N, "Nobs": Nobs, "error": str(e)}), flush=True)
                continue
            dP = torch.from_numpy(P).to(dev)
            dU = torch.zeros((B, 2 * N), dtype=torch.float64, device=dev)
            dY = torch.zeros_like(dU)
            dst = torch.zeros(B, dtype=torch.int32, device=dev)
            dstats = torch.zeros((B, 64), dtype=torch.uint8, device=dev)
            ms = []
            for _ in range(3):
                dU.zero_(); dY.zero_()
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                s.solve_batch_device(B, dP.data_ptr(), dU.data_ptr(), dY.data_ptr(), dst.data_ptr(), dstats.data_ptr(), 0)
                e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            best = min(ms)
            stats = np.frombuffer(dstats.cpu().numpy().tobytes(), dtype=pkg.STATS_DTYPE)
            ab = bench.algorithmic_bytes(N, Nobs, hc.Ndynobs)
            gbs = ab * B / (best * 1e-3) / 1e9
            print(json.dumps({"N": N, "Nobs": Nobs, "batch": B, "distinct": args.distinct, "ms": round(best, 3),
                              "solves_per_s": round(B / best * 1e3, 1), "us_per_solve": round(1e3 * best / B, 3),
                              "algorithmic_bytes_per_solve": ab, "hbm_gbs": round(gbs, 4), "hbm_frac": gbs / peak,
                              "inner_iterations_mean": float(stats["inner_iterations"].mean()),
                              "evals_per_solve": float((stats["n_cost_evals"] + stats["n_grad_evals"]).mean()),
                              "exit_status_counts": np.bincount(dst.cpu().numpy(), minlength=4).tolist()}), flush=True)
            s.close()


if __name__ == "__main__":
    main()
