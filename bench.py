#!/usr/bin/env python
"""bench.py — NMPC solves/sec (N=20, batched) on B200, next to the CPU restatement.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload config2|config3|config4|synthetic] [--batch B] [--extra light|full|off]

A "step" is one pass of the hot path (one batched solve launch) over one batch of
synthetic NMPC instances.  Headline workload = BASELINE.json configs[1]: B=4096 random
start/goal pairs on map complexity=3, N=20, first-step problems, cold start.
Multi-GPU (torchrun, one rank per GPU): the batch dimension shards with no data-path
collective — every rank solves its own B problems (weak scaling); the only NCCL traffic is
a broadcast of the static map / config table from rank 0 (and the gathers of the timings).

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (inputs already in
HBM); `e2e` goes through the public host-buffer API (NmpcSolver.solve_batch_into:
pinned host buffers, H2D + D2H inside the timed region).  `extra.configs` carries, from the
same run, the other BASELINE configs at their stated per-GPU size: config 3 (65 536 recorded
receding-horizon steps, weak), config 4 (this rank's 32 768-row shard of the 262 144 N=40
problems: strong split over 8 GPUs via sharding.shard_bounds, at fewer GPUs the same shard
size per GPU) and corner points of the config-5 grid, each with per-rank kernel times.
`--impl reference` times the CPU restatement of the reference's OpEn path (oracle/, all host
threads, built -O3 -march=native on the box) on the SAME batch as the headline — OpEn itself
cannot be installed here.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "nmpc_solves_per_sec"
UNIT = "solves/s"


def algorithmic_bytes(N, Nobs, Nd):
    """8*np + 8*2N (U0 in) + 8*2N (U out) + 48 (status/stats)."""
    npar = 20 + N + 3 * Nobs + 5 * Nd * N + 3 * N
    return 8 * npar + 2 * 8 * 2 * N + 48


def eval_flops(N, Nobs, Nd):
    """SURVEY.md §8d estimate of one psi forward evaluation (gradient ~3x)."""
    return N * (24 + 9 * Nobs + 16 * Nd + 18 * (N - 1)) + 10 * N


def kernel_source_hash():
    """sha256 over the CUDA sources + public header: ties a committed ncu capture to the kernel it was taken from."""
    from mpc_trajectory_generator_b200 import _build
    h = hashlib.sha256()
    for f in _build.sources():
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def load_traffic(workload, batch):
    """DRAM bytes per launch of the solve kernel from the committed `ncu --set full` capture
    (profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum) — only if the capture was taken from the
    kernel sources that are being benchmarked (source hash), otherwise null."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            t = json.load(f)
        if t.get("kernel_source_hash") != kernel_source_hash():
            return None
        for e in t["captures"]:
            if e["workload"] == workload and e["batch"] == batch:
                return e["dram_bytes_per_launch"]
    return None


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _closed_loop_rows(name, batch, seed, device):
    """config 3 / 4: `batch` recorded receding-horizon steps (map 11; config 4: N=40, smooth_velocity weights and
    bounds), each with the warm start the server held.  GPU arm: recorded with the fleet kernels; CPU arm: host loop +
    oracle (tests/test_gpu_more.py asserts both give the same rows).  Robots are added until `batch` rows exist."""
    from mpc_trajectory_generator_b200 import workloads
    from mpc_trajectory_generator_b200.host import assembly
    hc = assembly.HostConfig.smooth_velocity(N_hor=40) if name == "config4" else assembly.HostConfig.default()
    parts, rows, rnd = [], 0, 0
    robots = max(8, int(round(batch ** 0.5)))
    steps = max(1, min(256, -(-batch // robots)))
    while rows < batch and rnd < 8:
        if device is not None:
            import mpc_trajectory_generator_b200 as pkg
            s = pkg.NmpcSolver(workloads.solver_config_for(hc), device=device)
            rec = workloads.closed_loop_batch_device(s, hc, complexity=11, robots=robots, steps=steps, seed=seed + 1 + 97 * rnd)
            s.close()
        else:
            from oracle import oracle_c
            ocfg = oracle_c.default_config(**{k: getattr(workloads.solver_config_for(hc), k) for k in
                                              ("N_hor", "Nobs", "Ndynobs", "ang_vel_max", "ang_acc_max")})
            rec = workloads.closed_loop_batch(
                hc, lambda P, U0, Y0: oracle_c.solve_batch(ocfg, P, U0, Y0, nthreads=host_threads())[:3],
                complexity=11, robots=robots, steps=steps, seed=seed + 1 + 97 * rnd, sincos=oracle_c.sincos)
        parts.append(rec)
        rows += rec["P"].shape[0]
        rnd += 1
        robots = max(8, int(robots * max(0.25, min(1.0, (batch - rows) / max(rec["P"].shape[0], 1))) + 1))
    P = np.concatenate([r["P"] for r in parts])[:batch]
    U0 = np.concatenate([r["U0"] for r in parts])[:batch]
    Y0 = np.concatenate([r["Y0"] for r in parts])[:batch]
    desc = (f"configs[{2 if name == 'config3' else 3}]: {P.shape[0]} recorded receding-horizon steps "
            f"(map complexity=11, N={hc.N_hor}{', smooth_velocity.yaml weights / bounds' if name == 'config4' else ''}), "
            f"warm start = previous solution")
    return P, U0, Y0, hc, desc


def make_workload(name, batch, seed, N=20, device=None):
    """-> (P, U0, Y0, host_cfg, description)."""
    from mpc_trajectory_generator_b200 import workloads
    from mpc_trajectory_generator_b200.host import assembly
    hc = assembly.HostConfig.default(N_hor=N)
    if name in ("config3", "config4"):
        return _closed_loop_rows(name, batch, seed, device)
    if name == "config2":
        P, _ = workloads.first_step_batch(hc, complexity=3, B=batch, seed=seed)
        desc = f"configs[1]: batch={batch} random start/goal pairs, map complexity=3, N={N}, first step, cold start"
        return P, None, None, hc, desc
    if name == "synthetic":
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import nmpc_problems
        P = nmpc_problems.synth(N, hc.Nobs, hc.Ndynobs, batch, seed=seed, active=False)
        return P, None, None, hc, f"synthetic piecewise-linear references, batch={batch}, N={N}"
    raise SystemExit(f"unknown workload {name}")


def run_reference(args, rank, world):
    """CPU arm: the OpEn-equivalent restatement (oracle/) on the host cores, on the same batch as the headline."""
    if rank != 0:
        return
    from oracle import oracle_c
    oracle_c.build()
    native = oracle_c.use_native()      # -O3 -march=native, compiled on this host
    P, U0, Y0, hc, desc = make_workload(args.workload, args.batch, args.seed)
    cfg = oracle_c.default_config(N_hor=hc.N_hor, Nobs=hc.Nobs, Ndynobs=hc.Ndynobs, ang_vel_max=hc.ang_vel_max,
                                  ang_acc_max=hc.ang_acc_max)
    threads = host_threads()   # torchrun pins OMP_NUM_THREADS=1; the CPU arm uses every host core it may run on
    sample = P.shape[0] if args.ref_sample <= 0 else min(args.ref_sample, P.shape[0])
    Ps = P[:sample]
    U0s = None if U0 is None else U0[:sample]
    Y0s = None if Y0 is None else Y0[:sample]
    for _ in range(min(args.warmup, 1)):
        oracle_c.solve_batch(cfg, Ps[:min(64, sample)], None if U0s is None else U0s[:64],
                             None if Y0s is None else Y0s[:64], nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, _, st, _ = oracle_c.solve_batch(cfg, Ps, U0s, Y0s, nthreads=threads)
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    what = "the whole batch" if sample == P.shape[0] else f"first {sample} problems of the batch"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "batch_per_gpu": int(P.shape[0]), "N_hor": hc.N_hor, "Nobs": hc.Nobs,
                       "Ndynobs": hc.Ndynobs, "seed": args.seed, "sample": f"{what} per step"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{what}, {args.steps} passes, OpenMP over problems ({threads} threads); "
                                       f"OpEn-equivalent C restatement (oracle/nmpc_oracle.c, "
                                       f"{'-O3 -march=native' if native else '-O3 -march=x86-64-v3'}): the parity checker "
                                       f"doubling as the baseline — OpEn itself is not installable here",
                             "exit_status_counts": np.bincount(st, minlength=4).tolist()},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


class DeviceBatch:
    """One workload resident on the device + the event-timed step."""

    def __init__(self, solver, P, U0, Y0, dev, stream):
        import torch
        self.torch, self.solver, self.stream, self.dev = torch, solver, stream, dev
        self.B, self.n2 = P.shape[0], solver.n2
        self.dP = torch.from_numpy(P).to(dev)
        z = lambda: torch.zeros((self.B, self.n2), dtype=torch.float64, device=dev)  # noqa: E731
        self.dU0 = z() if U0 is None else torch.from_numpy(U0).to(dev)
        self.dY0 = z() if Y0 is None else torch.from_numpy(Y0).to(dev)
        self.dU, self.dY = torch.empty_like(self.dU0), torch.empty_like(self.dY0)
        self.dstatus = torch.zeros(self.B, dtype=torch.int32, device=dev)
        self.dstats = torch.zeros((self.B, 64), dtype=torch.uint8, device=dev)

    def step(self):
        torch = self.torch
        self.dU.copy_(self.dU0)
        self.dY.copy_(self.dY0)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        self.solver.solve_batch_device(self.B, self.dP.data_ptr(), self.dU.data_ptr(), self.dY.data_ptr(),
                                       self.dstatus.data_ptr(), self.dstats.data_ptr(), self.stream.cuda_stream)
        e1.record(self.stream)
        return e0, e1

    def stats(self):
        import mpc_trajectory_generator_b200 as pkg
        return np.frombuffer(self.dstats.cpu().numpy().tobytes(), dtype=pkg.STATS_DTYPE)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--ref-sample", type=int, default=0, help="problems per step for the CPU arm (0 = the whole batch)")
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extra", default="light", choices=["light", "full", "off"],
                    help="also time BASELINE configs 3, 4 and corner points of config 5 in this run (extra.configs)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the solver has no CPU fallback); "
                         "use --impl reference for the CPU arm")
    import torch.distributed as dist
    import mpc_trajectory_generator_b200 as pkg
    from mpc_trajectory_generator_b200 import workloads

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # rank r generates its own shard (weak scaling: per-GPU batch fixed)
    P, U0, Y0, hc, desc = make_workload(args.workload, args.batch, args.seed + rank, device=local_rank)
    B = P.shape[0]
    N, Nobs, Nd = hc.N_hor, hc.Nobs, hc.Ndynobs
    cfg = workloads.solver_config_for(hc)
    if world > 1:
        # the only collective on this path: the batch-invariant table (config scalars + cost weights)
        # goes out from rank 0 over NCCL; every rank builds its solver from what it received.
        from mpc_trajectory_generator_b200 import sharding
        cfg, weights = sharding.broadcast_static_table(cfg, P[0, 10:20], device=dev, src=0)
        P[:, 10:20] = np.asarray(weights)
    solver = pkg.NmpcSolver(cfg, device=local_rank)
    n2 = 2 * N

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.Stream(dev)   # the solve kernel, its input copies and the timing events share this stream
    torch.cuda.set_stream(stream)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(batch, steps, warmup):
        """W warm-up steps, then exactly K timed ones (barrier + synchronize on both sides, L2 flushed between steps,
        outside the event pairs); -> (per-step ms on this rank, launches)"""
        for _ in range(warmup):
            batch.step()
            flush.fill_(1)
        sync_all()
        l0 = batch.solver.launch_count
        pairs = []
        for _ in range(steps):
            pairs.append(batch.step())
            flush.fill_(1)
        sync_all()
        return [a.elapsed_time(b) for a, b in pairs], batch.solver.launch_count - l0

    def over_ranks(x):
        """-> (max over ranks, list of every rank's value)"""
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world == 1:
            return float(x), [float(x)]
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        v = [float(a[0]) for a in allv]
        return max(v), v

    # ---------------- device-resident arm (`value`) ----------------
    head = DeviceBatch(solver, P, U0, Y0, dev, stream)
    sampler = ClockSampler(local_rank)
    warm = max(args.warmup, 3)
    for _ in range(warm):
        head.step()
        flush.fill_(1)
    sync_all()
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    kernel_ms, launches = timed(head, args.steps, 0)
    t_wall = time.perf_counter() - t_wall0
    total_ms = float(sum(kernel_ms))
    clocks = sampler.stop() if rank == 0 else None
    status = head.dstatus.cpu().numpy()
    stats = head.stats()

    # ---------------- end-to-end arm (`e2e`): public API, pinned host buffers ----------------
    hP = torch.from_numpy(P).pin_memory()
    hU = torch.zeros((B, n2), dtype=torch.float64).pin_memory()
    hY = torch.zeros((B, n2), dtype=torch.float64).pin_memory()
    hstatus = torch.zeros(B, dtype=torch.int32).pin_memory()
    U0h = np.zeros((B, n2)) if U0 is None else U0
    Y0h = np.zeros((B, n2)) if Y0 is None else Y0

    def step_e2e():
        hU.numpy()[:] = U0h
        hY.numpy()[:] = Y0h
        t0 = time.perf_counter()
        solver.solve_batch_into(hP.numpy(), hU.numpy(), hY.numpy(), hstatus.numpy(), None)
        return time.perf_counter() - t0

    for _ in range(2):
        step_e2e()
    sync_all()
    e2e_s = 0.0
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        e2e_s += step_e2e()
        flush.fill_(1)
        torch.cuda.synchronize(dev)
    h2d = P.nbytes + 2 * B * n2 * 8
    d2h = 2 * B * n2 * 8 + B * 4

    total_ms_max, per_rank_ms = over_ranks(total_ms)
    e2e_ms_step, _ = over_ranks(e2e_s * 1e3 / e2e_steps)
    value = world * B * args.steps / (total_ms_max * 1e-3)
    e2e_value = world * B / (e2e_ms_step * 1e-3)

    # ---------------- the other BASELINE configs, same run (`extra.configs`) ----------------
    extra = []
    if args.extra != "off":
        full = args.extra == "full"
        specs = [("config3", dict(workload="config3", batch=65536, steps=3, scaling="weak")),
                 ("config4", dict(workload="config4", batch=32768, steps=2, scaling="strong at 8 GPUs: 262144 / 8 per GPU")),
                 ]
        corners = [(10, 10), (10, 200), (80, 10), (80, 200)] if not full else \
            [(n, o) for n in (10, 20, 40, 80) for o in (10, 50, 100, 200)]
        for name, sp in specs:
            t_gen = time.perf_counter()
            Pe, U0e, Y0e, hce, desce = make_workload(sp["workload"], sp["batch"], args.seed + rank, device=local_rank)
            gen_s = time.perf_counter() - t_gen
            se = solver if hce.N_hor == N and hce.Nobs == Nobs else pkg.NmpcSolver(workloads.solver_config_for(hce), device=local_rank)
            be = DeviceBatch(se, Pe, U0e, Y0e, dev, stream)
            ms, _ = timed(be, sp["steps"], 1)
            mx, per = over_ranks(sum(ms))
            ste = be.stats()
            Be = Pe.shape[0]
            ab = algorithmic_bytes(hce.N_hor, hce.Nobs, hce.Ndynobs)
            extra.append({"config": name, "workload": desce, "batch_per_gpu": int(Be), "steps": sp["steps"],
                          "scaling": sp["scaling"], "value": world * Be * sp["steps"] / (mx * 1e-3), "unit": UNIT,
                          "ms_per_step": mx / sp["steps"], "kernel_ms_per_step_by_rank": [p / sp["steps"] for p in per],
                          "hbm_frac": ab * Be * sp["steps"] / (sum(ms) * 1e-3) / 1e9 / load_peaks()[0],
                          "inner_iterations_mean": float(ste["inner_iterations"].mean()),
                          "exit_status_counts": np.bincount(be.dstatus.cpu().numpy(), minlength=4).tolist(),
                          "workload_generation_s": gen_s})
            del be
            if se is not solver:
                se.close()
        distinct, Bc = 256, 8192
        for (Nc, Oc) in corners:
            t_gen = time.perf_counter()
            P0, hcc = workloads.sweep_batch(Nc, Oc, B=distinct, seed=2 + rank)
            Bcc = Bc if Nc < 80 else Bc // 4
            Pc = np.ascontiguousarray(np.concatenate([P0] * (-(-Bcc // distinct)))[:Bcc])
            gen_s = time.perf_counter() - t_gen
            sc = pkg.NmpcSolver(workloads.solver_config_for(hcc), device=local_rank)
            bc = DeviceBatch(sc, Pc, None, None, dev, stream)
            ms, _ = timed(bc, 1, 1)
            mx, per = over_ranks(sum(ms))
            stc = bc.stats()
            ab = algorithmic_bytes(Nc, Oc, hcc.Ndynobs)
            extra.append({"config": "config5", "N_hor": Nc, "Nobs": Oc, "workload": f"configs[4] grid point: {distinct} "
                          f"first-step problems on map 11 tiled to {Bcc} per GPU, cold start", "batch_per_gpu": int(Bcc),
                          "steps": 1, "scaling": "weak", "value": world * Bcc / (mx * 1e-3), "unit": UNIT, "ms_per_step": mx,
                          "kernel_ms_per_step_by_rank": per, "algorithmic_bytes_per_solve": ab,
                          "hbm_frac": ab * Bcc / (sum(ms) * 1e-3) / 1e9 / load_peaks()[0],
                          "inner_iterations_mean": float(stc["inner_iterations"].mean()),
                          "exit_status_counts": np.bincount(bc.dstatus.cpu().numpy(), minlength=4).tolist(),
                          "workload_generation_s": gen_s})
            del bc
            sc.close()

        # closed-loop fleet stepping entirely on the device: one nmpc_fleet per rank,
        # 8 192 robots (256 distinct plans on map 11, tiled) x 10 receding-horizon steps after one warm-up step
        from mpc_trajectory_generator_b200.fleet import FleetPlan, NmpcFleet
        from mpc_trajectory_generator_b200.host import assembly as _asm
        t_gen = time.perf_counter()
        fhc = _asm.HostConfig.default()
        fsteps, frobots, fdistinct = 10, 8192, 256
        fplan = FleetPlan.from_scenarios(workloads.random_scenarios(fhc, 11, fdistinct, seed=5 + rank), max_steps=fsteps + 1)
        ftile = lambda a_: np.ascontiguousarray(np.concatenate([a_] * (frobots // fdistinct)))  # noqa: E731
        fbig = FleetPlan(ftile(fplan.n_ref), ftile(fplan.ref), ftile(fplan.n_vert), ftile(fplan.vert), ftile(fplan.start),
                         ftile(fplan.goal), fplan.brake_vel, fplan.brake_dist, fplan.weights, fplan.base_speed,
                         fplan.circle_radius)
        gen_s = time.perf_counter() - t_gen
        fsolver = pkg.NmpcSolver(workloads.solver_config_for(fhc), device=local_rank)
        fleet = NmpcFleet(fsolver, fbig)
        fleet.step(1)
        sync_all()
        fleet.step(fsteps)
        fms = fsolver.last_kernel_ms          # device time of the fsteps steps (events on the handle's stream)
        sync_all()
        mx, per = over_ranks(fms)
        fst = fleet.state()
        extra.append({"config": "fleet", "workload": f"{frobots} robots per GPU ({fdistinct} distinct plans on map 11, tiled), "
                      f"{fsteps} closed-loop steps in one nmpc_fleet_step call (assemble -> solve -> advance on the device)",
                      "batch_per_gpu": frobots, "steps": fsteps, "scaling": "weak", "value": world * frobots * fsteps / (mx * 1e-3),
                      "unit": "robot-steps/s", "ms_per_step": mx / fsteps, "kernel_ms_per_step_by_rank": [p_ / fsteps for p_ in per],
                      "exit_status_counts": np.bincount(fst["status"], minlength=4).tolist(),
                      "inner_iterations_mean": 0.0, "workload_generation_s": gen_s})
        fleet.close()
        fsolver.close()

    # single calls (the reference's own use: one robot, one solve per control period — BASELINE.md §3 (i)): problems of the
    # headline batch one at a time through the C ABI from host memory (cold start, helper warps next to the owner warp),
    # and the same problems on ONE host thread of the CPU restatement
    if args.extra != "off" and rank == 0:
        from oracle import oracle_c as _oc
        _oc.build()
        _oc.use_native()
        ocfg = _oc.default_config(N_hor=N, Nobs=Nobs, Ndynobs=Nd, ang_vel_max=hc.ang_vel_max, ang_acc_max=hc.ang_acc_max)
        pick = np.arange(0, B, max(1, B // 96))[:96]
        g_wall, g_kern, c_wall, its = [], [], [], []
        for i in pick:
            Pi = np.ascontiguousarray(P[i:i + 1])
            t0 = time.perf_counter()
            _, _, _, sti = solver.solve_batch(Pi)
            g_wall.append(1e3 * (time.perf_counter() - t0))
            g_kern.append(float(solver.last_kernel_ms))
            its.append(int(sti["inner_iterations"][0]))
            t0 = time.perf_counter()
            _oc.solve_batch(ocfg, Pi, nthreads=1)
            c_wall.append(1e3 * (time.perf_counter() - t0))

        def dist(v):
            v = np.asarray(v)
            return {"median": float(np.median(v)), "p90": float(np.percentile(v, 90)), "p99": float(np.percentile(v, 99)),
                    "mean": float(v.mean()), "max": float(v.max())}
        extra.append({"config": "single_call", "workload": f"{len(pick)} problems of the headline batch, one per call, cold start",
                      "unit": "ms per solve", "gpu_call_wall_ms": dist(g_wall), "gpu_kernel_ms": dist(g_kern),
                      "cpu_single_thread_ms": dist(c_wall), "inner_iterations": dist(its),
                      "api": "NmpcSolver.solve_batch -> nmpc_solve_batch (C ABI), B = 1, pageable host buffers",
                      "value": float(np.mean(g_wall)), "ms_per_step": float(np.mean(g_wall)), "batch_per_gpu": 1, "steps": len(pick),
                      "exit_status_counts": [], "inner_iterations_mean": float(np.mean(its)), "workload_generation_s": 0.0})

    if rank == 0:
        peak, peak_src = load_peaks()
        abytes = algorithmic_bytes(N, Nobs, Nd)
        ms_launch = total_ms / args.steps
        achieved = abytes * B / (ms_launch * 1e-3) / 1e9
        evals = float((stats["n_cost_evals"].astype(np.int64) + 3 * stats["n_grad_evals"].astype(np.int64)).sum())
        gflops = evals * eval_flops(N, Nobs, Nd) / (ms_launch * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "batch_per_gpu": B, "N_hor": N, "Nobs": Nobs, "Ndynobs": Nd,
                       "l2": "flushed between timed iterations (256 MiB write)", "seed": args.seed,
                       "parity": "bit-exact vs oracle/ (tests/test_gpu_parity.py); <= 1e-4 vs the serial-arithmetic "
                                 "build of the oracle on converged problems (tests/test_serial_pin.py); OpEn itself "
                                 "not runnable here"},
            "ms_per_solve": total_ms_max / args.steps / B,
            "kernel_ms_per_step_by_rank": [p / args.steps for p in per_rank_ms],
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h) * world,
                    "ms_per_step": e2e_ms_step, "api": "NmpcSolver.solve_batch_into -> nmpc_solve_batch (C ABI) on pinned host buffers; the kernel reads "
                           "the inputs from and writes the results to those buffers over PCIe in place (each byte once)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_traffic(args.workload, B), "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})",
                         "algorithmic_bytes_per_solve": abytes, "kernel_source_hash": kernel_source_hash(),
                         "note": "latency/FP64-ALU bound by construction (4128 B vs ~1e8 flop per solve); "
                                 "see fp64 block and profiles/"},
            "fp64": {"est_gflops": gflops, "peak_gflops_nominal": 148 * 64 * 2 * 1.965,
                     "frac": gflops / (148 * 64 * 2 * 1.965),
                     "evals_per_solve": float((stats["n_cost_evals"] + stats["n_grad_evals"]).mean()),
                     "inner_iterations_mean": float(stats["inner_iterations"].mean())},
            "exit_status_counts": np.bincount(status, minlength=4).tolist(),
            "clocks": clocks,
            "wall_s_timed_region": t_wall,
            "extra": {"configs": extra},
        }
        if not args.no_cpu_baseline and world >= 1:
            from oracle import oracle_c
            oracle_c.build()
            native = oracle_c.use_native()
            ocfg = oracle_c.default_config(N_hor=N, Nobs=Nobs, Ndynobs=Nd, ang_vel_max=hc.ang_vel_max,
                                           ang_acc_max=hc.ang_acc_max)
            threads = host_threads()   # torchrun pins OMP_NUM_THREADS=1; the CPU arm uses every host core it may run on
            t0 = time.perf_counter()
            oracle_c.solve_batch(ocfg, P[:64], None if U0 is None else U0[:64], None if Y0 is None else Y0[:64],
                                 nthreads=threads)
            rate = 64 / (time.perf_counter() - t0)
            sample = int(min(B, max(64, rate * args.cpu_baseline_seconds)))
            t0 = time.perf_counter()
            Uo, Yo, sto, _ = oracle_c.solve_batch(ocfg, P[:sample], None if U0 is None else U0[:sample],
                                                  None if Y0 is None else Y0[:sample], nthreads=threads)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": sample / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"first {sample} problems of rank 0's batch, one pass, OpenMP over "
                                              f"problems; OpEn-equivalent C restatement (oracle/nmpc_oracle.c, "
                                              f"{'-O3 -march=native' if native else '-O3 -march=x86-64-v3'}): the parity "
                                              f"checker doubling as the baseline",
                                    "ms_per_solve_per_core": 1e3 * dt * threads / sample}
            hU.numpy()[:] = U0h
            hY.numpy()[:] = Y0h
            solver.solve_batch_into(hP.numpy(), hU.numpy(), hY.numpy(), hstatus.numpy(), None)
            line["parity_check"] = {"sample": sample, "flags_equal": bool(np.array_equal(hstatus.numpy()[:sample], sto)),
                                    "bit_exact": bool(np.array_equal(hU.numpy()[:sample], Uo, equal_nan=True))}
        print(json.dumps(line))
    solver.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
