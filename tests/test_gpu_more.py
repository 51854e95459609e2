"""GPU tests added in round 2: BASELINE config 3's own setup, the device-recorded workload against the oracle-recorded
one, the time budget (NotConvergedOutOfTime), overlapping launches on several streams, the zero-copy path without a
multiplier buffer, fleet schedule bounds and the two-rank NCCL run of sharding.solve_sharded."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

import nmpc_problems as problems

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _config3_rows(oracle, robots=24, steps=14):
    """BASELINE config 3's setup at test size: robots on map 11, N=20, default weights, every receding-horizon step
    recorded with the warm start the server held (host loop + oracle)."""
    from mpc_trajectory_generator_b200 import workloads
    from mpc_trajectory_generator_b200.host import assembly
    hc = assembly.HostConfig.default()
    ocfg = oracle.default_config()
    rec = workloads.closed_loop_batch(hc, lambda P, U0, Y0: oracle.solve_batch(ocfg, P, U0, Y0)[:3], complexity=11,
                                      robots=robots, steps=steps, seed=2, sincos=oracle.sincos)
    return hc, rec


def test_config3_setup_parity(oracle, gpu_solver_factory):
    import mpc_trajectory_generator_b200 as pkg
    hc, rec = _config3_rows(oracle)
    s = gpu_solver_factory(pkg.NmpcConfig.default())
    U, Y, st, stats = s.solve_batch(rec["P"], rec["U0"], rec["Y0"])
    assert rec["P"].shape[0] >= 300 and np.abs(rec["U0"]).max() > 0      # warm-started rows
    assert np.array_equal(st, rec["status"])
    assert np.linalg.norm(U - rec["U"]) <= 1e-4 * np.linalg.norm(rec["U"])
    assert np.array_equal(U, rec["U"])


def test_device_recorded_workload_equals_oracle_recorded(oracle, gpu_solver_factory):
    """bench.py records config 3 / 4 with the fleet kernels on the GPU arm and with the host loop + oracle on the CPU
    arm: both recordings must be the same batch."""
    import mpc_trajectory_generator_b200 as pkg
    from mpc_trajectory_generator_b200 import workloads
    hc, rec = _config3_rows(oracle)
    s = gpu_solver_factory(pkg.NmpcConfig.default())
    dev = workloads.closed_loop_batch_device(s, hc, complexity=11, robots=24, steps=14, seed=2)
    for k in ("robot", "step", "status"):
        assert np.array_equal(dev[k], rec[k]), k
    for k in ("P", "U0", "Y0", "U"):
        assert np.array_equal(dev[k], rec[k]), k


def test_time_budget_ends_out_of_time(gpu_solver_factory):
    """max_duration_micros > 0 (OpEn's with_max_duration_micros, src/mpc/mpc_generator.py:9,186): solves that are still
    running when the budget is spent end NotConvergedOutOfTime with a usable (finite, feasible) reply; 0 = no limit."""
    import mpc_trajectory_generator_b200 as pkg
    P = problems.synth(20, 10, 3, 256, seed=11, active=True)
    ref = gpu_solver_factory(pkg.NmpcConfig.default())
    U0, _, st0, stats0 = ref.solve_batch(P)
    assert not np.any(st0 == 2)
    g = pkg.NmpcConfig.default(max_duration_micros=400)
    s = gpu_solver_factory(g)
    U, Y, st, stats = s.solve_batch(P)
    assert np.any(st == 2) and set(np.unique(st)) <= {0, 1, 2, 3}
    assert pkg.EXIT_STATUS_NAMES[2] == "NotConvergedOutOfTime"
    ok = st != 3
    lo = np.tile([g.lin_vel_min, -g.ang_vel_max], 20)
    hi = np.tile([g.lin_vel_max, g.ang_vel_max], 20)
    assert np.all(np.isfinite(U[ok])) and np.all(U[ok] >= lo) and np.all(U[ok] <= hi)
    assert np.all(stats["inner_iterations"][st == 2] <= stats0["inner_iterations"][st == 2])
    done = st == 0                       # what converged inside the budget is the unlimited solve's reply
    assert np.array_equal(U[done], U0[done])


def test_overlapping_launches_on_several_streams(oracle, gpu_solver_factory):
    """nmpc_solve_batch_device is asynchronous on the caller's stream: launches in flight at the same time (here six on
    three streams, more than the handle's ring of queue contexts) must not share a work-queue counter or the probe
    scratch."""
    import torch
    import mpc_trajectory_generator_b200 as pkg
    s = gpu_solver_factory(pkg.NmpcConfig.default())
    dev = torch.device("cuda", 0)
    streams = [torch.cuda.Stream(dev) for _ in range(3)]
    jobs = []
    for j in range(6):
        B = 2300 if j % 2 == 0 else 700          # > warp slots (probe + order) and < warp slots
        P = problems.synth(20, 10, 3, B, seed=300 + j, active=False)
        dP = torch.from_numpy(P).to(dev)
        dU = torch.zeros((B, 40), dtype=torch.float64, device=dev)
        dY = torch.zeros((B, 40), dtype=torch.float64, device=dev)
        dst = torch.zeros(B, dtype=torch.int32, device=dev)
        jobs.append((P, dP, dU, dY, dst))
    torch.cuda.synchronize(dev)
    for j, (P, dP, dU, dY, dst) in enumerate(jobs):
        st = streams[j % 3]
        s.solve_batch_device(P.shape[0], dP.data_ptr(), dU.data_ptr(), dY.data_ptr(), dst.data_ptr(), 0, st.cuda_stream)
    torch.cuda.synchronize(dev)
    ocfg = oracle.default_config()
    for P, dP, dU, dY, dst in jobs:
        Uo, Yo, sto, _ = oracle.solve_batch(ocfg, P)
        assert np.array_equal(dst.cpu().numpy(), sto)
        assert np.array_equal(dU.cpu().numpy(), Uo) and np.array_equal(dY.cpu().numpy(), Yo)


def test_zero_copy_without_multiplier_buffer(gpu_solver_factory):
    """nmpc_solve_batch on page-locked P / U / status with Y == NULL takes the in-place path (multipliers start at zero
    and their final state is dropped)."""
    import torch
    import mpc_trajectory_generator_b200 as pkg
    s = gpu_solver_factory(pkg.NmpcConfig.default())
    B = 200
    P = problems.synth(20, 10, 3, B, seed=5, active=True)
    U, Y, st, _ = s.solve_batch(P)
    hP = torch.from_numpy(P).pin_memory()
    hU = torch.zeros((B, 40), dtype=torch.float64).pin_memory()
    hst = torch.zeros(B, dtype=torch.int32).pin_memory()
    dp = C.POINTER(C.c_double)
    n0 = s.launch_count
    rc = s._lib.nmpc_solve_batch(s._h, B, hP.numpy().ctypes.data_as(dp), hU.numpy().ctypes.data_as(dp), None,
                                 hst.numpy().ctypes.data_as(C.POINTER(C.c_int32)), None)
    assert rc == 0 and s.launch_count == n0 + 1
    assert np.array_equal(hU.numpy(), U, equal_nan=True) and np.array_equal(hst.numpy(), st)


def test_fleet_refuses_to_run_past_its_obstacle_schedule(gpu_solver_factory):
    import mpc_trajectory_generator_b200 as pkg
    from mpc_trajectory_generator_b200 import NmpcError
    from mpc_trajectory_generator_b200.fleet import FleetPlan, NmpcFleet
    from mpc_trajectory_generator_b200.host import assembly
    hc = assembly.HostConfig.default()
    sc = assembly.Scenario(hc, assembly.load_maps()[12])
    with pytest.raises(NmpcError):
        FleetPlan.from_scenarios([sc])                      # moving obstacles: max_steps is required
    plan = FleetPlan.from_scenarios([sc], max_steps=4)
    s = gpu_solver_factory(pkg.NmpcConfig.default())
    fleet = NmpcFleet(s, plan)
    fleet.step(4)
    with pytest.raises(NmpcError):
        fleet.step(10)                                      # would read schedule rows that were never uploaded
    assert int(fleet.state()["t"][0]) == 4                  # nothing ran
    fleet.close()


def _nccl_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import mpc_trajectory_generator_b200 as pkg
    from mpc_trajectory_generator_b200 import sharding
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg0 = pkg.NmpcConfig.default()
    mine = cfg0 if rank == 0 else pkg.NmpcConfig.default(max_inner_iterations=7, ts=0.1)   # must be overwritten
    weights = problems.DEFAULT_WEIGHTS if rank == 0 else [0.0] * 10
    cfg, w = sharding.broadcast_static_table(mine, weights, device=torch.device("cuda", rank))
    assert cfg.as_dict() == cfg0.as_dict() and w == problems.DEFAULT_WEIGHTS
    P = problems.synth(20, 10, 3, 501, seed=9, active=False, weights=w)
    s = pkg.NmpcSolver(cfg, device=rank)

    def solve(Ps, U0, Y0):
        U, Y, st, _ = s.solve_batch(Ps, U0, Y0)
        return U, Y, st

    U, Y, st, (lo, hi) = sharding.solve_sharded(solve, P)
    assert (lo, hi) == sharding.shard_bounds(501, world, rank)
    if rank == 0:
        np.savez(out, U=U, Y=Y, st=st)
    s.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_nccl_matches_one_gpu(tmp_path, gpu_solver_factory):
    """sharding.solve_sharded with one process per GPU over NCCL (static table broadcast from rank 0, results
    all-gathered) equals the single-GPU solve of the whole batch."""
    import torch
    import torch.multiprocessing as mp
    import mpc_trajectory_generator_b200 as pkg
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = str(tmp_path / "nccl.npz")
    mp.spawn(_nccl_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    P = problems.synth(20, 10, 3, 501, seed=9, active=False, weights=problems.DEFAULT_WEIGHTS)
    s = gpu_solver_factory(pkg.NmpcConfig.default())
    U, Y, st, _ = s.solve_batch(P)
    assert np.array_equal(got["U"], U) and np.array_equal(got["Y"], Y) and np.array_equal(got["st"], st)


@pytest.mark.parametrize("N,Nobs,B", [(20, 10, 1), (20, 10, 5), (20, 10, 150), (20, 10, 1900), (20, 10, 4000), (40, 10, 40), (10, 10, 300)])
def test_helper_warps_do_not_change_a_bit(monkeypatch, oracle, gpu_solver_factory, N, Nobs, B):
    """Warps that have run out of problems evaluate the next batch of line-search trials for a warp of their CTA that is
    still solving (csrc/nmpc_device.cuh: PH_HELP, take_from_helper).  Who evaluates a trial must not matter: the same batch
    with the helper instantiation of the kernel, without it, and on the oracle — solutions, multipliers, flags and the
    iteration / evaluation counters all equal, from one problem (eleven idle warps next to the owner) to two waves."""
    import mpc_trajectory_generator_b200 as pkg
    cfg = pkg.NmpcConfig.default(N_hor=N, Nobs=Nobs)
    P = problems.synth(N, Nobs, 3, min(B, 320), seed=40 + B)
    P = np.ascontiguousarray(np.tile(P, (-(-B // len(P)), 1))[:B])
    out = {}
    for waves in ("0", "1000"):
        monkeypatch.setenv("NMPC_B200_HELP_MAX_WAVES", waves)
        s = gpu_solver_factory(cfg)
        out[waves] = s.solve_batch(P)
    for a, b in zip(out["0"][:3], out["1000"][:3]):
        assert np.array_equal(a, b, equal_nan=True)
    for k in ("inner_iterations", "outer_iterations", "n_grad_evals", "n_cost_evals"):
        assert np.array_equal(out["0"][3][k], out["1000"][3][k]), k
    n = min(B, 96)
    ocfg = oracle.default_config(N_hor=N, Nobs=Nobs)
    U, Y, st, stats = oracle.solve_batch(ocfg, P[:n])
    assert np.array_equal(st, out["1000"][2][:n]) and np.array_equal(U, out["1000"][0][:n], equal_nan=True)
