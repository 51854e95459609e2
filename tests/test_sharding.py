"""Multi-process path on CPU: world_size-2 gloo.  The batch shards contiguously, the static table is
broadcast from rank 0, and the gathered result equals the single-process solve.  The solve function here
is the CPU oracle (tests may use it); on the GPU box bench.py runs the same plumbing over NCCL."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

import nmpc_problems as problems
from mpc_trajectory_generator_b200 import sharding
from mpc_trajectory_generator_b200.solver import NmpcConfig


def test_shard_bounds_cover_batch():
    for B in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_static_table_roundtrip():
    cfg = NmpcConfig.default(N_hor=40, ang_vel_max=1.0, ang_acc_max=5.0)
    w = [1.0, 10.0, 0.0, 0.0, 0.0, 5.0, 0.2, 20.0, 8.0, 20.0]
    cfg2, w2 = sharding.unpack_static_table(sharding.pack_static_table(cfg, w))
    assert cfg2.as_dict() == cfg.as_dict() and w2 == w


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_c
    # rank 0 owns the true static table; the other rank starts from a wrong one and must receive it
    cfg0 = NmpcConfig.default(max_inner_iterations=30, max_outer_iterations=2)
    mine = cfg0 if rank == 0 else NmpcConfig.default(max_inner_iterations=7, ts=0.1)
    weights = problems.DEFAULT_WEIGHTS if rank == 0 else [0.0] * 10
    cfg, w = sharding.broadcast_static_table(mine, weights)
    assert cfg.as_dict() == cfg0.as_dict() and w == problems.DEFAULT_WEIGHTS
    P = problems.synth(20, 10, 3, 13, seed=4, active=False, weights=w)
    ocfg = oracle_c.default_config(max_inner_iterations=cfg.max_inner_iterations,
                                   max_outer_iterations=cfg.max_outer_iterations)

    def solve(Ps, U0, Y0):
        U, Y, st, _ = oracle_c.solve_batch(ocfg, Ps, U0, Y0, nthreads=1)
        return U, Y, st

    U, Y, st, (lo, hi) = sharding.solve_sharded(solve, P)
    assert (lo, hi) == sharding.shard_bounds(13, world, rank)
    if rank == 0:
        np.savez(out, U=U, Y=Y, st=st)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(tmp_path, oracle):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "sharded.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    ocfg = oracle.default_config(max_inner_iterations=30, max_outer_iterations=2)
    P = problems.synth(20, 10, 3, 13, seed=4, active=False, weights=problems.DEFAULT_WEIGHTS)
    U, Y, st, _ = oracle.solve_batch(ocfg, P)
    assert np.array_equal(got["U"], U) and np.array_equal(got["Y"], Y) and np.array_equal(got["st"], st)
