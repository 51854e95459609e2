"""Multi-GPU plumbing for the batched solve: the batch of independent NMPC instances is cut into
contiguous shards, one per rank (one process per GPU); there is NO data-path collective.  The only
collective is a broadcast of the batch-invariant table (bounds, ts, tolerances, cost weights) from
rank 0 — NCCL on GPUs, gloo in the CPU tests.

Each `mng.call` of the reference is self-contained (src/mpc/mpc_generator.py:206), which is what makes
the batch dimension embarrassingly parallel; a single robot's receding-horizon sequence stays serial.
"""
import numpy as np
import torch
import torch.distributed as dist

from .solver import NmpcConfig

_FLOAT_FIELDS = ["ts", "lin_vel_min", "lin_vel_max", "ang_vel_max", "lin_acc_min", "lin_acc_max", "ang_acc_max",
                 "tolerance", "initial_tolerance", "delta_tolerance", "inner_tolerance_update",
                 "penalty_update_factor", "initial_penalty", "sufficient_decrease_coeff"]
_INT_FIELDS = ["N_hor", "Nobs", "Ndynobs", "lbfgs_memory", "max_inner_iterations", "max_outer_iterations",
               "max_duration_micros"]


def shard_bounds(B, world, rank):
    """contiguous shard [lo, hi) of a batch of B problems for `rank` of `world` (sizes differ by <= 1)."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_static_table(cfg, weights):
    """batch-invariant data as one float64 vector: config scalars + the 10 cost weights of z0[10:20]."""
    vals = [float(getattr(cfg, k)) for k in _INT_FIELDS] + [float(getattr(cfg, k)) for k in _FLOAT_FIELDS]
    return torch.tensor(vals + [float(w) for w in weights], dtype=torch.float64)


def unpack_static_table(table):
    t = table.cpu().tolist()
    kw = {k: int(round(v)) for k, v in zip(_INT_FIELDS, t[:len(_INT_FIELDS)])}
    kw.update({k: v for k, v in zip(_FLOAT_FIELDS, t[len(_INT_FIELDS):len(_INT_FIELDS) + len(_FLOAT_FIELDS)])})
    weights = t[len(_INT_FIELDS) + len(_FLOAT_FIELDS):]
    return NmpcConfig.default(**kw), weights


def broadcast_static_table(cfg, weights, device=None, src=0):
    """rank `src` sends the static table; every rank returns (cfg, weights) built from what it received."""
    table = pack_static_table(cfg, weights)
    if device is not None:
        table = table.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(table, src=src)
    return unpack_static_table(table)


def solve_sharded(solve_fn, P, U0=None, Y0=None, gather=True):
    """Solve this rank's contiguous shard of P with `solve_fn(P, U0, Y0) -> (U, Y, status)`.
    With gather=True every rank returns the full (U, Y, status) (all_gather of the results only)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    B = P.shape[0]
    lo, hi = shard_bounds(B, world, rank)
    U, Y, st = solve_fn(P[lo:hi], None if U0 is None else U0[lo:hi], None if Y0 is None else Y0[lo:hi])
    if not gather or world == 1:
        return U, Y, st, (lo, hi)
    parts = [None] * world
    dist.all_gather_object(parts, (lo, hi, np.asarray(U), np.asarray(Y), np.asarray(st)))
    n2 = np.asarray(U).shape[1]
    Uf, Yf, sf = np.zeros((B, n2)), np.zeros((B, n2)), np.zeros(B, dtype=np.int32)
    for plo, phi, pu, py, ps in parts:
        Uf[plo:phi], Yf[plo:phi], sf[plo:phi] = pu, py, ps
    return Uf, Yf, sf, (lo, hi)
