"""CPU tests of the oracle itself (no GPU).  PARITY UNPINNED at the OpEn boundary: the
reference ships no golden vectors for the solve, so the C restatement (oracle/nmpc_oracle.c)
is pinned by (a) an independent torch-autograd restatement of the reference's CasADi graph
(oracle/oracle_np.py), (b) an independent plain-Python PANOC/ALM, (c) solver invariants, and
(d) the run recorded from the unmodified reference orchestrator (tests/golden/)."""
import math
import os
import warnings

import numpy as np
import pytest

import nmpc_problems as problems
from oracle import oracle_np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
warnings.filterwarnings("ignore")


def test_sincos_accuracy(oracle):
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-50, 50, 4000), rng.uniform(-1e-3, 1e-3, 200), [0.0, math.pi / 2, -math.pi, 1e5]])
    worst = 0.0
    for x in xs:
        s, c = oracle.sincos(x)
        worst = max(worst, abs(s - math.sin(x)) / np.spacing(max(abs(math.sin(x)), 1e-300)),
                    abs(c - math.cos(x)) / np.spacing(max(abs(math.cos(x)), 1e-300))) if abs(math.sin(x)) > 1e-3 and abs(math.cos(x)) > 1e-3 else worst
        assert abs(s - math.sin(x)) < 3e-16 and abs(c - math.cos(x)) < 3e-16
    assert worst <= 2.0, f"sincos off by {worst} ulp"
    s, c = oracle.sincos(float("nan"))
    assert math.isnan(s) and math.isnan(c)


@pytest.mark.parametrize("N,Nobs,Nd", [(20, 10, 3), (8, 2, 1), (40, 10, 3), (33, 0, 2), (70, 25, 3)])
def test_cost_and_gradient_vs_autograd(oracle, N, Nobs, Nd):
    """psi, grad psi, F1, F2 of the C oracle against the torch-float64 restatement of
    src/mpc/mpc_generator.py:70-171 (gradient by autograd)."""
    cfg = oracle.default_config(N_hor=N, Nobs=Nobs, Ndynobs=Nd)
    B = 6
    P = problems.synth(N, Nobs, Nd, B, seed=3 * N + Nobs)
    U = problems.random_controls(N, B, seed=1)
    rng = np.random.default_rng(2)
    Y = rng.normal(0, 2.0, (B, 2 * N))
    c = np.array([1.0, 5.0, 25.0, 125.0, 1.0, 625.0])
    psi, grad, F1, F2 = oracle.eval_batch(cfg, P, U, c, Y)
    cd = oracle_np.cfg_dict(cfg)
    for b in range(B):
        v, g, f1, f2 = oracle_np.eval_psi(U[b], P[b], c[b], Y[b], cd)
        assert abs(v - psi[b]) <= 1e-11 * abs(v)
        assert np.abs(g - grad[b]).max() <= 1e-10 * max(1.0, np.abs(g).max())
        assert np.abs(f1 - F1[b]).max() <= 1e-12
        assert np.abs(f2 - F2[b]).max() <= 1e-11
    assert F2.max() > 0, "the test must exercise active obstacle penalties"


def test_gradient_finite_difference(oracle):
    cfg = oracle.default_config()
    P = problems.synth(20, 10, 3, 2, seed=9, weights=problems.MIXED_WEIGHTS)
    U = problems.random_controls(20, 2, seed=4)
    Y = np.random.default_rng(1).normal(0, 1, (2, 40))
    c = np.array([5.0, 5.0])
    psi, grad, _, _ = oracle.eval_batch(cfg, P, U, c, Y)
    h = 1e-6
    for b in range(2):
        for i in range(0, 40, 7):
            Up, Um = U.copy(), U.copy()
            Up[b, i] += h; Um[b, i] -= h
            fd = (oracle.eval_batch(cfg, P, Up, c, Y)[0][b] - oracle.eval_batch(cfg, P, Um, c, Y)[0][b]) / (2 * h)
            assert abs(fd - grad[b, i]) <= 1e-4 * max(1.0, abs(grad[b, i]))


def test_solver_invariants(oracle):
    cfg = oracle.default_config()
    N = 20
    P = problems.synth(N, 10, 3, 96, seed=5, active=False)
    U, Y, st, stats = oracle.solve_batch(cfg, P)
    lo = np.tile([cfg.lin_vel_min, -cfg.ang_vel_max], N)
    hi = np.tile([cfg.lin_vel_max, cfg.ang_vel_max], N)
    assert np.all(U >= lo) and np.all(U <= hi)                 # the reply is the projected half step
    assert set(np.unique(st)) <= {0, 1}
    assert np.all(stats["outer_iterations"] >= 2)              # ALM criterion 1 needs a second outer iteration
    assert np.all(stats["outer_iterations"] <= cfg.max_outer_iterations)
    conv = st == 0
    assert conv.any()
    assert np.all(stats["delta_y_norm_over_c"][conv] <= cfg.delta_tolerance * 1.0000001)
    assert np.all(stats["f2_norm"][conv] <= cfg.delta_tolerance * 1.0000001)
    assert np.all(stats["last_norm_fpr"][conv] < cfg.tolerance)
    # acceleration constraints hold at converged points (F1 in C up to delta)
    _, _, F1, _ = oracle.eval_batch(cfg, P, U, 1.0)
    assert np.all(F1[conv, :N] <= cfg.lin_acc_max + 2e-4) and np.all(F1[conv, :N] >= cfg.lin_acc_min - 2e-4)
    # determinism and independence of batch composition / thread count
    U2, Y2, st2, _ = oracle.solve_batch(cfg, P[::-1].copy(), nthreads=3)
    assert np.array_equal(U2[::-1], U) and np.array_equal(st2[::-1], st)


def test_iteration_caps_and_flags(oracle):
    P = problems.synth(20, 10, 3, 8, seed=6, active=True)
    cfg = oracle.default_config(max_inner_iterations=5, max_outer_iterations=3)
    U, Y, st, stats = oracle.solve_batch(cfg, P)
    assert np.all(st == 1)                                       # NotConvergedIterations
    assert np.all(stats["outer_iterations"] == 3)
    assert np.all(stats["inner_iterations"] <= 3 * 5)
    bad = P.copy()
    bad[0, 0] = np.nan
    cfg = oracle.default_config()
    _, _, st, _ = oracle.solve_batch(cfg, bad[:1])
    assert st[0] == 3                                            # NotFiniteComputation -> is_ok() False in the reference


def test_c_solver_vs_python_solver(oracle):
    """Two independent implementations of the same OpEn control flow (C with the warp-ordered
    arithmetic contract; plain Python with libm, divisions, serial sums, autograd gradient) agree:
    same flags and outer iterations, iterates within 1e-3 relative after the same (truncated)
    iteration budget — rounding differences grow along a non-converged PANOC path, so this is a
    control-flow check, not a bit check."""
    cfg = oracle.default_config(max_inner_iterations=40, max_outer_iterations=3)
    P = problems.synth(20, 10, 3, 2, seed=21, active=False, weights=problems.SMOOTH_WEIGHTS)
    U, Y, st, stats = oracle.solve_batch(cfg, P)
    cd = oracle_np.cfg_dict(cfg)
    for b in range(2):
        r = oracle_np.panoc_alm_solve(P[b], cd)
        assert r["status"] == st[b]
        assert r["outer"] == stats["outer_iterations"][b]
        assert abs(r["inner"] - stats["inner_iterations"][b]) <= 2
        assert np.linalg.norm(r["u"] - U[b]) <= 1e-3 * max(1.0, np.linalg.norm(U[b]))


def test_recorded_reference_run_regression(oracle):
    """tests/golden/config1_run.npz was produced by the UNMODIFIED reference PathGenerator.run
    (configs/default.yaml, map complexity=1) calling an oracle-backed manager; re-solving the
    recorded parameter sequence with the server's warm start must reproduce every reply."""
    g = np.load(os.path.join(GOLD, "config1_run.npz"))
    cfg = oracle.default_config()
    u = np.zeros((1, 40)); y = np.zeros((1, 40))
    K = g["P"].shape[0]
    for k in range(0, K):
        U, Y, st, stats = oracle.solve_batch(cfg, g["P"][k:k + 1], u, y, nthreads=1)
        assert st[0] == g["status"][k] and stats["inner_iterations"][0] == g["inner"][k]
        assert np.array_equal(U[0], g["U"][k])
        u, y = U, Y
    # the recorded closed loop reached the goal of src/visibility/graphs.py:43 within the reference's tolerance
    assert abs(g["xx"][-1] - 19.0) < 0.05 and abs(g["xy"][-1] - 10.0) < 0.05


def test_lbfgs_compact_form_agrees_with_two_loop():
    """The L-BFGS direction can be evaluated in compact (Gram-matrix) form instead of the literal two-loop
    recursion (an experiment for the GPU, experiments/): same flags, converged solutions equal to solver
    tolerance.  Run in subprocesses because the switch is an environment variable read by the oracle."""
    import subprocess
    import sys
    code = ("import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r); import nmpc_problems as p; "
            "from oracle import oracle_c as oc; cfg = oc.default_config(); "
            "P = p.synth(20, 10, 3, 24, seed=5, active=False); U, Y, st, s = oc.solve_batch(cfg, P); "
            "np.save(sys.argv[1], U); np.save(sys.argv[2], st)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = code % (root, os.path.join(root, "tests"))
    out = {}
    for name, env in (("two_loop", {}), ("compact", {"NMPC_ORACLE_COMPACT": "1"})):
        fu, fs = f"/tmp/_lb_{name}_u.npy", f"/tmp/_lb_{name}_s.npy"
        subprocess.check_call([sys.executable, "-c", code, fu, fs], env=dict(os.environ, **env))
        out[name] = (np.load(fu), np.load(fs))
    (Ua, sa), (Ub, sb) = out["two_loop"], out["compact"]
    assert np.array_equal(sa, sb)
    conv = sa == 0
    assert conv.any()
    rel = np.linalg.norm(Ua[conv] - Ub[conv], axis=1) / np.linalg.norm(Ua[conv], axis=1)
    assert rel.max() < 1e-4
