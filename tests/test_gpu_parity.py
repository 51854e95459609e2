"""GPU parity: the CUDA path (through the C ABI) against the C oracle on the same seeded inputs.

PARITY UNPINNED at the third-party boundary: OpEn cannot be installed here, so "the oracle"
is our FP64 CPU restatement of OpEn's algorithm (oracle/nmpc_oracle.c), itself checked
against an independent torch-autograd restatement (tests/test_oracle.py).
Bars: exit flags element-wise equal; rel-L2(U_gpu, U_oracle) <= 1e-4 (north_star tolerance);
because kernel and oracle share one arithmetic contract the comparison is in fact bit-exact
and the tests also assert that."""
import numpy as np
import pytest

import nmpc_problems as problems

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4   # north_star: "trajectories within 1e-4 rel-L2"


def _cfgs(pkg, oracle, **kw):
    g = pkg.NmpcConfig.default(**kw)
    o = oracle.default_config(**kw)
    return g, o


@pytest.mark.parametrize("N,Nobs,Nd", [(20, 10, 3), (10, 10, 3), (40, 10, 3), (80, 50, 3), (20, 0, 0), (33, 7, 1)])
def test_eval_parity(oracle, gpu_solver_factory, N, Nobs, Nd):
    import mpc_trajectory_generator_b200 as pkg
    g, o = _cfgs(pkg, oracle, N_hor=N, Nobs=Nobs, Ndynobs=Nd)
    B = 64
    P = problems.synth(N, Nobs, Nd, B, seed=N + Nobs)
    U = problems.random_controls(N, B, seed=3)
    rng = np.random.default_rng(5)
    Y = rng.normal(0, 2.0, (B, 2 * N))
    c = 5.0 ** rng.integers(0, 5, B)
    s = gpu_solver_factory(g)
    psi, grad, F1, F2 = s.eval_batch(P, U, c, Y)
    psi_o, grad_o, F1_o, F2_o = oracle.eval_batch(o, P, U, c, Y)
    assert np.allclose(psi, psi_o, rtol=1e-12, atol=0)
    assert np.allclose(grad, grad_o, rtol=1e-9, atol=1e-9)
    assert np.array_equal(F1, F1_o)
    assert np.array_equal(F2, F2_o)
    assert np.array_equal(psi, psi_o), "psi not bit-exact"
    assert np.array_equal(grad, grad_o), "grad psi not bit-exact"


@pytest.mark.parametrize("N,Nobs,Nd,B,active", [(20, 10, 3, 192, False), (20, 10, 3, 96, True), (40, 10, 3, 48, False),
                                                (10, 10, 3, 64, False), (80, 50, 3, 8, False)])
def test_solve_parity(oracle, gpu_solver_factory, N, Nobs, Nd, B, active):
    import mpc_trajectory_generator_b200 as pkg
    g, o = _cfgs(pkg, oracle, N_hor=N, Nobs=Nobs, Ndynobs=Nd)
    P = problems.synth(N, Nobs, Nd, B, seed=11 + N, active=active)
    s = gpu_solver_factory(g)
    U, Y, st, stats = s.solve_batch(P)
    Uo, Yo, sto, statso = oracle.solve_batch(o, P)
    assert np.array_equal(st, sto), "exit flags differ"
    num = np.linalg.norm(U - Uo, axis=1)
    den = np.maximum(np.linalg.norm(Uo, axis=1), 1e-12)
    assert (num / den).max() <= REL_TOL
    assert np.linalg.norm(U - Uo) / np.linalg.norm(Uo) <= REL_TOL
    assert np.array_equal(stats["inner_iterations"], statso["inner_iterations"])
    assert np.array_equal(stats["outer_iterations"], statso["outer_iterations"])
    assert np.array_equal(U, Uo), "solution not bit-exact"
    assert np.array_equal(Y, Yo), "multipliers not bit-exact"
    assert np.array_equal(stats["n_grad_evals"], statso["n_grad_evals"])
    assert np.array_equal(stats["n_cost_evals"], statso["n_cost_evals"])
    lo = np.tile([g.lin_vel_min, -g.ang_vel_max], N)
    hi = np.tile([g.lin_vel_max, g.ang_vel_max], N)
    assert np.all(U >= lo) and np.all(U <= hi), "solution must lie in U (projected half step)"


def test_warm_start_call_sequence(oracle, gpu_solver_factory):
    """nmpc_call keeps (u, y) between calls like OpEn's TCP server; the oracle is driven with
    the same carried state."""
    import mpc_trajectory_generator_b200 as pkg
    g, o = _cfgs(pkg, oracle)
    P = problems.synth(20, 10, 3, 6, seed=77, active=False)
    s = gpu_solver_factory(g)
    s.reset_warm_start()
    u_prev = np.zeros((1, 40))
    y_prev = np.zeros((1, 40))
    for k in range(6):
        u, st, stats, ms = s.call(P[k])
        Uo, Yo, sto, _ = oracle.solve_batch(o, P[k:k + 1], u_prev, y_prev)
        assert st == sto[0]
        assert np.array_equal(u, Uo[0])
        u_prev, y_prev = Uo, Yo
