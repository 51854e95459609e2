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


def same(a, b):
    """bit-for-bit equality; a solve that ends NotFiniteComputation leaves NaNs in its reply on both sides"""
    return np.array_equal(a, b, equal_nan=True)


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


def test_nan_input_flags_not_finite(oracle, gpu_solver_factory):
    """a NaN in the parameters ends as NotFiniteComputation on both sides (reference: is_ok() False)."""
    import mpc_trajectory_generator_b200 as pkg
    g, o = _cfgs(pkg, oracle)
    P = problems.synth(20, 10, 3, 4, seed=8, active=False)
    P[1, 0] = np.nan
    s = gpu_solver_factory(g)
    U, Y, st, stats = s.solve_batch(P)
    Uo, Yo, sto, _ = oracle.solve_batch(o, P)
    assert st[1] == 3 and np.array_equal(st, sto)
    ok = st != 3
    assert np.array_equal(U[ok], Uo[ok])


def test_empty_and_single(oracle, gpu_solver_factory):
    import mpc_trajectory_generator_b200 as pkg
    g, o = _cfgs(pkg, oracle)
    s = gpu_solver_factory(g)
    U, Y, st, stats = s.solve_batch(np.zeros((0, 430)))
    assert U.shape == (0, 40) and st.shape == (0,)
    P = problems.synth(20, 10, 3, 1, seed=2, active=False)
    U, Y, st, stats = s.solve_batch(P)
    Uo, _, sto, _ = oracle.solve_batch(o, P)
    assert np.array_equal(U, Uo) and np.array_equal(st, sto)


def test_golden_reference_run_replay(gpu_solver_factory):
    """BASELINE config 1: the parameter sequence the UNMODIFIED reference PathGenerator.run assembled while it was
    driven by our oracle behind the manager (tests/golden/config1_run.npz: the parameters are the reference's, the
    replies are the oracle's) replayed through nmpc_call, whose handle keeps (u, y) between calls like OpEn's TCP
    server; every reply must equal the recorded one."""
    import os
    import mpc_trajectory_generator_b200 as pkg
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config1_run.npz"))
    s = gpu_solver_factory(pkg.NmpcConfig.default())
    s.reset_warm_start()
    for k in range(g["P"].shape[0]):
        u, st, stats, ms = s.call(g["P"][k])
        assert st == g["status"][k] and stats["inner_iterations"] == g["inner"][k]
        assert np.linalg.norm(u - g["U"][k]) <= REL_TOL * max(np.linalg.norm(g["U"][k]), 1e-12)
        assert np.array_equal(u, g["U"][k])


def test_manager_closed_loop_default_config():
    """The OptimizerTcpManager-shaped object drives a full receding-horizon run (our host mirror of
    src/path_generator.py:290-403) on map complexity=1 / configs/default.yaml and reaches the goal
    exactly like the run recorded from the reference orchestrator with the oracle behind it."""
    import os
    from mpc_trajectory_generator_b200.host import assembly, opengen_compat
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config1_run.npz"))
    hc = assembly.HostConfig.default()
    opengen_compat.configure(reference_config=hc)
    mng = opengen_compat.OptimizerTcpManager("mpc_build/navigation")
    mng.start()
    mng.ping()
    sc = assembly.Scenario(hc, assembly.load_maps()[1])
    done, k = False, 0
    while not done and k < 2500:
        resp = mng.call(list(sc.parameters()))
        assert resp.is_ok()
        sol = resp.get()
        assert sol.exit_status in ("Converged", "NotConvergedIterations") and sol.solve_time_ms > 0
        done = sc.apply(sol.solution)
        k += 1
    mng.kill()
    mng.kill()  # idempotent
    assert done and k == g["P"].shape[0]
    assert np.array_equal(np.array(sc.states[0::3]), g["xx"]) and np.array_equal(np.array(sc.states[1::3]), g["xy"])
    bad = opengen_compat.OptimizerTcpManager()
    bad.start()
    r = bad.call([0.0] * 7)
    assert not r.is_ok() and r.get().code == 3003
    bad.kill()


def test_full_size_properties(gpu_solver_factory):
    """BASELINE.json full size (B=4096, N=20): size-independent properties — every reply lies in U,
    flags are in the enum, the launch is deterministic, and a permutation of the batch permutes the
    replies (no cross-problem coupling through the work queue or shared memory)."""
    import mpc_trajectory_generator_b200 as pkg
    g = pkg.NmpcConfig.default()
    B = 4096
    P = problems.synth(20, 10, 3, B, seed=123, active=False)
    s = gpu_solver_factory(g)
    U, Y, st, stats = s.solve_batch(P)
    lo = np.tile([g.lin_vel_min, -g.ang_vel_max], 20)
    hi = np.tile([g.lin_vel_max, g.ang_vel_max], 20)
    assert np.all(U >= lo) and np.all(U <= hi)
    assert set(np.unique(st)) <= {0, 1}
    assert np.all(stats["exit_status"] == st)
    conv = st == 0
    assert conv.mean() > 0.3
    assert np.all(stats["last_norm_fpr"][conv] < g.tolerance)
    assert np.all(stats["f2_norm"][conv] <= g.delta_tolerance * 1.0000001)
    U2, Y2, st2, _ = s.solve_batch(P)
    assert np.array_equal(U, U2) and np.array_equal(st, st2)
    perm = np.random.default_rng(0).permutation(B)
    U3, Y3, st3, _ = s.solve_batch(P[perm])
    assert np.array_equal(U3, U[perm]) and np.array_equal(st3, st[perm]) and np.array_equal(Y3, Y[perm])


@pytest.mark.parametrize("N,Nobs", [(10, 50), (20, 200), (40, 100), (80, 10)])
def test_sweep_workload_parity(oracle, gpu_solver_factory, N, Nobs):
    """BASELINE config 5 grid points (horizon x static-obstacle slots), built by workloads.sweep_batch."""
    import mpc_trajectory_generator_b200 as pkg
    from mpc_trajectory_generator_b200 import workloads
    P, hc = workloads.sweep_batch(N, Nobs, B=10, seed=2)
    g, o = _cfgs(pkg, oracle, N_hor=N, Nobs=Nobs, Ndynobs=3)
    s = gpu_solver_factory(g)
    U, Y, st, stats = s.solve_batch(P)
    Uo, Yo, sto, statso = oracle.solve_batch(o, P)
    assert np.array_equal(st, sto)
    assert np.linalg.norm(U - Uo) <= REL_TOL * np.linalg.norm(Uo)
    assert np.array_equal(U, Uo) and np.array_equal(Y, Yo)
    assert np.array_equal(stats["inner_iterations"], statso["inner_iterations"])


def test_config2_workload_full_batch_parity(oracle, gpu_solver_factory):
    """BASELINE config 2 at its full size (B=4096 first-step problems on map 3): every reply against the oracle."""
    import mpc_trajectory_generator_b200 as pkg
    from mpc_trajectory_generator_b200 import workloads
    from mpc_trajectory_generator_b200.host import assembly
    hc = assembly.HostConfig.default()
    P, _ = workloads.first_step_batch(hc, complexity=3, B=4096, seed=0)
    g, o = _cfgs(pkg, oracle)
    s = gpu_solver_factory(g)
    U, Y, st, stats = s.solve_batch(P)
    Uo, Yo, sto, statso = oracle.solve_batch(o, P)
    assert np.array_equal(st, sto)
    num = np.linalg.norm(U - Uo, axis=1)
    assert (num / np.maximum(np.linalg.norm(Uo, axis=1), 1e-12)).max() <= REL_TOL
    assert np.array_equal(U, Uo) and np.array_equal(Y, Yo)
    assert np.array_equal(stats["n_grad_evals"], statso["n_grad_evals"])


@pytest.mark.parametrize("N,Nobs,B", [(20, 10, 1), (20, 10, 2), (20, 10, 3), (20, 10, 5), (20, 10, 13), (20, 10, 37),
                                      (20, 10, 900), (40, 10, 3), (40, 10, 11), (80, 50, 2), (80, 200, 4)])
def test_small_batches_with_idle_warps(oracle, gpu_solver_factory, N, Nobs, B):
    """Batches smaller than the machine (first wave only, CTAs with idle warps) with obstacles in the way (the
    penalty slow paths, Lipschitz halvings, exhausted line searches): replies, multipliers, iteration and evaluation
    counts against the oracle, twice in a row."""
    import mpc_trajectory_generator_b200 as pkg
    g, o = _cfgs(pkg, oracle, N_hor=N, Nobs=Nobs, Ndynobs=3)
    P = problems.synth(N, Nobs, 3, B, seed=1000 + B + N, active=True)
    s = gpu_solver_factory(g)
    Uo, Yo, sto, statso = oracle.solve_batch(o, P)
    for _ in range(2):
        U, Y, st, stats = s.solve_batch(P)
        assert np.array_equal(st, sto)
        assert same(U, Uo) and same(Y, Yo)
        for k in ("inner_iterations", "outer_iterations", "n_grad_evals", "n_cost_evals"):
            assert np.array_equal(stats[k], statso[k]), k


def test_pinned_host_buffers_in_place(gpu_solver_factory):
    """nmpc_solve_batch on page-locked buffers (the kernel reads / writes them in place) == the staged-copy path."""
    import torch
    import mpc_trajectory_generator_b200 as pkg
    g = pkg.NmpcConfig.default()
    B = 300
    P = problems.synth(20, 10, 3, B, seed=77, active=True)
    s = gpu_solver_factory(g)
    U, Y, st, stats = s.solve_batch(P)                      # pageable numpy arrays: staged copies
    hP = torch.from_numpy(P).pin_memory()
    hU = torch.zeros((B, 40), dtype=torch.float64).pin_memory()
    hY = torch.zeros((B, 40), dtype=torch.float64).pin_memory()
    hst = torch.zeros(B, dtype=torch.int32).pin_memory()
    s.solve_batch_into(hP.numpy(), hU.numpy(), hY.numpy(), hst.numpy(), None)
    assert np.array_equal(hU.numpy(), U) and np.array_equal(hY.numpy(), Y) and np.array_equal(hst.numpy(), st)
    assert s.last_kernel_ms > 0


def test_device_pointer_entry(oracle, gpu_solver_factory):
    """nmpc_solve_batch_device on torch tensors / torch's current stream (the path bench.py times)."""
    import torch
    import mpc_trajectory_generator_b200 as pkg
    g, o = _cfgs(pkg, oracle)
    B = 64
    P = problems.synth(20, 10, 3, B, seed=31, active=False)
    s = gpu_solver_factory(g)
    dev = torch.device("cuda", 0)
    dP = torch.from_numpy(P).to(dev)
    dU = torch.zeros((B, 40), dtype=torch.float64, device=dev)
    dY = torch.zeros((B, 40), dtype=torch.float64, device=dev)
    dst = torch.zeros(B, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev)
    n0 = s.launch_count
    s.solve_batch_device(B, dP.data_ptr(), dU.data_ptr(), dY.data_ptr(), dst.data_ptr(), 0, stream.cuda_stream)
    torch.cuda.synchronize(dev)
    assert s.launch_count == n0 + 1
    Uo, Yo, sto, _ = oracle.solve_batch(o, P)
    assert np.array_equal(dU.cpu().numpy(), Uo) and np.array_equal(dst.cpu().numpy(), sto)
