"""The reference's own benchmark and demo runs (maps 1-11 with configs/default.yaml, src/gen_runtime_plots.py:21-33;
map 12 with sinus_object=True, src/main.py:11-23) recorded from the UNMODIFIED PathGenerator.run with the manager
backed by the oracle (tests/golden/make_fixtures.py -> tests/golden/ref_runs.npz).  What these pin:

  * the reference-assembled parameter vectors of all twelve maps (moving ellipses on maps 2 and 12, up to 10 corner
    circles, the braking zone of long paths) reach the CPU oracle and the CUDA solver: every recorded solver call is a
    self-contained tuple (p, warm start u / y, reply), replayed as one batch;
  * the package's host mirror of the reference's parameter assembly (host/assembly.py, the thing the device fleet
    kernels are tested against) rebuilds every recorded parameter vector bit for bit when it walks the recorded runs.

(The solver replies in the recording are the oracle's: OpEn itself cannot run here — see oracle/nmpc_oracle.c.)"""
import os

import numpy as np
import pytest

from mpc_trajectory_generator_b200.host import assembly

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def runs():
    return np.load(os.path.join(GOLD, "ref_runs.npz"))


def test_recording_covers_the_reference_scripts(runs):
    maps = sorted(set(runs["map"].tolist()))
    assert maps == list(range(1, 13))
    summ = {int(r[0]): r for r in runs["summary"]}        # map, steps, kept, max circles, corner vertices, moving
    assert summ[2][5] == 1 and summ[12][5] == 1           # real moving ellipses on maps 2 and 12
    assert max(r[3] for r in summ.values()) == 10         # a step with all ten circle slots in use
    N, Nobs, Nd = 20, 10, 3
    ell = runs["P"][:, 20 + N + 3 * Nobs:20 + N + 3 * Nobs + 5 * Nd * N].reshape(-1, Nd, N, 5)
    m12 = runs["map"] == 12
    assert np.ptp(ell[m12][:, 0, :, 0]) > 1.0             # the sinus object moves through the horizon
    assert (runs["status"] == 1).sum() > 100 and (runs["status"] == 0).sum() > 100


def test_oracle_reproduces_the_recording(oracle, runs):
    cfg = oracle.default_config()
    U, Y, st, stats = oracle.solve_batch(cfg, runs["P"], runs["U0"], runs["Y0"])
    assert np.array_equal(st, runs["status"])
    assert np.array_equal(U, runs["U"]) and np.array_equal(Y, runs["Y"])
    assert np.array_equal(stats["inner_iterations"], runs["inner"])
    assert np.array_equal(stats["n_grad_evals"], runs["n_grad"]) and np.array_equal(stats["n_cost_evals"], runs["n_cost"])


@pytest.mark.parametrize("cx", list(range(1, 13)))
def test_host_mirror_rebuilds_every_recorded_parameter_vector(runs, cx):
    """Walk the recorded run of map cx with the controls the reference applied: at every kept step the mirror's
    parameter vector equals the one the unmodified reference assembled (src/path_generator.py:293-382)."""
    cfg = assembly.HostConfig.default()
    sc = assembly.Scenario(cfg, assembly.load_maps()[cx], sinus_object=(cx == 12))
    ctrl = runs["ctrl"][runs["ctrl_map"] == cx]
    sel = np.nonzero(runs["map"] == cx)[0]
    kept = {int(runs["step"][i]): i for i in sel}
    checked = 0
    for k in range(len(ctrl)):
        p = sc.parameters()        # stateful like the reference's loop (reference index, obstacle ring): every step
        if k in kept:
            assert np.array_equal(p, runs["P"][kept[k]]), f"map {cx} step {k}"
            checked += 1
        sc.apply(np.r_[ctrl[k], np.zeros(2 * cfg.N_hor - 2)])
    assert checked == len(sel)


@pytest.mark.gpu
def test_gpu_replays_the_recording_as_one_batch(gpu_solver_factory, runs):
    import mpc_trajectory_generator_b200 as pkg
    s = gpu_solver_factory(pkg.NmpcConfig.default())
    U, Y, st, stats = s.solve_batch(runs["P"], runs["U0"], runs["Y0"])
    assert np.array_equal(st, runs["status"])
    rel = np.linalg.norm(U - runs["U"], axis=1) / np.maximum(np.linalg.norm(runs["U"], axis=1), 1e-12)
    assert rel.max() <= 1e-4
    assert np.array_equal(U, runs["U"]) and np.array_equal(Y, runs["Y"])
    assert np.array_equal(stats["inner_iterations"], runs["inner"])
    assert np.array_equal(stats["n_grad_evals"], runs["n_grad"]) and np.array_equal(stats["n_cost_evals"], runs["n_cost"])


@pytest.mark.gpu
def test_gpu_call_sequence_on_the_moving_obstacle_map(gpu_solver_factory, runs):
    """nmpc_call (the mng.call replacement: the handle keeps u and y between calls) along the first recorded steps of
    map 12 — the steps are consecutive only at the start of the recording's stride, so each call is primed with the
    recorded warm start through a batch solve of the previous tuple."""
    import mpc_trajectory_generator_b200 as pkg
    s = gpu_solver_factory(pkg.NmpcConfig.default())
    sel = np.nonzero(runs["map"] == 12)[0][:12]
    for i in sel:
        U, Y, st, _ = s.solve_batch(runs["P"][i:i + 1], runs["U0"][i:i + 1], runs["Y0"][i:i + 1])
        assert st[0] == runs["status"][i] and np.array_equal(U[0], runs["U"][i]) and np.array_equal(Y[0], runs["Y"][i])
    s.reset_warm_start()
    i0 = np.nonzero((runs["map"] == 12) & (runs["step"] == 0))[0][0]
    u, st, stats, ms = s.call(runs["P"][i0])
    assert st == runs["status"][i0] and np.array_equal(u, runs["U"][i0])


@pytest.mark.gpu
@pytest.mark.parametrize("cx", list(range(1, 13)))
def test_fleet_kernels_walk_the_recorded_runs(oracle, gpu_solver_factory, runs, cx):
    """The reference's own run of map cx (its start / goal, its moving obstacles) stepped by the device fleet kernels
    (assemble -> solve with the persisted warm start -> plant step, all on the device): over the first 19 steps the
    device loop equals the host mirror + oracle loop bit for bit, and at every kept step of the recording (0, 3, ...,
    18) the parameter vector assembled ON THE DEVICE, the reply and the exit flag are the recorded ones bit for bit —
    the recording's vectors come from the unmodified src/path_generator.py:293-382 and its plant
    (src/mpc/mpc_generator.py:225-231, libm sin / cos; the solver's sincos is <= 2 ulp away and happens to round the
    same way along these steps; from step 21 of map 6 on the two drift apart at 1e-6, which is why this stops at 18)."""
    import mpc_trajectory_generator_b200 as pkg
    from mpc_trajectory_generator_b200 import workloads
    from mpc_trajectory_generator_b200.fleet import FleetPlan
    steps = 19
    hc = assembly.HostConfig.default()
    gmap = assembly.load_maps()[cx]
    dev_sc = assembly.Scenario(hc, gmap, sinus_object=(cx == 12))
    host_sc = assembly.Scenario(hc, gmap, sinus_object=(cx == 12))
    plan = FleetPlan.from_scenarios([dev_sc], max_steps=steps)
    solver = gpu_solver_factory(workloads.solver_config_for(hc))
    fleet = pkg.NmpcFleet(solver, plan, log_steps=steps)
    sel = np.nonzero(runs["map"] == cx)[0]
    kept = {int(runs["step"][i]): i for i in sel}
    ocfg = oracle.default_config()
    U = np.zeros((1, 2 * hc.N_hor))
    Y = np.zeros((1, 2 * hc.N_hor))
    checked = 0
    for k in range(steps):
        fleet.step(1)
        P, Ud, Yd = fleet.last()
        st = fleet.state()
        ph = host_sc.parameters()
        assert np.array_equal(P[0], ph), f"map {cx} step {k}: device and host mirror assemble different parameters"
        U, Y, sth, _ = oracle.solve_batch(ocfg, ph[None], U, Y)
        assert st["status"][0] == sth[0] and np.array_equal(Ud[0], U[0]) and np.array_equal(Yd[0], Y[0])
        if k in kept:
            i = kept[k]
            assert np.array_equal(P[0], runs["P"][i]), f"map {cx} step {k}: not the reference's parameter vector"
            assert np.array_equal(Ud[0], runs["U"][i]) and st["status"][0] == runs["status"][i]
            checked += 1
        finished = host_sc.apply(U[0], sincos=lambda th: oracle.sincos(th))
        assert np.array_equal(fleet.state()["state"][0], np.array(host_sc.states[-3:]))
        assert not finished
    assert checked == 7
    fleet.close()
