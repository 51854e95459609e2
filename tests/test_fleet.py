"""Fleet stepping: the device-side receding-horizon loop against the host mirror of the
reference's loop (host.assembly.Scenario — itself validated bit for bit against the unmodified reference's recorded
run, tests/test_host_assembly.py) driven by the C oracle.

Bars: the parameter vector every robot assembles on the device equals the host mirror's `parameters()` bit for bit at
every step; solutions, multipliers, exit flags, plant states and termination flags equal the host loop's bit for bit
when the host plant uses the solver's sincos (<= 2 ulp from libm; with libm the closed loops agree to ~1e-9 until a
non-converged solve amplifies the last-bit difference — reported, not asserted)."""
import numpy as np
import pytest

from mpc_trajectory_generator_b200 import workloads
from mpc_trajectory_generator_b200.fleet import FleetPlan
from mpc_trajectory_generator_b200.host import assembly


def _scenarios(complexity, n, seed, smooth=False, n_dyn=None, **cfgkw):
    hc = assembly.HostConfig.smooth_velocity(**cfgkw) if smooth else assembly.HostConfig.default(**cfgkw)
    if complexity in (2, 12):   # maps with dynamic obstacles: the map's own start/goal plus nearby variants
        gmap = dict(assembly.load_maps()[complexity])
        if n_dyn is not None:   # a map with fewer moving obstacles than the solver has slots
            gmap["dyn_obs"] = gmap["dyn_obs"][:n_dyn]
        rng = np.random.default_rng(seed)
        out = []
        env = assembly.Scenario.make_env(hc, gmap)
        while len(out) < n:
            s0 = list(gmap["start"])
            s0[0] += rng.uniform(-0.3, 0.3)
            s0[1] += rng.uniform(-0.3, 0.3)
            sc = assembly.Scenario(hc, gmap, s0, gmap["end"], env=env)
            if sc.ok:
                out.append(sc)
        return hc, out
    return hc, workloads.random_scenarios(hc, complexity, n, seed)


def test_plan_packing_cpu():
    hc, scs = _scenarios(3, 5, seed=4)
    plan = FleetPlan.from_scenarios(scs)
    assert plan.n_robots == 5 and plan.ref.shape[2] == 3 and plan.sched is None
    for b, s in enumerate(scs):
        n = plan.n_ref[b]
        assert n == len(s.x_ref)
        assert np.array_equal(plan.ref[b, :n, 0], s.x_ref) and np.array_equal(plan.ref[b, :n, 2], s.theta_ref)
        assert plan.n_vert[b] == len(s.vert)
        assert np.array_equal(plan.goal[b], s.end)
    assert plan.base_speed == hc.lin_vel_max * hc.throttle_ratio
    assert plan.circle_radius == hc.vehicle_width / 2 + hc.vehicle_margin


def test_no_circles_on_an_obstacle_free_map_cpu():
    """The reference fills the static-circle slots only when the map has obstacles (src/path_generator.py:295), even
    though find_original_vertices also returns corner vertices of a non-convex boundary: the fleet plan must not send
    them either."""
    hc = assembly.HostConfig.default()
    gmap = {"complexity": 99, "boundary": [[0.0, 0.0], [20.0, 0.0], [20.0, 20.0], [12.0, 20.0], [12.0, 8.0], [8.0, 8.0],
                                            [8.0, 20.0], [0.0, 20.0]],
            "obstacles": [], "start": [4.0, 16.0, -1.5], "end": [16.0, 16.0, 1.5], "dyn_obs": []}
    sc = assembly.Scenario(hc, gmap)
    assert sc.ok and len(sc.path) >= 3 and len(sc.vert) > 0      # the path bends around boundary corners
    plan = FleetPlan.from_scenarios([sc])
    assert plan.n_vert[0] == 0
    p = sc.parameters()
    assert not np.any(p[20 + hc.N_hor:20 + hc.N_hor + 3 * hc.Nobs])   # the host mirror sends no circles either


@pytest.mark.parametrize("steps,n_dyn", [(1, 3), (2, 3), (3, 3), (1, 1), (2, 2), (3, 1)])
def test_dynamic_schedule_matches_ring_cpu(steps, n_dyn):
    """The closed form the device uses against the reference's rotate-and-append ring (src/path_generator.py:306-316,
    restated literally by host.assembly.Scenario.parameters): real obstacle k, slot j at plant time t = schedule entry
    t + j; with unused slots, position q of the unused region holds obstacle 0's entry q + t - Lp once that is >= 0
    (the rotation of the flat list leaks the front of block 0 into the end of the list)."""
    hc, scs = _scenarios(12, 1, seed=0, n_dyn=n_dyn, num_steps_taken=steps)
    sc = scs[0]
    N, Nd = hc.N_hor, hc.Ndynobs
    iters = 30
    plan = FleetPlan.from_scenarios(scs, max_steps=iters)
    assert plan.num_steps_taken == steps and plan.n_dyn == n_dyn
    off = 20 + N + 3 * hc.Nobs
    Lp = (Nd - n_dyn) * N
    entry = lambda m, k: plan.sched_init[m, k] if m < N else plan.sched[m, k]  # noqa: E731
    phantom = np.array([0.0, 0.0, 1.0, 1.0, 0.0])
    for c in range(iters):
        t = c * steps
        p = sc.parameters()
        ring = p[off:off + 5 * Nd * N].reshape(Nd, N, 5)
        for k in range(Nd):
            for j in range(N):
                if k < n_dyn:
                    want = entry(t + j, k)
                else:
                    m = (k - n_dyn) * N + j + t - Lp
                    want = entry(m, 0) if m >= 0 else phantom
                assert np.array_equal(ring[k, j], want), (c, k, j)
        sc.apply(np.zeros(2 * N))   # any input: the ring does not depend on the robot


def _host_loop(oracle, hc, scs, steps, sincos):
    """reference loop on the host: parameters() -> oracle solve (persisted un-shifted warm start) -> apply"""
    ocfg = _oracle_cfg(oracle, hc)
    B, n2 = len(scs), 2 * hc.N_hor
    U = np.zeros((B, n2))
    Y = np.zeros((B, n2))
    live = np.ones(B, dtype=bool)
    hist = []
    for k in range(steps):
        ids = np.nonzero(live)[0]
        P = np.zeros((B, oracle.param_len(ocfg)))
        st = np.zeros(B, dtype=np.int32)
        if len(ids):
            P[ids] = np.stack([scs[i].parameters() for i in ids])
            U[ids], Y[ids], st[ids], _ = oracle.solve_batch(ocfg, P[ids], U[ids], Y[ids])
            for i in ids:
                if scs[i].apply(U[i], sincos=sincos):
                    live[i] = False
        hist.append(dict(P=P.copy(), U=U.copy(), Y=Y.copy(), status=st.copy(), live_before=ids.copy(),
                         state=np.array([s.states[-3:] for s in scs]), idx=np.array([s.idx for s in scs]),
                         done=(~live).astype(np.int32)))
    return hist


def _oracle_cfg(oracle, hc):
    return oracle.default_config(N_hor=hc.N_hor, Nobs=hc.Nobs, Ndynobs=hc.Ndynobs, ang_vel_max=hc.ang_vel_max,
                                 ang_acc_max=hc.ang_acc_max, lin_vel_min=hc.lin_vel_min, lin_vel_max=hc.lin_vel_max,
                                 lin_acc_min=hc.lin_acc_min, lin_acc_max=hc.lin_acc_max, ts=hc.ts)


@pytest.mark.gpu
@pytest.mark.parametrize("complexity,B,steps,smooth,N,taken,n_dyn",
                         [(3, 24, 12, False, 20, 1, None), (12, 6, 25, False, 20, 1, None), (1, 8, 10, False, 20, 1, None),
                          (11, 6, 8, True, 40, 1, None),    # BASELINE config 4's setup
                          (11, 6, 8, True, 40, 2, None),    # smooth_velocity.yaml as shipped: two controls per solve
                          (12, 4, 14, False, 20, 3, None),  # william_config.yaml's num_steps_taken with moving obstacles
                          (12, 4, 30, False, 20, 1, 1),     # fewer moving obstacles than slots: the leaking rotation
                          (2, 4, 12, False, 20, 2, 2)])
def test_fleet_steps_match_host_loop(oracle, gpu_solver_factory, complexity, B, steps, smooth, N, taken, n_dyn):
    import mpc_trajectory_generator_b200 as pkg
    hc, scs = _scenarios(complexity, B, seed=10 + complexity, smooth=smooth, n_dyn=n_dyn, N_hor=N, num_steps_taken=taken)
    plan = FleetPlan.from_scenarios(scs, max_steps=steps)
    solver = gpu_solver_factory(workloads.solver_config_for(hc))
    fleet = pkg.NmpcFleet(solver, plan, log_steps=steps)

    def sincos(th):
        s, c = oracle.sincos(th)
        return s, c
    hist = _host_loop(oracle, hc, scs, steps, sincos)
    for k in range(steps):
        fleet.step(1)
        P, U, Y = fleet.last()
        st = fleet.state()
        h = hist[k]
        ids = h["live_before"]
        assert np.array_equal(P[ids], h["P"][ids]), f"step {k}: assembled parameters differ"
        assert np.array_equal(st["status"][ids], h["status"][ids]), f"step {k}: exit flags differ"
        assert np.array_equal(U[ids], h["U"][ids]), f"step {k}: solutions differ"
        assert np.array_equal(Y[ids], h["Y"][ids]), f"step {k}: multipliers differ"
        assert np.array_equal(st["state"], h["state"]), f"step {k}: plant states differ"
        assert np.array_equal(st["idx"][ids], h["idx"][ids])
        assert np.array_equal(st["done"], h["done"])
    lg, n = fleet.log()
    assert np.array_equal(np.minimum(n, steps), np.minimum(st["t"], steps))   # the log holds `steps` plant steps
    b = 0
    assert np.array_equal(lg[b, :n[b], 0:3], np.array(scs[b].states[3:]).reshape(-1, 3)[:n[b]])
    fleet.close()


@pytest.mark.gpu
def test_fleet_run_to_goal_multi_step_launch(oracle, gpu_solver_factory):
    """n steps in one call == n calls of one step; a whole run on the reference's default scenario terminates at the
    goal like the host loop (same step count, same final state)."""
    import mpc_trajectory_generator_b200 as pkg
    hc = assembly.HostConfig.default()
    gmap = assembly.load_maps()[1]
    scs = [assembly.Scenario(hc, gmap)]
    plan = FleetPlan.from_scenarios(scs)
    solver = gpu_solver_factory(workloads.solver_config_for(hc))
    fa = pkg.NmpcFleet(solver, plan, log_steps=400)
    fb = pkg.NmpcFleet(solver, plan, log_steps=400)
    fa.step(40)
    for _ in range(40):
        fb.step(1)
    sa, sb = fa.state(), fb.state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    fa.step(260)
    sa = fa.state()
    assert sa["done"][0] == 1, "the default scenario reaches its goal within 300 steps"
    hist = _host_loop(oracle, hc, scs, int(sa["t"][0]), lambda th: oracle.sincos(th))
    assert hist[-1]["done"][0] == 1 and (len(hist) < 2 or hist[-2]["done"][0] == 0)
    assert np.array_equal(hist[-1]["state"][0], sa["state"][0])
    fa.close()
    fb.close()


@pytest.mark.gpu
@pytest.mark.parametrize("complexity", [3, 11])
def test_device_reference_sampler(gpu_solver_factory, complexity):
    """rough_ref on the device against the host mirror of src/mpc/mpc_generator.py:17-57.
    Same statements in the same order; hypot / atan2 are CUDA's instead of glibc's, so the bar is 1e-12 absolute on
    positions and headings (observed: last-bit differences) and identical sample counts."""
    import mpc_trajectory_generator_b200 as pkg
    hc, scs = _scenarios(complexity, 64, seed=30 + complexity)
    plan = FleetPlan.from_scenarios(scs)
    solver = gpu_solver_factory(workloads.solver_config_for(hc))
    fleet = pkg.NmpcFleet(solver, plan, sample_refs_on_device=True)
    ref, n = fleet.sample_refs(read_back=True)
    assert np.array_equal(n, plan.n_ref)
    for b in range(plan.n_robots):
        assert np.abs(ref[b, :n[b]] - plan.ref[b, :n[b]]).max() <= 1e-12
    # the fleet steps with the device-made references: first-step parameters agree with the host mirror to 1e-12
    fleet.step(1)
    P, _, _ = fleet.last()
    Ph = np.stack([s.parameters() for s in scs])
    assert np.abs(P - Ph).max() <= 1e-12
    fleet.close()


@pytest.mark.gpu
def test_large_fleet_longest_first_order(oracle, gpu_solver_factory):
    """A fleet larger than the GPU's warp slots: steps after the first hand the robots to the solve kernel longest-first
    (device counting sort on the previous step's iteration counts), the first one by the probe.  Scheduling must not
    change anything: every replica of a scenario follows its host loop bit for bit."""
    import mpc_trajectory_generator_b200 as pkg
    hc, scs = _scenarios(3, 16, seed=77)
    plan = FleetPlan.from_scenarios(scs)
    rep = 130                                   # 16 x 130 = 2080 robots > 148 SMs x 12 warps
    tile = lambda a: np.ascontiguousarray(np.concatenate([a] * rep))  # noqa: E731
    big = FleetPlan(tile(plan.n_ref), tile(plan.ref), tile(plan.n_vert), tile(plan.vert), tile(plan.start),
                    tile(plan.goal), plan.brake_vel, plan.brake_dist, plan.weights, plan.base_speed, plan.circle_radius)
    solver = gpu_solver_factory(workloads.solver_config_for(hc))
    fleet = pkg.NmpcFleet(solver, big)
    steps = 4
    hist = _host_loop(oracle, hc, scs, steps, lambda th: oracle.sincos(th))
    fleet.step(steps)
    P, U, Y = fleet.last()
    st = fleet.state()
    for r in range(rep):
        sl = slice(16 * r, 16 * (r + 1))
        assert np.array_equal(U[sl], hist[-1]["U"]) and np.array_equal(Y[sl], hist[-1]["Y"])
        assert np.array_equal(st["state"][sl], hist[-1]["state"]) and np.array_equal(P[sl], hist[-1]["P"])
        assert np.array_equal(st["status"][sl], hist[-1]["status"])
    fleet.close()
