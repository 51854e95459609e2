"""Host logic (no GPU): the package's mirror of the reference's parameter assembly against the
run recorded from the UNMODIFIED reference (tests/golden/make_fixtures.py), and the planner
substitute against the path the reference's own README picture shows."""
import os

import numpy as np

from mpc_trajectory_generator_b200.host import assembly, planner

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_maps_fixture():
    maps = assembly.load_maps()
    assert sorted(maps) == list(range(13))
    assert maps[1]["start"][:2] == [1.0, 5.0] and maps[1]["end"][:2] == [19.0, 10.0]   # src/visibility/graphs.py:43
    assert len(maps[3]["obstacles"]) == 6 and len(maps[11]["obstacles"]) == 4


def test_offset_polygon_miter():
    sq = [(3.0, 3.0), (3.0, 7.0), (7.0, 7.0), (7.0, 3.0)]          # clockwise, like the reference's obstacles
    out = planner.offset_polygon(sq, 0.5)
    assert planner.signed_area(out) > 0
    assert sorted(out) == sorted([(2.5, 2.5), (7.5, 2.5), (7.5, 7.5), (2.5, 7.5)])
    inn = planner.offset_polygon([(0, 0), (10, 0), (10, 10), (0, 10)], -0.5)
    assert sorted(inn) == sorted([(0.5, 0.5), (9.5, 0.5), (9.5, 9.5), (0.5, 9.5)])
    tri = planner.offset_polygon([(45.0, 15.0), (44.0, 20.0), (46.0, 20.0)], 0.5)   # acute apex: squared off
    assert len(tri) == 4


def test_astar_matches_reference_picture():
    """docs/example_image.png: (1,5) -> (4.5,15.5) -> (7.5,15.5) -> (11.5,12) -> (19,10), circles at
    the original vertices (5,15), (7,15), (12,12.5)."""
    cfg = assembly.HostConfig.default()
    s = assembly.Scenario(cfg, assembly.load_maps()[1])
    assert np.allclose(s.path, [(1, 5), (4.5, 15.5), (7.5, 15.5), (11.5, 12.0), (19, 10)], atol=1e-9)
    assert s.vert == [(5.0, 15.0), (7.0, 15.0), (12.0, 12.5)]


def test_helpers_match_reference():
    g = np.load(os.path.join(GOLD, "reference_helpers.npz"))
    cfg = assembly.HostConfig.default()
    bv, bd = assembly.brake_profile(cfg)
    assert np.array_equal(bv, g["brake_velocities"]) and np.array_equal(bd, g["brake_distances"])
    maps = assembly.load_maps()
    for cx in (1, 3, 11, 12):
        s = assembly.Scenario(cfg, maps[cx])
        assert np.array_equal(np.array(s.path), g[f"map{cx}_path"])
        assert np.array_equal(np.array(s.vert).reshape(-1, 2), g[f"map{cx}_vertices"])
        assert np.array_equal(np.array([s.x_ref, s.y_ref, s.theta_ref]).T, g[f"map{cx}_ref"])


def test_parameter_assembly_replays_reference_run():
    """Feed the recorded solutions back through our Scenario: every assembled parameter vector
    must equal the one the unmodified reference built (src/path_generator.py:378-379)."""
    g = np.load(os.path.join(GOLD, "config1_run.npz"))
    cfg = assembly.HostConfig.default()
    s = assembly.Scenario(cfg, assembly.load_maps()[1])
    K = g["P"].shape[0]
    for k in range(K):
        p = s.parameters()
        assert p.shape == (430,)
        assert np.array_equal(p, g["P"][k]), f"step {k}"
        done = s.apply(g["U"][k])
        assert done == (k == K - 1)
    assert np.array_equal(np.array(s.states[0::3]), g["xx"]) and np.array_equal(np.array(s.states[1::3]), g["xy"])
    assert np.array_equal(np.array(s.system_input[0::2]), g["uv"])


def test_dynamic_obstacle_ring_shapes():
    cfg = assembly.HostConfig.default()
    s = assembly.Scenario(cfg, assembly.load_maps()[12], sinus_object=True)
    p0 = s.parameters()
    s.apply(np.zeros(40))
    p1 = s.parameters()
    be = 20 + 20 + 30
    e0 = p0[be:be + 300].reshape(3, 20, 5)
    e1 = p1[be:be + 300].reshape(3, 20, 5)
    assert np.array_equal(e0[:, 1:, :], e1[:, :-1, :])          # ring rotated by one step
    assert np.all(e0[:, :, 2] > 0.5) and np.all(e0[:, :, 3] > 0.5)
