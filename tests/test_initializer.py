"""Initial-guess generation (reference: core/trajectory_initialization.py).  CPU only."""
import numpy as np
import pytest

from conftest import bench_yaml


def test_linear_and_default_initializers():
    from nlotrajectories_b200.initializer import DefaultInitializer, LinearInitializer
    X = LinearInitializer([0, 0, 0.785, 0, 0], [1, 1, 0.785, 0, 0], N=40).get_initial_guess()
    assert X.shape == (41, 5)                                  # N + 1 rows (trajectory_initialization.py:54-55)
    np.testing.assert_allclose(X[20], [0.5, 0.5, 0.785, 0, 0])
    assert DefaultInitializer().get_initial_guess() is None


@pytest.mark.parametrize("name", ["benchmark_3", "benchmark_4", "benchmark_6"])
def test_rrt_paths_are_collision_free_seeded_and_shaped_like_the_reference(name):
    from nlotrajectories_b200.config import Config
    from nlotrajectories_b200.initializer import RRTInitializer, rrt_multistart
    from nlotrajectories_b200.train import scene_sdf
    cfg = Config.load(bench_yaml(name))
    sdf = scene_sdf(cfg)
    b, s = cfg.body, cfg.solver
    hl, hw = 0.5 * b.length, 0.5 * b.width
    body = [(-hl, -hw), (-hl, hw), (hl, hw), (hl, -hw)]
    mk = lambda seed: RRTInitializer(s.N + 1, b.start_state, b.goal_state, s.dt, sdf, s.initializer.rrt_bounds, body_points=body,
                                     step_size=s.initializer.step_size, max_iter=s.initializer.max_iter, margin=s.initializer.margin, seed=seed)
    p = mk(3)
    assert abs(p.inflation - (max(hl, hw) + s.initializer.margin)) < 1e-12      # max_b |min(b_x, b_y)| + margin (:109-114)
    X = p.get_initial_guess()
    assert X.shape == (s.N + 1, len(b.start_state))            # exactly N rows for N = solver.N + 1 (run_benchmark.py:116)
    np.testing.assert_allclose(X[0, :2], b.start_state[:2], atol=1e-12)
    np.testing.assert_allclose(X[-1, :2], b.goal_state[:2], atol=1e-12)
    assert np.all(X[:, 2:] == 0)                                # only (x, y) are planned (:228-231)
    pos, parent = p.last_tree
    # every tree edge was checked against the inflated obstacles
    for i in range(1, len(pos) - 1):
        assert sdf(pos[i:i + 1, 0], pos[i:i + 1, 1])[0] >= p.inflation - 1e-12
    assert np.array_equal(mk(3).get_initial_guess(), X)         # seeded
    assert not np.array_equal(mk(4).get_initial_guess(), X)
    w = rrt_multistart(cfg, 3, first=5)
    n_X = len(b.start_state) * (s.N + 1)
    assert w.shape[0] == 3 and np.all(w[:, n_X:] == 0)          # U and slack start at zero (core/runner.py:106-108)
    assert np.array_equal(rrt_multistart(cfg, 1, first=6)[0], w[1])              # global start index = seed offset


def test_rrt_raises_like_the_reference_when_no_path_exists():
    from nlotrajectories_b200.initializer import RRTInitializer
    wall = lambda x, y: np.abs(x - 0.5) - 0.1                   # an infinite wall between start and goal
    p = RRTInitializer(11, [0, 0], [1, 0], 0.1, wall, [[0, -1], [1, 1]], rectangle=False, max_iter=200, seed=0)
    with pytest.raises(RuntimeError, match="RRT failed"):
        p.get_initial_guess()


@pytest.mark.parametrize("name", ["benchmark_3", "benchmark_5"])
def test_lifted_guess_zeroes_the_position_defects(name):
    """lift_path: heading and speed read off the planned path make the forward-Euler defects of the (x, y) rows vanish
    (up to the clipping of the controls), the pinned start state is restored."""
    import yaml
    from oracle import nlp_oracle as no
    from nlotrajectories_b200.config import Config
    from nlotrajectories_b200.initializer import rrt_multistart
    from conftest import bench_yaml
    cfg = Config.load(bench_yaml(name))
    spec = no.NlpSpec.from_yaml_dict(yaml.safe_load(open(bench_yaml(name))))
    w_flat = rrt_multistart(cfg, 2, lift=False).astype(np.float64)
    w_lift = rrt_multistart(cfg, 2, lift=True).astype(np.float64)
    np.testing.assert_allclose(w_lift[:, :spec.nx], np.tile(np.asarray(cfg.body.start_state, float), (2, 1)), atol=1e-6)
    X0, U0, _ = no.unpack(spec, w_flat)
    X1, U1, _ = no.unpack(spec, w_lift)
    np.testing.assert_allclose(X1[:, 1:, :2], X0[:, 1:, :2], atol=1e-6)        # the planned positions are kept
    def pos_defect(X, U):
        f, _, _ = no.dynamics_f(spec, X[:, :-1], U)
        return np.abs(X[:, 1:, :2] - X[:, :-1, :2] - spec.dt * f[..., :2])[:, 1:]   # knot 0 holds the pinned start heading
    assert pos_defect(X1, U1).max() < 0.2 * pos_defect(X0, U0).max()


def test_pooled_planners_match_the_serial_ones():
    """rrt_multistart plans on forked host processes; every start owns its seed, so the result does not depend on the pool."""
    from nlotrajectories_b200.config import Config
    from nlotrajectories_b200.initializer import rrt_multistart
    from conftest import bench_yaml
    cfg = Config.load(bench_yaml("benchmark_3"))
    a = rrt_multistart(cfg, 6, first=3, workers=1)
    b = rrt_multistart(cfg, 6, first=3, workers=3)
    c = rrt_multistart(cfg, 2, first=5, workers=1)
    assert np.array_equal(a, b)
    assert np.array_equal(a[2:4], c)                      # start i depends on first + i only: shards of a batch agree with the whole
