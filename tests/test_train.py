"""SDF training tool: exact SDFs without shapely, the reference's sampler, a tiny training run and the
weight export round trip (CPU only)."""
import numpy as np
import pytest

from conftest import bench_yaml


def test_exact_sdfs_match_closed_forms():
    from nlotrajectories_b200 import train as T
    rng = np.random.default_rng(0)
    x, y = rng.uniform(-1, 2, 5000), rng.uniform(-1, 2, 5000)
    # core/sdf/casadi.py:33-38
    np.testing.assert_allclose(T.sdf_circle(x, y, (0.5, 0.5), 0.2, 0.05), np.hypot(x - 0.5, y - 0.5) - 0.25)
    # a square given as a polygon equals the box formula (core/sdf/casadi.py:54-67 vs :135-148)
    sq = T.sdf_square(x, y, (0.5, 0.4), 0.6, 0.0)
    pg = T.sdf_polygon(x, y, [(0.2, 0.1), (0.8, 0.1), (0.8, 0.7), (0.2, 0.7)], 0.0)
    np.testing.assert_allclose(pg, sq, atol=1e-12)
    # polygon margin is subtracted from the signed distance (core/sdf/casadi.py:147)
    np.testing.assert_allclose(T.sdf_polygon(x, y, [(0, 0), (1, 0), (0, 1)], 0.1), T.sdf_polygon(x, y, [(0, 0), (1, 0), (0, 1)], 0.0) - 0.1)
    # known values for a triangle
    v = T.sdf_polygon(np.array([0.25, 2.0, -1.0]), np.array([0.25, 0.0, 0.0]), [(0, 0), (1, 0), (0, 1)])
    np.testing.assert_allclose(v, [-0.25, 1.0, 1.0], atol=1e-12)
    # elliptical half-ring: 2 x 15 arc points, first outer point at angle 0, last inner point at angle 0 (core/sdf/casadi.py:218-246)
    pts = T.elliptic_ring_points((0.5, 0.5), (0.4, 0.2), 0.05)
    assert len(pts) == 30
    np.testing.assert_allclose(pts[0], (0.9, 0.5)); np.testing.assert_allclose(pts[-1], (0.85, 0.5))
    with pytest.raises(ValueError):
        T.elliptic_ring_points((0, 0), (0.1, 0.05), 0.05)


@pytest.mark.parametrize("name", ["benchmark_3", "benchmark_4", "benchmark_6"])
def test_scene_sdf_and_sampler(name):
    from nlotrajectories_b200 import train as T
    from nlotrajectories_b200.config import Config
    cfg = Config.load(bench_yaml(name))
    f = T.scene_sdf(cfg)
    s0 = f(np.array([cfg.body.start_state[0]]), np.array([cfg.body.start_state[1]]))
    s1 = f(np.array([cfg.body.goal_state[0]]), np.array([cfg.body.goal_state[1]]))
    assert s0[0] > 0 and s1[0] > 0                       # start and goal lie outside the obstacles
    xs, ys = T.sample_points(f, (-0.5, 1.5), (-0.5, 1.5), 4000, margin=0.1, boundary_fraction=0.3, rng=np.random.default_rng(1))
    assert len(xs) == 4000 and np.all(np.abs(f(xs[-1200:], ys[-1200:])) < 0.1)          # the boundary-biased share
    xs2, _ = T.sample_points(f, (-0.5, 1.5), (-0.5, 1.5), 4000, margin=0.1, boundary_fraction=0.3, rng=np.random.default_rng(1))
    assert np.array_equal(xs, xs2)                       # seeded


def test_tiny_training_run_and_export_roundtrip(tmp_path):
    import torch
    from nlotrajectories_b200 import train as T
    from nlotrajectories_b200.config import Config
    from nlotrajectories_b200.sdf import SdfWeights
    from oracle import sdf_oracle as so
    cfg = Config.load(bench_yaml("benchmark_3"))
    net, info = T.train(cfg, n_samples=4000, epochs=3, batch_size=512, seed=0, device="cpu", verbose=False)
    assert info["epochs"] == 3 and np.isfinite(info["val_mse"])
    w = SdfWeights.from_state_dict("mlp", net.state_dict(), activation_function="ReLU")
    assert (w.hidden, w.n_hidden_mats) == (128, 1)       # 2 -> 128 -> 128 -> 1 (benchmark_3 model: hidden_dim 128, num_hidden_layers 2)
    w.save_npz(tmp_path / "w.npz")
    onet = so.from_npz(tmp_path / "w.npz")
    P = np.random.default_rng(2).uniform(-0.5, 1.5, (200, 2)).astype(np.float32)
    with torch.no_grad():
        ref = net(torch.from_numpy(P)).numpy()[:, 0]
    np.testing.assert_allclose(so.forward(onet.astype(np.float64), P.astype(np.float64)), ref, atol=2e-5)
    assert np.array_equal(SdfWeights.from_npz(tmp_path / "w.npz").blob, w.blob)
