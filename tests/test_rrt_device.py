"""Batched RRT - run here on torch's CPU device; the same tensor program runs on the GPU."""
import numpy as np
import pytest
import torch

from conftest import bench_yaml
from nlotrajectories_b200.config import Config
from nlotrajectories_b200.rrt_device import batched_rrt_trees, rrt_multistart_device, torch_scene_sdf
from nlotrajectories_b200.train import scene_sdf

CPU = torch.device("cpu")


@pytest.mark.parametrize("name", ["benchmark_1", "benchmark_3", "benchmark_4", "benchmark_5", "benchmark_6"])
def test_torch_scene_sdf_equals_the_numpy_one(name):
    cfg = Config.load(bench_yaml(name))
    Q = np.random.default_rng(3).uniform(-0.5, 1.5, (4000, 2))
    want = scene_sdf(cfg)(Q[:, 0], Q[:, 1])
    got = torch_scene_sdf(cfg, CPU)(torch.from_numpy(Q)).numpy()
    np.testing.assert_allclose(got, want, atol=1e-14)


def test_trees_are_collision_free_step_long_and_seeded():
    cfg = Config.load(bench_yaml("benchmark_6"))
    ini, b = cfg.solver.initializer, cfg.body
    sdf = torch_scene_sdf(cfg, CPU)
    infl = 0.05
    args = (sdf, b.start_state, b.goal_state, ini.rrt_bounds)
    pos, parent, final = batched_rrt_trees(*args, 6, [11, 12, 13, 14, 15, 16], ini.step_size, ini.max_iter, infl, device=CPU)
    exact = scene_sdf(cfg)
    paths = []
    for i in range(6):
        assert final[i] > 0
        idx, node = [], int(final[i])
        while node >= 0:
            idx.append(node); node = int(parent[i, node])
        path = pos[i, idx[::-1]]
        paths.append(path)
        np.testing.assert_allclose(path[0], b.start_state[:2]); np.testing.assert_allclose(path[-1], b.goal_state[:2])
        assert exact(path[:-1, 0], path[:-1, 1]).min() >= infl - 1e-12             # every tree node keeps the inflated clearance
        seg = np.linalg.norm(np.diff(path, axis=0), axis=1)
        np.testing.assert_allclose(seg[:-1], ini.step_size, rtol=1e-9)             # steer length; the last hop to the goal is shorter
        assert seg[-1] < ini.step_size
    assert not np.allclose(paths[0][:5], paths[1][:5])                              # one random stream per seed
    # a start depends on its own seed only
    pos2, parent2, final2 = batched_rrt_trees(*args, 2, [13, 14], ini.step_size, ini.max_iter, infl, device=CPU)
    assert final2[0] == final[2] and final2[1] == final[3]
    np.testing.assert_array_equal(pos2[0, :final2[0] + 1], pos[2, :final[2] + 1])


def test_multistart_guesses_have_the_reference_shape_and_clear_the_obstacles():
    cfg = Config.load(bench_yaml("benchmark_3"))
    w = rrt_multistart_device(cfg, 4, first=2, device=CPU)
    N, nx = cfg.solver.N, 5
    assert w.shape[0] == 4 and w.dtype == np.float32
    X = w[:, :nx * (N + 1)].reshape(4, N + 1, nx).astype(float)
    np.testing.assert_allclose(X[:, 0, :2], np.tile(cfg.body.start_state[:2], (4, 1)), atol=1e-6)
    np.testing.assert_allclose(X[:, -1, :2], np.tile(cfg.body.goal_state[:2], (4, 1)), atol=1e-6)
    assert np.all(X[:, :, 2:] == 0) and np.all(w[:, nx * (N + 1):] == 0)           # only (x, y) are planned (trajectory_initialization.py:228-231)
    assert scene_sdf(cfg)(X[..., 0], X[..., 1]).min() > 0
    again = rrt_multistart_device(cfg, 2, first=3, device=CPU)
    np.testing.assert_array_equal(again, w[1:3])


@pytest.mark.gpu
def test_trees_on_the_gpu_are_valid_and_feed_the_solver(library):
    """The same tensor program on CUDA: every start of benchmark_3 finds a path whose spline clears the obstacles, and the guesses
    go straight into the batched interior point."""
    from gpu_util import to_weights
    from oracle import sdf_oracle as so
    from conftest import GOLDEN
    from nlotrajectories_b200.problem import NlpProblem
    from nlotrajectories_b200.sdf import LearnedSDF
    from nlotrajectories_b200.solver import BatchedIPSolver, DeviceEvaluator
    cfg = Config.load(bench_yaml("benchmark_3"))
    w0 = rrt_multistart_device(cfg, 16, device=torch.device("cuda"), workers=1)
    N, nx = cfg.solver.N, 5
    X = w0[:, :nx * (N + 1)].reshape(16, N + 1, nx).astype(float)
    assert scene_sdf(cfg)(X[..., 0], X[..., 1]).min() > 0
    assert len({tuple(np.round(x[:, :2].ravel(), 6)) for x in X}) == 16          # sixteen different paths
    model = LearnedSDF(to_weights(so.from_npz(str(GOLDEN / "sdf_benchmark_3_relu128.npz"))))
    prob = NlpProblem.from_config(cfg, model)
    lb, ub = prob.bounds()
    res = BatchedIPSolver(DeviceEvaluator(prob), lb, ub, max_iter=200).solve(torch.from_numpy(w0.astype(np.float64)).cuda())
    usable = (res.converged | res.stalled).cpu().numpy() & (res.violation.cpu().numpy() <= 1e-4)
    assert usable.mean() >= 0.5
    assert abs(res.f.cpu().numpy()[usable].min() - 1.5321) < 5e-3
    model.close()
