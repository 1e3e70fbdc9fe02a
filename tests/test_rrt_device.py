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
@pytest.mark.parametrize("name", ["benchmark_3", "benchmark_4", "benchmark_5", "benchmark_6"])
def test_cuda_rrt_kernel_plans_the_same_paths_as_the_tensor_program(name, library):
    """nlo_rrt_paths (one warp per planner, csrc/rrt_kernels.cu) against the lock-step torch program on the CPU: same seeds, same
    draws, same trees - the extracted paths agree node for node (fp64 on both sides)."""
    from nlotrajectories_b200.rrt_device import _host_planner, cuda_rrt_paths
    cfg = Config.load(bench_yaml(name))
    ini, b = cfg.solver.initializer, cfg.body
    host, bounds = _host_planner(cfg)
    seeds = [2000 + i for i in range(24)]
    got = cuda_rrt_paths(cfg, seeds, b.start_state, b.goal_state, bounds, ini.step_size, ini.max_iter, host.inflation)
    pos, parent, final = batched_rrt_trees(torch_scene_sdf(cfg, CPU), b.start_state, b.goal_state, bounds, len(seeds), seeds, ini.step_size,
                                           ini.max_iter, host.inflation, device=CPU)
    same = 0
    for i in range(len(seeds)):
        if final[i] < 0:
            assert got[i] is None
            continue
        idx, node = [], int(final[i])
        while node >= 0:
            idx.append(node); node = int(parent[i, node])
        want = pos[i, idx[::-1]]
        assert got[i] is not None and got[i].shape == want.shape, (i, None if got[i] is None else got[i].shape, want.shape)
        np.testing.assert_allclose(got[i], want, atol=1e-9)
        same += 1
    assert same >= len(seeds) // 2


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["benchmark_4", "benchmark_6"])
def test_cuda_rrt_kernel_shortcuts_like_the_host_planner(name, library):
    """postprocess=1: corner midpoints + greedy shortcut inside the kernel against RRTInitializer's numpy versions on the same raw
    paths (decisions are comparisons of exact SDF values: a rare near-tie may differ, most paths must agree node for node)."""
    from nlotrajectories_b200.rrt_device import _host_planner, cuda_rrt_paths
    cfg = Config.load(bench_yaml(name))
    ini, b = cfg.solver.initializer, cfg.body
    host, bounds = _host_planner(cfg)
    seeds = [3000 + i for i in range(48)]
    args = (cfg, seeds, b.start_state, b.goal_state, bounds, ini.step_size, ini.max_iter, host.inflation)
    raw = cuda_rrt_paths(*args)
    post = cuda_rrt_paths(*args, postprocess=True)
    same = total = 0
    for r, p in zip(raw, post):
        if r is None:
            assert p is None
            continue
        want = host._shortcut(host._insert_intermediate_points(r))
        total += 1
        np.testing.assert_allclose(p[0], want[0]); np.testing.assert_allclose(p[-1], want[-1])
        same += int(p.shape == want.shape and np.allclose(p, want, atol=1e-9))
    assert total >= 24 and same >= 0.9 * total, (same, total)


@pytest.mark.gpu
def test_cuda_rrt_4096_benchmark_6_plans(library, capsys):
    """4,096 benchmark_6 starts planned on the GPU (tree search: one kernel) and post-processed on the host cores."""
    import time
    cfg = Config.load(bench_yaml("benchmark_6"))
    t0 = time.time()
    w0 = rrt_multistart_device(cfg, 4096, device=torch.device("cuda"))
    dt = time.time() - t0
    N, nx = cfg.solver.N, 7
    X = w0[:, :nx * (N + 1)].reshape(4096, N + 1, nx).astype(float)
    clear = scene_sdf(cfg)(X[..., 0], X[..., 1]).min(axis=1)
    with capsys.disabled():
        print(f"\n[rrt] 4,096 benchmark_6 plans in {dt:.2f} s (kernel + host post-processing); {(clear > 0).mean() * 100:.1f} % of the splines clear the obstacles")
    assert (clear > 0).mean() > 0.95
    assert len({tuple(np.round(x[::8, :2].ravel(), 5)) for x in X}) > 4000


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
