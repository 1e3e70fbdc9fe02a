"""SDF-quality metrics: numpy and torch paths against values produced by the reference's own
core/metrics.py (oracle/make_golden.py --metrics -> tests/golden/metrics_reference.npz), and compute_metrics on the GPU path."""
import numpy as np
import pytest

from conftest import GOLDEN, bench_yaml
from nlotrajectories_b200 import metrics as M


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN / "metrics_reference.npz")


def _ours(target, pred, X, Y, eps):
    return [M.mse(target, pred), M.iou(target, pred, 0.0), M.hausdorff(pred, target, X, Y, eps), M.chamfer(pred, target, X, Y, eps),
            M.surface_loss(target, pred, eps)]


@pytest.mark.parametrize("case", ["shifted", "wavy", "noisy", "far"])
@pytest.mark.parametrize("backend", ["numpy", "torch"])
def test_metrics_match_the_reference_module(gold, case, backend):
    X, Y, target, pred, eps = gold["X"], gold["Y"], gold["target"], gold["pred_" + case], float(gold["eps"])
    if backend == "torch":
        import torch
        X, Y, target, pred = (torch.from_numpy(a) for a in (X, Y, target, pred))
    got = _ours(target, pred, X, Y, eps)
    want = gold["vals_" + case]
    for g, w in zip(got, want):
        if np.isnan(w):
            assert g is None
        else:
            assert g == pytest.approx(w, rel=1e-12, abs=1e-14)


def test_corner_cases(gold):
    t = gold["target"]
    assert M.iou(t + 5.0, t + 6.0) == float(gold["iou_both_empty"]) == 1.0
    assert M.surface_loss(t + 5.0, t, 1e-2) is None
    with pytest.raises(ValueError):
        M.mse(t, t[:-1])
    with pytest.raises(ValueError):
        M.chamfer(t, t[:-1], gold["X"], gold["Y"])


def test_directed_nearest_blocks_agree_with_brute_force():
    rng = np.random.default_rng(0)
    A, B = rng.uniform(-1, 2, (700, 2)), rng.uniform(-1, 2, (900, 2))
    brute = np.sqrt(((A[:, None] - B[None]) ** 2).sum(-1)).min(1)
    np.testing.assert_allclose(M.directed_nearest(A, B, block=128), brute, rtol=1e-13)


@pytest.mark.gpu
def test_compute_metrics_on_the_gpu_path(library):
    """The reference's compute_metrics call (scripts/run_benchmark.py:34-46) for the trained benchmark_3 network: the learned field
    comes from the CUDA path; the numbers must agree with the same metrics of the oracle's field on the host."""
    from gpu_util import to_weights
    from oracle import sdf_oracle as so
    from nlotrajectories_b200.config import Config
    from nlotrajectories_b200.sdf import LearnedSDF
    from nlotrajectories_b200.train import scene_sdf
    net = so.from_npz(str(GOLDEN / "sdf_benchmark_3_relu128.npz"))
    model = LearnedSDF(to_weights(net))
    exact = scene_sdf(Config.load(bench_yaml("benchmark_3")))
    n = 300
    got = M.compute_metrics(model, exact, n_samples=n)
    x = np.linspace(-1, 2, n); X, Y = np.meshgrid(x, x)
    pred = so.value_jac(net.astype(np.float64), np.stack([X.ravel(), Y.ravel()], 1))[0].reshape(n, n)
    want = _ours(exact(X, Y), pred, X, Y, 1e-2)
    assert got[0] == pytest.approx(want[0], rel=1e-3)
    assert got[1] == pytest.approx(want[1], abs=2e-3)            # a few grid points sit within fp32 rounding of the zero level
    assert got[2] == pytest.approx(want[2], abs=2e-2) and got[3] == pytest.approx(want[3], rel=0.05)
    assert got[4] == pytest.approx(want[4], rel=1e-3)
    assert got[0] < 5e-3 and got[1] > 0.9                      # the grid reaches beyond the trained region
    full = M.compute_metrics(model, exact)                        # the reference's 1000^2 grid
    assert full[0] < 5e-3 and full[1] > 0.9 and full[3] < 0.02
    model.close()
