"""SDF-quality metrics of a learned obstacle model (reference: core/metrics.py:7-160 and ``compute_metrics``,
scripts/run_benchmark.py:34-46).

Same five numbers with the reference's definitions - mean squared error, occupancy IoU at a threshold, symmetric Chamfer and
Hausdorff distances between the ``|sdf| < eps`` grid points of the two fields, mean squared prediction on the target's surface
band - computed differently: the reference materialises the full (n_pred x n_target) distance matrix of the band points in
numpy (quadratic memory: tens of GB on the 1000^2 grid); here the two directed nearest-neighbour distance vectors are built
block by block, on the host with numpy or on the GPU with torch when the inputs are CUDA tensors.  ``compute_metrics`` evaluates
the learned field on the grid through the CUDA evaluation path (one value-only launch over the 10^6 grid points).

Return conventions kept: ``None`` when a band is empty (chamfer, hausdorff, surface_loss), IoU 1.0 for two empty occupancies,
``ValueError`` on shape mismatch.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def _is_torch(a) -> bool:
    return type(a).__module__.startswith("torch")


def _check(a, b):
    if tuple(a.shape) != tuple(b.shape):
        raise ValueError("Target and prediction must have the same shape.")


def mse(sdf_target, sdf_pred) -> float:
    _check(sdf_target, sdf_pred)
    d = sdf_target - sdf_pred
    return float((d * d).mean())


def iou(sdf_target, sdf_pred, threshold: float = 0.0) -> float:
    t, p = sdf_target < threshold, sdf_pred < threshold
    union = int((t | p).sum())
    if union == 0:
        return 1.0
    return int((t & p).sum()) / union


def surface_loss(sdf_target, sdf_pred, eps: float = 1e-2) -> Optional[float]:
    band = abs(sdf_target).reshape(-1) < eps
    if not bool(band.any()):
        return None
    v = sdf_pred.reshape(-1)[band]
    return float((v * v).mean())


def _band_points(sdf, X, Y, eps):
    keep = abs(sdf).reshape(-1) < eps
    if _is_torch(sdf):
        import torch
        return torch.stack([X.reshape(-1)[keep], Y.reshape(-1)[keep]], dim=1)
    return np.stack([np.asarray(X).reshape(-1)[keep], np.asarray(Y).reshape(-1)[keep]], axis=1)


def directed_nearest(A, B, block: int = 4096):
    """For every row of A (n, 2) the Euclidean distance to its nearest row of B (m, 2); O(block * m) memory."""
    if _is_torch(A):
        import torch
        out = torch.empty(A.shape[0], dtype=A.dtype, device=A.device)
        b2 = (B * B).sum(1)
        for i in range(0, A.shape[0], block):
            a = A[i:i + block]
            d2 = (a * a).sum(1)[:, None] - 2.0 * (a @ B.T) + b2[None, :]
            j = d2.argmin(1)                                    # refine the winner exactly: the expanded form cancels badly
            out[i:i + block] = (a - B[j]).norm(dim=1)
        return out
    out = np.empty(A.shape[0], dtype=np.result_type(A.dtype, B.dtype))
    b2 = (B * B).sum(1)
    for i in range(0, A.shape[0], block):
        a = A[i:i + block]
        d2 = (a * a).sum(1)[:, None] - 2.0 * (a @ B.T) + b2[None, :]
        j = d2.argmin(1)
        out[i:i + block] = np.linalg.norm(a - B[j], axis=1)
    return out


def _surface_distances(sdf_target, sdf_pred, X, Y, eps):
    _check(sdf_target, sdf_pred)
    P, T = _band_points(sdf_pred, X, Y, eps), _band_points(sdf_target, X, Y, eps)
    if P.shape[0] == 0 or T.shape[0] == 0:
        return None
    return directed_nearest(P, T), directed_nearest(T, P)


def chamfer(sdf_target, sdf_pred, X, Y, eps: float = 1e-2) -> Optional[float]:
    d = _surface_distances(sdf_target, sdf_pred, X, Y, eps)
    return None if d is None else float((d[0].mean() + d[1].mean()) / 2)


def hausdorff(sdf_target, sdf_pred, X, Y, eps: float = 1e-2) -> Optional[float]:
    d = _surface_distances(sdf_target, sdf_pred, X, Y, eps)
    return None if d is None else float(max(d[0].max(), d[1].max()))


def compute_metrics(model, exact_sdf: Callable, x_range: Tuple[float, float] = (-1, 2), y_range: Tuple[float, float] = (-1, 2),
                    n_samples: int = 1000, eps: float = 1e-2):
    """(mse, iou, hausdorff, chamfer, surface_loss) of a ``LearnedSDF`` against the exact field on an ``n_samples``^2 grid
    (scripts/run_benchmark.py:34-46).  The learned field is evaluated by the CUDA path, the band distances on the same GPU."""
    import torch
    x = np.linspace(x_range[0], x_range[1], n_samples)
    y = np.linspace(y_range[0], y_range[1], n_samples)
    X, Y = np.meshgrid(x, y)
    dev = torch.device("cuda", model.device)
    Xd = torch.from_numpy(X.reshape(-1).astype(np.float32)).to(dev)
    Yd = torch.from_numpy(Y.reshape(-1).astype(np.float32)).to(dev)
    pred = model.eval(Xd, Yd, want_jac=False)[0].double()
    target = torch.from_numpy(np.asarray(exact_sdf(X, Y), np.float64).reshape(-1)).to(dev)
    Xg, Yg = torch.from_numpy(X.reshape(-1)).to(dev), torch.from_numpy(Y.reshape(-1)).to(dev)
    return (mse(target, pred), iou(target, pred, 0.0), hausdorff(pred, target, Xg, Yg, eps), chamfer(pred, target, Xg, Yg, eps),
            surface_loss(target, pred, eps))
