"""Train the learned SDF of a benchmark scene and write the weight file the CUDA path consumes.

The reference retrains its network inside every ``run-benchmark`` call (scripts/run_benchmark.py:53-101,
core/sdf/l4casadi.py:77-228) and never saves it.  This module restates that step once, offline:

* exact signed distances of the YAML's obstacles without shapely (core/sdf/casadi.py:33-38 circle, :54-67 square,
  :135-148 polygon = distance to the boundary with the sign of containment, :193-248 elliptical half-ring as the
  polygon of its two sampled arcs, :381-383 union = min);
* the boundary-biased sampler (core/sdf/l4casadi.py:14-66) - including the reference's positional-argument slip that
  feeds ``boundary_fraction`` into ``margin`` (:118) - seeded here, where the reference is not;
* the loss: MSE + surface_loss_weight * mean(pred^2 on |target| < 1e-2) + eikonal_weight * mean((|grad| - 1)^2)
  (core/sdf/l4casadi.py:158-186, core/metrics.py:83-117), Adam lr 1e-3, 90/10 split, patience-10 early stopping;
* the model of the YAML's ``model`` section: ``mlp`` = l4casadi's naive MLP (2 -> H -> (L-1) x [H -> H] -> 1, activation by
  name), Kaiming-uniform weights and zero biases (core/sdf/l4casadi.py:69-74).

PyTorch does the optimisation (training is not on the evaluation hot path); the result is exported with
``SdfWeights.from_state_dict`` to ``.nlow`` / ``.npz``.

    python -m nlotrajectories_b200.train --config <yaml> --out sdf.nlow [--samples 200000] [--epochs 100] [--seed 0]
"""
from __future__ import annotations

import argparse
import copy
from pathlib import Path
from typing import List, Sequence

import numpy as np

from .config import Config


# ---- exact signed distances (numpy, vectorised) --------------------------------------------------------------------
def sdf_circle(x, y, center, radius, margin=0.0):
    return np.hypot(x - center[0], y - center[1]) - (radius + margin)


def sdf_square(x, y, center, size, margin=0.0):
    half = size / 2 + margin
    dx, dy = np.abs(x - center[0]) - half, np.abs(y - center[1]) - half
    return np.hypot(np.maximum(dx, 0), np.maximum(dy, 0)) + np.minimum(np.maximum(dx, dy), 0)


def sdf_polygon(x, y, points: Sequence[Sequence[float]], margin=0.0):
    """Distance to the polygon's boundary, negative inside (even-odd containment), minus the margin."""
    P = np.asarray(points, float)
    x = np.asarray(x, float); y = np.asarray(y, float)
    a = P
    b = np.roll(P, -1, axis=0)
    px, py = x[..., None], y[..., None]
    ex, ey = b[:, 0] - a[:, 0], b[:, 1] - a[:, 1]
    wx, wy = px - a[:, 0], py - a[:, 1]
    t = np.clip((wx * ex + wy * ey) / np.maximum(ex * ex + ey * ey, 1e-300), 0.0, 1.0)
    d = np.sqrt(((wx - t * ex) ** 2 + (wy - t * ey) ** 2).min(axis=-1))
    cond = (a[:, 1] > py) != (b[:, 1] > py)
    with np.errstate(divide="ignore", invalid="ignore"):
        xint = a[:, 0] + (py - a[:, 1]) * ex / np.where(ey == 0, 1.0, ey)
    inside = (np.sum(cond & (px < xint), axis=-1) % 2) == 1
    return np.where(inside, -d, d) - margin


def elliptic_ring_points(center, semi_axes, width, angle=np.pi, num_arc_points=15, rotation=0.0):
    """core/sdf/casadi.py:218-246: outer arc 0..angle, inner arc back, rotated and translated."""
    oa, ob = semi_axes
    ia, ib = oa - width, ob - width
    if ia <= 0 or ib <= 0:
        raise ValueError("Width too large for given semi-axes.")
    t = np.linspace(0.0, angle, num_arc_points)
    pts = [(oa * np.cos(ti), ob * np.sin(ti)) for ti in t] + [(ia * np.cos(ti), ib * np.sin(ti)) for ti in t[::-1]]
    c, s = np.cos(rotation), np.sin(rotation)
    return [(center[0] + px * c - py * s, center[1] + px * s + py * c) for px, py in pts]


def scene_sdf(cfg: Config):
    """Exact SDF of the YAML's obstacle list (MultiObstacle.sdf, core/sdf/casadi.py:381-383)."""
    fns = []
    for ob in cfg.obstacles:
        p = ob.params
        if ob.type == "circle":
            fns.append(lambda x, y, p=p: sdf_circle(x, y, p["center"], p["radius"], p.get("margin", 0.0)))
        elif ob.type == "square":
            fns.append(lambda x, y, p=p: sdf_square(x, y, p["center"], p["size"], p.get("margin", 0.0)))
        elif ob.type in ("polygon", "trapezoid"):
            fns.append(lambda x, y, p=p: sdf_polygon(x, y, p["points"], p.get("margin", 0.0)))
        elif ob.type == "elliptical_ring":
            pts = elliptic_ring_points(p["center"], p["semi_axes"], p["width"], p.get("angle", np.pi), p.get("num_arc_points", 15),
                                       p.get("rotation", 0.0))
            fns.append(lambda x, y, pts=pts, p=p: sdf_polygon(x, y, pts, p.get("margin", 0.0)))
        else:
            raise NotImplementedError(f"exact SDF of obstacle type {ob.type!r}")
    return lambda x, y: np.min(np.stack([f(x, y) for f in fns], axis=0), axis=0)


def sample_points(sdf, x_range, y_range, n_samples, margin=0.1, boundary_fraction=0.3, rng=None):
    """core/sdf/l4casadi.py:14-66 with a seeded generator."""
    rng = rng or np.random.default_rng(0)
    n_boundary = int(n_samples * boundary_fraction)
    xs = rng.uniform(*x_range, size=n_samples - n_boundary)
    ys = rng.uniform(*y_range, size=n_samples - n_boundary)
    bx: List[np.ndarray] = []; by: List[np.ndarray] = []
    got, tries = 0, 0
    while got < n_boundary and tries < 10 * max(n_boundary, 1):
        cx, cy = rng.uniform(*x_range, size=n_boundary), rng.uniform(*y_range, size=n_boundary)
        m = np.abs(sdf(cx, cy)) < margin
        bx.append(cx[m]); by.append(cy[m]); got += int(m.sum()); tries += 1
    if n_boundary:
        xs = np.concatenate([xs, np.concatenate(bx)[:n_boundary]]); ys = np.concatenate([ys, np.concatenate(by)[:n_boundary]])
    return xs, ys


# ---- model + training ------------------------------------------------------------------------------------------
def build_model(cfg: Config):
    import torch.nn as nn
    m = cfg.model
    if m.type != "mlp":
        raise NotImplementedError("train.py builds the benchmarks' model type (mlp); fourier / siren weights come from the reference's "
                                  "state_dict through SdfWeights.from_state_dict")
    act = {"relu": nn.ReLU, "tanh": nn.Tanh, "sigmoid": nn.Sigmoid, "leaky_relu": nn.LeakyReLU}[m.activation_function.lower()]

    class NaiveMLP(nn.Module):           # l4c.naive.MultiLayerPerceptron(2, H, 1, L, act): scripts/run_benchmark.py:65
        def __init__(self, H, L):
            super().__init__()
            self.input_layer = nn.Linear(2, H)
            self.hidden_layers = nn.ModuleList([nn.Linear(H, H) for _ in range(L - 1)])
            self.output_layer = nn.Linear(H, 1)
            self.act = act()

        def forward(self, x):
            x = self.act(self.input_layer(x))
            for layer in self.hidden_layers:
                x = self.act(layer(x))
            return self.output_layer(x)
    net = NaiveMLP(m.hidden_dim, m.num_hidden_layers)
    for layer in net.modules():          # core/sdf/l4casadi.py:69-74
        if isinstance(layer, nn.Linear):
            nn.init.kaiming_uniform_(layer.weight, nonlinearity="relu")
            nn.init.zeros_(layer.bias)
    return net


def train(cfg: Config, n_samples=200_000, epochs=100, batch_size=256, lr=1e-3, seed=0, x_range=(-0.5, 1.5), y_range=(-0.5, 1.5),
          patience=10, min_delta=1e-4, device=None, verbose=True):
    import torch
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    sdf = scene_sdf(cfg)
    m = cfg.model
    # the reference passes boundary_fraction positionally into `margin` (core/sdf/l4casadi.py:118): band = boundary_fraction, share 0.3
    xs, ys = sample_points(sdf, x_range, y_range, n_samples, margin=getattr(m, "boundary_fraction", 0.3), boundary_fraction=0.3, rng=rng)
    X = torch.tensor(np.stack([xs, ys], 1), dtype=torch.float32)
    Y = torch.tensor(sdf(xs, ys), dtype=torch.float32)[:, None]
    perm = torch.randperm(len(X))
    X, Y = X[perm].to(dev), Y[perm].to(dev)
    n_val = int(0.1 * len(X))
    Xv, Yv, Xt, Yt = X[:n_val], Y[:n_val], X[n_val:], Y[n_val:]
    net = build_model(cfg).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=lr)
    sw, ew = float(m.surface_loss_weight), float(m.eikonal_loss_weight)

    def loss_of(xb, yb, with_eik):
        xb = xb.requires_grad_(with_eik and ew > 0)
        pred = net(xb)
        loss = torch.mean((pred - yb) ** 2)
        if sw > 0:
            msk = yb.abs() < 1e-2
            if bool(msk.any()):
                loss = loss + sw * torch.mean(pred[msk] ** 2)
        if with_eik and ew > 0:
            g = torch.autograd.grad(pred, xb, torch.ones_like(pred), create_graph=True)[0]
            loss = loss + ew * torch.mean((torch.linalg.norm(g, dim=1) - 1.0) ** 2)
        return loss
    best, best_state, stale = float("inf"), copy.deepcopy(net.state_dict()), 0
    for ep in range(epochs):
        net.train()
        order = torch.randperm(len(Xt), device=dev)
        for i in range(0, len(Xt), batch_size):
            idx = order[i:i + batch_size]
            loss = loss_of(Xt[idx], Yt[idx], True)
            opt.zero_grad(); loss.backward(); opt.step()
        net.eval()
        with torch.no_grad():
            val = float(loss_of(Xv, Yv, False))
        if verbose and ep % 10 == 0:
            print(f"Epoch {ep:3d} - Val loss: {val:.6f}", flush=True)
        if val + min_delta < best:
            best, best_state, stale = val, copy.deepcopy(net.state_dict()), 0
        else:
            stale += 1
            if stale >= patience:
                if verbose:
                    print(f"Early stopping at epoch {ep:3d} (no improvement for {patience} epochs).", flush=True)
                break
    net.load_state_dict(best_state)
    net.eval()
    with torch.no_grad():
        mse = float(torch.mean((net(Xv) - Yv) ** 2))
    return net, {"val_mse": mse, "val_loss": best, "epochs": ep + 1, "samples": len(X)}


def main():
    from .sdf import SdfWeights
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True)
    ap.add_argument("--out", required=True, help=".nlow or .npz")
    ap.add_argument("--samples", type=int, default=None, help="default: the YAML's model.n_samples (scripts/run_benchmark.py:88-96)")
    ap.add_argument("--epochs", type=int, default=100)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--batch-size", type=int, default=256, help="the reference's value (core/sdf/l4casadi.py:89)")
    a = ap.parse_args()
    cfg = Config.load(Path(a.config))
    net, info = train(cfg, n_samples=a.samples if a.samples is not None else cfg.model.n_samples, epochs=a.epochs, seed=a.seed, batch_size=a.batch_size)
    w = SdfWeights.from_state_dict(cfg.model.type, net.state_dict(), activation_function=cfg.model.activation_function,
                                   omega_0=cfg.model.omega_0)
    (w.save_nlow if a.out.endswith(".nlow") else w.save_npz)(a.out)
    print(f"wrote {a.out}: {info}")


if __name__ == "__main__":
    main()
