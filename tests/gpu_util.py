"""Helpers shared by the GPU parity tests."""
import numpy as np

from oracle import sdf_oracle as so


def to_weights(net: so.SdfNet):
    from nlotrajectories_b200.sdf import SdfWeights
    return SdfWeights.pack(net.kind, net.W0, net.b0, net.hidden, net.w_out, net.b_out, net.act0, net.act, net.p0, net.p)


def kink_mask(net: so.SdfNet, P: np.ndarray, thr: float = 4e-6) -> np.ndarray:
    """Points within fp32 rounding of a ReLU / leaky-ReLU kink (measure-zero ties):
    their Jacobian is discontinuous, so they are excluded from the error statistic and counted."""
    n64 = net.astype(np.float64)
    _, pre = so._forward_all(n64, P.astype(np.float64))
    acts = [net.act0] + [net.act] * len(net.hidden)
    bad = np.zeros(P.shape[0], bool)
    for a, act in zip(pre, acts):
        if act in (so.ACT_RELU, so.ACT_LEAKY_RELU):
            scale = np.maximum(1.0, np.abs(a).max(axis=1, keepdims=True))
            bad |= (np.abs(a) < thr * scale).any(axis=1)
    return bad


def sample_points(n: int, seed: int = 1):
    rng = np.random.default_rng(seed)
    return rng.uniform(-0.5, 1.5, (n, 2)).astype(np.float32)


def sdf_row_ties(spec, net, w, n_g_before_sdf):
    """(P, n_g) mask of SDF constraint rows that touch a kink-adjacent footprint point (their Jacobian
    entries are discontinuous there); used to exclude exactly those entries from the error statistic."""
    from oracle import nlp_oracle as no
    P = w.shape[0]
    X = w[:, :spec.n_X].astype(np.float64).reshape(P, spec.N + 1, spec.nx)
    pts, _ = no._footprint(spec, X)
    nb = pts.shape[2]
    tie = kink_mask(net, pts.reshape(-1, 2).astype(np.float32)).reshape(P, spec.N + 1, nb)
    rows = np.zeros((P, spec.n_g), bool)
    if spec.sdf_rows_per_knot == 1:
        rows[:, n_g_before_sdf:n_g_before_sdf + spec.N + 1] = tie.any(axis=2)
    else:
        rows[:, n_g_before_sdf:n_g_before_sdf + (spec.N + 1) * nb] = tie.reshape(P, -1)
    return rows
