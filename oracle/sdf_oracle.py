"""CPU restatement of the learned signed-distance network (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows, in the reference tree:
* ``_l4c_generated/nn_sdf.cpp:57-104``      - the four external functions (value, jac, adj1, jac_adj1)
* TorchScript graphs inside ``_l4c_generated/*.pt`` (printed in SURVEY.md Appendix C)
* ``src/nlotrajectories/core/nn_architectures.py:8-100`` - FourierMLP / SIREN layer zoo
* ``src/nlotrajectories/scripts/run_benchmark.py:64-83`` - how the three model types are built
* l4casadi ``naive.MultiLayerPerceptron(2, H, 1, L, act)`` (third-party, unpinned; README.md:24):
  ``input_layer`` 2->H, ``L-1`` hidden H->H layers, ``output_layer`` H->1, activation after
  every layer but the last.

Unified network form used by oracle and CUDA kernels alike::

    a_0 = W0 p + b0          h_0 = phi0(a_0)            W0: (H, 2)
    a_l = W_l h_{l-1} + b_l  h_l = phi (a_l)  l=1..M    W_l: (H, H)   M = number of hidden matrices
    s   = w_out . h_M + b_out

    mlp     : phi0 = phi = act                       (act in RELU/TANH/SIGMOID/LEAKY_RELU)
    fourier : phi0 = scale*cos(a), phi = act         (nn_architectures.py:38, 69-71)
    siren   : phi0 = phi = sin(omega0 * a)           (nn_architectures.py:25-26, 95-99)

numpy only; ``dtype`` selects fp64 (truth) or fp32 (same arithmetic type as the reference's torch path).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np

# activation ids shared with include/nlo_b200.h
ACT_RELU, ACT_TANH, ACT_SIGMOID, ACT_LEAKY_RELU, ACT_SIN, ACT_COS_SCALE, ACT_IDENTITY = 0, 1, 2, 3, 4, 5, 6
LEAKY_SLOPE = 0.01  # torch.nn.functional.leaky_relu default (nn_architectures.py:51)

ACT_BY_NAME = {
    "relu": ACT_RELU, "tanh": ACT_TANH, "sigmoid": ACT_SIGMOID,
    "leaky_relu": ACT_LEAKY_RELU, "leakyrelu": ACT_LEAKY_RELU,
}


def act_id(name: str) -> int:
    key = name.replace("_", "").lower() if name.lower() not in ACT_BY_NAME else name.lower()
    if key not in ACT_BY_NAME:
        raise ValueError(f"Unsupported activation function: {name}")
    return ACT_BY_NAME[key]


@dataclass
class SdfNet:
    kind: str                      # "mlp" | "fourier" | "siren"
    W0: np.ndarray                 # (H, 2)
    b0: np.ndarray                 # (H,)
    hidden: List[Tuple[np.ndarray, np.ndarray]] = field(default_factory=list)  # [(W (H,H), b (H,))]
    w_out: np.ndarray = None       # (H,)
    b_out: float = 0.0
    act0: int = ACT_RELU
    act: int = ACT_RELU
    p0: float = 1.0                # parameter of act0 (fourier scale or omega0)
    p: float = 1.0                 # parameter of act  (omega0 for siren)

    @property
    def H(self) -> int:
        return int(self.W0.shape[0])

    def astype(self, dtype) -> "SdfNet":
        c = lambda a: np.asarray(a, dtype=dtype)
        return SdfNet(self.kind, c(self.W0), c(self.b0), [(c(W), c(b)) for W, b in self.hidden],
                      c(self.w_out), dtype(self.b_out), self.act0, self.act, self.p0, self.p)

    def flops_value_jac(self) -> int:
        """4(3H + M H^2)."""
        H, M = self.H, len(self.hidden)
        return 4 * (3 * H + M * H * H)


# ---------------------------------------------------------------------------------------------
# activations: value, first and second derivative with respect to the pre-activation
# ---------------------------------------------------------------------------------------------
def _phi(a, act, prm):
    if act == ACT_RELU:
        return np.maximum(a, 0)
    if act == ACT_TANH:
        return np.tanh(a)
    if act == ACT_SIGMOID:
        return 1.0 / (1.0 + np.exp(-a))
    if act == ACT_LEAKY_RELU:
        return np.where(a > 0, a, a * a.dtype.type(LEAKY_SLOPE))
    if act == ACT_SIN:
        return np.sin(a.dtype.type(prm) * a)
    if act == ACT_COS_SCALE:
        return np.cos(a) * a.dtype.type(prm)
    if act == ACT_IDENTITY:
        return a
    raise ValueError(act)


def _dphi(a, act, prm):
    one = a.dtype.type(1)
    if act == ACT_RELU:
        # torch threshold_backward(grad, relu_out, 0): passes grad where relu_out > 0 (strict)
        return (a > 0).astype(a.dtype)
    if act == ACT_TANH:
        t = np.tanh(a)
        return one - t * t
    if act == ACT_SIGMOID:
        s = one / (one + np.exp(-a))
        return s * (one - s)
    if act == ACT_LEAKY_RELU:
        return np.where(a > 0, one, a.dtype.type(LEAKY_SLOPE))
    if act == ACT_SIN:
        w = a.dtype.type(prm)
        return w * np.cos(w * a)
    if act == ACT_COS_SCALE:
        return -np.sin(a) * a.dtype.type(prm)
    if act == ACT_IDENTITY:
        return np.ones_like(a)
    raise ValueError(act)


def _d2phi(a, act, prm):
    one = a.dtype.type(1)
    if act in (ACT_RELU, ACT_LEAKY_RELU, ACT_IDENTITY):
        return np.zeros_like(a)
    if act == ACT_TANH:
        t = np.tanh(a)
        return -2 * t * (one - t * t)
    if act == ACT_SIGMOID:
        s = one / (one + np.exp(-a))
        return s * (one - s) * (one - 2 * s)
    if act == ACT_SIN:
        w = a.dtype.type(prm)
        return -(w * w) * np.sin(w * a)
    if act == ACT_COS_SCALE:
        return -np.cos(a) * a.dtype.type(prm)
    raise ValueError(act)


# ---------------------------------------------------------------------------------------------
# the four external functions, batched over points P: (n, 2)
# ---------------------------------------------------------------------------------------------
def _forward_all(net: SdfNet, P: np.ndarray):
    a = P @ net.W0.T + net.b0                      # nn_sdf.pt: mm + add
    pre = [a]
    h = _phi(a, net.act0, net.p0)
    for W, b in net.hidden:
        a = h @ W.T + b                            # addmm
        pre.append(a)
        h = _phi(a, net.act, net.p)
    s = h @ net.w_out + net.b_out                  # addmm_1
    return s, pre


def forward(net: SdfNet, P: np.ndarray) -> np.ndarray:
    """``nn_sdf`` (nn_sdf.cpp:57-60).  Returns s: (n,)."""
    return _forward_all(net, P)[0]


def _reverse(net: SdfNet, pre, seed):
    """g_l = ds/da_l, returned for every layer (index 0 = first layer)."""
    M = len(net.hidden)
    g = [None] * (M + 1)
    top_act, top_prm = (net.act, net.p) if M > 0 else (net.act0, net.p0)
    g[M] = (seed[:, None] * net.w_out[None, :]) * _dphi(pre[M], top_act, top_prm)
    for l in range(M, 0, -1):
        W = net.hidden[l - 1][0]
        a_prev = pre[l - 1]
        act, prm = (net.act, net.p) if l - 1 > 0 else (net.act0, net.p0)
        g[l - 1] = (g[l] @ W) * _dphi(a_prev, act, prm)
    return g


def value_jac(net: SdfNet, P: np.ndarray):
    """``nn_sdf`` + ``jac_nn_sdf`` (nn_sdf.cpp:57-70).  Returns (s (n,), J (n,2))."""
    s, pre = _forward_all(net, P)
    g = _reverse(net, pre, np.ones(P.shape[0], dtype=P.dtype))
    return s, g[0] @ net.W0


def adj1(net: SdfNet, P: np.ndarray, sbar: np.ndarray) -> np.ndarray:
    """``adj1_nn_sdf`` (nn_sdf.cpp:79-83): sbar * ds/dp.  Returns (n,2)."""
    _, pre = _forward_all(net, P)
    g = _reverse(net, pre, sbar.astype(P.dtype))
    return g[0] @ net.W0


def jac_adj1(net: SdfNet, P: np.ndarray, sbar: np.ndarray) -> np.ndarray:
    """``jac_adj1_nn_sdf`` (nn_sdf.cpp:91-104): d(adj1)/dp = sbar * Hessian(s).  Returns (n,2,2).

    Forward-over-reverse, the same structure as the traced graph in ``jac_adj1_nn_sdf.pt``.
    """
    n = P.shape[0]
    M = len(net.hidden)
    _, pre = _forward_all(net, P)
    g = _reverse(net, pre, sbar.astype(P.dtype))
    acts = [(net.act0, net.p0)] + [(net.act, net.p)] * M
    Hs = np.zeros((n, 2, 2), dtype=P.dtype)
    for d in range(2):
        # tangent of the forward pass in direction e_d
        adot = [np.broadcast_to(net.W0[:, d], (n, net.H)).astype(P.dtype)]
        hdot = _dphi(pre[0], *acts[0]) * adot[0]
        for l in range(1, M + 1):
            adot.append(hdot @ net.hidden[l - 1][0].T)
            hdot = _dphi(pre[l], *acts[l]) * adot[l]
        # tangent of the reverse pass
        gdot = (sbar[:, None].astype(P.dtype) * net.w_out[None, :]) * _d2phi(pre[M], *acts[M]) * adot[M]
        for l in range(M, 0, -1):
            W = net.hidden[l - 1][0]
            back = g[l] @ W
            gdot = (gdot @ W) * _dphi(pre[l - 1], *acts[l - 1]) + back * _d2phi(pre[l - 1], *acts[l - 1]) * adot[l - 1]
        Hs[:, :, d] = gdot @ net.W0
    return Hs


# ---------------------------------------------------------------------------------------------
# builders
# ---------------------------------------------------------------------------------------------
def synthetic_mlp(H: int, n_hidden_mats: int = 1, seed: int = 0, act: int = ACT_RELU, dtype=np.float32) -> SdfNet:
    """SURVEY.md section 8(d) synthetic sweep nets: W ~ N(0, 1/fan_in), small random bias.

    (Biases are drawn N(0, 0.1^2) rather than 0 so that bias handling is exercised by parity tests.)
    """
    rng = np.random.default_rng(seed)
    W0 = rng.standard_normal((H, 2)) / np.sqrt(2.0)
    b0 = 0.1 * rng.standard_normal(H)
    hidden = [(rng.standard_normal((H, H)) / np.sqrt(H), 0.1 * rng.standard_normal(H)) for _ in range(n_hidden_mats)]
    w_out = rng.standard_normal(H) / np.sqrt(H)
    b_out = 0.05
    net = SdfNet("mlp", W0, b0, hidden, w_out, b_out, act, act, 1.0, 1.0)
    return net.astype(dtype)


def synthetic_siren(H: int, n_hidden_mats: int = 1, omega0: float = 30.0, seed: int = 0, dtype=np.float32) -> SdfNet:
    """SIREN with the init of nn_architectures.py:16-23 (first layer U(-1/in,1/in), rest U(+-sqrt(6/in)/omega0))."""
    rng = np.random.default_rng(seed)
    W0 = rng.uniform(-0.5, 0.5, (H, 2))
    b0 = rng.uniform(-0.5, 0.5, H) / np.sqrt(2.0)
    bound = np.sqrt(6.0 / H) / omega0
    hidden = [(rng.uniform(-bound, bound, (H, H)), rng.uniform(-1, 1, H) / np.sqrt(H)) for _ in range(n_hidden_mats)]
    w_out = rng.uniform(-1, 1, H) / np.sqrt(H)
    net = SdfNet("siren", W0, b0, hidden, w_out, 0.0, ACT_SIN, ACT_SIN, omega0, omega0)
    return net.astype(dtype)


def synthetic_fourier(H: int, n_hidden_mats: int = 1, scale: float = 1.0, seed: int = 0, act: int = ACT_RELU,
                      dtype=np.float32) -> SdfNet:
    rng = np.random.default_rng(seed)
    W0 = rng.standard_normal((H, 2)) * scale            # nn_architectures.py:34 (stored in x out there)
    b0 = 0.1 * rng.standard_normal(H)
    hidden = [(rng.standard_normal((H, H)) / np.sqrt(H), 0.1 * rng.standard_normal(H)) for _ in range(n_hidden_mats)]
    w_out = rng.standard_normal(H) / np.sqrt(H)
    net = SdfNet("fourier", W0, b0, hidden, w_out, 0.02, ACT_COS_SCALE, act, scale, 1.0)
    return net.astype(dtype)


def from_npz(path) -> SdfNet:
    z = np.load(path, allow_pickle=False)
    M = int(z["n_hidden_mats"])
    hidden = [(z[f"W{l + 1}"], z[f"b{l + 1}"]) for l in range(M)]
    return SdfNet(str(z["kind"]), z["W0"], z["b0"], hidden, z["w_out"], z["b_out"].item(),
                  int(z["act0"]), int(z["act"]), float(z["p0"]), float(z["p"]))


def to_npz(net: SdfNet, path) -> None:
    d = dict(kind=np.array(net.kind), W0=net.W0, b0=net.b0, w_out=net.w_out, b_out=np.asarray(net.b_out),
             act0=np.int32(net.act0), act=np.int32(net.act), p0=np.float64(net.p0), p=np.float64(net.p),
             n_hidden_mats=np.int32(len(net.hidden)))
    for l, (W, b) in enumerate(net.hidden):
        d[f"W{l + 1}"] = W
        d[f"b{l + 1}"] = b
    np.savez(path, **d)
