"""Learned signed-distance network on the GPU: weight import and the evaluation handle.

Host-side mirror of the reference's learned-SDF objects:
* ``SdfWeights.from_module``       <- the torch models of core/nn_architectures.py:42-100 and l4casadi's
                                      naive MLP (scripts/run_benchmark.py:64-83)
* ``SdfWeights.from_torchscript``  <- the traced artefacts ``_l4c_generated/nn_sdf.pt``
* ``LearnedSDF``                   <- ``l4c.L4CasADi(model, device="cpu")`` (scripts/run_benchmark.py:100)
* ``NNObstacle``                   <- core/sdf/l4casadi.py:231-260 (numpy branch :242-246)

PyTorch is used only to read ``.pt`` files and to hold device memory; all arithmetic runs in
libnlo_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from pathlib import Path
from typing import Optional

import numpy as np

from . import lib as _lib

ACT_RELU, ACT_TANH, ACT_SIGMOID, ACT_LEAKY_RELU, ACT_SIN, ACT_COS_SCALE, ACT_IDENTITY = range(7)
KIND = {"mlp": 0, "fourier": 1, "siren": 2}
_ACT_NAMES = {"relu": ACT_RELU, "tanh": ACT_TANH, "sigmoid": ACT_SIGMOID, "leakyrelu": ACT_LEAKY_RELU}


def activation_id(name: str) -> int:
    key = name.replace("_", "").lower()
    if key not in _ACT_NAMES:
        # same failure the reference raises (core/nn_architectures.py:53-54)
        raise ValueError(f"Unsupported activation function: {name}")
    return _ACT_NAMES[key]


@dataclass
class SdfWeights:
    kind: str
    hidden: int
    n_hidden_mats: int
    act0: int
    act: int
    p0: float
    p: float
    blob: np.ndarray            # flat fp32, order documented in include/nlo_b200.h

    def desc(self) -> _lib.SdfDesc:
        return _lib.SdfDesc(KIND[self.kind], self.hidden, self.n_hidden_mats, self.act0, self.act, self.p0, self.p)

    @staticmethod
    def pack(kind, W0, b0, hidden, w_out, b_out, act0, act, p0=1.0, p=1.0) -> "SdfWeights":
        """W0 (H,2), b0 (H,), hidden [(W (H,H) rows = output neurons, b (H,))], w_out (H,), b_out scalar."""
        H = int(np.shape(W0)[0])
        parts = [np.asarray(W0, np.float32).reshape(H, 2).ravel(), np.asarray(b0, np.float32).reshape(H)]
        for W, b in hidden:
            parts += [np.asarray(W, np.float32).reshape(H, H).ravel(), np.asarray(b, np.float32).reshape(H)]
        parts += [np.asarray(w_out, np.float32).reshape(H), np.asarray([b_out], np.float32).reshape(1)]
        return SdfWeights(kind, H, len(hidden), int(act0), int(act), float(p0), float(p),
                          np.ascontiguousarray(np.concatenate(parts)))

    # ---- importers -------------------------------------------------------------------------------
    @staticmethod
    def from_npz(path) -> "SdfWeights":
        z = np.load(path, allow_pickle=False)
        M = int(z["n_hidden_mats"])
        return SdfWeights.pack(str(z["kind"]), z["W0"], z["b0"], [(z[f"W{l+1}"], z[f"b{l+1}"]) for l in range(M)],
                               z["w_out"], float(z["b_out"]), int(z["act0"]), int(z["act"]), float(z["p0"]), float(z["p"]))

    def save_npz(self, path) -> None:
        """The layout ``from_npz`` reads (also what oracle/sdf_oracle.py's to_npz / from_npz use)."""
        H, M = self.hidden, self.n_hidden_mats
        b = self.blob
        d = {"kind": np.array(self.kind), "n_hidden_mats": np.int64(M), "act0": np.int64(self.act0), "act": np.int64(self.act),
             "p0": np.float64(self.p0), "p": np.float64(self.p), "W0": b[:2 * H].reshape(H, 2), "b0": b[2 * H:3 * H]}
        o = 3 * H
        for l in range(M):
            d[f"W{l + 1}"] = b[o:o + H * H].reshape(H, H); d[f"b{l + 1}"] = b[o + H * H:o + H * H + H]; o += H * H + H
        d["w_out"] = b[o:o + H]; d["b_out"] = np.float64(b[o + H])
        np.savez(path, **d)

    @staticmethod
    def from_torchscript(path) -> "SdfWeights":
        """Traced FourierMLP artefact written by l4casadi (``nn_sdf.pt``): constants c0..c5 with
        h0 = cos(p @ c0 + c1) * scale ; h1 = relu(h0 @ c3 + c2) ; s = h1 @ c5 + c4."""
        import re
        import torch
        m = torch.jit.load(str(path), map_location="cpu")
        code, consts = m.code_with_constants
        c = {k: v.detach().float().numpy() for k, v in consts.const_mapping.items()}
        mt = re.search(r"torch\.mul\(cos, ([-0-9.e]+)\)", code)
        if mt is None or not {"c0", "c1", "c2", "c3", "c4", "c5"} <= set(c):
            raise ValueError(f"{path}: not a traced 2-layer FourierMLP graph")
        return SdfWeights.pack("fourier", c["c0"].T, c["c1"], [(c["c3"].T, c["c2"])], c["c5"][:, 0], float(c["c4"][0]),
                               ACT_COS_SCALE, ACT_RELU, float(mt.group(1)), 1.0)

    @staticmethod
    def from_state_dict(kind: str, sd: dict, activation_function: str = "ReLU", omega_0: float = 30.0,
                        scale: float = 1.0) -> "SdfWeights":
        """``state_dict`` of FourierMLP / SIREN (core/nn_architectures.py) or of l4casadi's naive MLP
        (``input_layer``, ``hidden_layers.i``, ``output_layer``)."""
        g = lambda k: np.asarray(sd[k].detach().cpu().float().numpy() if hasattr(sd[k], "detach") else sd[k], np.float32)
        if kind == "mlp":
            act = activation_id(activation_function)
            n = len({k.split(".")[1] for k in sd if k.startswith("hidden_layers.")})
            hidden = [(g(f"hidden_layers.{i}.weight"), g(f"hidden_layers.{i}.bias")) for i in range(n)]
            return SdfWeights.pack("mlp", g("input_layer.weight"), g("input_layer.bias"), hidden,
                                   g("output_layer.weight")[0], float(g("output_layer.bias")[0]), act, act)
        if kind == "fourier":
            act = activation_id(activation_function)
            n = len({k.split(".")[1] for k in sd if k.startswith("layers.")})
            hidden = [(g(f"layers.{i}.weight"), g(f"layers.{i}.bias")) for i in range(n)]
            return SdfWeights.pack("fourier", g("fourier.weights").T, g("fourier.bias"), hidden,
                                   g("output_layer.weight")[0], float(g("output_layer.bias")[0]), ACT_COS_SCALE, act, scale)
        if kind == "siren":
            n = len({k.split(".")[1] for k in sd if k.startswith("layers.")})
            hidden = [(g(f"layers.{i}.linear.weight"), g(f"layers.{i}.linear.bias")) for i in range(1, n)]
            return SdfWeights.pack("siren", g("layers.0.linear.weight"), g("layers.0.linear.bias"), hidden,
                                   g("output_layer.weight")[0], float(g("output_layer.bias")[0]), ACT_SIN, ACT_SIN,
                                   omega_0, omega_0)
        raise ValueError(f"Unsupported model type: {kind}")      # scripts/run_benchmark.py:83

    @staticmethod
    def load(path, **kw) -> "SdfWeights":
        """Dispatch on file type: .nlow (library format), .npz, TorchScript .pt, or a pickled state_dict .pt."""
        path = Path(path)
        if path.suffix == ".nlow":
            return SdfWeights.from_nlow(path)
        if path.suffix == ".npz":
            return SdfWeights.from_npz(path)
        import torch
        try:
            return SdfWeights.from_torchscript(path)
        except (RuntimeError, ValueError):
            sd = torch.load(str(path), map_location="cpu")
            return SdfWeights.from_state_dict(kw.pop("kind", "mlp"), sd, **kw)

    # ---- library file format ------------------------------------------------------------------------
    def save_nlow(self, path) -> None:
        L = _lib.load()
        d = self.desc()
        _lib.check(L.nlo_sdf_save(str(path).encode(), C.byref(d), self.blob.ctypes.data, self.blob.size))

    @staticmethod
    def from_nlow(path) -> "SdfWeights":
        raw = Path(path).read_bytes()
        if raw[:4] != b"NLOW":
            raise ValueError(f"{path} is not a .nlow file")
        d = _lib.SdfDesc.from_buffer_copy(raw[8:8 + C.sizeof(_lib.SdfDesc)])
        n = int(np.frombuffer(raw[40:48], np.uint64)[0])
        blob = np.frombuffer(raw[64:64 + 4 * n], np.float32).copy()
        kind = {v: k for k, v in KIND.items()}[d.kind]
        return SdfWeights(kind, d.hidden, d.n_hidden_mats, d.act0, d.act, d.p0, d.p, blob)


class LearnedSDF:
    """Device-resident network + evaluation entry points (value / Jacobian / adjoint / Hessian)."""

    def __init__(self, weights: SdfWeights, device: int = 0, precision: Optional[str] = None):
        self._L = _lib.load()
        _lib.require_gpu()
        self.weights = weights
        self.device = device
        h = C.c_void_p()
        d = weights.desc()
        _lib.check(self._L.nlo_sdf_create(C.byref(d), weights.blob.ctypes.data, weights.blob.size, device, C.byref(h)))
        self._h = h
        if precision is not None:
            self.set_precision(precision)

    def set_precision(self, precision: str) -> None:
        code = {"fp32": _lib.PREC_FP32_SIMT, "tc3xf16": _lib.PREC_TC_3XF16, "auto": _lib.PREC_AUTO}[precision]
        _lib.check(self._L.nlo_sdf_set_precision(self._h, code))

    @property
    def precision(self) -> str:
        return {0: "fp32", 1: "tc3xf16"}[self._L.nlo_sdf_get_precision(self._h)]

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None):
            self._L.nlo_sdf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- device tensors (torch, fp32, contiguous, on self.device) -------------------------------------
    def eval(self, x, y, sbar=None, want_jac: bool = True, out=None, stream=None):
        """Returns (s, jx, jy) torch tensors; jx, jy are None when ``want_jac`` is False."""
        import torch
        n = x.numel()
        assert x.is_cuda and y.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and y.is_contiguous()
        if out is None:
            s = torch.empty_like(x)
            jx = torch.empty_like(x) if want_jac else None
            jy = torch.empty_like(x) if want_jac else None
        else:
            s, jx, jy = out
        st = torch.cuda.current_stream(x.device).cuda_stream if stream is None else stream
        _lib.check(self._L.nlo_sdf_eval(self._h, x.data_ptr(), y.data_ptr(), _lib.ptr(sbar), n,
                                        _lib.ptr(s), _lib.ptr(jx), _lib.ptr(jy), st))
        return s, jx, jy

    def hess(self, x, y, sbar=None):
        import torch
        n = x.numel()
        hxx, hxy, hyy = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
        st = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(self._L.nlo_sdf_hess(self._h, x.data_ptr(), y.data_ptr(), _lib.ptr(sbar), n,
                                        hxx.data_ptr(), hxy.data_ptr(), hyy.data_ptr(), st))
        return hxx, hxy, hyy

    # ---- host buffers (numpy fp32) -----------------------------------------------------------------
    def eval_host(self, x: np.ndarray, y: np.ndarray, sbar: Optional[np.ndarray] = None, want_jac: bool = True, out=None):
        x = np.ascontiguousarray(x, np.float32); y = np.ascontiguousarray(y, np.float32)
        if sbar is not None:
            sbar = np.ascontiguousarray(sbar, np.float32)
        if out is None:
            s = np.empty_like(x)
            jx = np.empty_like(x) if want_jac else None
            jy = np.empty_like(x) if want_jac else None
        else:
            s, jx, jy = out
        _lib.check(self._L.nlo_sdf_eval_host(self._h, x.ctypes.data, y.ctypes.data, _lib.ptr(sbar), x.size,
                                             _lib.ptr(s), _lib.ptr(jx), _lib.ptr(jy)))
        return s, jx, jy

    def hess_host(self, x: np.ndarray, y: np.ndarray, sbar: Optional[np.ndarray] = None):
        x = np.ascontiguousarray(x, np.float32); y = np.ascontiguousarray(y, np.float32)
        if sbar is not None:
            sbar = np.ascontiguousarray(sbar, np.float32)
        hxx, hxy, hyy = np.empty_like(x), np.empty_like(x), np.empty_like(x)
        _lib.check(self._L.nlo_sdf_hess_host(self._h, x.ctypes.data, y.ctypes.data, _lib.ptr(sbar), x.size,
                                             hxx.ctypes.data, hxy.ctypes.data, hyy.ctypes.data))
        return hxx, hxy, hyy

    def bind_casadi(self, batch: int = 1) -> Path:
        """Make this model the one the CasADi externals ``nn_sdf`` ... evaluate; returns the library path to
        hand to ``casadi.external("nn_sdf", path)``."""
        _lib.check(self._L.nlo_casadi_set_batch(batch))
        _lib.check(self._L.nlo_casadi_bind(self._h))
        return _lib.LIB_PATH


class NNObstacle:
    """Mirror of core/sdf/l4casadi.py:231-260 for numeric inputs: same method names and shapes."""

    def __init__(self, obstacle, model: LearnedSDF):
        self.obstacle = obstacle
        self.model = model

    def sdf(self, x, y):
        return self.obstacle.sdf(x, y)

    def approximated_sdf(self, x, y):
        if isinstance(x, np.ndarray) and isinstance(y, np.ndarray):
            s, _, _ = self.model.eval_host(x.ravel(), y.ravel(), want_jac=False)
            return s.reshape(x.shape)
        raise TypeError("Inputs must be both NumPy arrays or both CasADi MX types.")   # l4casadi.py:257
