"""One NLP family (dynamics x footprint x horizon) bound to the CUDA library: sizes, bounds,
sparsity, initial guesses, and batched evaluation.

Host-side mirror of what ``RunBenchmark.run`` assembles symbolically (core/runner.py:44-108):
decision vector ``w = [vec(X); vec(U); slack]``, constraint rows in ``subject_to`` order, bounds in
Opti's canonical form.  The arithmetic runs in libnlo_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import lib as _lib
from .config import Config, DYNAMICS, SHAPES
from .sdf import LearnedSDF

DYN_DIMS = {"point_1st": (4, 2), "point_2nd": (4, 2), "unicycle": (3, 2), "unicycle_2nd": (5, 2),
            "ackermann": (4, 2), "ackermann_2nd": (7, 2)}     # core/dynamics.py state_dim()/control_dim()


class NlpProblem:
    def __init__(self, dynamics: str, shape: str, x0: Sequence[float], x_goal: Sequence[float], N: int, dt: float,
                 control_bounds: Sequence[Tuple[float, float]], use_slack: bool = False, slack_penalty: float = 1000.0,
                 use_smooth: bool = False, smooth_weight: float = 1000.0, enforce_heading: bool = True,
                 length: Optional[float] = None, width: Optional[float] = None, wheelbase: Optional[float] = None,
                 sdf: Optional[LearnedSDF] = None, circles: Sequence[Tuple[float, float, float, float]] = (),
                 device: int = 0):
        if dynamics not in DYNAMICS:
            raise ValueError(f"Unknown dynamics type: {dynamics}")          # core/config.py:49
        if shape not in SHAPES:
            raise ValueError(f"Unknown shape: {shape}")
        self.dynamics, self.shape = dynamics, shape
        self.nx, self.nu = DYN_DIMS[dynamics]
        self.x0 = np.asarray(x0, float); self.x_goal = np.asarray(x_goal, float)
        if self.x0.shape != (self.nx,) or self.x_goal.shape != (self.nx,):
            raise ValueError(f"start/goal state must have {self.nx} entries for {dynamics}")
        if len(control_bounds) < self.nu:
            raise ValueError(f"control_bounds needs {self.nu} (min, max) pairs")
        if shape != "dot" and (length is None or width is None):
            raise ValueError(f"{shape} footprint needs length and width")
        if dynamics.startswith("ackermann") and wheelbase is None:
            raise ValueError("ackermann dynamics need a wheelbase")
        self.N, self.dt = int(N), float(dt)
        self.control_bounds = [tuple(map(float, cb)) for cb in control_bounds]
        self.use_slack, self.use_smooth, self.enforce_heading = bool(use_slack), bool(use_smooth), bool(enforce_heading)
        self.slack_penalty = float(slack_penalty if slack_penalty is not None else 0.0)
        self.smooth_weight = float(smooth_weight)
        self.length, self.width, self.wheelbase = length, width, wheelbase
        self.sdf = sdf
        self.circles = list(circles)
        self.device = device
        self._L = _lib.load()
        _lib.require_gpu()
        d = _lib.NlpDesc()
        d.dynamics = DYNAMICS.index(dynamics); d.shape = SHAPES.index(shape); d.N = self.N
        d.use_slack, d.use_smooth, d.enforce_heading = int(self.use_slack), int(self.use_smooth), int(self.enforce_heading)
        d.sdf_mode = 0 if sdf is not None else 1
        if sdf is None and not self.circles:
            raise ValueError("either a learned SDF or analytic circles are required")
        d.n_circles = len(self.circles) if sdf is None else 0
        d.dt, d.slack_penalty, d.smooth_weight = self.dt, self.slack_penalty, self.smooth_weight
        d.length, d.width, d.wheelbase = float(length or 0.0), float(width or 0.0), float(wheelbase or 1.0)
        if len(self.circles) > _lib.NLO_MAX_CIRCLES:
            raise ValueError(f"at most {_lib.NLO_MAX_CIRCLES} analytic obstacles")
        for i, c in enumerate(self.circles):             # (cx, cy, radius | size, margin[, kind: 0 circle, 1 square])
            for q in range(4):
                d.circles[i][q] = float(c[q])
            d.obstacle_kind[i] = int(c[4]) if len(c) > 4 else 0
        h = C.c_void_p()
        _lib.check(self._L.nlo_nlp_create(C.byref(d), sdf.handle if sdf is not None else None, device, C.byref(h)))
        self._h = h
        self.n_w = int(self._L.nlo_nlp_n_w(h)); self.n_g = int(self._L.nlo_nlp_n_g(h))
        self.nnz = int(self._L.nlo_nlp_nnz_jac(h)); self.n_sdf_points = int(self._L.nlo_nlp_n_sdf_points(h))
        self.n_X = self.nx * (self.N + 1); self.n_U = self.nu * self.N

    # ---- construction from the reference's YAML --------------------------------------------------------
    @staticmethod
    def from_config(cfg: Config, sdf: Optional[LearnedSDF] = None, device: int = 0) -> "NlpProblem":
        b, s = cfg.body, cfg.solver
        learned = s.mode == "l4casadi"
        if learned and sdf is None:
            raise ValueError("solver.mode l4casadi needs a LearnedSDF (weights)")
        return NlpProblem(b.dynamic, b.shape, b.start_state, b.goal_state, s.N, s.dt, b.control_bounds, s.use_slack,
                          s.slack_penalty, s.use_smooth, s.smooth_weight, s.enforce_heading, b.length, b.width, b.wheelbase,
                          sdf=sdf if learned else None, circles=() if learned else cfg.circles(), device=device)

    def close(self):
        if getattr(self, "_h", None):
            self._L.nlo_nlp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- structure -----------------------------------------------------------------------------------
    @property
    def terminal_idx(self) -> List[int]:
        return list(range(self.nx)) if self.enforce_heading else [i for i in range(self.nx) if i != 2]   # runner.py:51-56

    def bounds(self):
        """(lbg, ubg) in Opti canonical form, fp64."""
        n_sdf = self.n_g - (self.nx + len(self.terminal_idx) + self.N * self.nx + ((self.N + 1) if self.use_slack else 0) + self.n_U)
        lb = [*self.x0, *[self.x_goal[i] for i in self.terminal_idx], *([0.0] * (self.N * self.nx))]
        ub = list(lb)
        if self.use_slack:
            lb += [0.0] * (self.N + 1); ub += [np.inf] * (self.N + 1)
        lb += [0.0] * n_sdf; ub += [np.inf] * n_sdf
        for i in range(self.nu):
            lb += [self.control_bounds[i][0]] * self.N; ub += [self.control_bounds[i][1]] * self.N
        return np.array(lb), np.array(ub)

    def jac_sparsity(self):
        """Compressed-column pattern of dg/dw: (colind[n_w+1], row[nnz])."""
        colind = np.zeros(self.n_w + 1, np.int32); row = np.zeros(self.nnz, np.int32)
        _lib.check(self._L.nlo_nlp_jac_sparsity(self._h, colind.ctypes.data_as(C.POINTER(C.c_int32)),
                                                row.ctypes.data_as(C.POINTER(C.c_int32))))
        return colind, row

    @property
    def nnz_hess(self) -> int:
        return int(self._L.nlo_nlp_nnz_hess(self._h))

    def hess_sparsity(self):
        """Compressed-column pattern of the upper triangle of the Hessian of the Lagrangian: (colind[n_w+1], row[nnz_hess])."""
        colind = np.zeros(self.n_w + 1, np.int32); row = np.zeros(self.nnz_hess, np.int32)
        _lib.check(self._L.nlo_nlp_hess_sparsity(self._h, colind.ctypes.data_as(C.POINTER(C.c_int32)),
                                                 row.ctypes.data_as(C.POINTER(C.c_int32))))
        return colind, row

    def eval_hess_device(self, w, lam, sigma=None, out=None, P: Optional[int] = None, stream=None):
        """sigma * hess f + sum_r lam_r hess g_r for every problem (what IPOPT's eval_h needs; core/runner.py:113-125 keeps
        the exact Hessian).  w (n_w, ld), lam (n_g, ld), sigma (P,) or None (= 1) are torch fp32 CUDA tensors; returns the
        values (nnz_hess, ld) on ``hess_sparsity()``."""
        import torch
        ld = w.shape[1]
        P = ld if P is None else P
        if out is None:
            out = torch.empty((self.nnz_hess, ld), dtype=torch.float32, device=w.device)
        st = torch.cuda.current_stream(w.device).cuda_stream if stream is None else stream
        _lib.check(self._L.nlo_nlp_hess(self._h, w.data_ptr(), _lib.ptr(sigma), lam.data_ptr(), P, ld, out.data_ptr(), st))
        return out

    # ---- initial guesses ---------------------------------------------------------------------------------
    def linear_guess(self) -> np.ndarray:
        """``LinearInitializer`` (core/trajectory_initialization.py:54-55) packed as w (U = slack = 0)."""
        w = np.zeros(self.n_w)
        w[:self.n_X] = np.linspace(self.x0, self.x_goal, self.N + 1).reshape(-1)
        return w

    def multistart_guess(self, P: int, first: int = 0, seed0: int = 1234) -> np.ndarray:
        """Seeded multi-start family; problem i of the global batch uses default_rng(seed0+i).
        Returns problem-major (P, n_w) fp32 for global indices first..first+P-1."""
        t = np.linspace(0.0, 1.0, self.N + 1)
        base = self.x0[None, :] + t[:, None] * (self.x_goal - self.x0)[None, :]
        d = self.x_goal[:2] - self.x0[:2]
        nrm = np.array([-d[1], d[0]]) / (np.linalg.norm(d) + 1e-12)
        bump = np.sin(np.pi * t)[:, None] * nrm[None, :]
        w = np.zeros((P, self.n_w), np.float32)
        X = np.empty((self.N + 1, self.nx))
        for i in range(P):
            rng = np.random.default_rng(seed0 + first + i)
            amp = rng.uniform(-0.4, 0.4)
            X[:] = base
            X[:, :2] += amp * bump + rng.normal(0.0, 0.01, (self.N + 1, 2))
            w[i, :self.n_X] = X.reshape(-1)
        return w

    # ---- evaluation -----------------------------------------------------------------------------------------
    def eval_device(self, w, g=None, jac=None, f=None, grad_f=None, P: Optional[int] = None, stream=None):
        """Device-resident SoA evaluation: w is a torch fp32 CUDA tensor of shape (n_w, ld); outputs
        (n_g, ld), (nnz, ld), (P,), (n_w, ld) or None.  Asynchronous on the current torch stream."""
        import torch
        ld = w.shape[1]
        P = ld if P is None else P
        st = torch.cuda.current_stream(w.device).cuda_stream if stream is None else stream
        _lib.check(self._L.nlo_nlp_eval(self._h, w.data_ptr(), P, ld, _lib.ptr(g), _lib.ptr(jac), _lib.ptr(f), _lib.ptr(grad_f), st))

    def eval_dynamics_device(self, w, g=None, jac=None, P: Optional[int] = None, stream=None):
        """K2 alone: the Euler defect rows of g and their Jacobian values (same SoA layouts as ``eval_device``)."""
        import torch
        ld = w.shape[1]
        P = ld if P is None else P
        st = torch.cuda.current_stream(w.device).cuda_stream if stream is None else stream
        _lib.check(self._L.nlo_nlp_eval_dynamics(self._h, w.data_ptr(), P, ld, _lib.ptr(g), _lib.ptr(jac), st))

    def reserve(self, P: int) -> None:
        """Size the handle's device scratch for P problems up front (later evaluations with <= P problems allocate nothing)."""
        _lib.check(self._L.nlo_nlp_reserve(self._h, P))

    def alloc_outputs(self, P: int, device=None):
        import torch
        dev = torch.device("cuda", self.device) if device is None else device
        mk = lambda r: torch.empty((r, P), dtype=torch.float32, device=dev)
        return mk(self.n_g), mk(self.nnz), torch.empty(P, dtype=torch.float32, device=dev), mk(self.n_w)

    def eval_host(self, w: np.ndarray, want=("g", "jac", "f", "grad_f"), out=None):
        """Host problem-major evaluation: w (P, n_w) fp32 -> dict of (P, n_g), (P, nnz), (P,), (P, n_w)."""
        w = np.ascontiguousarray(w, np.float32)
        P = w.shape[0]
        res = out if out is not None else {}
        shapes = {"g": (P, self.n_g), "jac": (P, self.nnz), "f": (P,), "grad_f": (P, self.n_w)}
        for k in want:
            if k not in res:
                res[k] = np.empty(shapes[k], np.float32)
        _lib.check(self._L.nlo_nlp_eval_host(self._h, w.ctypes.data, P, _lib.ptr(res.get("g")), _lib.ptr(res.get("jac")),
                                             _lib.ptr(res.get("f")), _lib.ptr(res.get("grad_f"))))
        return res

    # ---- compact host form: only what varies with w travels back over PCIe ----------------------------------------
    def compact_layout(self) -> dict:
        """Index lists of ``nlo_nlp_compact_layout``: which rows of g / non-zeros of dg/dw / entries of grad f vary with w
        (``g_var_rows``, ``jac_var_nz``, ``grad_var_idx``) and how everything else follows (``g_copy_rows/vars``,
        ``jac_const_nz/val``, ``grad_lin_idx/coef``)."""
        if getattr(self, "_compact", None) is None:
            cnt = _lib.CompactCounts()
            _lib.check(self._L.nlo_nlp_compact_counts(self._h, C.byref(cnt)))
            i32 = lambda n: np.zeros(int(n), np.int32)
            f32 = lambda n: np.zeros(int(n), np.float32)
            lay = dict(g_var_rows=i32(cnt.n_g_var), g_copy_rows=i32(cnt.n_g_copy), g_copy_vars=i32(cnt.n_g_copy),
                       jac_var_nz=i32(cnt.n_jac_var), jac_const_nz=i32(cnt.n_jac_const), jac_const_val=f32(cnt.n_jac_const),
                       grad_var_idx=i32(cnt.n_grad_var), grad_lin_idx=i32(cnt.n_grad_lin), grad_lin_coef=f32(cnt.n_grad_lin))
            order = ("g_var_rows", "g_copy_rows", "g_copy_vars", "jac_var_nz", "jac_const_nz", "jac_const_val", "grad_var_idx",
                     "grad_lin_idx", "grad_lin_coef")
            _lib.check(self._L.nlo_nlp_compact_layout(self._h, *[lay[k].ctypes.data for k in order]))
            self._compact = lay
        return self._compact

    def eval_host_compact(self, w: np.ndarray, want=("g", "jac", "f", "grad_f"), out=None):
        """Host problem-major evaluation returning only the varying entries: dict of ``g`` (P, n_g_var), ``jac`` (P, n_jac_var),
        ``f`` (P,), ``grad_f`` (P, n_grad_var); ``expand_compact`` rebuilds the full arrays."""
        w = np.ascontiguousarray(w, np.float32)
        P = w.shape[0]
        lay = self.compact_layout()
        res = out if out is not None else {}
        shapes = {"g": (P, len(lay["g_var_rows"])), "jac": (P, len(lay["jac_var_nz"])), "f": (P,), "grad_f": (P, len(lay["grad_var_idx"]))}
        for k in want:
            if k not in res:
                res[k] = np.empty(shapes[k], np.float32)
        _lib.check(self._L.nlo_nlp_eval_host_compact(self._h, w.ctypes.data, P, _lib.ptr(res.get("g")), _lib.ptr(res.get("jac")),
                                                     _lib.ptr(res.get("f")), _lib.ptr(res.get("grad_f"))))
        return res

    def expand_compact(self, w: np.ndarray, compact: dict) -> dict:
        """Full (P, n_g) / (P, nnz) / (P, n_w) arrays from a compact result and the decision vectors it was evaluated at
        (what a per-problem solver does once for the constant parts, then per call for the varying ones)."""
        w = np.ascontiguousarray(w, np.float32)
        P = w.shape[0]
        lay = self.compact_layout()
        res = {}
        if "g" in compact:
            g = np.empty((P, self.n_g), np.float32)
            g[:, lay["g_var_rows"]] = compact["g"]
            g[:, lay["g_copy_rows"]] = w[:, lay["g_copy_vars"]]
            res["g"] = g
        if "jac" in compact:
            jac = np.empty((P, self.nnz), np.float32)
            jac[:, lay["jac_var_nz"]] = compact["jac"]
            jac[:, lay["jac_const_nz"]] = lay["jac_const_val"][None, :]
            res["jac"] = jac
        if "grad_f" in compact:
            gr = np.zeros((P, self.n_w), np.float32)
            gr[:, lay["grad_var_idx"]] = compact["grad_f"]
            gr[:, lay["grad_lin_idx"]] = lay["grad_lin_coef"][None, :] * w[:, lay["grad_lin_idx"]]
            res["grad_f"] = gr
        if "f" in compact:
            res["f"] = compact["f"]
        return res

    def jac_tvec(self, jac, y, add=None, out=None, P: Optional[int] = None):
        """out = add + J^T y per problem (SoA tensors: jac (nnz, ld), y (n_g, ld), add/out (n_w, ld))."""
        import torch
        ld = jac.shape[1]
        P = ld if P is None else P
        if out is None:
            out = torch.empty((self.n_w, ld), dtype=torch.float32, device=jac.device)
        _lib.check(self._L.nlo_nlp_jac_tvec(self._h, jac.data_ptr(), y.data_ptr(), _lib.ptr(add), P, ld, out.data_ptr(),
                                            torch.cuda.current_stream(jac.device).cuda_stream))
        return out

    def violation(self, g, lbg, ubg, P: Optional[int] = None):
        import torch
        ld = g.shape[1]
        P = ld if P is None else P
        v = torch.empty(P, dtype=torch.float32, device=g.device)
        _lib.check(self._L.nlo_nlp_violation(self._h, g.data_ptr(), lbg.data_ptr(), ubg.data_ptr(), P, ld, v.data_ptr(),
                                             torch.cuda.current_stream(g.device).cuda_stream))
        return v
