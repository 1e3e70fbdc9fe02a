"""Batched benchmark runner: the reference's ``RunBenchmark`` (core/runner.py:9-153) with the NLP
evaluated for P independent multi-start problems on the GPU instead of one symbolic Opti instance.

Constructor arguments keep the reference's names and meaning; ``dynamics`` / ``geometry`` are the YAML
names (``body.dynamic`` / ``body.shape``) instead of CasADi expression builders, ``sdf_func`` is a
``LearnedSDF`` (solver.mode l4casadi) or a list of analytic obstacles ``(cx, cy, radius | size, margin[, kind])`` with kind 0 = circle,
1 = square (solver.mode casadi).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from .problem import NlpProblem
from .sdf import LearnedSDF


class RunBenchmark:
    def __init__(self, dynamics: str, geometry: str, x0, x_goal, N: int, dt: float, sdf_func,
                 control_bounds: Sequence = ((-1.0, 1.0), (-1.0, 1.0)), use_slack: bool = False, slack_penalty: float = 1000,
                 use_smooth: bool = False, smooth_weight: float = 1000, initializer=None, enforce_heading: bool = True,
                 solver_type: str = "ipopt", length: Optional[float] = None, width: Optional[float] = None,
                 wheelbase: Optional[float] = None, device: int = 0):
        if solver_type not in ("ipopt", "sqpmethod"):
            raise ValueError(f"Unsupported solver type: {solver_type}")      # core/runner.py:131
        learned = isinstance(sdf_func, LearnedSDF)
        self.problem = NlpProblem(dynamics, geometry, x0, x_goal, N, dt, control_bounds, use_slack, slack_penalty, use_smooth,
                                  smooth_weight, enforce_heading, length, width, wheelbase,
                                  sdf=sdf_func if learned else None, circles=() if learned else sdf_func, device=device)
        self.initializer = initializer
        self.solver_type = solver_type

    def initial_guess(self, P: int = 1, first: int = 0) -> np.ndarray:
        """(P, n_w): the initializer's X (core/runner.py:106-108; U and slack start at 0) for P == 1 with an
        initializer, otherwise the seeded multi-start family."""
        if self.initializer is not None and P == 1:
            X = self.initializer.get_initial_guess()
            w = np.zeros((1, self.problem.n_w), np.float32)
            if X is not None:
                w[0, :self.problem.n_X] = np.asarray(X, np.float32).reshape(-1)
            return w
        return self.problem.multistart_guess(P, first)

    def evaluate(self, w: np.ndarray):
        """g, dg/dw values (CCS order), f, grad f for a (P, n_w) batch held on the host."""
        return self.problem.eval_host(w)

    def run(self, P: int = 64, first: int = 0, max_iter: int = 300, tol: float = 1e-4, verbose: bool = False, dense: bool = False,
            accept_stalled: bool = True):
        """The reference's ``run()`` (core/runner.py:44-153) for a batch of P multi-start problems: every start is solved by the
        interior point (tol 1e-4 and the exact Hessian, like the IPOPT options at core/runner.py:112-125) - the device solver of the
        CUDA library, or the dense torch solver with ``dense=True`` - and the best start is returned in the reference's shape:
        ``(X_opt (nx, N+1), U_opt (nu, N), result, X_init (N+1, nx), status)``.

        status: the reference returns "success" only when IPOPT converges and "failed" with the last iterate otherwise
        (core/runner.py:134-148).  Here "success" means the returned start converged to tol; with ``accept_stalled`` (default) a
        start that is feasible to tol with a stationary objective but a KKT error above tol - what a piecewise-linear (ReLU) SDF
        leaves at its kinks - also counts, which is a deliberate deviation: pass ``accept_stalled=False`` for the reference's rule."""
        import torch
        from .solver import BatchedIPSolver, DeviceEvaluator, DeviceIPSolver
        pr = self.problem
        w0 = self.initial_guess(P, first).astype(np.float64)
        lb, ub = pr.bounds()
        if dense:
            dev = torch.device("cuda", pr.device)
            res = BatchedIPSolver(DeviceEvaluator(pr), lb, ub, tol=tol, max_iter=max_iter, verbose=verbose).solve(torch.from_numpy(w0).to(dev))
        else:
            res = DeviceIPSolver(pr, max_problems=P, tol=tol, max_iter=max_iter, verbose=verbose).solve(w0)
        f = res.f.cpu().numpy(); ok = res.converged.cpu().numpy(); viol = res.violation.cpu().numpy()
        usable = ok | (res.stalled.cpu().numpy() & (viol <= tol)) if accept_stalled else ok
        score = np.where(usable, f, f + 1e3 * (1.0 + viol))         # usable starts first, by objective
        best = int(np.argmin(score))
        w = res.w[best].cpu().numpy()
        X_opt = w[:pr.n_X].reshape(pr.N + 1, pr.nx).T.copy()
        U_opt = w[pr.n_X:pr.n_X + pr.n_U].reshape(pr.N, pr.nu).T.copy()
        X_init = w0[best, :pr.n_X].reshape(pr.N + 1, pr.nx)
        res.best = best
        return X_opt, U_opt, res, X_init, ("success" if usable[best] else "failed")
