"""Batched benchmark runner: the reference's ``RunBenchmark`` (core/runner.py:9-153) with the NLP
evaluated for P independent multi-start problems on the GPU instead of one symbolic Opti instance.

Constructor arguments keep the reference's names and meaning; ``dynamics`` / ``geometry`` are the YAML
names (``body.dynamic`` / ``body.shape``) instead of CasADi expression builders, ``sdf_func`` is a
``LearnedSDF`` (solver.mode l4casadi) or a list of circles (solver.mode casadi).
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

    def run(self):
        raise NotImplementedError(
            "the IPOPT solve of core/runner.py:112-133 is the caller of this hot path and is not part of it; "
            "use evaluate() / NlpProblem.eval_device() as the callbacks of a solver")
