"""Batched multi-start NLP solver on the GPU: the caller on top of the evaluation hot path.

The reference hands each problem to IPOPT (core/runner.py:112-133: tol 1e-4, max_iter 1000).  IPOPT does not
exist in this environment and is inherently one-problem-at-a-time, so the batched path solves all P starts at
once with a bound-constrained augmented Lagrangian whose inner problems are minimised by a batched L-BFGS:

    L_rho(w, lam) = f(w) + sum_r [ (y_r^2 - lam_r^2) / (2 rho) ],   y = rho * (z - clip(z, lbg, ubg)),  z = g(w) + lam/rho
    grad_w L_rho  = grad f + J^T y                                   (J^T y: nlo_nlp_jac_tvec)
    lam <- y after every inner solve; rho grows where the violation stalls.

Every iteration is one call of the evaluation hot path (g, nnz(dg/dw), f, grad f for the whole batch) plus
vector updates; all state is structure-of-arrays [variable][problem] and stays on the device.  Equalities are
rows with lbg == ubg; one-sided rows have an infinite bound.  This is a first-order method: it reaches the
reference's tolerance (1e-4) on constraint violation and stationarity, not IPOPT's iteration counts.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np


@dataclass
class SolveResult:
    w: "object"            # torch (n_w, P) SoA
    f: "object"            # (P,)
    violation: "object"    # (P,) max bound violation of g
    stationarity: "object" # (P,) inf-norm of grad_w L
    outer_iterations: int
    evaluations: int


class BatchedALSolver:
    def __init__(self, prob, rho0: float = 10.0, rho_max: float = 1e4, outer_iters: int = 40, inner_iters: int = 100,
                 memory: int = 12, tol: float = 1e-4, verbose: bool = False):
        self.prob = prob
        self.rho0, self.rho_max = rho0, rho_max
        self.outer_iters, self.inner_iters, self.m = outer_iters, inner_iters, memory
        self.tol, self.verbose = tol, verbose
        self.evals = 0

    # ---- one evaluation of the augmented Lagrangian and its gradient -------------------------------------------
    def _al(self, w, lam, rho, bufs):
        import torch
        g, jac, f, grad = bufs
        self.prob.eval_device(w, g, jac, f, grad)
        self.evals += 1
        z = g + lam / rho
        y = rho * (z - torch.minimum(torch.maximum(z, self.lb), self.ub))
        val = f + ((y * y - lam * lam).sum(dim=0)) / (2.0 * rho)
        gl = self.prob.jac_tvec(jac, y, add=grad)
        return val, gl, y

    def solve(self, w0, lam0=None) -> SolveResult:
        import torch
        prob = self.prob
        dev = w0.device
        P = w0.shape[1]
        big = 1e30
        lb, ub = prob.bounds()
        self.lb = torch.from_numpy(np.clip(lb, -big, big).astype(np.float32)).to(dev)[:, None]
        self.ub = torch.from_numpy(np.clip(ub, -big, big).astype(np.float32)).to(dev)[:, None]
        w = w0.clone()
        lam = torch.zeros((prob.n_g, P), device=dev) if lam0 is None else lam0.clone()
        rho = torch.full((P,), self.rho0, device=dev)
        bufs = prob.alloc_outputs(P, dev)
        bufs_trial = prob.alloc_outputs(P, dev)
        m = self.m
        S = torch.zeros((m, prob.n_w, P), device=dev)
        Y = torch.zeros((m, prob.n_w, P), device=dev)
        R = torch.zeros((m, P), device=dev)
        prev_viol = torch.full((P,), float("inf"), device=dev)
        outer = 0
        for outer in range(1, self.outer_iters + 1):
            val, gl, y = self._al(w, lam, rho, bufs)
            S.zero_(); Y.zero_(); R.zero_()
            n_hist = 0
            step0 = 1.0 / torch.clamp(gl.abs().amax(dim=0), min=1.0)          # first step: conservative
            for it in range(self.inner_iters):
                gnorm = gl.abs().amax(dim=0)
                if bool((gnorm < self.tol * 0.1).all()):
                    break
                # two-loop recursion, batched over problems
                q = gl.clone()
                alphas = []
                order = [(n_hist - 1 - j) % m for j in range(min(n_hist, m))]
                for idx in order:
                    a = R[idx] * (S[idx] * q).sum(dim=0)
                    q -= a * Y[idx]
                    alphas.append(a)
                if n_hist > 0:
                    last = (n_hist - 1) % m
                    yy = (Y[last] * Y[last]).sum(dim=0)
                    gamma = torch.where(yy > 0, (S[last] * Y[last]).sum(dim=0) / torch.clamp(yy, min=1e-30), step0)
                else:
                    gamma = step0
                q *= gamma
                for idx, a in zip(reversed(order), reversed(alphas)):
                    b = R[idx] * (Y[idx] * q).sum(dim=0)
                    q += (a - b) * S[idx]
                d = -q
                slope = (gl * d).sum(dim=0)
                bad = slope >= 0                                                # not a descent direction: steepest descent
                if bool(bad.any()):
                    d = torch.where(bad[None, :], -gl * step0, d)
                    slope = (gl * d).sum(dim=0)
                # Armijo backtracking, batched: every problem keeps its own step
                t = torch.ones(P, device=dev)
                done = gnorm < self.tol * 0.1
                w_new, val_new, gl_new, y_new = w, val, gl, y
                acc_w = w.clone(); acc_val = val.clone(); acc_gl = gl.clone(); acc_y = y.clone()
                for ls in range(12):
                    wt = w + t * d
                    vt, gt, yt = self._al(wt, lam, rho, bufs_trial)
                    ok = (vt <= val + 1e-4 * t * slope) & ~done
                    if bool(ok.any()):
                        sel = ok[None, :]
                        acc_w = torch.where(sel, wt, acc_w); acc_gl = torch.where(sel, gt, acc_gl); acc_y = torch.where(sel, yt, acc_y)
                        acc_val = torch.where(ok, vt, acc_val)
                        done = done | ok
                    if bool(done.all()):
                        break
                    t = torch.where(done, t, t * 0.5)
                s_vec = acc_w - w
                y_vec = acc_gl - gl
                sy = (s_vec * y_vec).sum(dim=0)
                good = sy > 1e-10 * (y_vec * y_vec).sum(dim=0).clamp(min=1e-30)
                slot = n_hist % m
                S[slot] = torch.where(good[None, :], s_vec, torch.zeros_like(s_vec))
                Y[slot] = torch.where(good[None, :], y_vec, torch.zeros_like(y_vec))
                R[slot] = torch.where(good, 1.0 / sy.clamp(min=1e-30), torch.zeros_like(sy))
                n_hist += 1
                w, val, gl, y = acc_w, acc_val, acc_gl, acc_y
            # multiplier / penalty update
            g = bufs[0]
            prob.eval_device(w, g, None, bufs[2], None); self.evals += 1
            viol = torch.clamp(torch.maximum(self.lb - g, g - self.ub), min=0).amax(dim=0)
            z = g + lam / rho
            lam = rho * (z - torch.minimum(torch.maximum(z, self.lb), self.ub))
            stall = viol > 0.25 * prev_viol
            rho = torch.where(stall & (viol > self.tol), torch.clamp(rho * 4.0, max=self.rho_max), rho)
            prev_viol = viol
            if self.verbose:
                print(f"[AL] outer {outer:2d} evals {self.evals:5d} f med {bufs[2].median().item():.5f} viol max {viol.max().item():.2e} "
                      f"med {viol.median().item():.2e} |gradL| max {gl.abs().amax(dim=0).max().item():.2e} rho max {rho.max().item():.0e}")
            if bool((viol < self.tol).all()) and bool((gl.abs().amax(dim=0) < self.tol).all()):
                break
        f = bufs[2].clone()
        return SolveResult(w, f, viol, gl.abs().amax(dim=0), outer, self.evals)
