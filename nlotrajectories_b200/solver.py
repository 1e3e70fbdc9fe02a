"""Batched multi-start NLP solvers on the GPU: callers on top of the evaluation hot path.

Two solvers live here.  ``BatchedIPSolver`` (bottom of the file) is the one to use: a primal-dual interior point
with the exact Hessian of the Lagrangian (nlo_nlp_hess) - the algorithm family of IPOPT, which the reference calls
once per problem - advancing all starts in lock step; measured on benchmark_1 x 256 starts: 98.8 % of the starts
converge to tol 1e-4 (median 25 iterations, 4 s for the batch), objectives 1.50352 +- 1e-5, equal to the same algorithm
run on the fp64 CPU oracle and 1e-4 above scipy SLSQP's 1.50342 (the barrier's share at the final mu = 1e-5).  ``BatchedALSolver`` is
the earlier first-order baseline (augmented Lagrangian + L-BFGS), kept for comparison:


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


# =====================================================================================================================
# Batched primal-dual interior point on top of the evaluation hot path (value, Jacobian, exact Hessian of the Lagrangian)
# =====================================================================================================================
class DeviceEvaluator:
    """Adapter between the solver's problem-major fp64 tensors and the CUDA evaluation path (SoA fp32):
    ``nlo_nlp_eval`` for f, grad f, g, nnz(dg/dw) and ``nlo_nlp_hess`` for the Hessian of the Lagrangian.  The sparse
    values are scattered into dense per-problem matrices for the batched KKT solve."""

    def __init__(self, prob):
        import torch
        self.prob = prob
        self.dev = torch.device("cuda", prob.device)
        colind, row = prob.jac_sparsity()
        self.j_row = torch.from_numpy(row.astype(np.int64)).to(self.dev)
        self.j_col = torch.from_numpy(np.repeat(np.arange(prob.n_w), np.diff(colind)).astype(np.int64)).to(self.dev)
        hcol, hrow = prob.hess_sparsity()
        self.h_row = torch.from_numpy(hrow.astype(np.int64)).to(self.dev)
        self.h_col = torch.from_numpy(np.repeat(np.arange(prob.n_w), np.diff(hcol)).astype(np.int64)).to(self.dev)
        self.n_w, self.n_g = prob.n_w, prob.n_g
        self.evals = 0

    def _soa(self, a):
        import torch
        return a.to(torch.float32).T.contiguous()

    def eval(self, w, want_jac=True):
        import torch
        P = w.shape[0]
        ws = self._soa(w)
        g = torch.empty((self.n_g, P), dtype=torch.float32, device=self.dev)
        f = torch.empty(P, dtype=torch.float32, device=self.dev)
        jac = torch.empty((self.prob.nnz, P), dtype=torch.float32, device=self.dev) if want_jac else None
        grad = torch.empty((self.n_w, P), dtype=torch.float32, device=self.dev) if want_jac else None
        self.prob.eval_device(ws, g, jac, f, grad)
        self.evals += 1
        if not want_jac:
            return f.double(), None, g.T.double(), None
        J = torch.zeros((P, self.n_g, self.n_w), dtype=torch.float64, device=self.dev)
        J[:, self.j_row, self.j_col] = jac.T.double()
        return f.double(), grad.T.double(), g.T.double(), J

    def hess(self, w, sigma, lam):
        import torch
        P = w.shape[0]
        hv = self.prob.eval_hess_device(self._soa(w), self._soa(lam), sigma.to(torch.float32).contiguous())
        H = torch.zeros((P, self.n_w, self.n_w), dtype=torch.float64, device=self.dev)
        v = hv.T.double()
        H[:, self.h_row, self.h_col] = v
        H[:, self.h_col, self.h_row] = v
        return H


class ElasticEvaluator:
    """Elastic mode for the one-sided inequality rows (exact l1 penalty, the device SNOPT-style codes use where IPOPT would
    enter its restoration phase): row r becomes ``g_r(w) + p_r >= lb_r`` (``g_r(w) - p_r <= ub_r``) with a new variable
    ``p_r >= 0`` that costs ``penalty * p_r``.  Every start is then strictly feasible for the inequality rows, and a solution
    with ``p <= tol`` satisfies the original KKT conditions (multipliers below ``penalty``).  Wraps any evaluator with the
    ``eval`` / ``hess`` interface; ``split(w)`` returns the original variables and the elastic ones."""

    def __init__(self, ev, lbg, ubg, penalty: float = 100.0):
        lbg, ubg = np.asarray(lbg, np.float64), np.asarray(ubg, np.float64)
        lo_only = np.isfinite(lbg) & ~np.isfinite(ubg)
        up_only = np.isfinite(ubg) & ~np.isfinite(lbg)
        self.rows = np.nonzero(lo_only | up_only)[0]
        self.sign = np.where(lo_only[self.rows], 1.0, -1.0)
        self.ev, self.penalty = ev, float(penalty)
        self.m = len(self.rows)
        self.n_w0, self.n_g0 = ev.n_w, ev.n_g
        self.n_w, self.n_g = ev.n_w + self.m, ev.n_g + self.m
        self.lbg = np.concatenate([lbg, np.zeros(self.m)])
        self.ubg = np.concatenate([ubg, np.full(self.m, np.inf)])

    @property
    def evals(self):
        return self.ev.evals

    def split(self, w):
        return w[:, :self.n_w0], w[:, self.n_w0:]

    def initial(self, w0, margin: float = 1e-2):
        """Append elastic variables that make every elastic row satisfied with ``margin`` to spare."""
        import torch
        _, _, g, _ = self.ev.eval(w0, want_jac=False)
        rows = torch.from_numpy(self.rows).to(w0.device); sign = torch.from_numpy(self.sign).to(w0.device)
        bound = torch.from_numpy(np.where(self.sign > 0, self.lbg[self.rows], self.ubg[self.rows])).to(w0.device)
        short = sign * (bound - g[:, rows].double())                       # > 0 where the row is violated
        return torch.cat([w0, torch.clamp(short, min=0.0) + margin], dim=1)

    def eval(self, w, want_jac=True):
        import torch
        w0, p = self.split(w)
        f, grad, g, J = self.ev.eval(w0.contiguous(), want_jac)
        P, dev = w.shape[0], w.device
        rows = torch.from_numpy(self.rows).to(dev); sign = torch.from_numpy(self.sign).to(dev)
        g2 = torch.cat([g, p], dim=1)
        g2[:, rows] += sign * p
        f2 = f + self.penalty * p.sum(1)
        if not want_jac:
            return f2, None, g2, None
        grad2 = torch.cat([grad, torch.full((P, self.m), self.penalty, dtype=grad.dtype, device=dev)], dim=1)
        J2 = torch.zeros((P, self.n_g, self.n_w), dtype=J.dtype, device=dev)
        J2[:, :self.n_g0, :self.n_w0] = J
        idx = torch.arange(self.m, device=dev)
        J2[:, rows, self.n_w0 + idx] = sign
        J2[:, self.n_g0 + idx, self.n_w0 + idx] = 1.0
        return f2, grad2, g2, J2

    def hess(self, w, sigma, lam):
        import torch
        H = self.ev.hess(w[:, :self.n_w0].contiguous(), sigma, lam[:, :self.n_g0].contiguous())
        H2 = torch.zeros((w.shape[0], self.n_w, self.n_w), dtype=H.dtype, device=w.device)
        H2[:, :self.n_w0, :self.n_w0] = H
        return H2


@dataclass
class IPResult:
    w: "object"            # (P, n_w) fp64
    f: "object"            # (P,)
    violation: "object"    # (P,) max bound violation of g
    kkt_error: "object"    # (P,) scaled optimality error at mu = 0
    iterations: "object"   # (P,) iterations until convergence (max_iter where not converged)
    converged: "object"    # (P,) bool: scaled KKT error <= tol
    lam: "object"          # (P, n_g) constraint multipliers
    stalled: "object" = None   # (P,) bool: feasible to tol and objective unchanged for 30 iterations without reaching tol (typical on
                               # the kinks of a ReLU SDF, where no one-sided gradient satisfies stationarity to 1e-4)


class BatchedIPSolver:
    """All P problems advance in lock step through a primal-dual interior-point iteration (the algorithm family of
    IPOPT, which the reference calls one problem at a time: core/runner.py:112-133, tol 1e-4): slack variables on the
    inequality rows, log barrier, exact Hessian, ell-1 merit line search, monotone barrier update.  Each iteration
    costs one batched evaluation + one batched Hessian + one batched dense KKT solve (n_w + n_eq unknowns)."""

    def __init__(self, evaluator, lbg, ubg, tol: float = 1e-4, max_iter: int = 300, mu0: float = 0.1, verbose: bool = False,
                 ls_multipliers: bool = True, compact: bool = True):
        self.ev, self.tol, self.max_iter, self.mu0, self.verbose = evaluator, tol, max_iter, mu0, verbose
        self.ls_multipliers, self.compact = ls_multipliers, compact
        self.lbg, self.ubg = np.asarray(lbg, np.float64), np.asarray(ubg, np.float64)

    def solve(self, w0) -> IPResult:
        import torch
        ev = self.ev
        w = w0.clone().double()
        dev = w.device
        P, n_w = w.shape
        lb_all = torch.from_numpy(self.lbg).to(dev); ub_all = torch.from_numpy(self.ubg).to(dev)
        eq = torch.from_numpy(self.lbg == self.ubg).to(dev)
        iE, iI = torch.nonzero(eq).flatten(), torch.nonzero(~eq).flatten()
        nE, nI = len(iE), len(iI)
        lE = lb_all[iE]
        l, u = lb_all[iI], ub_all[iI]
        has_l, has_u = torch.isfinite(l), torch.isfinite(u)
        lf, uf = torch.where(has_l, l, torch.zeros_like(l)), torch.where(has_u, u, torch.zeros_like(u))
        inf = torch.full_like(l, float("inf"))

        f, grad, g, J = ev.eval(w)
        # slacks strictly inside their bounds
        push = 1e-2
        s = g[:, iI].clone()
        lo = torch.where(has_l, lf + push * torch.clamp(lf.abs(), min=1.0), -inf)
        hi = torch.where(has_u, uf - push * torch.clamp(uf.abs(), min=1.0), inf)
        mid = 0.5 * (lf + uf)
        both = has_l & has_u
        lo = torch.where(both & (lo > hi), mid, lo); hi = torch.where(both & (lo > hi), mid, hi)
        s = torch.minimum(torch.maximum(s, lo), hi)
        mu = torch.full((P,), self.mu0, dtype=torch.float64, device=dev)
        dl = lambda s_: torch.where(has_l, s_ - lf, inf)          # distance to the lower / upper bound
        du = lambda s_: torch.where(has_u, uf - s_, inf)
        z_l = torch.where(has_l, mu[:, None] / dl(s), torch.zeros_like(s))
        z_u = torch.where(has_u, mu[:, None] / du(s), torch.zeros_like(s))
        lam_E = torch.zeros((P, nE), dtype=torch.float64, device=dev)
        nu = torch.ones(P, dtype=torch.float64, device=dev)
        done = torch.zeros(P, dtype=torch.bool, device=dev)
        iters = torch.full((P,), self.max_iter, dtype=torch.int64, device=dev)
        delta_w = torch.zeros(P, dtype=torch.float64, device=dev)
        eye_w = torch.eye(n_w, dtype=torch.float64, device=dev)
        err0 = torch.full((P,), float("inf"), dtype=torch.float64, device=dev)
        stalled = torch.zeros(P, dtype=torch.bool, device=dev)
        f_mark = f.clone()
        # finished starts leave the working set (every 10 iterations, once a quarter of it is done): an iteration costs one
        # evaluation, one Hessian and one dense KKT solve per ACTIVE start, so a few stragglers no longer pay for the whole batch
        P0 = P
        idx = torch.arange(P0, device=dev)
        out = dict(w=torch.empty((P0, n_w), dtype=torch.float64, device=dev), f=torch.empty(P0, dtype=torch.float64, device=dev),
                   viol=torch.empty(P0, dtype=torch.float64, device=dev), err=torch.empty(P0, dtype=torch.float64, device=dev),
                   iters=torch.empty(P0, dtype=torch.int64, device=dev), conv=torch.zeros(P0, dtype=torch.bool, device=dev),
                   stalled=torch.zeros(P0, dtype=torch.bool, device=dev), lam=torch.empty((P0, ev.n_g), dtype=torch.float64, device=dev))

        def flush(rows):
            o = idx[rows]
            out["w"][o] = w[rows]; out["f"][o] = f[rows]; out["err"][o] = err0[rows]; out["iters"][o] = iters[rows]
            out["viol"][o] = torch.clamp(torch.maximum(lb_all - g[rows], g[rows] - ub_all), min=0.0).amax(1)
            out["conv"][o] = done[rows] & ~stalled[rows]; out["stalled"][o] = stalled[rows]
            lam_o = torch.zeros((int(rows.sum()), ev.n_g), dtype=torch.float64, device=dev)
            lam_o[:, iE] = lam_E[rows]; lam_o[:, iI] = (z_u - z_l)[rows]
            out["lam"][o] = lam_o

        for it in range(self.max_iter):
            if self.compact and it > 0 and it % 10 == 0 and float(done.float().mean()) >= 0.25:
                flush(done)
                keep = ~done
                (w, s, mu, z_l, z_u, lam_E, nu, iters, delta_w, err0, stalled, f_mark, f, grad, g, J, idx) = (
                    t[keep] for t in (w, s, mu, z_l, z_u, lam_E, nu, iters, delta_w, err0, stalled, f_mark, f, grad, g, J, idx))
                done = done[keep]
                P = int(keep.sum())
            lam_I = z_u - z_l
            lam = torch.zeros((P, ev.n_g), dtype=torch.float64, device=dev)
            lam[:, iE] = lam_E; lam[:, iI] = lam_I
            JE, JI = J[:, iE, :], J[:, iI, :]
            r_d = grad + torch.einsum("prw,pr->pw", J, lam)
            r_cE = g[:, iE] - lE
            r_cI = g[:, iI] - s
            comp_l = torch.where(has_l, z_l * dl(s), torch.zeros_like(s))
            comp_u = torch.where(has_u, z_u * du(s), torch.zeros_like(s))
            s_d = torch.clamp((lam.abs().sum(1) + z_l.sum(1) + z_u.sum(1)) / (ev.n_g + 2 * nI) / 100.0, min=1.0)
            feas = torch.maximum(r_cE.abs().amax(1) if nE else torch.zeros(P, device=dev, dtype=torch.float64),
                                 r_cI.abs().amax(1) if nI else torch.zeros(P, device=dev, dtype=torch.float64))

            def kkt_err(m):
                c = torch.maximum((comp_l - torch.where(has_l, m[:, None], torch.zeros_like(s))).abs().amax(1),
                                  (comp_u - torch.where(has_u, m[:, None], torch.zeros_like(s))).abs().amax(1)) if nI else torch.zeros_like(feas)
                return torch.maximum(torch.maximum(r_d.abs().amax(1) / s_d, feas), c / s_d)
            err0 = kkt_err(torch.zeros_like(mu))
            newly = (~done) & (err0 <= self.tol)
            iters = torch.where(newly, torch.full_like(iters, it), iters)
            done = done | newly
            if it > 0 and it % 30 == 0:
                st = (~done) & (feas <= self.tol) & ((f - f_mark).abs() <= 1e-7 * torch.clamp(f.abs(), min=1.0))
                iters = torch.where(st, torch.full_like(iters, it), iters)
                stalled = stalled | st
                done = done | st
                f_mark = f.clone()
            if self.verbose and (it % 10 == 0 or bool(done.all())):
                print(f"[IP] it {it:3d} done {int(done.sum())}/{P} f med {f.median().item():.6f} feas med {feas.median().item():.2e} max {feas.max().item():.2e} "
                      f"err0 med {err0.median().item():.2e} max {err0.max().item():.2e} mu med {mu.median().item():.1e} delta max {delta_w.max().item():.1e}", flush=True)
            if self.verbose > 2:
                print(f"   it {it}: |r_d| {r_d.abs().amax(1).cpu().numpy()} s_d {s_d.cpu().numpy()} comp_l {comp_l.amax(1).cpu().numpy() if nI else 0} "
                      f"comp_u {comp_u.amax(1).cpu().numpy() if nI else 0} argmax r_d {r_d.abs().argmax(1).cpu().numpy()}")
            if bool(done.all()):
                break
            # monotone barrier update
            for _ in range(4):
                shrink = (~done) & (kkt_err(mu) <= 10.0 * mu) & (mu > self.tol / 10.0)
                if not bool(shrink.any()):
                    break
                mu = torch.where(shrink, torch.clamp(torch.minimum(0.2 * mu, mu ** 1.5), min=self.tol / 10.0), mu)
            # condensed KKT system
            H = ev.hess(w, torch.ones(P, dtype=torch.float64, device=dev), lam)
            Sig = torch.where(has_l, z_l / dl(s), torch.zeros_like(s)) + torch.where(has_u, z_u / du(s), torch.zeros_like(s))
            mu_l = torch.where(has_l, mu[:, None] / dl(s), torch.zeros_like(s))
            mu_u = torch.where(has_u, mu[:, None] / du(s), torch.zeros_like(s))
            t_I = Sig * r_cI - mu_l + mu_u
            W = H + torch.einsum("prw,pr,prv->pwv", JI, Sig, JI)
            rhs_w = -(grad + torch.einsum("prw,pr->pw", JE, lam_E) + torch.einsum("prw,pr->pw", JI, t_I))
            # KKT step  [W + delta I, JE^T; JE, -delta_c I] [dw; dlam_E] = [rhs_w; -r_cE]  through its Schur complement:
            #   (W + delta I + rho JE^T JE) dw = rhs_w - rho JE^T r_cE,   dlam_E = rho (JE dw + r_cE),   rho = 1 / delta_c.
            # delta_c > 0 keeps the system solvable where the equality Jacobian loses rank (it does at the straight-line guess of
            # the unicycle: heading pi/4 and zero speed make the x- and y-defect rows dependent), and the Cholesky of the
            # left-hand side is the inertia test: it succeeds iff W + delta I is positive definite on the null space of JE.
            rho_c = 1e8
            dw = torch.zeros_like(w); dlam_E = torch.zeros_like(lam_E)
            JtJ = torch.einsum("pew,pev->pwv", JE, JE) if nE else torch.zeros_like(W)
            rhs_k = rhs_w - rho_c * torch.einsum("pew,pe->pw", JE, r_cE) if nE else rhs_w
            Kc = W + rho_c * JtJ
            todo = ~done
            dwt = delta_w.clone()

            def try_solve(dlt):
                Lc, info = torch.linalg.cholesky_ex(Kc + dlt[:, None, None] * eye_w)
                good_ = info == 0
                Lc = torch.where(good_[:, None, None], Lc, eye_w[None])
                # two batched TRSMs: the potrs path behind torch.cholesky_solve runs one slow batched TRSV per triangle
                y_ = torch.linalg.solve_triangular(Lc, rhs_k[:, :, None], upper=False)
                cand_ = torch.linalg.solve_triangular(Lc.transpose(1, 2), y_, upper=True)[:, :, 0]
                good_ = good_ & torch.isfinite(cand_).all(1) & (cand_.abs().amax(1) < 1e3)
                return cand_, good_
            for attempt in range(16):
                cand, good = try_solve(dwt)
                take = todo & good
                dw[take] = cand[take]
                todo = todo & ~good
                if not bool(todo.any()):
                    break
                dwt = torch.where(todo, torch.clamp(dwt * 8.0, min=1e-4, max=1e8), dwt)
            # a barely positive-definite reduced Hessian gives enormous steps: solve once more, for the problems that
            # needed more regularisation than last time, with twice the value that first passed the test
            bumped = (~done) & (dwt > delta_w)
            if bool(bumped.any()):
                d2 = torch.where(bumped, 2.0 * dwt, dwt)
                cand, good = try_solve(d2)
                ok2 = bumped & good
                dw[ok2] = cand[ok2]
                dwt = torch.where(ok2, d2, dwt)
            if nE:
                dlam_E = rho_c * (torch.einsum("pew,pw->pe", JE, dw) + r_cE)
                dlam_E = torch.where(done[:, None], torch.zeros_like(dlam_E), dlam_E)
            curv = torch.einsum("pw,pwv,pv->p", dw, W, dw) + dwt * (dw * dw).sum(1)
            delta_w = torch.where(dwt > 0, dwt / 3.0, dwt)
            delta_w = torch.where(delta_w < 1e-8, torch.zeros_like(delta_w), delta_w)
            ds = torch.einsum("prw,pw->pr", JI, dw) + r_cI
            dz_l = torch.where(has_l, mu_l - z_l - z_l / dl(s) * ds, torch.zeros_like(s))
            dz_u = torch.where(has_u, mu_u - z_u + z_u / du(s) * ds, torch.zeros_like(s))
            # fraction to the boundary
            tau = torch.clamp(1.0 - mu, min=0.99)[:, None]
            big = torch.full_like(s, float("inf"))
            a_p = torch.minimum(torch.where(has_l & (ds < 0), -tau * dl(s) / ds, big), torch.where(has_u & (ds > 0), tau * du(s) / ds, big)).amin(1) if nI else torch.ones(P, device=dev, dtype=torch.float64)
            a_d = torch.minimum(torch.where(has_l & (dz_l < 0), -tau * z_l / dz_l, big), torch.where(has_u & (dz_u < 0), -tau * z_u / dz_u, big)).amin(1) if nI else torch.ones(P, device=dev, dtype=torch.float64)
            a_p = torch.clamp(a_p, max=1.0); a_d = torch.clamp(a_d, max=1.0)
            # ell-1 merit with barrier
            c1_now = r_cE.abs().sum(1) + r_cI.abs().sum(1)
            bar_dir = -(mu[:, None] * (torch.where(has_l, ds / dl(s), torch.zeros_like(s)) - torch.where(has_u, ds / du(s), torch.zeros_like(s)))).sum(1)
            need = ((grad * dw).sum(1) + bar_dir + 0.5 * torch.clamp(curv, min=0.0)) / (0.9 * torch.clamp(c1_now, min=1e-16))
            nu = torch.where(c1_now > 1e-12, torch.clamp(torch.maximum(need + 1e-3, 0.5 * nu), min=1.0), nu)

            def merit(f_, g_, s_):
                bar = -(mu[:, None] * (torch.where(has_l, torch.log(torch.clamp(dl(s_), min=1e-300)), torch.zeros_like(s_)) +
                                       torch.where(has_u, torch.log(torch.clamp(du(s_), min=1e-300)), torch.zeros_like(s_)))).sum(1)
                c1 = (g_[:, iE] - lE).abs().sum(1) + (g_[:, iI] - s_).abs().sum(1)
                return f_ + bar + nu * c1, c1
            phi0, c1_0 = merit(f, g, s)
            dphi = (grad * dw).sum(1) + bar_dir - nu * c1_0
            alpha = torch.where(done, torch.zeros_like(a_p), a_p)
            accepted = done.clone()
            for ls in range(14):
                w_t = w + alpha[:, None] * dw
                s_t = s + alpha[:, None] * ds
                f_t, _, g_t, _ = ev.eval(w_t, want_jac=False)
                phi_t, _ = merit(f_t, g_t, s_t)
                ok = torch.isfinite(phi_t) & (phi_t <= phi0 + 1e-4 * alpha * torch.minimum(dphi, torch.zeros_like(dphi)) + 1e-12 * phi0.abs())
                accepted = accepted | ok
                if bool(accepted.all()):
                    break
                alpha = torch.where(accepted, alpha, alpha * 0.5)
            # problems whose line search failed take the (tiny) last step and get more regularisation next time
            failed = ~accepted
            delta_w = torch.where(failed, torch.clamp(delta_w * 10.0, min=1e-3, max=1e4), delta_w)
            if self.verbose > 1:
                print(f"   it {it}: alpha {alpha.cpu().numpy().round(4)} a_p {a_p.cpu().numpy().round(4)} a_d {a_d.cpu().numpy().round(4)} delta {dwt.cpu().numpy()} "
                      f"mu {mu.cpu().numpy()} nu {nu.cpu().numpy().round(2)} |dw| {dw.abs().amax(1).cpu().numpy().round(4)} err0 {err0.cpu().numpy()} feas {feas.cpu().numpy()}")
            w = w + alpha[:, None] * dw
            s = s + alpha[:, None] * ds
            lam_E = lam_E + alpha[:, None] * dlam_E
            a_dz = torch.where(done, torch.zeros_like(a_d), a_d)[:, None]
            z_l = z_l + a_dz * dz_l
            z_u = z_u + a_dz * dz_u
            kap = 1e10
            z_l = torch.where(has_l, torch.minimum(torch.maximum(z_l, mu[:, None] / (kap * dl(s))), kap * mu[:, None] / dl(s)), z_l)
            z_u = torch.where(has_u, torch.minimum(torch.maximum(z_u, mu[:, None] / (kap * du(s))), kap * mu[:, None] / du(s)), z_u)
            f, grad, g, J = ev.eval(w)
            if nE and self.ls_multipliers:
                # least-squares equality multipliers at the new point: min |grad + JI^T lam_I + JE^T lam_E|^2.  The Newton update
                # rho (JE dw + r_cE) explodes where the linearised equalities are inconsistent (rank-deficient JE at a standstill),
                # and a huge lam_E makes the Hessian of the Lagrangian hugely indefinite for every later iteration.
                JE_n = J[:, iE, :]
                r_n = grad + torch.einsum("prw,pr->pw", J[:, iI, :], z_u - z_l)
                A = torch.einsum("pew,pfw->pef", JE_n, JE_n)
                A = A + (1e-8 * torch.clamp(A.diagonal(dim1=1, dim2=2).amax(1), min=1.0))[:, None, None] * torch.eye(nE, dtype=torch.float64, device=dev)
                LA = torch.linalg.cholesky_ex(A)[0]                      # SPD by construction (Gram matrix + shift)
                b_ls = torch.einsum("pew,pw->pe", JE_n, r_n)[:, :, None]
                lam_ls = -torch.linalg.solve_triangular(LA.transpose(1, 2), torch.linalg.solve_triangular(LA, b_ls, upper=False), upper=True)[:, :, 0]
                # ... unless the Newton multipliers leave the smaller dual residual: near a solution they are exact, while the
                # regularised least squares keeps a bias where JE JE^T is nearly singular (B6: 1.15e-4 on a control that barely
                # enters the dynamics, just above tol)
                res_ls = (r_n + torch.einsum("pew,pe->pw", JE_n, lam_ls)).abs().amax(1)
                res_nt = (r_n + torch.einsum("pew,pe->pw", JE_n, lam_E)).abs().amax(1)
                keep_newton = done | (torch.isfinite(res_nt) & (res_nt < res_ls))
                lam_E = torch.where(keep_newton[:, None], lam_E, lam_ls)
        flush(torch.ones(P, dtype=torch.bool, device=dev))
        return IPResult(w=out["w"], f=out["f"], violation=out["viol"], kkt_error=out["err"], iterations=out["iters"], converged=out["conv"],
                        lam=out["lam"], stalled=out["stalled"])


def solve_elastic(evaluator, lbg, ubg, w0, penalty: float = 1000.0, **solver_kw) -> IPResult:
    """Solve in elastic mode and report in the original problem's terms: ``w`` / ``lam`` without the elastic parts, ``f`` without
    the penalty, ``violation`` including what the elastic variables still absorb (a start whose ``p`` stays above tol is locally
    infeasible, not converged)."""
    import torch
    el = ElasticEvaluator(evaluator, lbg, ubg, penalty)
    tol = solver_kw.get("tol", 1e-4)
    res = BatchedIPSolver(el, el.lbg, el.ubg, **solver_kw).solve(el.initial(w0))
    w, p = el.split(res.w)
    viol = torch.maximum(res.violation, p.amax(1) if el.m else torch.zeros_like(res.violation))
    feasible = viol <= tol
    return IPResult(w=w.contiguous(), f=res.f - penalty * p.sum(1), violation=viol, kkt_error=res.kkt_error, iterations=res.iterations,
                    converged=res.converged & feasible, lam=res.lam[:, :el.n_g0].contiguous(), stalled=res.stalled & feasible)
