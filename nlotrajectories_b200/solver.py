"""Batched multi-start NLP solvers on the GPU: the callers on top of the evaluation hot path.

The reference hands each problem to IPOPT (core/runner.py:112-133: tol 1e-4, exact Hessian, max_iter 1000).  IPOPT does not
exist in this environment and is one-problem-at-a-time by construction; here all P multi-start problems advance in lock step
through a primal-dual interior point with the exact Hessian of the Lagrangian (``nlo_nlp_hess``) - the algorithm family of IPOPT.

* ``DeviceIPSolver`` - the solver to use: the whole iteration lives in the CUDA library (``csrc/ip_solver.cu``, C ABI
  ``nlo_ip_*``).  Per problem the condensed KKT matrix is block tridiagonal in stage order (x_k, u_k, slack_k: at most 10
  unknowns per block), so its Cholesky costs ~10^5 flops instead of the 727^3 / 3 of a dense factorisation, and batches of
  4,096 (benchmark_4) or 65,536 (benchmark_6) starts fit one GPU.
* ``BatchedIPSolver`` - the same algorithm in torch with dense fp64 linear algebra (``torch.linalg``): the readable reference
  of the iteration, usable with any evaluator (the CPU oracle in the tests); limited to a few hundred starts.

Both produce the same iterates on the same inputs (``tests/test_ip_device.py``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np


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


class DeviceIPSolver:
    """The interior point of ``BatchedIPSolver`` run entirely by the CUDA library (``nlo_ip_*``): evaluation, Hessian,
    block-tridiagonal KKT factorisation, line search and multiplier updates stay on the device for the whole solve."""

    def __init__(self, prob, max_problems: int, tol: float = 1e-4, max_iter: int = 300, mu0: float = 0.1, ls_multipliers: bool = True,
                 compact: bool = True, verbose: bool = False):
        import ctypes as C
        from . import lib as _lib
        self.prob = prob
        self._L = _lib.load()
        lb, ub = prob.bounds()
        self._lb = np.ascontiguousarray(lb, np.float64); self._ub = np.ascontiguousarray(ub, np.float64)
        self.opt = _lib.IpOptions(tol, int(max_iter), mu0, int(ls_multipliers), int(compact), int(verbose))
        h = C.c_void_p()
        _lib.check(self._L.nlo_ip_create(prob._h, self._lb.ctypes.data, self._ub.ctypes.data, int(max_problems), C.byref(h)))
        self._h = h
        self.stats = {}

    def close(self):
        if getattr(self, "_h", None):
            self._L.nlo_ip_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self, w0) -> IPResult:
        """w0: (P, n_w) starts (numpy or torch, any float type).  Returns an ``IPResult`` of CPU tensors."""
        import ctypes as C
        import torch
        from . import lib as _lib
        if hasattr(w0, "detach"):
            w0 = w0.detach().cpu().numpy()
        w0 = np.ascontiguousarray(w0, np.float64)
        P, n_w = w0.shape
        if n_w != self.prob.n_w:
            raise ValueError(f"starts have {n_w} columns, the problem has {self.prob.n_w} decision variables")
        w = np.empty_like(w0); lam = np.empty((P, self.prob.n_g)); f = np.empty(P); viol = np.empty(P); err = np.empty(P)
        iters = np.empty(P, np.int32); status = np.empty(P, np.int32)
        st = _lib.IpStats()
        _lib.check(self._L.nlo_ip_solve(self._h, w0.ctypes.data, P, C.byref(self.opt), w.ctypes.data, f.ctypes.data, viol.ctypes.data,
                                        err.ctypes.data, iters.ctypes.data, status.ctypes.data, lam.ctypes.data, C.byref(st)))
        self.stats = {k: int(getattr(st, k)) for k in ("iterations", "evaluations", "hessians", "trials", "compactions", "trial_problems")}
        self.stats.update(kkt_problems=int(st.kkt_problems), kkt_retries=int(st.kkt_retries), kkt_retry_hist=[int(v) for v in st.kkt_retry_hist])
        self.stats["phase_ms"] = dict(zip(("evaluation", "residual", "hessian", "kkt", "step", "line_search", "update", "multipliers", "compaction_output"),
                                          (round(float(v), 2) for v in st.phase_ms)))
        t = torch.from_numpy
        return IPResult(w=t(w), f=t(f), violation=t(viol), kkt_error=t(err), iterations=t(iters.astype(np.int64)), converged=t(status == 1),
                        lam=t(lam), stalled=t(status == 2))

    def kkt_step(self, jac, hess, omega, rhs, delta_in):
        """One regularised Newton step of the condensed KKT system on device tensors (variable-major): jac (nnz, ld) / hess
        (nnz_hess, ld) fp32, omega (n_g, ld) / rhs (n_w, ld) / delta_in (P,) fp64.  Returns (dw (n_w, ld), delta_out (P,))."""
        import torch
        from . import lib as _lib
        ld, P = jac.shape[1], delta_in.numel()
        dw = torch.zeros((self.prob.n_w, ld), dtype=torch.float64, device=jac.device)
        d_out = torch.empty(P, dtype=torch.float64, device=jac.device)
        _lib.check(self._L.nlo_ip_kkt_step(self._h, jac.data_ptr(), hess.data_ptr(), omega.data_ptr(), rhs.data_ptr(), delta_in.data_ptr(), P, ld,
                                           dw.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream(jac.device).cuda_stream))
        return dw, d_out
