// Batched primal-dual interior point on the device: per-problem bodies of every kernel and the iteration
// loop that drives them.  The reference hands each problem to IPOPT (core/runner.py:112-133: tol 1e-4, exact Hessian,
// max_iter 1000); here all P multi-start problems advance in lock step, every problem keeps its own barrier parameter,
// regularisation, multipliers and step lengths, and the linear algebra exploits the stage structure of the transcription:
// in stage order (x_k, u_k, slack_k) the condensed KKT matrix  H + JI^T Sigma JI + rho JE^T JE + delta I  is block
// tridiagonal with blocks of nx + nu + 1 <= 10 unknowns, so a factorisation costs ~10^5 flops instead of 727^3 / 3.
//
// Everything here is plain C++ that compiles for the device and for the host: the CUDA kernels in ip_solver.cu are thin
// wrappers around these bodies (one thread per problem, structure-of-arrays: element (r, problem p) at base[r * ld + p], so
// consecutive threads touch consecutive addresses), and tests/tools/ip_host_emul.cu runs the same bodies and the same loop
// on the CPU against the numpy oracle, which pins the solver logic without a GPU.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define IP_HD __host__ __device__ __forceinline__
#else
#define IP_HD inline
#endif

#define IP_RHO_C 1e8          // 1 / delta_c: weight of the equality rows in the condensed system
#define IP_KAPPA 1e10         // multiplier safeguard (IPOPT's kappa_sigma)

// ---- tables: identical for all problems -----------------------------------------------------------------------------------
struct IpTables {
  int n_w, n_g, nnz, nnzh, nE, nI;
  const int* rkind;                   // [n_g] 0 = equality row (lb == ub), 1 = inequality row
  const int* ridx;                    // [n_g] index within the equality / inequality list
  const double* lb; const double* ub; // [n_g] (+-inf allowed)
  const int* colind; const int* row;  // dg/dw, compressed columns (values in this order)
  const int* rptr; const int* rnz; const int* rcol;   // the same pattern by rows: rnz = position in the column order
  const int* hcolind; const int* hrow;                // Hessian of the Lagrangian, upper triangle, compressed columns
};

// ---- per-problem state (double, SoA) -----------------------------------------------------------------------------------------
struct IpState {
  size_t ld;
  double *w, *s, *zl, *zu, *lamE;                   // [n_w] [nI] [nI] [nI] [nE] rows
  double *mu, *nu, *delta_w, *err0, *f_mark;        // one row each
  int *iters, *stalled, *orig, *done;
};

struct IpWork {
  size_t ld;
  // evaluation at the current point (fp32, written by nlo_nlp_eval / nlo_nlp_hess)
  float *f, *grad, *g, *jac, *hess;
  float *w32, *lam32;
  // trial point of the line search
  float *wt32, *ft, *gt;
  double *omega, *v, *rhs, *dw, *dw_alt, *ds, *dzl, *dzu, *dlamE, *lam_ls;
  double *dwt, *alpha, *alpha_d, *phi0, *dphi, *eps_ls, *viol;
  int *accepted;
  int *counters;                      // [0] problems done, [1] line-search rejections of the current trial
  int *ls_list[2];                    // problems still in the line search: the list a trial reads and the one it appends to
};

struct IpOut {                        // final results, indexed by the ORIGINAL problem index
  size_t ld;
  double *w, *lam;                    // [n_w] [n_g] rows
  double *f, *viol, *err;
  int *iters, *status;                // status: 1 converged, 2 stalled-feasible, 0 neither
};

IP_HD bool ip_finite(double x) { return x - x == 0.0; }
IP_HD double ip_max(double a, double b) { return a > b ? a : b; }     // NaN-propagating like torch.maximum is not needed here
IP_HD double ip_min(double a, double b) { return a < b ? a : b; }

// Every body below is written for a TEAM of workers per problem: worker `part` of `nparts` takes the rows / columns
// part, part + nparts, ... of each loop, partial results meet in the team's reducer R, and everything that follows a reduction is
// computed redundantly by all workers, while only worker 0 stores per-problem scalars.  On the device a team is the W warps of a
// thread block whose 32 lanes are 32 consecutive problems (so every global access stays coalesced along the problem index) and R
// reduces through shared memory; on the host (tests) a team is one worker and R is the identity.  One thread per problem left the
// GPU idle: at 4,096 problems the row loops (~1,000 dependent iterations) ran on 32 warps in total.
struct IpSolo {                         // team of one
  IP_HD int part() const { return 0; }
  IP_HD int nparts() const { return 1; }
  IP_HD double sum(double v) { return v; }
  IP_HD double max(double v) { return v; }
  IP_HD double min(double v) { return v; }
  IP_HD bool any(bool v) { return v; }
  IP_HD void sync() {}
};

// ---- initial slacks and multipliers -----------------------------------------------------------------------------------------
template <class R>
IP_HD void ip_init_body(const IpTables& T, const IpState& S, const IpWork& W, size_t p, double mu0, int max_iter, R& red) {
  const size_t ld = S.ld;
  const double push = 1e-2;
  for (int r = red.part(); r < T.n_g; r += red.nparts()) {
    if (T.rkind[r] == 0) { S.lamE[(size_t)T.ridx[r] * ld + p] = 0.0; continue; }
    const int i = T.ridx[r];
    const double l = T.lb[r], u = T.ub[r];
    const bool hl = ip_finite(l), hu = ip_finite(u);
    const double lf = hl ? l : 0.0, uf = hu ? u : 0.0;
    double lo = hl ? lf + push * ip_max(fabs(lf), 1.0) : -HUGE_VAL;
    double hi = hu ? uf - push * ip_max(fabs(uf), 1.0) : HUGE_VAL;
    const double mid = 0.5 * (lf + uf);
    if (hl && hu && lo > hi) lo = mid;
    if (hl && hu && lo > hi) hi = mid;
    double sv = (double)W.g[(size_t)r * ld + p];
    sv = ip_min(ip_max(sv, lo), hi);
    S.s[(size_t)i * ld + p] = sv;
    S.zl[(size_t)i * ld + p] = hl ? mu0 / (sv - lf) : 0.0;
    S.zu[(size_t)i * ld + p] = hu ? mu0 / (uf - sv) : 0.0;
  }
  if (red.part() == 0) {
    S.mu[p] = mu0; S.nu[p] = 1.0; S.delta_w[p] = 0.0; S.err0[p] = HUGE_VAL; S.f_mark[p] = (double)W.f[p];
    S.iters[p] = max_iter; S.stalled[p] = 0; S.done[p] = 0;
  }
}

// ---- residuals, convergence test, barrier update, weights and right-hand side of the condensed system ------------------------
// Returns 1 when the problem is done after this test.
template <class R>
IP_HD int ip_residual_body(const IpTables& T, const IpState& S, const IpWork& W, size_t p, int it, double tol, R& red) {
  const size_t ld = S.ld;
  int done = S.done[p];
  double mu = S.mu[p];
  const double f_mark = S.f_mark[p];
  double feas = 0.0, sum_abs = 0.0, sum_z = 0.0, cmax = -HUGE_VAL, cmin = HUGE_VAL;
  bool any_comp = false;
  for (int r = red.part(); r < T.n_g; r += red.nparts()) {
    const double gr = (double)W.g[(size_t)r * ld + p];
    double lam, rc;
    if (T.rkind[r] == 0) {
      lam = S.lamE[(size_t)T.ridx[r] * ld + p];
      rc = gr - T.lb[r];
    } else {
      const size_t q = (size_t)T.ridx[r] * ld + p;
      const double sv = S.s[q], zl = S.zl[q], zu = S.zu[q];
      lam = zu - zl;
      rc = gr - sv;
      sum_z += zl + zu;
      if (ip_finite(T.lb[r])) { const double c = zl * (sv - T.lb[r]); cmax = ip_max(cmax, c); cmin = ip_min(cmin, c); any_comp = true; }
      if (ip_finite(T.ub[r])) { const double c = zu * (T.ub[r] - sv); cmax = ip_max(cmax, c); cmin = ip_min(cmin, c); any_comp = true; }
    }
    feas = ip_max(feas, fabs(rc));
    sum_abs += fabs(lam);
    W.v[(size_t)r * ld + p] = lam;
    W.lam32[(size_t)r * ld + p] = (float)lam;
  }
  feas = red.max(feas); sum_abs = red.sum(sum_abs); sum_z = red.sum(sum_z); cmax = red.max(cmax); cmin = red.min(cmin); any_comp = red.any(any_comp);
  red.sync();                                           // the multipliers written above are read column-wise below
  // dual residual  grad f + J^T lam
  double rd = 0.0;
  for (int c = red.part(); c < T.n_w; c += red.nparts()) {
    double acc = (double)W.grad[(size_t)c * ld + p];
    for (int z = T.colind[c]; z < T.colind[c + 1]; ++z) acc += (double)W.jac[(size_t)z * ld + p] * W.v[(size_t)T.row[z] * ld + p];
    rd = ip_max(rd, fabs(acc));
  }
  rd = red.max(rd);
  red.sync();                                           // ... and overwritten row-wise further down
  const double s_d = ip_max((sum_abs + sum_z) / (double)(T.n_g + 2 * T.nI) / 100.0, 1.0);
  // max_i |comp_i - m| over the bounded slack sides = max(cmax - m, m - cmin)
#define IP_KKT_ERR(m) ip_max(ip_max(rd / s_d, feas), (any_comp ? ip_max(0.0, ip_max(cmax - (m), (m) - cmin)) : 0.0) / s_d)
  const double err0 = IP_KKT_ERR(0.0);
  int iters = -1, stalled = 0;
  if (!done && err0 <= tol) { iters = it; done = 1; }
  const double f = (double)W.f[p];
  const bool mark = it > 0 && it % 30 == 0;
  if (mark && !done && feas <= tol && fabs(f - f_mark) <= 1e-7 * ip_max(fabs(f), 1.0)) { iters = it; stalled = 1; done = 1; }
  // monotone barrier update
  if (!done)
    for (int q = 0; q < 4; ++q) {
      if (!(IP_KKT_ERR(mu) <= 10.0 * mu && mu > tol / 10.0)) break;
      mu = ip_max(ip_min(0.2 * mu, pow(mu, 1.5)), tol / 10.0);
    }
#undef IP_KKT_ERR
  if (red.part() == 0) {
    S.err0[p] = err0; S.done[p] = done; S.mu[p] = mu;
    if (iters >= 0) S.iters[p] = iters;
    if (stalled) S.stalled[p] = 1;
    if (mark) S.f_mark[p] = f;
  }
  // weights of the rows in the condensed matrix and the vector v with  rhs = -(grad f + J^T v)
  for (int r = red.part(); r < T.n_g; r += red.nparts()) {
    const double gr = (double)W.g[(size_t)r * ld + p];
    double om, v;
    if (T.rkind[r] == 0) {
      om = IP_RHO_C;
      v = S.lamE[(size_t)T.ridx[r] * ld + p] + IP_RHO_C * (gr - T.lb[r]);
    } else {
      const size_t q = (size_t)T.ridx[r] * ld + p;
      const double sv = S.s[q];
      om = 0.0; v = 0.0;
      if (ip_finite(T.lb[r])) { const double d = sv - T.lb[r]; om += S.zl[q] / d; v -= mu / d; }
      if (ip_finite(T.ub[r])) { const double d = T.ub[r] - sv; om += S.zu[q] / d; v += mu / d; }
      v += om * (gr - sv);
    }
    W.omega[(size_t)r * ld + p] = om;
    W.v[(size_t)r * ld + p] = v;
  }
  red.sync();
  for (int c = red.part(); c < T.n_w; c += red.nparts()) {
    double acc = (double)W.grad[(size_t)c * ld + p];
    for (int z = T.colind[c]; z < T.colind[c + 1]; ++z) acc += (double)W.jac[(size_t)z * ld + p] * W.v[(size_t)T.row[z] * ld + p];
    W.rhs[(size_t)c * ld + p] = -acc;
  }
  return done;
}

// ---- the step: multiplier / slack directions, fraction to the boundary, merit parameters --------------------------------------
template <class R>
IP_HD void ip_step_body(const IpTables& T, const IpState& S, const IpWork& W, size_t p, R& red) {
  const size_t ld = S.ld;
  const double mu = S.mu[p];
  const int done = S.done[p];
  const double tau = ip_max(1.0 - mu, 0.99);
  double a_p = HUGE_VAL, a_d = HUGE_VAL, c1 = 0.0, bar_dir = 0.0, bar0 = 0.0, curv = 0.0;
  for (int r = red.part(); r < T.n_g; r += red.nparts()) {
    double jd = 0.0;
    for (int t = T.rptr[r]; t < T.rptr[r + 1]; ++t) jd += (double)W.jac[(size_t)T.rnz[t] * ld + p] * W.dw[(size_t)T.rcol[t] * ld + p];
    const double gr = (double)W.g[(size_t)r * ld + p];
    if (T.rkind[r] == 0) {
      const double rc = gr - T.lb[r];
      c1 += fabs(rc);
      W.dlamE[(size_t)T.ridx[r] * ld + p] = done ? 0.0 : IP_RHO_C * (jd + rc);
    } else {
      const size_t q = (size_t)T.ridx[r] * ld + p;
      const double sv = S.s[q], zl = S.zl[q], zu = S.zu[q];
      const double rc = gr - sv, ds = jd + rc;
      c1 += fabs(rc);
      curv += W.omega[(size_t)r * ld + p] * jd * jd;
      double dzl = 0.0, dzu = 0.0;
      if (ip_finite(T.lb[r])) {
        const double d = sv - T.lb[r];
        dzl = mu / d - zl - zl / d * ds;
        if (ds < 0.0) a_p = ip_min(a_p, -tau * d / ds);
        if (dzl < 0.0) a_d = ip_min(a_d, -tau * zl / dzl);
        bar_dir -= mu * ds / d;
        bar0 -= mu * log(ip_max(d, 1e-300));
      }
      if (ip_finite(T.ub[r])) {
        const double d = T.ub[r] - sv;
        dzu = mu / d - zu + zu / d * ds;
        if (ds > 0.0) a_p = ip_min(a_p, tau * d / ds);
        if (dzu < 0.0) a_d = ip_min(a_d, -tau * zu / dzu);
        bar_dir += mu * ds / d;
        bar0 -= mu * log(ip_max(d, 1e-300));
      }
      W.ds[q] = ds; W.dzl[q] = dzl; W.dzu[q] = dzu;
    }
  }
  // dw^T H dw over the upper triangle, grad f . dw, |dw|^2
  double gd = 0.0, d2 = 0.0;
  for (int c = red.part(); c < T.n_w; c += red.nparts()) {
    const double dc = W.dw[(size_t)c * ld + p];
    gd += (double)W.grad[(size_t)c * ld + p] * dc;
    d2 += dc * dc;
    double acc = 0.0;
    for (int z = T.hcolind[c]; z < T.hcolind[c + 1]; ++z) {
      const int rr = T.hrow[z];
      acc += (rr == c ? 1.0 : 2.0) * (double)W.hess[(size_t)z * ld + p] * W.dw[(size_t)rr * ld + p];
    }
    curv += acc * dc;
  }
  a_p = ip_min(red.min(a_p), 1.0); a_d = ip_min(red.min(a_d), 1.0);
  c1 = red.sum(c1); bar_dir = red.sum(bar_dir); bar0 = red.sum(bar0); curv = red.sum(curv); gd = red.sum(gd); d2 = red.sum(d2);
  if (red.part() != 0) return;
  const double dwt = W.dwt[p];
  curv += dwt * d2;
  double dlt = dwt > 0.0 ? dwt / 3.0 : dwt;
  if (dlt < 1e-8) dlt = 0.0;
  S.delta_w[p] = dlt;
  double nu = S.nu[p];
  const double need = (gd + bar_dir + 0.5 * ip_max(curv, 0.0)) / (0.9 * ip_max(c1, 1e-16));
  if (c1 > 1e-12) nu = ip_max(ip_max(need + 1e-3, 0.5 * nu), 1.0);
  S.nu[p] = nu;
  W.phi0[p] = (double)W.f[p] + bar0 + nu * c1;
  W.dphi[p] = gd + bar_dir - nu * c1;
  W.alpha[p] = done ? 0.0 : a_p;
  W.alpha_d[p] = a_d;
  W.accepted[p] = done;
}

// ---- line search ------------------------------------------------------------------------------------------------------------------
// Armijo test of the l1 merit function at the trial point w + a dw, s + a ds, evaluated in column q_col of the trial batch
// (wt32 -> gt, ft).  Collective over the team; every worker gets the verdict.  `skip`: nothing to test (the loops are skipped).
template <class R>
IP_HD bool ip_merit_ok(const IpTables& T, const IpState& S, const IpWork& W, size_t p, size_t q_col, double a, bool skip, R& red) {
  const size_t ld = S.ld;
  const double mu = S.mu[p];
  double c1 = 0.0, bar = 0.0;
  if (!skip)
    for (int r = red.part(); r < T.n_g; r += red.nparts()) {
      const double gr = (double)W.gt[(size_t)r * ld + q_col];
      if (T.rkind[r] == 0) { c1 += fabs(gr - T.lb[r]); continue; }
      const size_t q = (size_t)T.ridx[r] * ld + p;
      const double st = S.s[q] + a * W.ds[q];
      c1 += fabs(gr - st);
      if (ip_finite(T.lb[r])) bar -= mu * log(ip_max(st - T.lb[r], 1e-300));
      if (ip_finite(T.ub[r])) bar -= mu * log(ip_max(T.ub[r] - st, 1e-300));
    }
  c1 = red.sum(c1); bar = red.sum(bar);                  // (collectives: reached by every worker of every team, tested or not)
  const double phi = (double)W.ft[q_col] + bar + S.nu[p] * c1;
  const double phi0 = W.phi0[p];
  return ip_finite(phi) && phi <= phi0 + 1e-4 * a * ip_min(W.dphi[p], 0.0) + 1e-12 * fabs(phi0);
}

// One backtracking trial: returns 1 if the problem is still not accepted (its step length is halved for the next trial).
// The trial batch holds only the problems still searching: column q_col belongs to problem p.  Worker 0 updates accepted / alpha.
template <class R>
IP_HD int ip_merit_body(const IpTables& T, const IpState& S, const IpWork& W, size_t p, size_t q_col, R& red) {
  const int was_accepted = W.accepted[p];
  const double a = W.alpha[p];
  const bool ok = ip_merit_ok(T, S, W, p, q_col, a, was_accepted != 0, red);
  red.sync();                                            // every worker has read accepted / alpha
  if (was_accepted) return 0;
  if (red.part() == 0) { if (ok) W.accepted[p] = 1; else W.alpha[p] = 0.5 * a; }
  return ok ? 0 : 1;
}

#define IP_LS_TRIALS 14       // backtracking trials per iteration (the first at the fraction-to-the-boundary step)

// ---- take the step ------------------------------------------------------------------------------------------------------------
template <class R>
IP_HD void ip_update_body(const IpTables& T, const IpState& S, const IpWork& W, size_t p, R& red) {
  const size_t ld = S.ld;
  const int done = S.done[p];
  if (red.part() == 0 && !W.accepted[p]) S.delta_w[p] = ip_min(ip_max(S.delta_w[p] * 10.0, 1e-3), 1e4);
  const double a = W.alpha[p], adz = done ? 0.0 : W.alpha_d[p], mu = S.mu[p];
  for (int c = red.part(); c < T.n_w; c += red.nparts()) {
    const double wn = S.w[(size_t)c * ld + p] + a * W.dw[(size_t)c * ld + p];
    S.w[(size_t)c * ld + p] = wn;
    W.w32[(size_t)c * ld + p] = (float)wn;
  }
  for (int r = red.part(); r < T.n_g; r += red.nparts()) {
    if (T.rkind[r] == 0) { const size_t q = (size_t)T.ridx[r] * ld + p; S.lamE[q] += a * W.dlamE[q]; continue; }
    const size_t q = (size_t)T.ridx[r] * ld + p;
    const double sv = S.s[q] + a * W.ds[q];
    S.s[q] = sv;
    double zl = S.zl[q] + adz * W.dzl[q], zu = S.zu[q] + adz * W.dzu[q];
    if (ip_finite(T.lb[r])) { const double d = sv - T.lb[r]; zl = ip_min(ip_max(zl, mu / (IP_KAPPA * d)), IP_KAPPA * mu / d); }
    if (ip_finite(T.ub[r])) { const double d = T.ub[r] - sv; zu = ip_min(ip_max(zu, mu / (IP_KAPPA * d)), IP_KAPPA * mu / d); }
    S.zl[q] = zl; S.zu[q] = zu;
  }
}

// ---- least-squares equality multipliers:  min | grad f + JI^T lam_I + JE^T lam_E |^2  -------------------------------------------
// r_n = grad f + JI^T (zu - zl)  (kept in W.rhs),  right-hand side -JE r_n (kept in W.v, indexed by equality),  shift of the Gram matrix
template <class R>
IP_HD void ip_lsq_prep_body(const IpTables& T, const IpState& S, const IpWork& W, size_t p, R& red) {
  const size_t ld = S.ld;
  for (int c = red.part(); c < T.n_w; c += red.nparts()) {
    double acc = (double)W.grad[(size_t)c * ld + p];
    for (int z = T.colind[c]; z < T.colind[c + 1]; ++z) {
      const int r = T.row[z];
      if (T.rkind[r] == 1) { const size_t q = (size_t)T.ridx[r] * ld + p; acc += (double)W.jac[(size_t)z * ld + p] * (S.zu[q] - S.zl[q]); }
    }
    W.rhs[(size_t)c * ld + p] = acc;
  }
  red.sync();
  double amax = 0.0;
  for (int r = red.part(); r < T.n_g; r += red.nparts()) {
    if (T.rkind[r] != 0) continue;
    double b = 0.0, d = 0.0;
    for (int t = T.rptr[r]; t < T.rptr[r + 1]; ++t) {
      const double j = (double)W.jac[(size_t)T.rnz[t] * ld + p];
      b += j * W.rhs[(size_t)T.rcol[t] * ld + p];
      d += j * j;
    }
    W.v[(size_t)T.ridx[r] * ld + p] = -b;
    amax = ip_max(amax, d);
  }
  amax = red.max(amax);
  if (red.part() == 0) W.eps_ls[p] = 1e-8 * ip_max(amax, 1.0);
}

// keep the Newton multipliers where they leave the smaller dual residual, else take the least-squares ones
template <class R>
IP_HD void ip_lsq_choose_body(const IpTables& T, const IpState& S, const IpWork& W, size_t p, R& red) {
  const size_t ld = S.ld;
  double res_ls = 0.0, res_nt = 0.0;
  bool nt_bad = false;
  for (int c = red.part(); c < T.n_w; c += red.nparts()) {
    double a_ls = W.rhs[(size_t)c * ld + p], a_nt = a_ls;
    for (int z = T.colind[c]; z < T.colind[c + 1]; ++z) {
      const int r = T.row[z];
      if (T.rkind[r] == 0) {
        const size_t q = (size_t)T.ridx[r] * ld + p;
        const double j = (double)W.jac[(size_t)z * ld + p];
        a_ls += j * W.lam_ls[q]; a_nt += j * S.lamE[q];
      }
    }
    res_ls = ip_max(res_ls, fabs(a_ls));
    if (!ip_finite(a_nt)) nt_bad = true;
    res_nt = ip_max(res_nt, fabs(a_nt));
  }
  res_ls = red.max(res_ls); res_nt = red.max(res_nt); nt_bad = red.any(nt_bad);
  red.sync();                                            // every worker has read lamE
  const bool keep_newton = S.done[p] || (!nt_bad && res_nt < res_ls);
  if (!keep_newton)
    for (int e = red.part(); e < T.nE; e += red.nparts()) S.lamE[(size_t)e * ld + p] = W.lam_ls[(size_t)e * ld + p];
}

// ---- results of problem p -> outputs at its original index ---------------------------------------------------------------------
template <class R>
IP_HD void ip_flush_body(const IpTables& T, const IpState& S, const IpWork& W, const IpOut& O, size_t p, R& red) {
  const size_t ld = S.ld, o = (size_t)S.orig[p];
  double viol = 0.0;
  for (int r = red.part(); r < T.n_g; r += red.nparts()) {
    const double gr = (double)W.g[(size_t)r * ld + p];
    viol = ip_max(viol, ip_max(T.lb[r] - gr, gr - T.ub[r]));
    const size_t q = (size_t)T.ridx[r] * ld + p;
    O.lam[(size_t)r * O.ld + o] = T.rkind[r] == 0 ? S.lamE[q] : S.zu[q] - S.zl[q];
  }
  for (int c = red.part(); c < T.n_w; c += red.nparts()) O.w[(size_t)c * O.ld + o] = S.w[(size_t)c * ld + p];
  viol = red.max(viol);
  if (red.part() == 0) {
    O.f[o] = (double)W.f[p]; O.viol[o] = viol; O.err[o] = S.err0[p]; O.iters[o] = S.iters[p];
    O.status[o] = S.stalled[p] ? 2 : (S.done[p] ? 1 : 0);
  }
}

// =====================================================================================================================
// Block-tridiagonal symmetric positive definite systems (one per problem)
// =====================================================================================================================
// Unknowns are grouped into nb blocks of NS (dummies pad the short ones); block k couples only with blocks k-1 and k+1 and
// only the first NXR rows of a sub-diagonal block can be non-zero.  Storage per block k, slot-major SoA (slot q of problem p
// at base[(k * SL + q) * ld + p]):
//   matrix  K: D_k lower triangle (row-major, ND = NS (NS+1) / 2 slots) | O_k = K[block k+1, block k] rows 0..NXR-1 (NXR * NS slots)
//   factor  L: L_kk lower triangle | L_{k+1,k} rows 0..NXR-1 | y_k (NS slots)
// Every slot of K is a sum of terms given by a table (CSR by slot; the same for all problems):
//   kind 0: hess[a]    kind 1: omega[c] * jac[a] * jac[b]    kind 2: jac[a] * jac[b]    kind 3: the constant 1 (dummy diagonal)
struct BtTables {
  int nb, NS, NXR, SLK, SLL;
  const int* var;        // [nb * NS] unknown index of (block, local row) or -1 for a dummy
  const int* term_ptr;   // [nb * SLK + 1]
  const int* terms;      // 4 ints per term: kind, a, b, c
};

IP_HD double bt_slot_value(const BtTables& B, int slot, const float* jac, const float* hess, const double* omega, size_t ld, size_t p) {
  double acc = 0.0;
  for (int t = B.term_ptr[slot]; t < B.term_ptr[slot + 1]; ++t) {
    const int kind = B.terms[4 * t], a = B.terms[4 * t + 1], b = B.terms[4 * t + 2], c = B.terms[4 * t + 3];
    if (kind == 0) acc += (double)hess[(size_t)a * ld + p];
    else if (kind == 1) acc += omega[(size_t)c * ld + p] * (double)jac[(size_t)a * ld + p] * (double)jac[(size_t)b * ld + p];
    else if (kind == 2) acc += (double)jac[(size_t)a * ld + p] * (double)jac[(size_t)b * ld + p];
    else acc += 1.0;
  }
  return acc;
}

#define BT_LI(i, j) ((i) * ((i) + 1) / 2 + (j))

// Staging of the blocks the sequential sweep reads next.  On the device the matrix (forward sweep) and factor (backward sweep) blocks
// of the NEXT stage are pulled into the thread's shared-memory column with cp.async while the current stage is computed: with one
// thread per problem nothing else hides the ~1 us of a global load, and register pressure keeps the compiler from batching ~100
// independent loads per stage on its own.  On the host (tests) the blocks are read in place.
template <int NSLOT>
struct BtStage {
#if defined(__CUDA_ARCH__)
  double* buf; int ss;                                        // NSLOT doubles, element stride ss
  IP_HD void fetch(const double* src, size_t ld, int n) {     // slots 0..n-1 of a block whose slot q lives at src[q * ld]
    for (int q = 0; q < n; ++q) {
      const unsigned dst = (unsigned)__cvta_generic_to_shared(buf + (size_t)q * ss);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src + (size_t)q * ld) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  IP_HD void wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
  IP_HD double get(int q) const { return buf[(size_t)q * ss]; }
#else
  const double* src; size_t ld;
  double* buf; int ss;
  IP_HD void fetch(const double* s, size_t l, int) { src = s; ld = l; }
  IP_HD void wait() {}
  IP_HD double get(int q) const { return src[(size_t)q * ld]; }
#endif
};

// One factor + solve attempt with `delta` added to the diagonal of every real unknown.  sc: per-thread scratch of
// 2 * (NXR * NS + ND) + NS doubles with element stride ss (shared memory on the device).  Writes x (indexed by unknown).  Returns true
// when every pivot was positive and the solution is finite with max |x| < xmax.
template <int NS, int NXR>
IP_HD bool bt_solve_attempt(const BtTables& B, const double* __restrict__ Ksrc, double* __restrict__ Lf, const double* __restrict__ rhs_src,
                            double* __restrict__ x, size_t ld, size_t p_src, size_t p, double delta, double xmax, double* sc, int ss) {
  // the matrix and the right-hand side are read from column p_src, the factor and the solution are written to column p (several
  // attempts at one problem with different delta run side by side, each in its own column)
  const double* K = Ksrc + p_src - p;          // so that K[... * ld + p] addresses column p_src
  const double* rhs = rhs_src + p_src - p;
  constexpr int ND = NS * (NS + 1) / 2, NO = NXR * NS, SLK = ND + NO, SLL = ND + NO + NS;
  double* sLo = sc;                      // L_{k+1,k} rows
  double* sSn = sc + (size_t)NO * ss;    // Schur complement of the next diagonal block
  BtStage<SLL> stage;                    // the next stage's blocks
  stage.buf = sc + (size_t)(NO + ND) * ss; stage.ss = ss;
  const int nb = B.nb;
  bool ok = true;
  double L[ND], y[NS], dinv[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    const int u = B.var[i];
#pragma unroll
    for (int j = 0; j <= i; ++j) L[BT_LI(i, j)] = K[(size_t)BT_LI(i, j) * ld + p];
    if (u >= 0) { L[BT_LI(i, i)] += delta; y[i] = rhs[(size_t)u * ld + p]; } else y[i] = 0.0;
  }
  for (int k = 0; k < nb; ++k) {
    // O_k (rows NXR x NS, slots ND.. of block k) and D_{k+1} (slots 0..ND-1 of block k + 1) are contiguous in K: one fetch
    if (k < nb - 1) stage.fetch(K + ((size_t)k * SLK + ND) * ld + p, ld, NO + ND);
    // Cholesky of the current block in place (row-wise Crout) and y = L^-1 b
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      double d = L[BT_LI(j, j)];
#pragma unroll
      for (int m = 0; m < j; ++m) d -= L[BT_LI(j, m)] * L[BT_LI(j, m)];
      if (!(d > 0.0) || !ip_finite(d)) { ok = false; d = 1.0; }
      const double sd = sqrt(d);
      dinv[j] = 1.0 / sd;
      L[BT_LI(j, j)] = sd;
#pragma unroll
      for (int i = j + 1; i < NS; ++i) {
        double v = L[BT_LI(i, j)];
#pragma unroll
        for (int m = 0; m < j; ++m) v -= L[BT_LI(i, m)] * L[BT_LI(j, m)];
        L[BT_LI(i, j)] = v * dinv[j];
      }
    }
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      double v = y[i];
#pragma unroll
      for (int m = 0; m < i; ++m) v -= L[BT_LI(i, m)] * y[m];
      y[i] = v * dinv[i];
    }
    double* Lk = Lf + (size_t)k * SLL * ld + p;
#pragma unroll
    for (int q = 0; q < ND; ++q) Lk[(size_t)q * ld] = L[q];
#pragma unroll
    for (int i = 0; i < NS; ++i) Lk[(size_t)(ND + NO + i) * ld] = y[i];
    if (!ok) { stage.wait(); return false; }         // the inertia test failed: the caller retries with a larger delta
    if (k == nb - 1) break;
    // next diagonal block and right-hand side
    stage.wait();
    double bn[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      const int u = B.var[(k + 1) * NS + i];
#pragma unroll
      for (int j = 0; j <= i; ++j) sSn[(size_t)BT_LI(i, j) * ss] = stage.get(NO + BT_LI(i, j));
      if (u >= 0) { sSn[(size_t)BT_LI(i, i) * ss] += delta; bn[i] = rhs[(size_t)u * ld + p]; } else bn[i] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < NXR; ++i) {
      // row i of L_{k+1,k} = O_k L_kk^-T
      double lo[NS];
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        double v = stage.get(i * NS + j);
#pragma unroll
        for (int m = 0; m < j; ++m) v -= lo[m] * L[BT_LI(j, m)];
        lo[j] = v * dinv[j];
      }
      double by = 0.0;
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        sLo[(size_t)(i * NS + j) * ss] = lo[j];
        Lk[(size_t)(ND + i * NS + j) * ld] = lo[j];
        by += lo[j] * y[j];
      }
      bn[i] -= by;
#pragma unroll
      for (int i2 = 0; i2 <= i; ++i2) {
        double dot = 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j) dot += lo[j] * sLo[(size_t)(i2 * NS + j) * ss];
        sSn[(size_t)BT_LI(i, i2) * ss] -= dot;
      }
    }
#pragma unroll
    for (int q = 0; q < ND; ++q) L[q] = sSn[(size_t)q * ss];
#pragma unroll
    for (int i = 0; i < NS; ++i) y[i] = bn[i];
  }
  // backward pass  x_k = L_kk^-T (y_k - L_{k+1,k}^T x_{k+1}); the factor blocks of stage k - 1 arrive while stage k is computed
  double xn[NS];
  double amax = 0.0;
  bool fin = true;
#pragma unroll
  for (int i = 0; i < NS; ++i) xn[i] = 0.0;
  stage.fetch(Lf + (size_t)(nb - 1) * SLL * ld + p, ld, SLL);
  for (int k = nb - 1; k >= 0; --k) {
    stage.wait();
    double t[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) t[i] = stage.get(ND + NO + i);
    if (k < nb - 1) {
#pragma unroll
      for (int i = 0; i < NXR; ++i)
#pragma unroll
        for (int j = 0; j < NS; ++j) t[j] -= stage.get(ND + i * NS + j) * xn[i];
    }
#pragma unroll
    for (int q = 0; q < ND; ++q) L[q] = stage.get(q);
    if (k > 0) stage.fetch(Lf + (size_t)(k - 1) * SLL * ld + p, ld, SLL);
#pragma unroll
    for (int i = NS - 1; i >= 0; --i) {
      double v = t[i];
#pragma unroll
      for (int m = i + 1; m < NS; ++m) v -= L[BT_LI(m, i)] * xn[m];
      xn[i] = v / L[BT_LI(i, i)];
    }
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      const int u = B.var[k * NS + i];
      if (u >= 0) {
        x[(size_t)u * ld + p] = xn[i];
        if (!ip_finite(xn[i])) fin = false;
        amax = ip_max(amax, fabs(xn[i]));
      }
    }
  }
  return ok && fin && amax < xmax;
}

// delta of attempt a (0 = the value that worked last time) of the regularisation search below
IP_HD double bt_ladder_delta(double delta0, int a) {
  double d = delta0;
  for (int i = 0; i < a; ++i) d = ip_min(ip_max(d * 8.0, 1e-4), 1e8);
  return d;
}

// Newton step of the condensed KKT system with the inertia-correcting regularisation of the reference solver:
// delta starts at the value that worked last time, grows x8 (from 1e-4) until the factorisation succeeds (<= 16 attempts), and a
// problem that needed more than last time is solved once more with twice the value that first passed.
template <int NS, int NXR>
IP_HD void bt_kkt_body(const BtTables& B, const double* K, double* Lf, const double* rhs, double* dw, double* dw_alt, size_t ld, size_t p,
                       int n_unknown, double delta0, int skip, double* dwt_out, double* sc, int ss) {
  if (skip) {
    for (int c = 0; c < n_unknown; ++c) dw[(size_t)c * ld + p] = 0.0;
    *dwt_out = delta0;
    return;
  }
  double dwt = delta0;
  bool good = false;
  for (int attempt = 0; attempt < 16; ++attempt) {
    good = bt_solve_attempt<NS, NXR>(B, K, Lf, rhs, dw, ld, p, p, dwt, 1e3, sc, ss);
    if (good) break;
    dwt = ip_min(ip_max(dwt * 8.0, 1e-4), 1e8);
  }
  if (!good)
    for (int c = 0; c < n_unknown; ++c) dw[(size_t)c * ld + p] = 0.0;
  if (dwt > delta0) {
    const double d2 = 2.0 * dwt;
    if (bt_solve_attempt<NS, NXR>(B, K, Lf, rhs, dw_alt, ld, p, p, d2, 1e3, sc, ss)) {
      for (int c = 0; c < n_unknown; ++c) dw[(size_t)c * ld + p] = dw_alt[(size_t)c * ld + p];
      dwt = d2;
    }
  }
  *dwt_out = dwt;
}

// =====================================================================================================================
// The iteration, written against a backend X that runs each body for problems 0..P-1 (CUDA kernels or host loops)
// =====================================================================================================================
struct IpOptions { double tol; int max_iter; double mu0; int ls_multipliers; int compact; int verbose; };
struct IpStats { int iterations, evaluations, hessians, trials, compactions; long long trial_problems; };

// Backtracking line search, one trial after the other: after the first trial only the problems whose step was refused are evaluated
// again.  (The CUDA backend evaluates all remaining step lengths of the refused problems in ONE batch instead - same verdicts.)
template <class X>
int ip_line_search_sequential(X& x, size_t P, IpStats* st) {
  size_t n_ls = P;
  for (int ls = 0; ls < IP_LS_TRIALS; ++ls) {
    size_t rejected = 0;
    if (x.trial(n_ls, ls, &rejected)) return 1;
    ++st->trials; st->trial_problems += (long long)n_ls;
    if (rejected == 0) break;
    n_ls = rejected;
  }
  return 0;
}

template <class X>
int ip_solve_loop(X& x, size_t P0, const IpOptions& opt, IpStats* stats) {
  size_t P = P0;
  IpStats st = {0, 0, 0, 0, 0, 0};
  if (x.eval_full(P)) return 1;
  ++st.evaluations;
  if (x.init(P, opt.mu0, opt.max_iter)) return 1;
  size_t n_done = 0;
  for (int it = 0; it < opt.max_iter; ++it) {
    if (opt.compact && it > 0 && it % 10 == 0 && 4 * n_done >= P) {
      // finished problems leave the working set: an iteration costs an evaluation, a Hessian and a factorisation per ACTIVE problem
      size_t newP = 0;
      if (x.compact(P, &newP)) return 1;
      P = newP; n_done = 0; ++st.compactions;
      if (P == 0) break;
      if (x.eval_full(P)) return 1;
      ++st.evaluations;
    }
    if (x.residual(P, it, opt.tol, &n_done)) return 1;
    st.iterations = it + 1;
    if (opt.verbose) x.report(P, it, n_done);
    if (n_done == P) break;
    if (x.hessian(P)) return 1;
    ++st.hessians;
    if (x.kkt_solve(P)) return 1;
    if (x.step(P)) return 1;
    if (x.line_search(P, &st)) return 1;
    if (x.update(P)) return 1;
    if (x.eval_full(P)) return 1;
    ++st.evaluations;
    if (opt.ls_multipliers && x.lsq_multipliers(P)) return 1;
  }
  if (P > 0 && x.flush_all(P)) return 1;
  if (stats) *stats = st;
  return 0;
}
