// CPU emulation of the device interior-point solver: the SAME per-problem bodies (csrc/ip_core.cuh), table builders
// (csrc/ip_tables.hpp) and iteration loop that the CUDA kernels of csrc/ip_solver.cu wrap, run as plain loops over the problems,
// with the NLP evaluation supplied by the caller (the numpy oracle, through ctypes callbacks).  TEST INFRASTRUCTURE ONLY: it pins
// the solver logic and the block-tridiagonal linear algebra on a machine without a GPU; it is not part of libnlo_b200.so and
// nothing in the product loads it.     g++ -O2 -shared -fPIC -o libip_emul.so ip_host_emul.cpp
#include "../../nlotrajectories_b200/csrc/ip_core.cuh"
#include "../../nlotrajectories_b200/csrc/ip_tables.hpp"
#include <cstdio>
#include <cstring>
#include <vector>

extern "C" {
typedef int (*ip_eval_cb)(const float* w32, size_t P, size_t ld, float* g, float* jac, float* f, float* grad);
typedef int (*ip_hess_cb)(const float* w32, const float* lam32, size_t P, size_t ld, float* hess);
}

namespace {

struct Emul {
  IpHostTables HT; BtHost kkt_h, lsq_h;
  IpTables T; BtTables KB, LB;
  size_t ld;
  std::vector<double> state[2], f64, K, L, out;
  std::vector<int> istate[2], i32, iout;
  std::vector<float> f32;
  IpState S; IpWork W; IpOut O;
  size_t state_rows; int cur;
  ip_eval_cb eval; ip_hess_cb hess;

  void bind_state(int which) {
    double* b = state[which].data();
    S.ld = ld;
    S.w = b; b += (size_t)T.n_w * ld; S.s = b; b += (size_t)T.nI * ld; S.zl = b; b += (size_t)T.nI * ld; S.zu = b; b += (size_t)T.nI * ld;
    S.lamE = b; b += (size_t)T.nE * ld; S.mu = b; b += ld; S.nu = b; b += ld; S.delta_w = b; b += ld; S.err0 = b; b += ld; S.f_mark = b;
    int* ib = istate[which].data();
    S.iters = ib; S.stalled = ib + ld; S.orig = ib + 2 * ld; S.done = ib + 3 * ld;
    cur = which;
  }
};

void bind_bt(BtTables& B, const BtHost& h) {
  B.nb = h.nb; B.NS = h.NS; B.NXR = h.NXR; B.SLK = h.SLK; B.SLL = h.SLL;
  B.var = h.var.data(); B.term_ptr = h.term_ptr.data(); B.terms = h.terms.data();
}

template <int NS, int NXR>
void kkt_all(Emul& e, size_t P, const double* delta_in, const int* skip, double* dwt_out) {
  double sc[2 * (NXR * NS + NS * (NS + 1) / 2) + NS];
  for (size_t p = 0; p < P; ++p) {
    double dwt;
    bt_kkt_body<NS, NXR>(e.KB, e.K.data(), e.L.data(), e.W.rhs, e.W.dw, e.W.dw_alt, e.ld, p, e.T.n_w, delta_in[p], skip ? skip[p] : 0, &dwt, sc, 1);
    dwt_out[p] = dwt;
  }
}
template <int NS, int NXR>
void spd_all(Emul& e, size_t P) {
  double sc[2 * (NXR * NS + NS * (NS + 1) / 2) + NS];
  for (size_t p = 0; p < P; ++p)
    if (!bt_solve_attempt<NS, NXR>(e.LB, e.K.data(), e.L.data(), e.W.v, e.W.lam_ls, e.ld, p, p, e.W.eps_ls[p], HUGE_VAL, sc, 1))
      for (int q = 0; q < e.T.nE; ++q) e.W.lam_ls[(size_t)q * e.ld + p] = e.S.lamE[(size_t)q * e.ld + p];
}

#define IP_KKT_SIZES(X) X(5, 3) X(6, 3) X(6, 4) X(7, 4) X(7, 5) X(8, 5) X(9, 7) X(10, 7)
#define IP_LSQ_SIZES(X) X(3, 3) X(4, 4) X(5, 5) X(7, 7)

struct HostBackend {
  Emul& e;
  int verbose;
  IpSolo solo;
  int eval_full(size_t P) { return e.eval(e.W.w32, P, e.ld, e.W.g, e.W.jac, e.W.f, e.W.grad); }
  int init(size_t P, double mu0, int max_iter) { for (size_t p = 0; p < P; ++p) ip_init_body(e.T, e.S, e.W, p, mu0, max_iter, solo); return 0; }
  int residual(size_t P, int it, double tol, size_t* n_done) {
    size_t n = 0;
    for (size_t p = 0; p < P; ++p) n += ip_residual_body(e.T, e.S, e.W, p, it, tol, solo);
    *n_done = n;
    return 0;
  }
  int hessian(size_t P) { return e.hess(e.W.w32, e.W.lam32, P, e.ld, e.W.hess); }
  void assemble(const BtTables& B, size_t P) {
    const int n_slots = B.nb * B.SLK;
    for (int s = 0; s < n_slots; ++s)
      for (size_t p = 0; p < P; ++p) e.K[(size_t)s * e.ld + p] = bt_slot_value(B, s, e.W.jac, e.W.hess, e.W.omega, e.ld, p);
  }
  int kkt_solve(size_t P) {
    assemble(e.KB, P);
#define IP_CASE(NS_, NXR_) if (e.KB.NS == NS_ && e.KB.NXR == NXR_) { kkt_all<NS_, NXR_>(e, P, e.S.delta_w, e.S.done, e.W.dwt); return 0; }
    IP_KKT_SIZES(IP_CASE)
#undef IP_CASE
    return 1;
  }
  int step(size_t P) { for (size_t p = 0; p < P; ++p) ip_step_body(e.T, e.S, e.W, p, solo); return 0; }
  int trial(size_t n, int ls, size_t* rejected) {
    const int* list = ls > 0 ? e.W.ls_list[(ls + 1) & 1] : nullptr;
    int* next = e.W.ls_list[ls & 1];
    for (int c = 0; c < e.T.n_w; ++c)
      for (size_t q = 0; q < n; ++q) {
        const size_t p = list ? (size_t)list[q] : q;
        e.W.wt32[(size_t)c * e.ld + q] = (float)(e.S.w[(size_t)c * e.ld + p] + e.W.alpha[p] * e.W.dw[(size_t)c * e.ld + p]);
      }
    if (e.eval(e.W.wt32, n, e.ld, e.W.gt, nullptr, e.W.ft, nullptr)) return 1;
    size_t m = 0;
    for (size_t q = 0; q < n; ++q) {
      const size_t p = list ? (size_t)list[q] : q;
      if (ip_merit_body(e.T, e.S, e.W, p, q, solo)) next[m++] = (int)p;
    }
    *rejected = m;
    return 0;
  }
  int line_search(size_t P, IpStats* st) { return ip_line_search_sequential(*this, P, st); }
  int update(size_t P) { for (size_t p = 0; p < P; ++p) ip_update_body(e.T, e.S, e.W, p, solo); return 0; }
  int lsq_multipliers(size_t P) {
    if (e.T.nE == 0) return 0;
    for (size_t p = 0; p < P; ++p) ip_lsq_prep_body(e.T, e.S, e.W, p, solo);
    assemble(e.LB, P);
    bool hit = false;
#define IP_CASE(NS_, NXR_) if (!hit && e.LB.NS == NS_) { spd_all<NS_, NXR_>(e, P); hit = true; }
    IP_LSQ_SIZES(IP_CASE)
#undef IP_CASE
    if (!hit) return 1;
    for (size_t p = 0; p < P; ++p) ip_lsq_choose_body(e.T, e.S, e.W, p, solo);
    return 0;
  }
  int flush_all(size_t P) { for (size_t p = 0; p < P; ++p) ip_flush_body(e.T, e.S, e.W, e.O, p, solo); return 0; }
  int compact(size_t P, size_t* newP) {
    std::vector<int> keep;
    for (size_t p = 0; p < P; ++p) { if (e.S.done[p]) ip_flush_body(e.T, e.S, e.W, e.O, p, solo); else keep.push_back((int)p); }
    *newP = keep.size();
    if (keep.empty()) return 0;
    const int other = 1 - e.cur;
    for (size_t r = 0; r < e.state_rows; ++r)
      for (size_t q = 0; q < keep.size(); ++q) e.state[other][r * e.ld + q] = e.state[e.cur][r * e.ld + keep[q]];
    for (size_t r = 0; r < 4; ++r)
      for (size_t q = 0; q < keep.size(); ++q) e.istate[other][r * e.ld + q] = e.istate[e.cur][r * e.ld + keep[q]];
    e.bind_state(other);
    for (int c = 0; c < e.T.n_w; ++c)
      for (size_t q = 0; q < keep.size(); ++q) e.W.w32[(size_t)c * e.ld + q] = (float)e.S.w[(size_t)c * e.ld + q];
    return 0;
  }
  void report(size_t P, int it, size_t n_done) { if (verbose) fprintf(stderr, "[ip_emul] it %3d  active %zu  done %zu\n", it, P, n_done); }
};

void setup(Emul& e, int n_w, int n_g, const int* jcolind, const int* jrow, int nnzh, const int* hcolind, const int* hrow, const double* lb,
           const double* ub, const IpStages& stg, size_t P, std::string* err) {
  ip_build_tables(n_w, n_g, jcolind, jrow, nnzh, hcolind, hrow, lb, ub, &e.HT);
  if (!ip_build_kkt_system(e.HT, stg, &e.kkt_h)) { *err = e.kkt_h.error; return; }
  if (e.HT.nE > 0 && !ip_build_lsq_system(e.HT, stg, &e.lsq_h)) { *err = e.lsq_h.error; return; }
  const IpHostTables& H = e.HT;
  IpTables& T = e.T;
  T.n_w = H.n_w; T.n_g = H.n_g; T.nnz = H.nnz; T.nnzh = H.nnzh; T.nE = H.nE; T.nI = H.nI;
  T.rkind = H.rkind.data(); T.ridx = H.ridx.data(); T.lb = H.lb.data(); T.ub = H.ub.data(); T.colind = H.colind.data(); T.row = H.row.data();
  T.rptr = H.rptr.data(); T.rnz = H.rnz.data(); T.rcol = H.rcol.data(); T.hcolind = H.hcolind.data(); T.hrow = H.hrow.data();
  bind_bt(e.KB, e.kkt_h); bind_bt(e.LB, e.lsq_h);
  const size_t ld = e.ld = P;
  e.state_rows = (size_t)T.n_w + 3 * (size_t)T.nI + T.nE + 5;
  for (int b = 0; b < 2; ++b) { e.state[b].assign(e.state_rows * ld, 0.0); e.istate[b].assign(4 * ld, 0); }
  const size_t f32_rows = 1 + (size_t)T.n_w + T.n_g + T.nnz + T.nnzh + T.n_w + T.n_g + T.n_w + 1 + T.n_g;
  const size_t f64_rows = (size_t)T.n_g * 2 + T.n_w * 3 + T.nI * 3 + T.nE * 2 + 6;
  e.f32.assign(f32_rows * ld, 0.f); e.f64.assign(f64_rows * ld, 0.0); e.i32.assign(3 * ld + 8, 0);
  e.K.assign(std::max((size_t)e.kkt_h.nb * e.kkt_h.SLK, (size_t)e.lsq_h.nb * e.lsq_h.SLK) * ld, 0.0);
  e.L.assign(std::max((size_t)e.kkt_h.nb * e.kkt_h.SLL, (size_t)e.lsq_h.nb * e.lsq_h.SLL) * ld, 0.0);
  e.out.assign(((size_t)T.n_w + T.n_g + 3) * ld, 0.0); e.iout.assign(2 * ld, 0);
  IpWork& W = e.W;
  W.ld = ld;
  float* f = e.f32.data();
  W.f = f; f += ld; W.grad = f; f += (size_t)T.n_w * ld; W.g = f; f += (size_t)T.n_g * ld; W.jac = f; f += (size_t)T.nnz * ld;
  W.hess = f; f += (size_t)T.nnzh * ld; W.w32 = f; f += (size_t)T.n_w * ld; W.lam32 = f; f += (size_t)T.n_g * ld;
  W.wt32 = f; f += (size_t)T.n_w * ld; W.ft = f; f += ld; W.gt = f;
  double* d = e.f64.data();
  W.omega = d; d += (size_t)T.n_g * ld; W.v = d; d += (size_t)T.n_g * ld; W.rhs = d; d += (size_t)T.n_w * ld; W.dw = d; d += (size_t)T.n_w * ld;
  W.dw_alt = d; d += (size_t)T.n_w * ld; W.ds = d; d += (size_t)T.nI * ld; W.dzl = d; d += (size_t)T.nI * ld; W.dzu = d; d += (size_t)T.nI * ld;
  W.dlamE = d; d += (size_t)T.nE * ld; W.lam_ls = d; d += (size_t)T.nE * ld;
  W.dwt = d; d += ld; W.alpha = d; d += ld; W.alpha_d = d; d += ld; W.phi0 = d; d += ld; W.dphi = d; d += ld; W.eps_ls = d; W.viol = nullptr;
  W.accepted = e.i32.data(); W.ls_list[0] = e.i32.data() + ld; W.ls_list[1] = e.i32.data() + 2 * ld; W.counters = e.i32.data() + 3 * ld;
  IpOut& O = e.O;
  O.ld = ld; O.w = e.out.data(); O.lam = O.w + (size_t)T.n_w * ld; O.f = O.lam + (size_t)T.n_g * ld; O.viol = O.f + ld; O.err = O.viol + ld;
  O.iters = e.iout.data(); O.status = e.iout.data() + ld;
  e.bind_state(0);
}

char g_err[512] = "";

}  // namespace

extern "C" {

const char* ip_emul_last_error() { return g_err; }

// Solve P problems.  w0 / w_out: [P][n_w] doubles; lam_out [P][n_g] or NULL.  stats: iterations, evaluations, hessians, trials, compactions.
int ip_emul_solve(int n_w, int n_g, const int* jcolind, const int* jrow, int nnzh, const int* hcolind, const int* hrow, const double* lb,
                  const double* ub, int N, int nx, int nu, int use_slack, int n_term, int g_off_dyn, const double* w0, size_t P, double tol,
                  int max_iter, double mu0, int ls_multipliers, int compact, int verbose, ip_eval_cb eval, ip_hess_cb hess, double* w_out,
                  double* f_out, double* viol_out, double* err_out, int* iters_out, int* status_out, double* lam_out, int* stats_out) {
  Emul e;
  std::string err;
  IpStages stg = {N, nx, nu, use_slack, n_term, g_off_dyn};
  setup(e, n_w, n_g, jcolind, jrow, nnzh, hcolind, hrow, lb, ub, stg, P, &err);
  if (!err.empty()) { snprintf(g_err, sizeof(g_err), "%s", err.c_str()); return 1; }
  e.eval = eval; e.hess = hess;
  for (size_t p = 0; p < P; ++p) {
    e.S.orig[p] = (int)p;
    for (int c = 0; c < n_w; ++c) { e.S.w[(size_t)c * P + p] = w0[p * n_w + c]; e.W.w32[(size_t)c * P + p] = (float)w0[p * n_w + c]; }
  }
  HostBackend x{e, verbose};
  IpOptions opt = {tol, max_iter, mu0, ls_multipliers, compact, verbose};
  IpStats st;
  if (ip_solve_loop(x, P, opt, &st)) { snprintf(g_err, sizeof(g_err), "solver loop failed"); return 1; }
  for (size_t p = 0; p < P; ++p) {
    for (int c = 0; c < n_w; ++c) w_out[p * n_w + c] = e.O.w[(size_t)c * P + p];
    if (lam_out) for (int r = 0; r < n_g; ++r) lam_out[p * n_g + r] = e.O.lam[(size_t)r * P + p];
    f_out[p] = e.O.f[p]; viol_out[p] = e.O.viol[p]; err_out[p] = e.O.err[p]; iters_out[p] = e.O.iters[p]; status_out[p] = e.O.status[p];
  }
  if (stats_out) { stats_out[0] = st.iterations; stats_out[1] = st.evaluations; stats_out[2] = st.hessians; stats_out[3] = st.trials; stats_out[4] = st.compactions; }
  return 0;
}

// (H + J^T diag(omega) J + delta I) dw = rhs per problem through the block-tridiagonal path: all arrays variable-major [rows][P]
int ip_emul_kkt_step(int n_w, int n_g, const int* jcolind, const int* jrow, int nnzh, const int* hcolind, const int* hrow, const double* lb,
                     const double* ub, int N, int nx, int nu, int use_slack, int n_term, int g_off_dyn, size_t P, const float* jac,
                     const float* hess, const double* omega, const double* rhs, const double* delta_in, double* dw, double* delta_out) {
  Emul e;
  std::string err;
  IpStages stg = {N, nx, nu, use_slack, n_term, g_off_dyn};
  setup(e, n_w, n_g, jcolind, jrow, nnzh, hcolind, hrow, lb, ub, stg, P, &err);
  if (!err.empty()) { snprintf(g_err, sizeof(g_err), "%s", err.c_str()); return 1; }
  memcpy(e.W.jac, jac, (size_t)e.T.nnz * P * sizeof(float));
  memcpy(e.W.hess, hess, (size_t)e.T.nnzh * P * sizeof(float));
  memcpy(e.W.omega, omega, (size_t)n_g * P * sizeof(double));
  memcpy(e.W.rhs, rhs, (size_t)n_w * P * sizeof(double));
  HostBackend x{e, 0};
  x.assemble(e.KB, P);
  bool hit = false;
#define IP_CASE(NS_, NXR_) if (!hit && e.KB.NS == NS_ && e.KB.NXR == NXR_) { kkt_all<NS_, NXR_>(e, P, delta_in, nullptr, delta_out); hit = true; }
  IP_KKT_SIZES(IP_CASE)
#undef IP_CASE
  if (!hit) { snprintf(g_err, sizeof(g_err), "no kernel for block size %d / %d", e.KB.NS, e.KB.NXR); return 1; }
  memcpy(dw, e.W.dw, (size_t)n_w * P * sizeof(double));
  return 0;
}

}  // extern "C"
