// Host-side construction of the solver's index tables (plain C++; shared by ip_solver.cu and the CPU emulation used in tests).
#pragma once
#include <algorithm>
#include <cmath>
#include <map>
#include <string>
#include <vector>

struct IpHostTables {
  int n_w = 0, n_g = 0, nnz = 0, nnzh = 0, nE = 0, nI = 0;
  std::vector<int> rkind, ridx, colind, row, rptr, rnz, rcol, hcolind, hrow;
  std::vector<double> lb, ub;
};

// Stage structure of the transcription (core/runner.py:44-103): w = [x_0 .. x_N ; u_0 .. u_{N-1} ; slack_0 .. slack_N],
// g = [init (nx) ; terminal (n_term) ; Euler defects (N * nx) ; ...].
struct IpStages { int N, nx, nu, use_slack, n_term, g_off_dyn; };

inline void ip_build_tables(int n_w, int n_g, const int* jcolind, const int* jrow, int nnzh, const int* hcolind, const int* hrow,
                            const double* lb, const double* ub, IpHostTables* T) {
  T->n_w = n_w; T->n_g = n_g; T->nnz = jcolind[n_w]; T->nnzh = nnzh;
  T->colind.assign(jcolind, jcolind + n_w + 1); T->row.assign(jrow, jrow + T->nnz);
  T->hcolind.assign(hcolind, hcolind + n_w + 1); T->hrow.assign(hrow, hrow + nnzh);
  T->lb.assign(lb, lb + n_g); T->ub.assign(ub, ub + n_g);
  T->rkind.resize(n_g); T->ridx.resize(n_g);
  T->nE = T->nI = 0;
  for (int r = 0; r < n_g; ++r) {
    const bool eq = lb[r] == ub[r];
    T->rkind[r] = eq ? 0 : 1;
    T->ridx[r] = eq ? T->nE++ : T->nI++;
  }
  T->rptr.assign(n_g + 1, 0);
  for (int z = 0; z < T->nnz; ++z) T->rptr[jrow[z] + 1]++;
  for (int r = 0; r < n_g; ++r) T->rptr[r + 1] += T->rptr[r];
  T->rnz.resize(T->nnz); T->rcol.resize(T->nnz);
  std::vector<int> fill(T->rptr.begin(), T->rptr.end() - 1);
  for (int c = 0; c < n_w; ++c)
    for (int z = jcolind[c]; z < jcolind[c + 1]; ++z) { const int q = fill[jrow[z]]++; T->rnz[q] = z; T->rcol[q] = c; }
}

struct BtHost {
  int nb = 0, NS = 0, NXR = 0, SLK = 0, SLL = 0;
  std::vector<int> var, term_ptr, terms;
  std::string error;
};

namespace ip_detail {
struct BtBuilder {
  BtHost* B;
  std::vector<int> blk, loc;                       // unknown -> (block, local row)
  std::map<int, std::vector<int>> slot_terms;      // slot -> flat terms
  bool add(int ua, int ub, int kind, int a, int b, int c) {
    int ka = blk[ua], ia = loc[ua], kb = blk[ub], ib = loc[ub];
    const int ND = B->NS * (B->NS + 1) / 2;
    int slot;
    if (ka == kb) {
      const int i = std::max(ia, ib), j = std::min(ia, ib);
      slot = ka * B->SLK + i * (i + 1) / 2 + j;
    } else {
      if (ka > kb) { std::swap(ka, kb); std::swap(ia, ib); }
      if (kb != ka + 1) { B->error = "coupling between blocks " + std::to_string(ka) + " and " + std::to_string(kb) + " is not tridiagonal"; return false; }
      if (ib >= B->NXR) { B->error = "sub-diagonal row " + std::to_string(ib) + " beyond the " + std::to_string(B->NXR) + " coupled rows"; return false; }
      slot = ka * B->SLK + ND + ib * B->NS + ia;
    }
    auto& v = slot_terms[slot];
    v.push_back(kind); v.push_back(a); v.push_back(b); v.push_back(c);
    return true;
  }
  void finish() {
    const int n_slots = B->nb * B->SLK;
    B->term_ptr.assign(n_slots + 1, 0);
    B->terms.clear();
    for (int s = 0; s < n_slots; ++s) {
      auto it = slot_terms.find(s);
      if (it != slot_terms.end()) B->terms.insert(B->terms.end(), it->second.begin(), it->second.end());
      B->term_ptr[s + 1] = (int)B->terms.size() / 4;
    }
  }
};
}  // namespace ip_detail

// Condensed KKT matrix  H + J^T diag(omega) J  in stage order: block k = (x_k, u_k, slack_k); u_N is a dummy.
inline bool ip_build_kkt_system(const IpHostTables& T, const IpStages& S, BtHost* B) {
  const int ns = S.nx + S.nu + (S.use_slack ? 1 : 0);
  const int n_X = S.nx * (S.N + 1), n_U = S.nu * S.N;
  B->nb = S.N + 1; B->NS = ns; B->NXR = S.nx;
  const int ND = ns * (ns + 1) / 2;
  B->SLK = ND + B->NXR * ns; B->SLL = B->SLK + ns;
  B->var.assign((size_t)B->nb * ns, -1);
  ip_detail::BtBuilder bb; bb.B = B; bb.blk.assign(T.n_w, -1); bb.loc.assign(T.n_w, -1);
  for (int k = 0; k <= S.N; ++k) {
    for (int i = 0; i < S.nx; ++i) B->var[(size_t)k * ns + i] = k * S.nx + i;
    if (k < S.N) for (int i = 0; i < S.nu; ++i) B->var[(size_t)k * ns + S.nx + i] = n_X + k * S.nu + i;
    if (S.use_slack) B->var[(size_t)k * ns + S.nx + S.nu] = n_X + n_U + k;
    for (int i = 0; i < ns; ++i) {
      const int u = B->var[(size_t)k * ns + i];
      if (u >= 0) { bb.blk[u] = k; bb.loc[u] = i; }
      else bb.slot_terms[k * B->SLK + i * (i + 1) / 2 + i] = {3, 0, 0, 0};
    }
  }
  for (int u = 0; u < T.n_w; ++u) if (bb.blk[u] < 0) { B->error = "decision variable without a stage"; return false; }
  for (int c = 0; c < T.n_w; ++c)
    for (int z = T.hcolind[c]; z < T.hcolind[c + 1]; ++z) if (!bb.add(T.hrow[z], c, 0, z, 0, 0)) return false;
  for (int r = 0; r < T.n_g; ++r)
    for (int a = T.rptr[r]; a < T.rptr[r + 1]; ++a)
      for (int b = a; b < T.rptr[r + 1]; ++b) if (!bb.add(T.rcol[a], T.rcol[b], 1, T.rnz[a], T.rnz[b], r)) return false;
  bb.finish();
  return true;
}

// Gram matrix JE JE^T of the equality rows in constraint-stage order: block 0 = x_0 pin, block 1 + k = Euler defects of interval k,
// block N + 1 = terminal pin (padded with dummies).  Unknown index = index within the equality list.
inline bool ip_build_lsq_system(const IpHostTables& T, const IpStages& S, BtHost* B) {
  const int ns = S.nx;
  B->nb = S.N + 2; B->NS = ns; B->NXR = ns;
  const int ND = ns * (ns + 1) / 2;
  B->SLK = ND + ns * ns; B->SLL = B->SLK + ns;
  B->var.assign((size_t)B->nb * ns, -1);
  ip_detail::BtBuilder bb; bb.B = B; bb.blk.assign(T.nE, -1); bb.loc.assign(T.nE, -1);
  for (int r = 0; r < T.n_g; ++r) {
    if (T.rkind[r] != 0) continue;
    int k, i;
    if (r < S.nx) { k = 0; i = r; }
    else if (r < S.nx + S.n_term) { k = S.N + 1; i = r - S.nx; }
    else if (r >= S.g_off_dyn && r < S.g_off_dyn + S.N * S.nx) { k = 1 + (r - S.g_off_dyn) / S.nx; i = (r - S.g_off_dyn) % S.nx; }
    else { B->error = "equality row " + std::to_string(r) + " outside the init / terminal / defect blocks"; return false; }
    const int e = T.ridx[r];
    bb.blk[e] = k; bb.loc[e] = i; B->var[(size_t)k * ns + i] = e;
  }
  for (int k = 0; k < B->nb; ++k)
    for (int i = 0; i < ns; ++i) if (B->var[(size_t)k * ns + i] < 0) bb.slot_terms[k * B->SLK + i * (i + 1) / 2 + i] = {3, 0, 0, 0};
  for (int c = 0; c < T.n_w; ++c)
    for (int a = T.colind[c]; a < T.colind[c + 1]; ++a) {
      if (T.rkind[T.row[a]] != 0) continue;
      for (int b = a; b < T.colind[c + 1]; ++b) {
        if (T.rkind[T.row[b]] != 0) continue;
        if (!bb.add(T.ridx[T.row[a]], T.ridx[T.row[b]], 2, a, b, 0)) return false;
      }
    }
  bb.finish();
  return true;
}
