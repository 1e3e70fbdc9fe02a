// K2-K5: batched NLP evaluation for P independent problems, structure-of-arrays (variable-major).
//
// Reference (src/nlotrajectories/...):
//   core/runner.py:44-103        decision vector / constraint order / objective
//   core/dynamics.py:33-148      f(x,u) of the six models        -> K2 Euler defects + banded Jacobian values
//   core/geometry.py:59-144      footprint transform, SDF rows   -> K3
//   core/utils.py:18-33          soft_min (alpha = 10, un-stabilised log-sum-exp)
//   core/sdf/casadi.py:27-45,377-390  circle SDF + soft-min union -> K5
//   core/runner.py:80-98         objective                        -> K4
//
// Layout: element (v, problem i) of w / g / jac / grad_f lives at [v*ld + i]; consecutive threads
// handle consecutive problems so every global access is a fully coalesced 128-byte line.
// dg/dw values are written straight to their compressed-column slot through a small index table
// (nzmap: emission order -> CCS position) that is identical for all problems.
#include "nlo_common.cuh"
#include "nlp_internal.cuh"
#include <vector>
#include <algorithm>
#include <numeric>

// ---- structural tables -----------------------------------------------------
namespace {

// A / B: structural (row, column) pairs of df/dx, df/du; Ac / Bc: 1 where the entry is the constant 1 for every (x, u)
// (Dyn<>::eval below writes exactly 1.f there), 0 where it varies.
struct DynTable { int nx, nu, nA, nB; int A[10][2]; int B[4][2]; int Ac[10]; int Bc[4]; };
const DynTable kDyn[6] = {
    /* point_1st     */ {4, 2, 0, 2, {{0, 0}}, {{0, 0}, {1, 1}}, {0}, {1, 1}},
    /* point_2nd     */ {4, 2, 2, 2, {{0, 2}, {1, 3}}, {{2, 0}, {3, 1}}, {1, 1}, {1, 1}},
    /* unicycle      */ {3, 2, 2, 3, {{0, 2}, {1, 2}}, {{0, 0}, {1, 0}, {2, 1}}, {0, 0}, {0, 0, 1}},
    /* unicycle_2nd  */ {5, 2, 5, 2, {{0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 4}}, {{3, 0}, {4, 1}}, {0, 0, 0, 0, 1}, {1, 1}},
    /* ackermann     */ {4, 2, 3, 4, {{0, 2}, {1, 2}, {2, 3}}, {{0, 0}, {1, 0}, {2, 0}, {3, 1}}, {0, 0, 0}, {0, 0, 0, 1}},
    /* ackermann_2nd */ {7, 2, 10, 3, {{0, 2}, {0, 4}, {1, 2}, {1, 4}, {2, 3}, {2, 4}, {3, 6}, {4, 3}, {4, 4}, {4, 6}},
                         {{4, 0}, {5, 0}, {6, 1}}, {0, 0, 0, 0, 0, 0, 1, 0, 0, 0}, {0, 1, 1}},
};

}  // namespace

int nlo_nlp_build_layout(const nlo_nlp_desc* d, NlpDev* L, std::vector<int>* rows_ccs, std::vector<int>* cols_ccs,
                         std::vector<int>* nzmap, std::vector<int>* copy_row, std::vector<int>* copy_var,
                         std::vector<int>* copy_emit, std::vector<float>* const_ccs) {
  if (d->dynamics > 5) return nlo_fail("unknown dynamics id %u", d->dynamics);
  if (d->shape > 2) return nlo_fail("unknown shape id %u", d->shape);
  if (d->N < 1) return nlo_fail("N must be >= 1");
  const DynTable& T = kDyn[d->dynamics];
  NlpDev& l = *L;
  l.dyn = d->dynamics; l.shape = d->shape; l.N = d->N; l.nx = T.nx; l.nu = T.nu;
  l.use_slack = d->use_slack != 0; l.use_smooth = d->use_smooth != 0; l.enforce_heading = d->enforce_heading != 0;
  l.sdf_mode = d->sdf_mode; l.n_circles = d->n_circles;
  if (l.sdf_mode == NLO_SDF_CIRCLES && (l.n_circles < 1 || l.n_circles > NLO_MAX_CIRCLES))
    return nlo_fail("analytic mode needs 1..%d circles", NLO_MAX_CIRCLES);
  if (l.shape != NLO_SHAPE_DOT && T.nx < 3) return nlo_fail("polygon footprint needs a heading state");
  l.dt = d->dt; l.slack_penalty = d->slack_penalty; l.smooth_weight = d->smooth_weight; l.wheelbase = d->wheelbase;
  for (int c = 0; c < NLO_MAX_CIRCLES; ++c) {
    for (int q = 0; q < 4; ++q) l.circles[c][q] = d->circles[c][q];
    if (d->obstacle_kind[c] > NLO_OBST_SQUARE) return nlo_fail("unknown analytic obstacle kind %u", d->obstacle_kind[c]);
    l.okind[c] = (int)d->obstacle_kind[c];
  }
  const float hl = 0.5f * d->length, hw = 0.5f * d->width;
  if (l.shape == NLO_SHAPE_DOT) { l.nb = 1; l.bx[0] = l.by[0] = 0.f; }
  else if (l.shape == NLO_SHAPE_RECTANGLE) {            // core/geometry.py:125-135
    l.nb = 4; const float bx[4] = {-hl, -hl, hl, hl}, by[4] = {-hw, hw, hw, -hw};
    for (int c = 0; c < 4; ++c) { l.bx[c] = bx[c]; l.by[c] = by[c]; }
  } else {                                               // core/geometry.py:138-144
    l.nb = 3; const float bx[3] = {hl, -hl, -hl}, by[3] = {0.f, hw, -hw};
    for (int c = 0; c < 3; ++c) { l.bx[c] = bx[c]; l.by[c] = by[c]; }
  }
  const int N = l.N, nx = l.nx, nu = l.nu;
  l.n_X = nx * (N + 1); l.n_U = nu * N; l.n_w = l.n_X + l.n_U + (l.use_slack ? N + 1 : 0);
  l.rows_per_knot = (l.shape == NLO_SHAPE_DOT || l.use_slack) ? 1 : l.nb;
  l.nnz_sdf_row = (l.shape == NLO_SHAPE_DOT) ? 2 : (l.use_slack ? 4 : 3);
  int n_term = 0; for (int i = 0; i < nx; ++i) if (l.enforce_heading || i != 2) ++n_term;
  l.n_term = n_term;
  l.nA_off = 0; for (int a = 0; a < T.nA; ++a) if (T.A[a][0] != T.A[a][1]) ++l.nA_off;
  l.nnz_dyn = nx + nx + l.nA_off + T.nB;
  l.g_off_term = nx; l.g_off_dyn = nx + n_term; l.g_off_slack = l.g_off_dyn + N * nx;
  l.g_off_sdf = l.g_off_slack + (l.use_slack ? N + 1 : 0);
  l.g_off_ctrl = l.g_off_sdf + (N + 1) * l.rows_per_knot;
  l.n_g = l.g_off_ctrl + nu * N;
  // emission order == oracle/nlp_oracle.py::jac_pattern
  std::vector<int> rows, cols;
  std::vector<float> cval;                     // per emission: the value when it is the same for every w, NAN when it varies
  auto iX = [&](int i, int k) { return k * nx + i; };
  auto iU = [&](int i, int k) { return l.n_X + k * nu + i; };
  auto iS = [&](int k) { return l.n_X + l.n_U + k; };
  int r = 0;
  auto add_copy = [&](int row, int var) { copy_row->push_back(row); copy_var->push_back(var); copy_emit->push_back((int)rows.size()); rows.push_back(row); cols.push_back(var); cval.push_back(1.f); };
  for (int i = 0; i < nx; ++i) add_copy(r++, iX(i, 0));
  for (int i = 0; i < nx; ++i) if (l.enforce_heading || i != 2) add_copy(r++, iX(i, N));
  l.e_off_dyn = (int)rows.size();
  for (int k = 0; k < N; ++k) {
    for (int i = 0; i < nx; ++i) { rows.push_back(r + i); cols.push_back(iX(i, k + 1)); cval.push_back(1.f); }
    for (int i = 0; i < nx; ++i) {
      bool diag_varies = false;
      for (int a = 0; a < T.nA; ++a) if (T.A[a][0] == i && T.A[a][1] == i) diag_varies = true;
      rows.push_back(r + i); cols.push_back(iX(i, k)); cval.push_back(diag_varies ? NAN : -1.f);
    }
    for (int a = 0; a < T.nA; ++a) if (T.A[a][0] != T.A[a][1]) { rows.push_back(r + T.A[a][0]); cols.push_back(iX(T.A[a][1], k)); cval.push_back(T.Ac[a] ? -l.dt * 1.f : NAN); }
    for (int b = 0; b < T.nB; ++b) { rows.push_back(r + T.B[b][0]); cols.push_back(iU(T.B[b][1], k)); cval.push_back(T.Bc[b] ? -l.dt * 1.f : NAN); }
    r += nx;
  }
  if (l.use_slack) for (int k = 0; k <= N; ++k) add_copy(r++, iS(k));
  l.e_off_sdf = (int)rows.size();
  for (int k = 0; k <= N; ++k) {
    if (l.shape == NLO_SHAPE_DOT) { rows.push_back(r); cols.push_back(iX(0, k)); rows.push_back(r); cols.push_back(iX(1, k)); ++r; cval.push_back(NAN); cval.push_back(NAN); }
    else if (l.use_slack) {
      rows.push_back(r); cols.push_back(iX(0, k)); rows.push_back(r); cols.push_back(iX(1, k));
      rows.push_back(r); cols.push_back(iX(2, k)); rows.push_back(r); cols.push_back(iS(k)); ++r;
      cval.push_back(NAN); cval.push_back(NAN); cval.push_back(NAN); cval.push_back(1.f);
    } else {
      for (int c = 0; c < l.nb; ++c) { for (int q = 0; q < 3; ++q) { rows.push_back(r); cols.push_back(iX(q, k)); cval.push_back(NAN); } ++r; }
    }
  }
  for (int i = 0; i < nu; ++i) for (int k = 0; k < N; ++k) add_copy(r++, iU(i, k));
  if (r != l.n_g) return nlo_fail("internal: row count %d != n_g %d", r, l.n_g);
  l.nnz = (int)rows.size();
  std::vector<int> perm(l.nnz);
  std::iota(perm.begin(), perm.end(), 0);
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return cols[a] != cols[b] ? cols[a] < cols[b] : rows[a] < rows[b]; });
  nzmap->assign(l.nnz, 0); rows_ccs->resize(l.nnz); cols_ccs->resize(l.nnz);
  for (int pos = 0; pos < l.nnz; ++pos) { (*nzmap)[perm[pos]] = pos; (*rows_ccs)[pos] = rows[perm[pos]]; (*cols_ccs)[pos] = cols[perm[pos]]; }
  if ((int)cval.size() != l.nnz) return nlo_fail("internal: constant table size %zu != nnz %d", cval.size(), l.nnz);
  if (const_ccs) { const_ccs->assign(l.nnz, NAN); for (int pos = 0; pos < l.nnz; ++pos) (*const_ccs)[pos] = cval[perm[pos]]; }
  return 0;
}

// ---- K2: Euler defects + Jacobian values ---------------------------------------------------------
namespace {

template <int DYN> struct Dyn;
// Each model: f(x,u) and the A/B structural values in the order of kDyn (row-major over the listed pairs).
template <> struct Dyn<NLO_DYN_POINT_1ST> { static constexpr int nx = 4, nu = 2, nA = 0, nB = 2;
  __device__ static void eval(const float* x, const float* u, float, float* f, float* A, float* B) {
    f[0] = u[0]; f[1] = u[1]; f[2] = 0.f; f[3] = 0.f; B[0] = 1.f; B[1] = 1.f; } };
template <> struct Dyn<NLO_DYN_POINT_2ND> { static constexpr int nx = 4, nu = 2, nA = 2, nB = 2;
  __device__ static void eval(const float* x, const float* u, float, float* f, float* A, float* B) {
    f[0] = x[2]; f[1] = x[3]; f[2] = u[0]; f[3] = u[1]; A[0] = 1.f; A[1] = 1.f; B[0] = 1.f; B[1] = 1.f; } };
template <> struct Dyn<NLO_DYN_UNICYCLE> { static constexpr int nx = 3, nu = 2, nA = 2, nB = 3;
  __device__ static void eval(const float* x, const float* u, float, float* f, float* A, float* B) {
    float s, c; sincosf(x[2], &s, &c); const float v = u[0];
    f[0] = v * c; f[1] = v * s; f[2] = u[1]; A[0] = -v * s; A[1] = v * c; B[0] = c; B[1] = s; B[2] = 1.f; } };
template <> struct Dyn<NLO_DYN_UNICYCLE_2ND> { static constexpr int nx = 5, nu = 2, nA = 5, nB = 2;
  __device__ static void eval(const float* x, const float* u, float, float* f, float* A, float* B) {
    float s, c; sincosf(x[2], &s, &c); const float v = x[3];
    f[0] = v * c; f[1] = v * s; f[2] = x[4]; f[3] = u[0]; f[4] = u[1];
    A[0] = -v * s; A[1] = c; A[2] = v * c; A[3] = s; A[4] = 1.f; B[0] = 1.f; B[1] = 1.f; } };
template <> struct Dyn<NLO_DYN_ACKERMANN> { static constexpr int nx = 4, nu = 2, nA = 3, nB = 4;
  __device__ static void eval(const float* x, const float* u, float L, float* f, float* A, float* B) {
    float s, c; sincosf(x[2], &s, &c); const float t = tanf(x[3]), v = u[0], iL = 1.f / L;
    f[0] = v * c; f[1] = v * s; f[2] = v * t * iL; f[3] = u[1];
    A[0] = -v * s; A[1] = v * c; A[2] = v * (1.f + t * t) * iL; B[0] = c; B[1] = s; B[2] = t * iL; B[3] = 1.f; } };
// core/dynamics.py:131-148 including the slot quirk: v = x[4], psi_dot = x[6]
template <> struct Dyn<NLO_DYN_ACKERMANN_2ND> { static constexpr int nx = 7, nu = 2, nA = 10, nB = 3;
  __device__ static void eval(const float* x, const float* u, float L, float* f, float* A, float* B) {
    float s, c; sincosf(x[2], &s, &c);
    const float psi = x[3], v = x[4], pd = x[6], a = u[0], t = tanf(psi), iL = 1.f / L;
    const float q = 1.f / (1.f + psi * psi), sec2 = 1.f + t * t;
    f[0] = v * c; f[1] = v * s; f[2] = v * t * iL; f[3] = pd; f[4] = (pd * q * v + t * a) * iL; f[5] = a; f[6] = u[1];
    A[0] = -v * s; A[1] = c; A[2] = v * c; A[3] = s; A[4] = v * sec2 * iL; A[5] = t * iL; A[6] = 1.f;
    A[7] = (a * sec2 - 2.f * psi * v * pd * q * q) * iL; A[8] = pd * q * iL; A[9] = v * q * iL;
    B[0] = t * iL; B[1] = 1.f; B[2] = 1.f; } };

// The kernels below are written as device bodies that take their row / knot index k (uniform per block); one fused launch
// (nlp_phase0_kernel) runs them all, selecting the role by blockIdx.y range.
template <int DYN>
__device__ __forceinline__ void nlp_dyn_body(const NlpDev& L, const float* __restrict__ w, size_t P, size_t ld,
                                             float* __restrict__ g, float* __restrict__ jac, const int k) {
  using D = Dyn<DYN>;
  constexpr int nx = D::nx, nu = D::nu;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    float x[nx], xn[nx], u[nu], f[nx], A[D::nA > 0 ? D::nA : 1], B[D::nB];
#pragma unroll
    for (int i = 0; i < nx; ++i) { x[i] = w[(size_t)(k * nx + i) * ld + p]; xn[i] = w[(size_t)((k + 1) * nx + i) * ld + p]; }
#pragma unroll
    for (int i = 0; i < nu; ++i) u[i] = w[(size_t)(L.n_X + k * nu + i) * ld + p];
    D::eval(x, u, L.wheelbase, f, A, B);
    if (g) {
#pragma unroll
      for (int i = 0; i < nx; ++i) g[(size_t)(L.g_off_dyn + k * nx + i) * ld + p] = xn[i] - fmaf(L.dt, f[i], x[i]);
    }
    if (jac) {
      const int* __restrict__ nz = L.nzmap + L.e_off_dyn + k * L.nnz_dyn;
      int e = 0;
#pragma unroll
      for (int i = 0; i < nx; ++i) jac[(size_t)nz[e++] * ld + p] = 1.f;
      float diag[nx];
#pragma unroll
      for (int i = 0; i < nx; ++i) diag[i] = -1.f;
#pragma unroll
      for (int a = 0; a < D::nA; ++a) if (kDynA<DYN>(a, 0) == kDynA<DYN>(a, 1)) diag[kDynA<DYN>(a, 0)] = -1.f - L.dt * A[a];
#pragma unroll
      for (int i = 0; i < nx; ++i) jac[(size_t)nz[e++] * ld + p] = diag[i];
#pragma unroll
      for (int a = 0; a < D::nA; ++a) if (kDynA<DYN>(a, 0) != kDynA<DYN>(a, 1)) jac[(size_t)nz[e++] * ld + p] = -L.dt * A[a];
#pragma unroll
      for (int b = 0; b < D::nB; ++b) jac[(size_t)nz[e++] * ld + p] = -L.dt * B[b];
    }
  }
}

// ---- rows that are plain copies of a variable (init, terminal, slack >= 0, control box): dg/dw = 1 ----
__device__ __forceinline__ void nlp_copy_rows_body(const NlpDev& L, const float* __restrict__ w, size_t P, size_t ld,
                                                   float* __restrict__ g, float* __restrict__ jac, const int r) {
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    if (g) g[(size_t)L.copy_row[r] * ld + p] = w[(size_t)L.copy_var[r] * ld + p];
    if (jac) jac[(size_t)L.copy_nz[r] * ld + p] = 1.f;
  }
}

// ---- K3a: footprint points (core/geometry.py:78-83) ------------------------------------------------------
__device__ __forceinline__ void nlp_points_body(const NlpDev& L, const float* __restrict__ w, size_t P, size_t ld,
                                                float* __restrict__ px, float* __restrict__ py, const int k) {
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    const float x = w[(size_t)(k * L.nx + 0) * ld + p], y = w[(size_t)(k * L.nx + 1) * ld + p];
    if (L.shape == NLO_SHAPE_DOT) { px[(size_t)k * P + p] = x; py[(size_t)k * P + p] = y; continue; }
    float s, c; sincosf(w[(size_t)(k * L.nx + 2) * ld + p], &s, &c);
    for (int b = 0; b < L.nb; ++b) {
      px[(size_t)(k * L.nb + b) * P + p] = x + c * L.bx[b] - s * L.by[b];
      py[(size_t)(k * L.nb + b) * P + p] = y + s * L.bx[b] + c * L.by[b];
    }
  }
}

// ---- K3b: SDF constraint rows + chain rule to the pose (core/geometry.py:63-67, 107-117) --------------------
// Rows of knot k of problem p.  point(b, sv, gx, gy) delivers the SDF value / gradient of footprint point b (from the scratch of the
// learned SDF, or computed on the spot for analytic obstacles); sn, cs = sin / cos of the heading.  The footprint loop is unrolled to
// its bound of four with the soft-min accumulated on the fly: with run-time trip counts the per-point arrays lived in local memory
// (96 B of stack per thread; the benchmark_5 rows kernel took 276 us for 41 MB of traffic).
template <typename F>
__device__ __forceinline__ void nlp_rows_emit(const NlpDev& L, const float* __restrict__ w, size_t P, size_t ld, float* __restrict__ g,
                                              float* __restrict__ jac, const int k, const size_t p, const float sn, const float cs, F&& point) {
  const int* __restrict__ nz = L.nzmap + L.e_off_sdf + k * L.rows_per_knot * L.nnz_sdf_row;
  if (L.shape == NLO_SHAPE_DOT) {
    float sv, gx, gy;
    point(0, sv, gx, gy);
    if (g) g[(size_t)(L.g_off_sdf + k) * ld + p] = sv;
    if (jac) { jac[(size_t)nz[0] * ld + p] = gx; jac[(size_t)nz[1] * ld + p] = gy; }
    return;
  }
  if (L.use_slack) {
    // soft_min + slack (core/utils.py:28-31: un-stabilised)
    float sum = 0.f, rx = 0.f, ry = 0.f, rt = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      if (b < L.nb) {
        float sv, gx, gy;
        point(b, sv, gx, gy);
        const float dpx = -sn * L.bx[b] - cs * L.by[b], dpy = cs * L.bx[b] - sn * L.by[b];
        const float e = expf(-NLO_ALPHA * sv);
        sum += e; rx += e * gx; ry += e * gy; rt += e * (gx * dpx + gy * dpy);
      }
    }
    const float inv = 1.f / sum;
    if (g) g[(size_t)(L.g_off_sdf + k) * ld + p] = -logf(sum) / NLO_ALPHA + w[(size_t)(L.n_X + L.n_U + k) * ld + p];
    if (jac) {
      jac[(size_t)nz[0] * ld + p] = rx * inv; jac[(size_t)nz[1] * ld + p] = ry * inv; jac[(size_t)nz[2] * ld + p] = rt * inv;
      jac[(size_t)nz[3] * ld + p] = 1.f;
    }
  } else {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      if (b < L.nb) {
        float sv, gx, gy;
        point(b, sv, gx, gy);
        const float dpx = -sn * L.bx[b] - cs * L.by[b], dpy = cs * L.bx[b] - sn * L.by[b];
        if (g) g[(size_t)(L.g_off_sdf + k * L.nb + b) * ld + p] = sv;
        if (jac) {
          jac[(size_t)nz[3 * b + 0] * ld + p] = gx; jac[(size_t)nz[3 * b + 1] * ld + p] = gy;
          jac[(size_t)nz[3 * b + 2] * ld + p] = gx * dpx + gy * dpy;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) nlp_sdf_rows_kernel(NlpDev L, const float* __restrict__ w, size_t P, size_t ld,
                                                           const float* __restrict__ s, const float* __restrict__ jx,
                                                           const float* __restrict__ jy, float* __restrict__ g, float* __restrict__ jac) {
  const int k = blockIdx.y;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    float sn = 0.f, cs = 1.f;
    if (L.shape != NLO_SHAPE_DOT) sincosf(w[(size_t)(k * L.nx + 2) * ld + p], &sn, &cs);
    nlp_rows_emit(L, w, P, ld, g, jac, k, p, sn, cs, [&](int b, float& sv, float& gx, float& gy) {
      const size_t q = (size_t)(k * L.nb + b) * P + p;
      sv = s[q]; gx = jx[q]; gy = jy[q];
    });
  }
}

// ---- K3 + K5 in one pass for the analytic obstacles (solver.mode casadi): footprint points, soft-min union of circles / squares and
// the constraint rows of knot k, without any scratch between them (benchmarks 1, 2, 5) -----------------------------------------------
__device__ __forceinline__ void nlp_analytic_rows_body(const NlpDev& L, const float* __restrict__ w, size_t P, size_t ld,
                                                       float* __restrict__ g, float* __restrict__ jac, const int k) {
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    const float x = w[(size_t)(k * L.nx + 0) * ld + p], y = w[(size_t)(k * L.nx + 1) * ld + p];
    float sn = 0.f, cs = 1.f;
    if (L.shape != NLO_SHAPE_DOT) sincosf(w[(size_t)(k * L.nx + 2) * ld + p], &sn, &cs);
    nlp_rows_emit(L, w, P, ld, g, jac, k, p, sn, cs, [&](int b, float& sv, float& gx, float& gy) {
      const float px = L.shape == NLO_SHAPE_DOT ? x : x + cs * L.bx[b] - sn * L.by[b];
      const float py = L.shape == NLO_SHAPE_DOT ? y : y + sn * L.bx[b] + cs * L.by[b];
      const NloJet1 u = nlo_union_jet1(L.n_circles, L.okind, L.circles, px, py);
      sv = u.v; gx = u.dx; gy = u.dy;
    });
  }
}

// ---- K4: objective and gradient (core/runner.py:80-98) -----------------------------------------------------
// Objective: a block is 32 problems x 8 knot ranges (problem index fastest -> every load is one 128-byte line per warp);
// the 8 partial sums of a problem meet in shared memory.
// `sub` of `n_sub` block rows share the problems: with one row the role was 256 blocks walking 8 problem groups each - the tail of the launch.
__device__ __forceinline__ void nlp_obj_body(const NlpDev& L, const float* __restrict__ w, size_t P, size_t ld, float* __restrict__ f,
                                             const int sub, const int n_sub) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, c = threadIdx.x >> 5;
  for (size_t p0 = ((size_t)blockIdx.x + (size_t)sub * gridDim.x) * 32; p0 < P; p0 += (size_t)gridDim.x * n_sub * 32) {
    const size_t p = p0 + tx;
    float acc = 0.f;
    if (p < P) {
      const int per = (L.N + 7) / 8;
      const int k0 = c * per + 1, k1 = min(L.N, (c + 1) * per);
      if (k0 <= k1) {
        float x0 = w[(size_t)((k0 - 1) * L.nx) * ld + p], y0 = w[(size_t)((k0 - 1) * L.nx + 1) * ld + p];
#pragma unroll 4
        for (int k = k0; k <= k1; ++k) {
          const float x1 = w[(size_t)(k * L.nx) * ld + p], y1 = w[(size_t)(k * L.nx + 1) * ld + p];
          const float dx = x1 - x0, dy = y1 - y0;
          acc += sqrtf(dx * dx + dy * dy + NLO_EPS_PATH);
          x0 = x1; y0 = y1;
        }
      }
      if (L.use_slack) {
        const int pers = (L.N + 8) / 8;
        float ss = 0.f;
        for (int k = c * pers; k < min(L.N + 1, (c + 1) * pers); ++k) { const float v = w[(size_t)(L.n_X + L.n_U + k) * ld + p]; ss = fmaf(v, v, ss); }
        acc = fmaf(L.slack_penalty, ss, acc);
      }
      if (L.use_smooth) {
        const int nq = (L.N - 1) * L.nu, perq = (nq + 7) / 8;
        float uu = 0.f;
#pragma unroll 4
        for (int q = c * perq; q < min(nq, (c + 1) * perq); ++q) { const float v = w[(size_t)(L.n_X + q) * ld + p]; uu = fmaf(v, v, uu); }
        acc = fmaf(L.smooth_weight, uu, acc);
      }
    }
    part[c][tx] = acc;
    __syncthreads();
    if (c == 0 && p < P) {
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) t += part[q][tx];      // fixed order: deterministic
      f[p] = t;
    }
    __syncthreads();
  }
}

// Gradient: blockIdx.y = knot (uniform per block, no index division); a thread writes every gradient entry that belongs
// to its (problem, knot): the nx state entries, the nu controls of interval k and the knot's slack.
__device__ __forceinline__ void nlp_grad_body(const NlpDev& L, const float* __restrict__ w, size_t P, size_t ld, float* __restrict__ grad,
                                              const int k) {
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    const float xk = w[(size_t)(k * L.nx) * ld + p], yk = w[(size_t)(k * L.nx + 1) * ld + p];
    float gx = 0.f, gy = 0.f;
    if (k > 0) {
      const float dx = xk - w[(size_t)((k - 1) * L.nx) * ld + p], dy = yk - w[(size_t)((k - 1) * L.nx + 1) * ld + p];
      const float r = rsqrtf(dx * dx + dy * dy + NLO_EPS_PATH);
      gx += dx * r; gy += dy * r;
    }
    if (k < L.N) {
      const float dx = w[(size_t)((k + 1) * L.nx) * ld + p] - xk, dy = w[(size_t)((k + 1) * L.nx + 1) * ld + p] - yk;
      const float r = rsqrtf(dx * dx + dy * dy + NLO_EPS_PATH);
      gx -= dx * r; gy -= dy * r;
    }
    grad[(size_t)(k * L.nx) * ld + p] = gx;
    grad[(size_t)(k * L.nx + 1) * ld + p] = gy;
    for (int i = 2; i < L.nx; ++i) grad[(size_t)(k * L.nx + i) * ld + p] = 0.f;
    if (k < L.N) {
      for (int i = 0; i < L.nu; ++i) {
        const size_t v = (size_t)(L.n_X + k * L.nu + i);
        grad[v * ld + p] = (L.use_smooth && k < L.N - 1) ? 2.f * L.smooth_weight * w[v * ld + p] : 0.f;
      }
    }
    if (L.use_slack) {
      const size_t v = (size_t)(L.n_X + L.n_U + k);
      grad[v * ld + p] = 2.f * L.slack_penalty * w[v * ld + p];
    }
  }
}

// ---- everything that precedes the SDF evaluation, in ONE launch: blockIdx.y selects the role and its row -----------------
template <int DYN>
__global__ void __launch_bounds__(256, 8) nlp_phase0_kernel(NlpDev L, const float* __restrict__ w, size_t P, size_t ld, float* __restrict__ g,
                                                         float* __restrict__ jac, float* __restrict__ px, float* __restrict__ py,
                                                         float* __restrict__ f, float* __restrict__ grad, int n_dyn, int n_copy, int n_pts,
                                                         int n_grad, int n_obj) {
  int r = blockIdx.y;
  // the objective first (block rows are dispatched in order): its threads walk a tenth of a trajectory each, and as the last rows of the
  // grid they ran on after everything else had finished
  if (r < n_obj) { nlp_obj_body(L, w, P, ld, f, r, n_obj); return; }
  r -= n_obj;
  if (r < n_dyn) { nlp_dyn_body<DYN>(L, w, P, ld, g, jac, r); return; }
  r -= n_dyn;
  if (r < n_copy) { nlp_copy_rows_body(L, w, P, ld, g, jac, r); return; }
  r -= n_copy;
  if (r < n_pts) { nlp_points_body(L, w, P, ld, px, py, r); return; }
  r -= n_pts;
  if (r < n_grad) { nlp_grad_body(L, w, P, ld, grad, r); return; }
}
// analytic obstacles (solver.mode casadi): footprint points, union SDF and constraint rows of one knot per block row - its own launch,
// because the jets of the square obstacles need far more than the 32 registers the light roles of the phase-0 launch are held to
// (inside that launch they cost every role its occupancy: benchmark_1 x 65,536 went from 0.18 to 0.25 ms)
__global__ void __launch_bounds__(256) nlp_analytic_rows_kernel(NlpDev L, const float* __restrict__ w, size_t P, size_t ld,
                                                                float* __restrict__ g, float* __restrict__ jac) {
  nlp_analytic_rows_body(L, w, P, ld, g, jac, blockIdx.y);
}
__global__ void __launch_bounds__(256) nlp_points_kernel(NlpDev L, const float* __restrict__ w, size_t P, size_t ld,
                                                         float* __restrict__ px, float* __restrict__ py) {
  nlp_points_body(L, w, P, ld, px, py, blockIdx.y);
}

// ---- max bound violation per problem (best-of selection) ----------------------------------------------------
__global__ void __launch_bounds__(256) nlp_violation_kernel(int n_g, const float* __restrict__ g, const float* __restrict__ lb,
                                                            const float* __restrict__ ub, size_t P, size_t ld, float* __restrict__ viol) {
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    float m = 0.f;
    for (int r = 0; r < n_g; ++r) {
      const float v = g[(size_t)r * ld + p];
      m = fmaxf(m, fmaxf(lb[r] - v, v - ub[r]));
    }
    viol[p] = m;
  }
}

// ---- out = J^T y  (one thread per (problem, column); the column's non-zeros are contiguous in CCS order) ------------
__global__ void __launch_bounds__(256) nlp_jtv_kernel(int n_w, const int* __restrict__ colind, const int* __restrict__ row,
                                                      const float* __restrict__ jac, const float* __restrict__ y, size_t P, size_t ld,
                                                      const float* __restrict__ add, float* __restrict__ out) {
  const size_t total = (size_t)n_w * P;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx / P);
    const size_t p = idx - (size_t)c * P;
    float acc = add ? add[(size_t)c * ld + p] : 0.f;
    for (int z = colind[c]; z < colind[c + 1]; ++z) acc = fmaf(jac[(size_t)z * ld + p], y[(size_t)row[z] * ld + p], acc);
    out[(size_t)c * ld + p] = acc;
  }
}

// ---- [P][rows] <-> [rows][ld] ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n_in_rows,
                                                        size_t n_in_cols, size_t ld_in, size_t ld_out) {
  // in: [n_in_rows][ld_in] (n_in_cols valid) -> out: [n_in_cols][ld_out]
  __shared__ float tile[32][33];
  const size_t tiles_c = (n_in_cols + 31) / 32, tiles_r = (n_in_rows + 31) / 32;
  for (size_t tidx = blockIdx.x; tidx < tiles_c * tiles_r; tidx += gridDim.x) {
    const size_t tr = tidx / tiles_c, tc = tidx - tr * tiles_c;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
      const size_t r = tr * 32 + j, c = tc * 32 + tx;
      tile[j][tx] = (r < n_in_rows && c < n_in_cols) ? in[r * ld_in + c] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      const size_t c = tc * 32 + j, r = tr * 32 + tx;
      if (c < n_in_cols && r < n_in_rows) out[c * ld_out + r] = tile[tx][j];
    }
    __syncthreads();
  }
}

// ---- gather + transpose: out[p][c] = in[idx[c]][p]  (variable-major rows picked by an index list -> problem-major records) ----
__global__ void __launch_bounds__(256) pack_rows_kernel(const float* __restrict__ in, size_t ld_in, size_t P, const int* __restrict__ idx,
                                                        int n_out, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const size_t tiles_c = ((size_t)n_out + 31) / 32, tiles_p = (P + 31) / 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (size_t tidx = blockIdx.x; tidx < tiles_c * tiles_p; tidx += gridDim.x) {
    const size_t tp = tidx / tiles_c, tc = tidx - tp * tiles_c;
    for (int j = ty; j < 32; j += 8) {
      const size_t c = tc * 32 + j, p = tp * 32 + tx;
      tile[j][tx] = (c < (size_t)n_out && p < P) ? in[(size_t)idx[c] * ld_in + p] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      const size_t p = tp * 32 + j, c = tc * 32 + tx;
      if (p < P && c < (size_t)n_out) out[p * (size_t)n_out + c] = tile[tx][j];
    }
    __syncthreads();
  }
}

inline int grid_for(size_t total, int threads, int sm_count) {
  size_t want = (total + threads - 1) / threads;
  size_t cap = (size_t)sm_count * 8;
  return (int)std::max<size_t>(1, std::min(want, cap));
}

}  // namespace

int nlo_nlp_launch_assembly(nlo_nlp* p, const NlpScratch& sc, const float* w, size_t P, size_t ld, float* g, float* jac, float* f,
                            float* grad_f, cudaStream_t st, int phase, bool fused_rows) {
  const NlpDev& L = p->L;
  const int sm = p->sm_count;
  if (phase == 0) {
    const bool gj = g || jac;
    const bool analytic = L.sdf_mode == NLO_SDF_CIRCLES;
    const int n_dyn = gj ? L.N : 0, n_copy = gj ? L.n_copy : 0, n_pts = (gj && !analytic && !fused_rows) ? L.N + 1 : 0, n_grad = grad_f ? L.N + 1 : 0, n_obj = f ? 8 : 0;
    const unsigned rows = (unsigned)(n_dyn + n_copy + n_pts + n_grad + n_obj);
    if (rows) {
      const dim3 grid((unsigned)std::min<size_t>((P + 255) / 256, 4096), rows);
      switch (L.dyn) {
#define NLO_CASE(D) case D: nlp_phase0_kernel<D><<<grid, 256, 0, st>>>(L, w, P, ld, g, jac, sc.px, sc.py, f, grad_f, n_dyn, n_copy, n_pts, n_grad, n_obj); break;
        NLO_CASE(NLO_DYN_POINT_1ST) NLO_CASE(NLO_DYN_POINT_2ND) NLO_CASE(NLO_DYN_UNICYCLE)
        NLO_CASE(NLO_DYN_UNICYCLE_2ND) NLO_CASE(NLO_DYN_ACKERMANN) NLO_CASE(NLO_DYN_ACKERMANN_2ND)
#undef NLO_CASE
      }
      NLO_CHECK_LAUNCH();
    }
    if (gj && analytic) {                                    // footprint + union SDF + rows in one pass, no scratch
      nlp_analytic_rows_kernel<<<dim3((unsigned)std::min<size_t>((P + 255) / 256, 4096), (unsigned)(L.N + 1)), 256, 0, st>>>(L, w, P, ld, g, jac);
      NLO_CHECK_LAUNCH();
    }
  } else {
    if ((g || jac) && L.sdf_mode != NLO_SDF_CIRCLES) {       // (analytic obstacles: nlp_analytic_rows_kernel wrote the rows in phase 0)
      nlp_sdf_rows_kernel<<<dim3((unsigned)std::min<size_t>((P + 255) / 256, 4096), (unsigned)(L.N + 1)), 256, 0, st>>>(L, w, P, ld, sc.s, sc.jx, sc.jy, g, jac);
      NLO_CHECK_LAUNCH();
    }
  }
  return 0;
}

// K2 alone (Euler defects + their Jacobian values): the HBM-bound kernel the roofline report times in isolation
int nlo_nlp_launch_dynamics(const NlpDev& L, const float* w, size_t P, size_t ld, float* g, float* jac, cudaStream_t st) {
  const dim3 grid((unsigned)std::min<size_t>((P + 255) / 256, 4096), (unsigned)L.N);
  switch (L.dyn) {
#define NLO_CASE(D) case D: nlp_phase0_kernel<D><<<grid, 256, 0, st>>>(L, w, P, ld, g, jac, nullptr, nullptr, nullptr, nullptr, L.N, 0, 0, 0, 0); break;
    NLO_CASE(NLO_DYN_POINT_1ST) NLO_CASE(NLO_DYN_POINT_2ND) NLO_CASE(NLO_DYN_UNICYCLE)
    NLO_CASE(NLO_DYN_UNICYCLE_2ND) NLO_CASE(NLO_DYN_ACKERMANN) NLO_CASE(NLO_DYN_ACKERMANN_2ND)
#undef NLO_CASE
  }
  NLO_CHECK_LAUNCH();
  return 0;
}

int nlo_nlp_launch_points(const NlpDev& L, const float* w, size_t P, size_t ld, float* px, float* py, cudaStream_t st) {
  nlp_points_kernel<<<dim3((unsigned)std::min<size_t>((P + 255) / 256, 4096), (unsigned)(L.N + 1)), 256, 0, st>>>(L, w, P, ld, px, py);
  NLO_CHECK_LAUNCH();
  return 0;
}

int nlo_launch_violation(int n_g, const float* g, const float* lb, const float* ub, size_t P, size_t ld, float* viol, int sm, cudaStream_t st) {
  nlp_violation_kernel<<<grid_for(P, 128, sm), 128, 0, st>>>(n_g, g, lb, ub, P, ld, viol);
  NLO_CHECK_LAUNCH();
  return 0;
}

int nlo_launch_jtv(int n_w, const int* colind, const int* row, const float* jac, const float* y, size_t P, size_t ld, const float* add,
                   float* out, int sm, cudaStream_t st) {
  nlp_jtv_kernel<<<grid_for((size_t)n_w * P, 256, sm), 256, 0, st>>>(n_w, colind, row, jac, y, P, ld, add, out);
  NLO_CHECK_LAUNCH();
  return 0;
}

int nlo_launch_pack_rows(const float* in, size_t ld_in, size_t P, const int* idx, int n_out, float* out, int sm, cudaStream_t st) {
  if (P == 0 || n_out == 0) return 0;
  const size_t tiles = (((size_t)n_out + 31) / 32) * ((P + 31) / 32);
  pack_rows_kernel<<<(int)std::min<size_t>(tiles, (size_t)sm * 16), 256, 0, st>>>(in, ld_in, P, idx, n_out, out);
  NLO_CHECK_LAUNCH();
  return 0;
}

int nlo_launch_transpose(const float* in, float* out, size_t n_in_rows, size_t n_in_cols, size_t ld_in, size_t ld_out, int sm, cudaStream_t st) {
  if (n_in_rows == 0 || n_in_cols == 0) return 0;
  const size_t tiles = ((n_in_cols + 31) / 32) * ((n_in_rows + 31) / 32);
  const int grid = (int)std::min<size_t>(tiles, (size_t)sm * 16);
  transpose_kernel<<<grid, 256, 0, st>>>(in, out, n_in_rows, n_in_cols, ld_in, ld_out);
  NLO_CHECK_LAUNCH();
  return 0;
}
