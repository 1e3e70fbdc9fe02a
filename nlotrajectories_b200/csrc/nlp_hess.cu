// Hessian of the Lagrangian  sigma * hess f(w) + sum_r lam_r * hess g_r(w)  for P independent problems.
//
// IPOPT's default is the exact Hessian (core/runner.py:113-125 sets no hessian_approximation); CasADi assembles it from
// second derivatives of the same expressions this file restates:
//   core/dynamics.py:59-148       second derivatives of f(x,u) (unicycle / Ackermann families; point masses are linear)
//   core/geometry.py:78-117       footprint transform: d2 p / d theta2, chain rule of s(p(x, y, theta))
//   core/utils.py:28-31           soft_min: sum_b om_b H_b - alpha (sum_b om_b d_b d_b^T - m m^T)
//   _l4c_generated/nn_sdf.cpp:88-104  jac_adj1_nn_sdf = the 2x2 Hessian of the learned SDF (K1b, sdf_simt.cu)
//   core/sdf/casadi.py:33-41,385-386  analytic circles + soft-min union (mode: casadi)
//   core/runner.py:80-98          path length (couples neighbouring knots), slack and control penalties
// Output: values of the structural non-zeros of the UPPER triangle in compressed-column order
// [upstream-memory: CasADi declares nlp_hess_l as "triu:hess:gamma:x:x"], SoA [nnz_h][ld] like everything else.
//
// One thread per (problem, knot): it owns every Hessian entry whose row variable belongs to knot k - the within-knot
// block of (x_k, u_k), the slack diagonal and the coupling block (x,y)_k x (x,y)_{k+1} - so each entry is written exactly
// once (no atomics, no zero-fill).  A knot has 19 emission slots; hmap[k*19 + slot] is the CCS position or -1.
#include "nlo_common.cuh"
#include "nlp_internal.cuh"
#include <algorithm>
#include <map>
#include <vector>

namespace {

constexpr int HS = NLO_HESS_SLOTS;     // 0-5 pose/objective block (00,01,02,11,12,22); 6-11 extra dynamics pairs; 12-13 control
                                       // diagonal; 14 slack diagonal; 15-18 coupling (x_k,x_k+1) (x_k,y_k+1) (y_k,x_k+1) (y_k,y_k+1)

// extra (beyond the pose block) structurally non-zero pairs of d2 f / dz2, z = (x, u), in slot order 6, 7, ...
struct DynPairs { int n; int p[6][2]; };
const DynPairs kDynH[6] = {
    /* point_1st     */ {0, {{0, 0}}},
    /* point_2nd     */ {0, {{0, 0}}},
    /* unicycle      */ {1, {{2, 3}}},
    /* unicycle_2nd  */ {1, {{2, 3}}},
    /* ackermann     */ {3, {{2, 4}, {3, 3}, {3, 4}}},
    /* ackermann_2nd */ {6, {{2, 4}, {3, 3}, {3, 4}, {3, 6}, {3, 7}, {4, 6}}},
};
const bool kDynHasThetaTheta[6] = {false, false, true, true, true, true};

// sum_i lam_i d2 f_i / dz_a dz_b, scaled by sc, added into the slots
template <int DYN> struct DynH { __device__ static void add(const float*, const float*, float, const float*, float, float*) {} };
template <> struct DynH<NLO_DYN_UNICYCLE> {
  __device__ static void add(const float* x, const float* u, float, const float* l, float sc, float* h) {
    float s, c; sincosf(x[2], &s, &c); const float v = u[0];
    h[5] += sc * (-v * (l[0] * c + l[1] * s)); h[6] += sc * (-l[0] * s + l[1] * c); } };
template <> struct DynH<NLO_DYN_UNICYCLE_2ND> {
  __device__ static void add(const float* x, const float* u, float, const float* l, float sc, float* h) {
    float s, c; sincosf(x[2], &s, &c); const float v = x[3];
    h[5] += sc * (-v * (l[0] * c + l[1] * s)); h[6] += sc * (-l[0] * s + l[1] * c); } };
template <> struct DynH<NLO_DYN_ACKERMANN> {
  __device__ static void add(const float* x, const float* u, float L, const float* l, float sc, float* h) {
    float s, c; sincosf(x[2], &s, &c); const float v = u[0], t = tanf(x[3]), sec2 = 1.f + t * t, iL = 1.f / L;
    h[5] += sc * (-v * (l[0] * c + l[1] * s)); h[6] += sc * (-l[0] * s + l[1] * c);
    h[7] += sc * (l[2] * v * 2.f * t * sec2 * iL); h[8] += sc * (l[2] * sec2 * iL); } };
template <> struct DynH<NLO_DYN_ACKERMANN_2ND> {       // with the reference's slot quirk: v = x[4], psi_dot = x[6]
  __device__ static void add(const float* x, const float* u, float L, const float* l, float sc, float* h) {
    float s, c; sincosf(x[2], &s, &c);
    const float psi = x[3], v = x[4], pd = x[6], a = u[0], t = tanf(psi), sec2 = 1.f + t * t, iL = 1.f / L;
    const float q = 1.f / (1.f + psi * psi), dq = -2.f * psi * q * q, d2q = q * q * (8.f * psi * psi * q - 2.f);
    h[5] += sc * (-v * (l[0] * c + l[1] * s)); h[6] += sc * (-l[0] * s + l[1] * c);
    h[7] += sc * ((l[2] * v * 2.f * t * sec2 + l[4] * (pd * v * d2q + a * 2.f * t * sec2)) * iL);
    h[8] += sc * ((l[2] * sec2 + l[4] * pd * dq) * iL);
    h[9] += sc * (l[4] * v * dq * iL); h[10] += sc * (l[4] * sec2 * iL); h[11] += sc * (l[4] * q * iL); } };

template <int DYN, int NX, int NU>
__global__ void __launch_bounds__(256) nlp_hess_kernel(NlpDev L, const int* __restrict__ hmap, const float* __restrict__ w,
                                                       const float* __restrict__ sigma, const float* __restrict__ lam, size_t P, size_t ld,
                                                       const float* __restrict__ s, const float* __restrict__ jx, const float* __restrict__ jy,
                                                       const float* __restrict__ hxx, const float* __restrict__ hxy, const float* __restrict__ hyy,
                                                       float* __restrict__ hess) {
  const int k = blockIdx.y;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    float x[NX], u[NU], h[HS];
#pragma unroll
    for (int e = 0; e < HS; ++e) h[e] = 0.f;
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = w[(size_t)(k * NX + i) * ld + p];
#pragma unroll
    for (int i = 0; i < NU; ++i) u[i] = k < L.N ? w[(size_t)(L.n_X + k * NU + i) * ld + p] : 0.f;
    const float sig = sigma ? sigma[p] : 1.f;
    // ---- objective: path length, slack and control penalties (core/runner.py:80-98) ---------------------------------
    if (k > 0) {
      const float dx = x[0] - w[(size_t)((k - 1) * NX) * ld + p], dy = x[1] - w[(size_t)((k - 1) * NX + 1) * ld + p];
      const float r2 = dx * dx + dy * dy + NLO_EPS_PATH, ir3 = sig * rsqrtf(r2) / r2;
      h[0] += (r2 - dx * dx) * ir3; h[1] -= dx * dy * ir3; h[3] += (r2 - dy * dy) * ir3;
    }
    if (k < L.N) {
      const float dx = w[(size_t)((k + 1) * NX) * ld + p] - x[0], dy = w[(size_t)((k + 1) * NX + 1) * ld + p] - x[1];
      const float r2 = dx * dx + dy * dy + NLO_EPS_PATH, ir3 = sig * rsqrtf(r2) / r2;
      const float m00 = (r2 - dx * dx) * ir3, m01 = -dx * dy * ir3, m11 = (r2 - dy * dy) * ir3;
      h[0] += m00; h[1] += m01; h[3] += m11;
      h[15] = -m00; h[16] = -m01; h[17] = -m01; h[18] = -m11;
      if (L.use_smooth && k < L.N - 1) { h[12] = 2.f * L.smooth_weight * sig; h[13] = h[12]; }
      // ---- Euler defects: g = x_{k+1} - x_k - dt f(x_k, u_k) ----------------------------------------------------------
      float ld_[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) ld_[i] = lam[(size_t)(L.g_off_dyn + k * NX + i) * ld + p];
      DynH<DYN>::add(x, u, L.wheelbase, ld_, -L.dt, h);
    }
    if (L.use_slack) h[14] = 2.f * L.slack_penalty * sig;
    // ---- SDF rows (core/geometry.py:63-67, 107-117) ------------------------------------------------------------------
    if (L.shape == NLO_SHAPE_DOT) {
      const size_t q = (size_t)k * P + p;
      const float lm = lam[(size_t)(L.g_off_sdf + k) * ld + p];
      h[0] += lm * hxx[q]; h[1] += lm * hxy[q]; h[3] += lm * hyy[q];
    } else {
      float sn, cs; sincosf(x[2], &sn, &cs);
      float Hb[4][6], d3[4][3], sv[4];
      for (int b = 0; b < L.nb; ++b) {
        const size_t q = (size_t)(k * L.nb + b) * P + p;
        const float gx = jx[q], gy = jy[q], a = hxx[q], bxy = hxy[q], c = hyy[q];
        const float rx = cs * L.bx[b] - sn * L.by[b], ry = sn * L.bx[b] + cs * L.by[b];     // p - (x, y)
        const float tx = -ry, ty = rx;                                                     // dp / dtheta
        const float ax = a * tx + bxy * ty, ay = bxy * tx + c * ty;
        sv[b] = s[q];
        d3[b][0] = gx; d3[b][1] = gy; d3[b][2] = gx * tx + gy * ty;
        Hb[b][0] = a; Hb[b][1] = bxy; Hb[b][2] = ax; Hb[b][3] = c; Hb[b][4] = ay;
        Hb[b][5] = tx * ax + ty * ay - gx * rx - gy * ry;                                   // d2p/dtheta2 = -(p - (x, y))
      }
      if (L.use_slack) {
        const float lm = lam[(size_t)(L.g_off_sdf + k) * ld + p];
        float e[4], sum = 0.f;
        for (int b = 0; b < L.nb; ++b) { e[b] = expf(-NLO_ALPHA * sv[b]); sum += e[b]; }
        const float inv = 1.f / sum;
        float m[3] = {0.f, 0.f, 0.f}, acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int b = 0; b < L.nb; ++b) {
          const float om = e[b] * inv;
          m[0] += om * d3[b][0]; m[1] += om * d3[b][1]; m[2] += om * d3[b][2];
          acc[0] += om * (Hb[b][0] - NLO_ALPHA * d3[b][0] * d3[b][0]); acc[1] += om * (Hb[b][1] - NLO_ALPHA * d3[b][0] * d3[b][1]);
          acc[2] += om * (Hb[b][2] - NLO_ALPHA * d3[b][0] * d3[b][2]); acc[3] += om * (Hb[b][3] - NLO_ALPHA * d3[b][1] * d3[b][1]);
          acc[4] += om * (Hb[b][4] - NLO_ALPHA * d3[b][1] * d3[b][2]); acc[5] += om * (Hb[b][5] - NLO_ALPHA * d3[b][2] * d3[b][2]);
        }
        h[0] += lm * (acc[0] + NLO_ALPHA * m[0] * m[0]); h[1] += lm * (acc[1] + NLO_ALPHA * m[0] * m[1]);
        h[2] += lm * (acc[2] + NLO_ALPHA * m[0] * m[2]); h[3] += lm * (acc[3] + NLO_ALPHA * m[1] * m[1]);
        h[4] += lm * (acc[4] + NLO_ALPHA * m[1] * m[2]); h[5] += lm * (acc[5] + NLO_ALPHA * m[2] * m[2]);
      } else {
        for (int b = 0; b < L.nb; ++b) {
          const float lm = lam[(size_t)(L.g_off_sdf + k * L.nb + b) * ld + p];
#pragma unroll
          for (int e = 0; e < 6; ++e) h[e] += lm * Hb[b][e];
        }
      }
    }
    const int* __restrict__ hm = hmap + k * HS;
#pragma unroll
    for (int e = 0; e < HS; ++e) { const int pos = hm[e]; if (pos >= 0) hess[(size_t)pos * ld + p] = h[e]; }
  }
}

// analytic circles / squares + soft-min union: value, gradient and Hessian (core/sdf/casadi.py:33-41, 69-115, 385-386)
__global__ void __launch_bounds__(256) nlp_circles_hess_kernel(NlpDev L, const float* __restrict__ px, const float* __restrict__ py, size_t n,
                                                               float* __restrict__ s, float* __restrict__ jx, float* __restrict__ jy,
                                                               float* __restrict__ hxx, float* __restrict__ hxy, float* __restrict__ hyy) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const NloJet u = nlo_union_jet(L.n_circles, L.okind, L.circles, px[i], py[i]);
    s[i] = u.v; jx[i] = u.dx; jy[i] = u.dy; hxx[i] = u.dxx; hxy[i] = u.dxy; hyy[i] = u.dyy;
  }
}

}  // namespace

// structural pattern (upper triangle, CCS) and the per-knot slot table; same construction as oracle/nlp_oracle.py::hess_pattern
int nlo_nlp_build_hess_layout(const NlpDev& l, std::vector<int>* rows, std::vector<int>* cols, std::vector<int>* hmap) {
  const int nx = l.nx, nu = l.nu, N = l.N;
  auto iX = [&](int i, int k) { return k * nx + i; };
  auto iU = [&](int i, int k) { return l.n_X + k * nu + i; };
  auto iS = [&](int k) { return l.n_X + l.n_U + k; };
  auto zi = [&](int a, int k) { return a < nx ? iX(a, k) : iU(a - nx, k); };
  const DynPairs& D = kDynH[l.dyn];
  const int pose[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
  // slot -> (row, col) per knot, or (-1, -1)
  std::vector<std::pair<int, int>> slot((size_t)(N + 1) * HS, {-1, -1});
  for (int k = 0; k <= N; ++k) {
    auto* sl = &slot[(size_t)k * HS];
    for (int e = 0; e < 6; ++e) {
      const int a = pose[e][0], b = pose[e][1];
      const bool in_obj = (a < 2 && b < 2);
      const bool in_pose = (l.shape != NLO_SHAPE_DOT);
      const bool in_dyn = (e == 5 && k < N && kDynHasThetaTheta[l.dyn]);
      if (in_obj || in_pose || in_dyn) sl[e] = {iX(a, k), iX(b, k)};
    }
    if (k < N) {
      for (int e = 0; e < D.n; ++e) sl[6 + e] = {zi(D.p[e][0], k), zi(D.p[e][1], k)};
      if (l.use_smooth && k < N - 1) for (int j = 0; j < nu && j < 2; ++j) sl[12 + j] = {iU(j, k), iU(j, k)};
      sl[15] = {iX(0, k), iX(0, k + 1)}; sl[16] = {iX(0, k), iX(1, k + 1)};
      sl[17] = {iX(1, k), iX(0, k + 1)}; sl[18] = {iX(1, k), iX(1, k + 1)};
    }
    if (l.use_slack) sl[14] = {iS(k), iS(k)};
  }
  std::vector<std::pair<int, int>> ent;     // (col, row) for CCS sorting
  for (auto& rc : slot) if (rc.first >= 0) { if (rc.first > rc.second) return nlo_fail("internal: lower-triangle Hessian slot"); ent.push_back({rc.second, rc.first}); }
  std::sort(ent.begin(), ent.end());
  if (std::adjacent_find(ent.begin(), ent.end()) != ent.end()) return nlo_fail("internal: duplicate Hessian slot");
  std::map<std::pair<int, int>, int> pos;
  rows->clear(); cols->clear();
  for (size_t i = 0; i < ent.size(); ++i) { pos[ent[i]] = (int)i; rows->push_back(ent[i].second); cols->push_back(ent[i].first); }
  hmap->assign(slot.size(), -1);
  for (size_t i = 0; i < slot.size(); ++i) if (slot[i].first >= 0) (*hmap)[i] = pos[{slot[i].second, slot[i].first}];
  return 0;
}

int nlo_nlp_launch_circles_hess(const NlpDev& L, const float* px, const float* py, size_t n, float* s, float* jx, float* jy,
                                float* hxx, float* hxy, float* hyy, int sm, cudaStream_t st) {
  const size_t want = (n + 255) / 256, cap = (size_t)sm * 8;
  nlp_circles_hess_kernel<<<(unsigned)std::max<size_t>(1, std::min(want, cap)), 256, 0, st>>>(L, px, py, n, s, jx, jy, hxx, hxy, hyy);
  NLO_CHECK_LAUNCH();
  return 0;
}

int nlo_nlp_launch_hess(const NlpDev& L, const int* hmap, const float* w, const float* sigma, const float* lam, size_t P, size_t ld,
                        const float* s, const float* jx, const float* jy, const float* hxx, const float* hxy, const float* hyy,
                        float* hess, cudaStream_t st) {
  const dim3 grid((unsigned)std::min<size_t>((P + 255) / 256, 4096), (unsigned)(L.N + 1));
  switch (L.dyn) {
#define NLO_CASE(D, NX, NU) case D: nlp_hess_kernel<D, NX, NU><<<grid, 256, 0, st>>>(L, hmap, w, sigma, lam, P, ld, s, jx, jy, hxx, hxy, hyy, hess); break;
    NLO_CASE(NLO_DYN_POINT_1ST, 4, 2) NLO_CASE(NLO_DYN_POINT_2ND, 4, 2) NLO_CASE(NLO_DYN_UNICYCLE, 3, 2)
    NLO_CASE(NLO_DYN_UNICYCLE_2ND, 5, 2) NLO_CASE(NLO_DYN_ACKERMANN, 4, 2) NLO_CASE(NLO_DYN_ACKERMANN_2ND, 7, 2)
#undef NLO_CASE
    default: return nlo_fail("unknown dynamics id %d", L.dyn);
  }
  NLO_CHECK_LAUNCH();
  return 0;
}
