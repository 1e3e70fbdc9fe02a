// Internal layout descriptor of one NLP family (see nlp_kernels.cu / capi.cu).
#pragma once
#include "nlo_common.cuh"
#include <vector>

#define NLO_HOST_LANES 4
#define NLO_HESS_SLOTS 19        // Hessian emission slots per knot (nlp_hess.cu)

struct NlpDev {                 // passed by value to kernels
  int dyn, shape, N, nx, nu, nb;
  int use_slack, use_smooth, enforce_heading, sdf_mode, n_circles;
  int rows_per_knot, nnz_sdf_row, n_term, nA_off, nnz_dyn;
  int n_X, n_U, n_w, n_g, nnz;
  int g_off_term, g_off_dyn, g_off_slack, g_off_sdf, g_off_ctrl;
  int e_off_dyn, e_off_sdf;     // emission offsets of the dynamics / SDF blocks in nzmap
  int n_copy;
  float dt, slack_penalty, smooth_weight, wheelbase;
  float bx[4], by[4];
  float circles[NLO_MAX_CIRCLES][4];
  int okind[NLO_MAX_CIRCLES];
  const int* nzmap;             // device: emission index -> compressed-column position
  const int* copy_row;          // device: rows of g that are copies of a variable
  const int* copy_var;
  const int* copy_nz;
};

struct NlpScratch {             // footprint points and SDF outputs for cap_P problems: (N+1)*nb*cap_P floats each
  size_t cap_P;
  float *px, *py, *s, *jx, *jy;
};
struct NlpLane {                // one pipeline lane of nlo_nlp_eval_host
  cudaStream_t stream;
  size_t cap_P;                 // problems per chunk the buffers below hold
  float *d_in, *d_w;            // problem-major staging of w, and its SoA transpose
  float *d_g, *d_jac, *d_f, *d_grad;          // SoA outputs
  float *d_og, *d_ojac, *d_ograd;             // problem-major staging of the outputs (full or compact records)
  NlpScratch scratch;
};

struct nlo_nlp {
  nlo_nlp_desc desc;
  NlpDev L;
  nlo_sdf_model* model;
  int device, sm_count;
  std::vector<int> rows_ccs, cols_ccs;
  int* d_tables;                // nzmap | copy_row | copy_var | copy_nz | colind (n_w+1) | row (nnz)
  const int* d_colind; const int* d_row;
  NlpScratch scratch;           // SDF point / value scratch of the device entry point
  // Hessian of the Lagrangian (nlp_hess.cu)
  std::vector<int> hrows_ccs, hcols_ccs;   // structural pattern, upper triangle, compressed-column order
  int* d_hmap;                  // device: [N+1][NLO_HESS_SLOTS] emission slot -> CCS position (or -1)
  float* d_hs[3]; size_t hs_cap_P;          // SDF Hessian scratch (hxx, hxy, hyy) for hs_cap_P problems
  float* zc; size_t zc_cap;                 // pinned, device-mapped staging of the small-batch host entry point (floats)
  // compact host form (nlo_nlp_eval_host_compact): only what varies with w travels back
  std::vector<int> cg_rows, cj_nz, cgr_idx;           // varying rows of g / non-zeros of dg/dw / entries of grad f
  std::vector<int> cg_copy_row, cg_copy_var;          // g[row] = w[var]
  std::vector<int> cj_const_nz; std::vector<float> cj_const_val;     // dg/dw[nz] = value for every w
  std::vector<int> cgr_lin_idx; std::vector<float> cgr_lin_coef;     // grad f[idx] = coef * w[idx]; every other non-varying entry is 0
  int* d_compact;                                     // device: cg_rows | cj_nz | cgr_idx
  NlpLane lane[NLO_HOST_LANES];  // host entry point: lanes (streams) process alternating chunks of problems
};

// ---- analytic obstacles (solver.mode casadi): value, gradient and Hessian w.r.t. (x, y) -----------------------------------
// A small second-order jet pushes the reference's smooth square SDF (core/sdf/casadi.py:69-115: soft |.|, soft max / min with
// eps = 1e-6) through the chain rule; circles (:33-41) are closed form.
struct NloJet {
  float v, dx, dy, dxx, dxy, dyy;
  __device__ __forceinline__ static NloJet cst(float c) { return {c, 0.f, 0.f, 0.f, 0.f, 0.f}; }
};
__device__ __forceinline__ NloJet operator+(const NloJet& a, const NloJet& b) { return {a.v + b.v, a.dx + b.dx, a.dy + b.dy, a.dxx + b.dxx, a.dxy + b.dxy, a.dyy + b.dyy}; }
__device__ __forceinline__ NloJet operator-(const NloJet& a, const NloJet& b) { return {a.v - b.v, a.dx - b.dx, a.dy - b.dy, a.dxx - b.dxx, a.dxy - b.dxy, a.dyy - b.dyy}; }
__device__ __forceinline__ NloJet operator*(const NloJet& a, float c) { return {a.v * c, a.dx * c, a.dy * c, a.dxx * c, a.dxy * c, a.dyy * c}; }
__device__ __forceinline__ NloJet operator*(const NloJet& a, const NloJet& b) {
  return {a.v * b.v, a.dx * b.v + a.v * b.dx, a.dy * b.v + a.v * b.dy, a.dxx * b.v + 2.f * a.dx * b.dx + a.v * b.dxx,
          a.dxy * b.v + a.dx * b.dy + a.dy * b.dx + a.v * b.dxy, a.dyy * b.v + 2.f * a.dy * b.dy + a.v * b.dyy};
}
__device__ __forceinline__ NloJet nlo_jsqrt(const NloJet& a) {
  const float r = sqrtf(a.v), f1 = 0.5f / r, f2 = -0.25f / (r * a.v);
  return {r, f1 * a.dx, f1 * a.dy, f2 * a.dx * a.dx + f1 * a.dxx, f2 * a.dx * a.dy + f1 * a.dxy, f2 * a.dy * a.dy + f1 * a.dyy};
}
__device__ __forceinline__ NloJet nlo_obstacle_jet(int kind, const float* c, float x, float y) {
  if (kind == NLO_OBST_CIRCLE) {
    const float dx = x - c[0], dy = y - c[1], d = sqrtf(dx * dx + dy * dy), id = 1.f / d, nx = dx * id, ny = dy * id;
    return {d - (c[2] + c[3]), nx, ny, (1.f - nx * nx) * id, -nx * ny * id, (1.f - ny * ny) * id};
  }
  const NloJet X{x - c[0], 1.f, 0.f, 0.f, 0.f, 0.f}, Y{y - c[1], 0.f, 1.f, 0.f, 0.f, 0.f};
  const NloJet eps = NloJet::cst(1e-6f), half = NloJet::cst(0.5f * c[2] + c[3]), zero = NloJet::cst(0.f);
  const NloJet d_x = nlo_jsqrt(X * X + eps) - half, d_y = nlo_jsqrt(Y * Y + eps) - half;
  auto smax = [&](const NloJet& a, const NloJet& b) { return (a + b + nlo_jsqrt((a - b) * (a - b) + eps)) * 0.5f; };
  auto smin = [&](const NloJet& a, const NloJet& b) { return (a + b - nlo_jsqrt((a - b) * (a - b) + eps)) * 0.5f; };
  const NloJet xo = smax(d_x, zero), yo = smax(d_y, zero);
  return nlo_jsqrt(xo * xo + yo * yo) + smin(smax(d_x, d_y), zero);
}
// First-order twin of the jet above (value and gradient only) for the constraint rows: the compiler does not drop the second-order
// fields of NloJet from the smooth square's chain of products and square roots, and they are three quarters of its arithmetic
// (the benchmark_5 rows kernel - 4 footprint points x 3 obstacles per thread - was bound by exactly that arithmetic).
struct NloJet1 {
  float v, dx, dy;
  __device__ __forceinline__ static NloJet1 cst(float c) { return {c, 0.f, 0.f}; }
};
__device__ __forceinline__ NloJet1 operator+(const NloJet1& a, const NloJet1& b) { return {a.v + b.v, a.dx + b.dx, a.dy + b.dy}; }
__device__ __forceinline__ NloJet1 operator-(const NloJet1& a, const NloJet1& b) { return {a.v - b.v, a.dx - b.dx, a.dy - b.dy}; }
__device__ __forceinline__ NloJet1 operator*(const NloJet1& a, float c) { return {a.v * c, a.dx * c, a.dy * c}; }
__device__ __forceinline__ NloJet1 operator*(const NloJet1& a, const NloJet1& b) { return {a.v * b.v, a.dx * b.v + a.v * b.dx, a.dy * b.v + a.v * b.dy}; }
__device__ __forceinline__ NloJet1 nlo_jsqrt(const NloJet1& a) {
  const float ir = rsqrtf(a.v), f1 = 0.5f * ir;       // one MUFU instead of an IEEE square root and an IEEE division (2 ulp; a.v >= eps > 0)
  return {a.v * ir, f1 * a.dx, f1 * a.dy};
}
__device__ __forceinline__ NloJet1 nlo_obstacle_jet1(int kind, const float* c, float x, float y) {
  if (kind == NLO_OBST_CIRCLE) {
    const float dx = x - c[0], dy = y - c[1], q = dx * dx + dy * dy, id = rsqrtf(q);
    return {q * id - (c[2] + c[3]), dx * id, dy * id};
  }
  const NloJet1 X{x - c[0], 1.f, 0.f}, Y{y - c[1], 0.f, 1.f};
  const NloJet1 eps = NloJet1::cst(1e-6f), half = NloJet1::cst(0.5f * c[2] + c[3]), zero = NloJet1::cst(0.f);
  const NloJet1 d_x = nlo_jsqrt(X * X + eps) - half, d_y = nlo_jsqrt(Y * Y + eps) - half;
  auto smax = [&](const NloJet1& a, const NloJet1& b) { return (a + b + nlo_jsqrt((a - b) * (a - b) + eps)) * 0.5f; };
  auto smin = [&](const NloJet1& a, const NloJet1& b) { return (a + b - nlo_jsqrt((a - b) * (a - b) + eps)) * 0.5f; };
  const NloJet1 xo = smax(d_x, zero), yo = smax(d_y, zero);
  return nlo_jsqrt(xo * xo + yo * yo) + smin(smax(d_x, d_y), zero);
}
__device__ __forceinline__ NloJet1 nlo_union_jet1(int n, const int* kinds, const float (*circles)[4], float x, float y) {
  float sum = 0.f, gx = 0.f, gy = 0.f;
  for (int q = 0; q < n; ++q) {
    const NloJet1 j = nlo_obstacle_jet1(kinds[q], circles[q], x, y);
    const float e = expf(-NLO_ALPHA * j.v);
    sum += e; gx += e * j.dx; gy += e * j.dy;
  }
  const float inv = 1.f / sum;
  return {-logf(sum) / NLO_ALPHA, gx * inv, gy * inv};
}

// soft-min union (core/sdf/casadi.py:385-386, core/utils.py:28-31) of the obstacles' jets
__device__ __forceinline__ NloJet nlo_union_jet(int n, const int* kinds, const float (*circles)[4], float x, float y) {
  float sum = 0.f, gx = 0.f, gy = 0.f, a = 0.f, b = 0.f, c = 0.f;
  for (int q = 0; q < n; ++q) {
    const NloJet j = nlo_obstacle_jet(kinds[q], circles[q], x, y);
    const float e = expf(-NLO_ALPHA * j.v);
    sum += e; gx += e * j.dx; gy += e * j.dy;
    a += e * (j.dxx - NLO_ALPHA * j.dx * j.dx); b += e * (j.dxy - NLO_ALPHA * j.dx * j.dy); c += e * (j.dyy - NLO_ALPHA * j.dy * j.dy);
  }
  const float inv = 1.f / sum, mx = gx * inv, my = gy * inv;
  return {-logf(sum) / NLO_ALPHA, mx, my, a * inv + NLO_ALPHA * mx * mx, b * inv + NLO_ALPHA * mx * my, c * inv + NLO_ALPHA * my * my};
}

// structural A pairs per model, usable in device code after unrolling
template <int DYN>
__host__ __device__ constexpr int kDynA(int a, int q) {
  constexpr int p1[2][2] = {{0, 2}, {1, 3}};
  constexpr int un[2][2] = {{0, 2}, {1, 2}};
  constexpr int u2[5][2] = {{0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 4}};
  constexpr int ak[3][2] = {{0, 2}, {1, 2}, {2, 3}};
  constexpr int a2[10][2] = {{0, 2}, {0, 4}, {1, 2}, {1, 4}, {2, 3}, {2, 4}, {3, 6}, {4, 3}, {4, 4}, {4, 6}};
  return DYN == NLO_DYN_POINT_2ND ? p1[a < 2 ? a : 0][q]
       : DYN == NLO_DYN_UNICYCLE ? un[a < 2 ? a : 0][q]
       : DYN == NLO_DYN_UNICYCLE_2ND ? u2[a < 5 ? a : 0][q]
       : DYN == NLO_DYN_ACKERMANN ? ak[a < 3 ? a : 0][q]
       : DYN == NLO_DYN_ACKERMANN_2ND ? a2[a < 10 ? a : 0][q]
       : (q == 0 ? 0 : 1);
}

int nlo_nlp_build_layout(const nlo_nlp_desc* d, NlpDev* L, std::vector<int>* rows_ccs, std::vector<int>* cols_ccs,
                         std::vector<int>* nzmap, std::vector<int>* copy_row, std::vector<int>* copy_var,
                         std::vector<int>* copy_emit, std::vector<float>* const_ccs = nullptr);
int nlo_launch_pack_rows(const float* in, size_t ld_in, size_t P, const int* idx, int n_out, float* out, int sm, cudaStream_t st);
// phase 0: everything before the SDF evaluation (defects, copy rows, footprint points / circles, f, grad f)
// phase 1: SDF rows (needs p->d_s/d_jx/d_jy)
int nlo_nlp_launch_assembly(nlo_nlp* p, const NlpScratch& sc, const float* w, size_t P, size_t ld, float* g, float* jac, float* f,
                            float* grad_f, cudaStream_t st, int phase, bool fused_rows = false);
int nlo_nlp_build_hess_layout(const NlpDev& L, std::vector<int>* rows, std::vector<int>* cols, std::vector<int>* hmap);
int nlo_nlp_launch_dynamics(const NlpDev& L, const float* w, size_t P, size_t ld, float* g, float* jac, cudaStream_t st);
int nlo_nlp_launch_points(const NlpDev& L, const float* w, size_t P, size_t ld, float* px, float* py, cudaStream_t st);
int nlo_nlp_launch_circles_hess(const NlpDev& L, const float* px, const float* py, size_t n, float* s, float* jx, float* jy,
                                float* hxx, float* hxy, float* hyy, int sm, cudaStream_t st);
int nlo_nlp_launch_hess(const NlpDev& L, const int* hmap, const float* w, const float* sigma, const float* lam, size_t P, size_t ld,
                        const float* s, const float* jx, const float* jy, const float* hxx, const float* hxy, const float* hyy,
                        float* hess, cudaStream_t st);
int nlo_launch_violation(int n_g, const float* g, const float* lb, const float* ub, size_t P, size_t ld, float* viol, int sm, cudaStream_t st);
int nlo_launch_jtv(int n_w, const int* colind, const int* row, const float* jac, const float* y, size_t P, size_t ld, const float* add,
                   float* out, int sm, cudaStream_t st);
int nlo_launch_transpose(const float* in, float* out, size_t n_in_rows, size_t n_in_cols, size_t ld_in, size_t ld_out, int sm, cudaStream_t st);
