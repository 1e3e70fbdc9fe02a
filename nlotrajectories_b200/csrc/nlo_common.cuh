// Shared declarations for libnlo_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <math.h>
#include "../../include/nlo_b200.h"

#define NLO_LEAKY_SLOPE 0.01f   // torch.nn.functional.leaky_relu default (core/nn_architectures.py:51)
#define NLO_ALPHA 10.0f         // soft_min sharpness (core/utils.py:18)
#define NLO_EPS_PATH 1e-8f      // core/runner.py:82

// ---- error plumbing (capi.cu) -----------------------------------------------------------------
int nlo_fail(const char* fmt, ...);
void nlo_count_launch(unsigned n = 1);
#define NLO_CUDA(call)                                                                            \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess)                                                                       \
      return nlo_fail("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
  } while (0)
#define NLO_CHECK_LAUNCH()                                                                        \
  do {                                                                                            \
    cudaError_t e__ = cudaGetLastError();                                                         \
    if (e__ != cudaSuccess)                                                                       \
      return nlo_fail("%s:%d: kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
    nlo_count_launch();                                                                           \
  } while (0)

// ---- model ------------------------------------------------------------------------------------
#define NLO_STREAM_SLOTS 8
struct SdfNetDev {           // passed by value to kernels
  const float* w;            // flat blob on device (layout in nlo_b200.h)
  int H, M;
  int act0, act;
  float p0, p;
  __host__ __device__ int off_W0() const { return 0; }
  __host__ __device__ int off_b0() const { return 2 * H; }
  __host__ __device__ int off_W(int l) const { return 3 * H + (l - 1) * (H * H + H); }   // l = 1..M
  __host__ __device__ int off_b(int l) const { return off_W(l) + H * H; }
  __host__ __device__ int off_wout() const { return 3 * H + M * (H * H + H); }
  __host__ __device__ int off_bout() const { return off_wout() + H; }
  __host__ __device__ int count() const { return off_bout() + 1; }
};

struct nlo_sdf_model {
  nlo_sdf_desc desc;
  int device;
  int sm_count;
  int prec;                  // resolved NLO_PREC_*
  unsigned long long uid;    // process-unique id (owner tag of per-device constant memory)
  float* d_w;                // fp32 blob
  float* d_wt;               // transposed copies of the hidden matrices: wt[l][k][j] = W_{l+1}[j][k]
  size_t n_w;
  // tensor-path operand images (built on demand by sdf_tc.cu)
  void* d_tc;                // W1 split into fp16 hi | lo images in UMMA core-matrix order
  size_t tc_bytes;           // bytes of the images; NLO_STREAM_SLOTS x {tile counter, finished-CTA counter} follow them (zeroed once;
                             // the last CTA of a launch resets its pair, so no memset is needed per launch)
  float tc_params[12];       // TcParams of sdf_tc.cu (scales and bounds)
  float tc_const[776];       // TcConst of sdf_tc.cu (small vectors handed to the kernel as a __grid_constant__ parameter)
  float* h_deep;             // deep tensor path (sdf_tc_deep.cu): host blob, TcDeepParams (32 floats) | TcDeepConst
  // scratch for the host-buffer entry points
  float* d_io; size_t io_cap;        // device staging
  float* h_io; size_t h_cap;         // pinned staging
  // Per-stream slots: one model is evaluated from several streams at once (the two lanes of nlo_nlp_eval_host, the model's own
  // stream, caller streams), so everything a launch mutates on the device - the activation workspace of the FP32 paths and the
  // dynamic tile counter of the tensor path - is owned by the stream that launches it (slot = nlo_model_stream_slot()).
  cudaStream_t slot_stream[NLO_STREAM_SLOTS]; int n_slots;
  float* d_ws[NLO_STREAM_SLOTS]; size_t ws_cap[NLO_STREAM_SLOTS];   // global-memory activation scratch (large nets / hessian)
  cudaStream_t stream;               // private stream for host-buffer entry points
  SdfNetDev net() const {
    SdfNetDev n; n.w = d_w; n.H = (int)desc.hidden; n.M = (int)desc.n_hidden_mats;
    n.act0 = (int)desc.act0; n.act = (int)desc.act; n.p0 = desc.p0; n.p = desc.p; return n;
  }
};

// ---- activations ------------------------------------------------------------------------------
__device__ __forceinline__ float nlo_phi(float a, int act, float prm) {
  switch (act) {
    case NLO_ACT_RELU: return fmaxf(a, 0.f);
    case NLO_ACT_TANH: return tanhf(a);
    case NLO_ACT_SIGMOID: return 1.f / (1.f + expf(-a));
    case NLO_ACT_LEAKY_RELU: return a > 0.f ? a : a * NLO_LEAKY_SLOPE;
    case NLO_ACT_SIN: return sinf(prm * a);
    case NLO_ACT_COS_SCALE: return cosf(a) * prm;
    default: return a;
  }
}
// value and first derivative
__device__ __forceinline__ void nlo_phi_d(float a, int act, float prm, float& v, float& d) {
  switch (act) {
    case NLO_ACT_RELU: v = fmaxf(a, 0.f); d = a > 0.f ? 1.f : 0.f; break;
    case NLO_ACT_TANH: { float t = tanhf(a); v = t; d = 1.f - t * t; } break;
    case NLO_ACT_SIGMOID: { float s = 1.f / (1.f + expf(-a)); v = s; d = s * (1.f - s); } break;
    case NLO_ACT_LEAKY_RELU: v = a > 0.f ? a : a * NLO_LEAKY_SLOPE; d = a > 0.f ? 1.f : NLO_LEAKY_SLOPE; break;
    case NLO_ACT_SIN: { float s, c; sincosf(prm * a, &s, &c); v = s; d = prm * c; } break;
    case NLO_ACT_COS_SCALE: { float s, c; sincosf(a, &s, &c); v = prm * c; d = -prm * s; } break;
    default: v = a; d = 1.f; break;
  }
}
// first and second derivative
__device__ __forceinline__ void nlo_phi_d2(float a, int act, float prm, float& d, float& d2) {
  switch (act) {
    case NLO_ACT_RELU: d = a > 0.f ? 1.f : 0.f; d2 = 0.f; break;
    case NLO_ACT_TANH: { float t = tanhf(a); d = 1.f - t * t; d2 = -2.f * t * d; } break;
    case NLO_ACT_SIGMOID: { float s = 1.f / (1.f + expf(-a)); d = s * (1.f - s); d2 = d * (1.f - 2.f * s); } break;
    case NLO_ACT_LEAKY_RELU: d = a > 0.f ? 1.f : NLO_LEAKY_SLOPE; d2 = 0.f; break;
    case NLO_ACT_SIN: { float s, c; sincosf(prm * a, &s, &c); d = prm * c; d2 = -prm * prm * s; } break;
    case NLO_ACT_COS_SCALE: { float s, c; sincosf(a, &s, &c); d = -prm * s; d2 = -prm * c; } break;
    default: d = 1.f; d2 = 0.f; break;
  }
}

// ---- SFU sine / cosine behind an exact-enough range reduction (tensor path, sdf_tc.cu) -----------------------
// The Fourier-feature and SIREN layers (core/nn_architectures.py:8-39) evaluate cos / sin at |a| up to ~10^2
// (shipped model: |W| up to 24).  sinf / cosf cost ~30-50 instructions each; here the argument is reduced to
// [-pi, pi] with two FMAs against 2*pi = hi + lo (absolute error ~2e-7: one rounding of a result of magnitude <= pi)
// and handed to MUFU.SIN / MUFU.COS (absolute error 2^-21.4 on that interval).  Both are below the error the
// reference's own fp32 evaluation of `a` carries at that magnitude (ulp(50) = 3.8e-6).
__device__ __forceinline__ float nlo_reduce_2pi(float a) {
  const float kk = fmaf(a, 0.15915494309189535f, 12582912.f);     // round-to-nearest integer via the 1.5*2^23 trick
  const float k = kk - 12582912.f;
  const float r = fmaf(k, -6.2831854820251465f, a);               // fp32(2*pi)
  return fmaf(k, 1.7484555e-07f, r);                              // fp32(2*pi) - 2*pi
}
// No branch to a library slow path: the two-FMA reduction is exact enough while the rounded quotient fits the 1.5 * 2^23 trick,
// i.e. up to |a| ~ 2^22 * 2 pi = 2.6e7 - far beyond where an fp32 argument still carries a phase (ulp(2.6e7) = 2 rad), so any fp32
// evaluation, the reference's included, is noise there.  (A conditional call inside the fully unrolled tile bodies cost every
// Fourier / SIREN kernel a 300-500 byte stack frame of caller-saved registers.)
__device__ __forceinline__ void nlo_sincos_fast(float a, float& s, float& c) {
  const float r = nlo_reduce_2pi(a);
  s = __sinf(r); c = __cosf(r);
}
__device__ __forceinline__ float nlo_cos_fast(float a) { return __cosf(nlo_reduce_2pi(a)); }
__device__ __forceinline__ float nlo_sin_fast(float a) { return __sinf(nlo_reduce_2pi(a)); }
// tanh and the logistic function through MUFU.EX2 / MUFU.RCP (6 instructions instead of libdevice's ~25 with branches):
//   tanh(a) = 1 - 2 / (1 + e^{2a}),  sigmoid(a) = 1 / (1 + e^{-a}).  __expf is good to ~2 ulp of the result, the reciprocal to 1 ulp;
// absolute error <= ~2e-7 everywhere (saturation: e^{2a} -> inf gives exactly 1, -> 0 gives exactly -1).
__device__ __forceinline__ float nlo_rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }   // MUFU.RCP, 1 ulp
__device__ __forceinline__ float nlo_tanh_fast(float a) { return fmaf(-2.f, nlo_rcp_fast(1.f + __expf(2.f * a)), 1.f); }
__device__ __forceinline__ float nlo_sigmoid_fast(float a) { return nlo_rcp_fast(1.f + __expf(-a)); }
// activations of the tensor path: identical to nlo_phi / nlo_phi_d except for the SFU forms above
__device__ __forceinline__ float nlo_phi_tc(float a, int act, float prm) {
  switch (act) {
    case NLO_ACT_SIN: return nlo_sin_fast(prm * a);
    case NLO_ACT_COS_SCALE: return nlo_cos_fast(a) * prm;
    case NLO_ACT_TANH: return nlo_tanh_fast(a);
    case NLO_ACT_SIGMOID: return nlo_sigmoid_fast(a);
    default: return nlo_phi(a, act, prm);
  }
}
__device__ __forceinline__ void nlo_phi_d_tc(float a, int act, float prm, float& v, float& d) {
  switch (act) {
    case NLO_ACT_SIN: { float s, c; nlo_sincos_fast(prm * a, s, c); v = s; d = prm * c; } break;
    case NLO_ACT_COS_SCALE: { float s, c; nlo_sincos_fast(a, s, c); v = prm * c; d = -prm * s; } break;
    case NLO_ACT_TANH: { const float t = nlo_tanh_fast(a); v = t; d = fmaf(-t, t, 1.f); } break;
    case NLO_ACT_SIGMOID: { const float g = nlo_sigmoid_fast(a); v = g; d = g * (1.f - g); } break;
    default: nlo_phi_d(a, act, prm, v, d); break;
  }
}
// second derivative from the value v = phi(a) and first derivative d = phi'(a) (no further transcendental)
__device__ __forceinline__ float nlo_phi_d2_from_vd(int act, float prm, float v, float d) {
  switch (act) {
    case NLO_ACT_TANH: return -2.f * v * d;
    case NLO_ACT_SIGMOID: return d * (1.f - 2.f * v);
    case NLO_ACT_SIN: return -prm * prm * v;
    case NLO_ACT_COS_SCALE: return -v;
    default: return 0.f;
  }
}

// ---- kernels' host-side launchers (one per .cu) -------------------------------------------------
// capi.cu: index of the per-stream slot of `st` in the model (claims a free one; when all are taken by other streams the device is
// drained and the table starts over).  < 0 on error.
int nlo_model_stream_slot(nlo_sdf_model* m, cudaStream_t st);
// sdf_simt.cu
int nlo_sdf_simt_launch(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n,
                        float* s, float* jx, float* jy, cudaStream_t st);
int nlo_sdf_simt_hess_launch(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n,
                             float* hxx, float* hxy, float* hyy, cudaStream_t st);
// sdf_tc.cu
bool nlo_sdf_tc_supported(const nlo_sdf_desc* d);
int nlo_sdf_tc_prepare(nlo_sdf_model* m, const float* weights_host);
// K3 fused into K1 (sdf_tc.cu, ReLU / ReLU H = 128 form): the hard SDF rows of an NLP batch straight from the poses in w
bool nlo_sdf_tc_rows_supported(const nlo_sdf_model* m);
int nlo_sdf_tc_rows_launch(nlo_sdf_model* m, const float* w, size_t P, size_t ld, int n_knots, int nx, int nb, const float* bx, const float* by,
                           float* g_rows, float* jac, const int* nz, int nz_per_knot, cudaStream_t st);
int nlo_sdf_tc_launch(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n,
                      float* s, float* jx, float* jy, cudaStream_t st);
bool nlo_sdf_tc_hess_supported(const nlo_sdf_model* m);
int nlo_sdf_tc_hess_launch(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n,
                           float* s, float* jx, float* jy, float* hxx, float* hxy, float* hyy, cudaStream_t st);
// sdf_tc_hess.cu: value + Jacobian + Hessian of one-hidden-matrix networks with a smooth hidden activation, as GEMMs
bool nlo_sdf_tc_hess_gemm_supported(const nlo_sdf_model* m);
int nlo_sdf_tc_hess_gemm_launch(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n, float* s, float* jx, float* jy,
                                float* hxx, float* hxy, float* hyy, cudaStream_t st);
// sdf_tc_deep.cu: two or three H x H matrices (reached through nlo_sdf_tc_supported / _prepare / _launch)
bool nlo_sdf_tc_deep_supported(const nlo_sdf_desc* d);
int nlo_sdf_tc_deep_prepare(nlo_sdf_model* m, const float* weights_host);
int nlo_sdf_tc_deep_launch(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n,
                           float* s, float* jx, float* jy, cudaStream_t st);
// sdf_tc256.cu (reached through the three nlo_sdf_tc_* entry points above)
bool nlo_sdf_tc256_supported(const nlo_sdf_desc* d);
int nlo_sdf_tc256_prepare(nlo_sdf_model* m, const float* weights_host);
int nlo_sdf_tc256_launch(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n,
                         float* s, float* jx, float* jy, cudaStream_t st);
