// K1 (FP32 SIMT form): fused learned-SDF value + Jacobian / adjoint, and K1b: Hessian.
//
// Replaces the per-point TorchScript calls behind _l4c_generated/nn_sdf.cpp:57-104
// (nn_sdf, jac_nn_sdf, adj1_nn_sdf, jac_adj1_nn_sdf) for the whole layer zoo of
// core/nn_architectures.py:8-100 and l4casadi's naive MLP (scripts/run_benchmark.py:64-83).
//
// One thread owns one point.  The whole weight blob lives in shared memory for the lifetime of a
// persistent CTA; every thread keeps its activations in a private shared-memory column
// (element k of thread t at [k*blockDim + t]: conflict-free).  Forward and reverse passes both read
// the row-major W_l[out][in] with 128-bit warp-broadcast loads:
//   forward  z[j]  = sum_k W[j][k] h[k]      (chunk of JB outputs in registers, k in steps of 4)
//   reverse  g'[k] = sum_j W[j][k] g[j]      (chunk of JB inputs  in registers, j in steps of 1)
// so no transposed copy is needed.  This is the general path (any H, any depth, any activation);
// the tcgen05 path in sdf_tc.cu takes the H x H contractions when the shape allows.
#include "nlo_common.cuh"

namespace {

constexpr int JB = 32;   // register chunk

template <bool S_SMEM>
struct Scratch {
  float* base; int stride; int H;
  __device__ __forceinline__ float& at(int slot, int k) const { return base[(size_t)(slot * H + k) * stride]; }
};

// ---- forward dense layer: out_pre[j] = b[j] + sum_k W[j][k] in[k] ---------------------------------
// epilogue(j, a) is called with the pre-activation of every output neuron.
template <class In, class Epi>
__device__ __forceinline__ void dense_fwd(const float* __restrict__ W, const float* __restrict__ b, int H, In in, Epi epi) {
  for (int j0 = 0; j0 < H; j0 += JB) {
    float acc[JB];
#pragma unroll
    for (int jj = 0; jj < JB; ++jj) acc[jj] = (j0 + jj < H) ? b[j0 + jj] : 0.f;
    for (int k = 0; k < H; k += 4) {
      const float h0 = in(k), h1 = in(k + 1), h2 = in(k + 2), h3 = in(k + 3);
#pragma unroll
      for (int jj = 0; jj < JB; ++jj) {
        if (j0 + jj < H) {
          const float4 w4 = *reinterpret_cast<const float4*>(W + (size_t)(j0 + jj) * H + k);
          acc[jj] = fmaf(w4.x, h0, acc[jj]);
          acc[jj] = fmaf(w4.y, h1, acc[jj]);
          acc[jj] = fmaf(w4.z, h2, acc[jj]);
          acc[jj] = fmaf(w4.w, h3, acc[jj]);
        }
      }
    }
#pragma unroll
    for (int jj = 0; jj < JB; ++jj)
      if (j0 + jj < H) epi(j0 + jj, acc[jj]);
  }
}

// ---- reverse dense layer: out[k] = sum_j W[j][k] g[j] -------------------------------------------------
template <class In, class Epi>
__device__ __forceinline__ void dense_bwd(const float* __restrict__ W, int H, In g, Epi epi) {
  for (int k0 = 0; k0 < H; k0 += JB) {
    float acc[JB];
#pragma unroll
    for (int kk = 0; kk < JB; ++kk) acc[kk] = 0.f;
    for (int j = 0; j < H; ++j) {
      const float gj = g(j);
#pragma unroll
      for (int kk = 0; kk < JB; kk += 4) {
        if (k0 + kk < H) {
          const float4 w4 = *reinterpret_cast<const float4*>(W + (size_t)j * H + k0 + kk);
          acc[kk + 0] = fmaf(w4.x, gj, acc[kk + 0]);
          acc[kk + 1] = fmaf(w4.y, gj, acc[kk + 1]);
          acc[kk + 2] = fmaf(w4.z, gj, acc[kk + 2]);
          acc[kk + 3] = fmaf(w4.w, gj, acc[kk + 3]);
        }
      }
    }
#pragma unroll
    for (int kk = 0; kk < JB; ++kk)
      if (k0 + kk < H) epi(k0 + kk, acc[kk]);
  }
}

// Scratch slots (value/Jacobian kernel), M >= 1:
//   0            h_0
//   2l-1, 2l     a_l, h_l          l = 1..M-1     (a_l later overwritten in place by g_l)
//   2M-1         g_M
__host__ __device__ inline int vj_slots(int M) { return M == 0 ? 0 : 2 * M; }

template <bool W_SMEM, bool S_SMEM>
__global__ void __launch_bounds__(256) sdf_simt_kernel(SdfNetDev net, const float* __restrict__ x, const float* __restrict__ y,
                                                       const float* __restrict__ sbar, size_t n, float* __restrict__ s_out,
                                                       float* __restrict__ jx_out, float* __restrict__ jy_out, float* __restrict__ ws) {
  extern __shared__ __align__(16) float smem[];
  const int H = net.H, M = net.M;
  const int T = blockDim.x, t = threadIdx.x;
  const float* wts = net.w;
  float* sm_scratch = smem;
  if (W_SMEM) {
    const int cnt = net.count();
    for (int i = t; i < cnt; i += T) smem[i] = net.w[i];
    wts = smem;
    sm_scratch = smem + ((cnt + 3) & ~3);
    __syncthreads();
  }
  const int slots = vj_slots(M);
  Scratch<S_SMEM> sc;
  sc.H = H; sc.stride = T;
  sc.base = S_SMEM ? (sm_scratch + t) : (ws + (size_t)blockIdx.x * slots * H * T + t);

  const float* W0 = wts + net.off_W0();
  const float* b0 = wts + net.off_b0();
  const float* wout = wts + net.off_wout();
  const float bout = wts[net.off_bout()];
  const bool want_jac = (jx_out != nullptr) || (jy_out != nullptr);

  for (size_t i = (size_t)blockIdx.x * T + t; i < n; i += (size_t)gridDim.x * T) {
    const float px = x[i], py = y[i];
    const float seed = sbar ? sbar[i] : 1.f;
    float s = bout, jx = 0.f, jy = 0.f;
    if (M == 0) {
      for (int k = 0; k < H; ++k) {
        const float a = fmaf(W0[2 * k], px, fmaf(W0[2 * k + 1], py, b0[k]));
        float v, d; nlo_phi_d(a, net.act0, net.p0, v, d);
        s = fmaf(wout[k], v, s);
        const float g = seed * wout[k] * d;
        jx = fmaf(g, W0[2 * k], jx); jy = fmaf(g, W0[2 * k + 1], jy);
      }
    } else {
      // layer 0
      for (int k = 0; k < H; ++k) {
        const float a = fmaf(W0[2 * k], px, fmaf(W0[2 * k + 1], py, b0[k]));
        sc.at(0, k) = nlo_phi(a, net.act0, net.p0);
      }
      // hidden layers 1..M
      for (int l = 1; l <= M; ++l) {
        const float* W = wts + net.off_W(l);
        const float* b = wts + net.off_b(l);
        const int in_slot = (l == 1) ? 0 : 2 * (l - 1);
        auto in = [&](int k) { return sc.at(in_slot, k); };
        if (l < M) {
          dense_fwd(W, b, H, in, [&](int j, float a) {
            sc.at(2 * l - 1, j) = a;
            sc.at(2 * l, j) = nlo_phi(a, net.act, net.p);
          });
        } else {
          dense_fwd(W, b, H, in, [&](int j, float a) {
            float v, d; nlo_phi_d(a, net.act, net.p, v, d);
            s = fmaf(wout[j], v, s);
            sc.at(2 * M - 1, j) = seed * wout[j] * d;      // g_M
          });
        }
      }
      if (want_jac) {
        // reverse: g_{l-1} = (W_l^T g_l) * phi'(a_{l-1})
        for (int l = M; l >= 1; --l) {
          const float* W = wts + net.off_W(l);
          const int g_slot = 2 * l - 1;                    // g_l lives where a_l was (or the g_M slot)
          auto g = [&](int j) { return sc.at(g_slot, j); };
          if (l > 1) {
            const int a_slot = 2 * (l - 1) - 1;
            dense_bwd(W, H, g, [&](int k, float v) {
              float d, d2; nlo_phi_d2(sc.at(a_slot, k), net.act, net.p, d, d2);
              sc.at(a_slot, k) = v * d;
            });
          } else {
            dense_bwd(W, H, g, [&](int k, float v) {
              const float a = fmaf(W0[2 * k], px, fmaf(W0[2 * k + 1], py, b0[k]));
              float d, d2; nlo_phi_d2(a, net.act0, net.p0, d, d2);
              const float g0 = v * d;
              jx = fmaf(g0, W0[2 * k], jx); jy = fmaf(g0, W0[2 * k + 1], jy);
            });
          }
        }
      }
    }
    if (s_out) s_out[i] = s;
    if (jx_out) jx_out[i] = jx;
    if (jy_out) jy_out[i] = jy;
  }
}

// ---- Hessian (K1b): forward-over-reverse, the structure of jac_adj1_nn_sdf.pt --------------------------
// Per-thread scratch in global memory (L2-resident columns), layers l = 0..M:
//   slot 5l+0 h_l   5l+1 phi'(a_l)   5l+2 phi''(a_l)   5l+3 adot_l^x   5l+4 adot_l^y
//   then 6 slots: g, gdot^x, gdot^y (ping) and (pong)
__host__ __device__ inline int hess_slots(int M) { return 5 * (M + 1) + 6; }

__global__ void __launch_bounds__(128) sdf_hess_kernel(SdfNetDev net, const float* __restrict__ x, const float* __restrict__ y,
                                                       const float* __restrict__ sbar, size_t n, float* __restrict__ hxx,
                                                       float* __restrict__ hxy, float* __restrict__ hyy, float* __restrict__ ws) {
  const int H = net.H, M = net.M, T = blockDim.x, t = threadIdx.x;
  const float* wts = net.w;
  const int slots = hess_slots(M);
  Scratch<false> sc; sc.H = H; sc.stride = T;
  sc.base = ws + (size_t)blockIdx.x * slots * H * T + t;
  const float* W0 = wts + net.off_W0();
  const float* b0 = wts + net.off_b0();
  const float* wout = wts + net.off_wout();
  const int GP = 5 * (M + 1);
  for (size_t i = (size_t)blockIdx.x * T + t; i < n; i += (size_t)gridDim.x * T) {
    const float px = x[i], py = y[i];
    const float seed = sbar ? sbar[i] : 1.f;
    for (int k = 0; k < H; ++k) {
      const float a = fmaf(W0[2 * k], px, fmaf(W0[2 * k + 1], py, b0[k]));
      float d, d2; nlo_phi_d2(a, net.act0, net.p0, d, d2);
      sc.at(0, k) = nlo_phi(a, net.act0, net.p0); sc.at(1, k) = d; sc.at(2, k) = d2;
      sc.at(3, k) = W0[2 * k]; sc.at(4, k) = W0[2 * k + 1];
    }
    for (int l = 1; l <= M; ++l) {
      const float* W = wts + net.off_W(l);
      const float* b = wts + net.off_b(l);
      const int pi = 5 * (l - 1), po = 5 * l;
      for (int j = 0; j < H; ++j) {
        float a = b[j], ax = 0.f, ay = 0.f;
        for (int k = 0; k < H; ++k) {
          const float w = W[(size_t)j * H + k];
          const float d1 = sc.at(pi + 1, k);
          a = fmaf(w, sc.at(pi, k), a);
          ax = fmaf(w, d1 * sc.at(pi + 3, k), ax);
          ay = fmaf(w, d1 * sc.at(pi + 4, k), ay);
        }
        float d, d2; nlo_phi_d2(a, net.act, net.p, d, d2);
        sc.at(po, j) = nlo_phi(a, net.act, net.p); sc.at(po + 1, j) = d; sc.at(po + 2, j) = d2;
        sc.at(po + 3, j) = ax; sc.at(po + 4, j) = ay;
      }
    }
    // top of the reverse pass
    int cur = GP, nxt = GP + 3;
    for (int j = 0; j < H; ++j) {
      const int po = 5 * M;
      const float sw = seed * wout[j];
      sc.at(cur, j) = sw * sc.at(po + 1, j);
      sc.at(cur + 1, j) = sw * sc.at(po + 2, j) * sc.at(po + 3, j);
      sc.at(cur + 2, j) = sw * sc.at(po + 2, j) * sc.at(po + 4, j);
    }
    for (int l = M; l >= 1; --l) {
      const float* W = wts + net.off_W(l);
      const int pi = 5 * (l - 1);
      for (int k = 0; k < H; ++k) {
        float back = 0.f, bx = 0.f, by = 0.f;
        for (int j = 0; j < H; ++j) {
          const float w = W[(size_t)j * H + k];
          back = fmaf(w, sc.at(cur, j), back);
          bx = fmaf(w, sc.at(cur + 1, j), bx);
          by = fmaf(w, sc.at(cur + 2, j), by);
        }
        const float d1 = sc.at(pi + 1, k), d2 = sc.at(pi + 2, k);
        sc.at(nxt, k) = back * d1;
        sc.at(nxt + 1, k) = fmaf(bx, d1, back * d2 * sc.at(pi + 3, k));
        sc.at(nxt + 2, k) = fmaf(by, d1, back * d2 * sc.at(pi + 4, k));
      }
      const int tmp = cur; cur = nxt; nxt = tmp;
    }
    float vxx = 0.f, vxy = 0.f, vyy = 0.f;
    for (int k = 0; k < H; ++k) {
      const float gx = sc.at(cur + 1, k), gy = sc.at(cur + 2, k);
      vxx = fmaf(gx, W0[2 * k], vxx);        // d(adj_x)/dx
      vxy = fmaf(gy, W0[2 * k], vxy);        // d(adj_x)/dy  (== d(adj_y)/dx)
      vyy = fmaf(gy, W0[2 * k + 1], vyy);
    }
    if (hxx) hxx[i] = vxx;
    if (hxy) hxy[i] = vxy;
    if (hyy) hyy[i] = vyy;
  }
}

int ensure_ws(nlo_sdf_model* m, size_t bytes) {
  if (m->ws_cap >= bytes) return 0;
  if (m->d_ws) cudaFree(m->d_ws);
  m->d_ws = nullptr; m->ws_cap = 0;
  NLO_CUDA(cudaMalloc(&m->d_ws, bytes));
  m->ws_cap = bytes;
  return 0;
}

}  // namespace

int nlo_sdf_simt_launch(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n,
                        float* s, float* jx, float* jy, cudaStream_t st) {
  if (n == 0) return 0;
  SdfNetDev net = m->net();
  const int H = net.H, M = net.M;
  if (H % 4 != 0) return nlo_fail("hidden width must be a multiple of 4 (got %d)", H);
  const size_t max_smem = 227 * 1024;
  const size_t w_bytes = (size_t)((net.count() + 3) & ~3) * sizeof(float);
  const size_t per_thread = (size_t)vj_slots(M) * H * sizeof(float);
  // choose: weights in smem if they leave room for >= 64 scratch columns
  bool w_smem = w_bytes + 64 * per_thread <= max_smem;
  size_t avail = max_smem - (w_smem ? w_bytes : 0);
  int T = per_thread ? (int)(avail / per_thread) : 256;
  bool s_smem = T >= 64;
  if (!s_smem) T = 128;
  T = T > 256 ? 256 : (T / 32) * 32;
  const size_t smem = (w_smem ? w_bytes : 0) + (s_smem ? (size_t)T * per_thread : 0);
  int blocks_per_sm = 1;
  if (smem > 0) { blocks_per_sm = (int)(max_smem / (smem + 1024)); if (blocks_per_sm < 1) blocks_per_sm = 1; if (blocks_per_sm > 8) blocks_per_sm = 8; }
  else blocks_per_sm = 8;
  size_t want = (n + T - 1) / T;
  size_t cap = (size_t)m->sm_count * blocks_per_sm;
  int grid = (int)(want < cap ? want : cap);
  float* ws = nullptr;
  if (!s_smem) {
    if (ensure_ws(m, (size_t)grid * T * per_thread)) return 1;
    ws = m->d_ws;
  }
#define NLO_LAUNCH(WS, SS)                                                                                         \
  do {                                                                                                             \
    auto kfn = sdf_simt_kernel<WS, SS>;                                                                            \
    NLO_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));               \
    kfn<<<grid, T, smem, st>>>(net, x, y, sbar, n, s, jx, jy, ws);                                                 \
  } while (0)
  if (w_smem && s_smem) NLO_LAUNCH(true, true);
  else if (!w_smem && s_smem) NLO_LAUNCH(false, true);
  else if (w_smem && !s_smem) NLO_LAUNCH(true, false);
  else NLO_LAUNCH(false, false);
#undef NLO_LAUNCH
  NLO_CHECK_LAUNCH();
  return 0;
}

int nlo_sdf_simt_hess_launch(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n,
                             float* hxx, float* hxy, float* hyy, cudaStream_t st) {
  if (n == 0) return 0;
  SdfNetDev net = m->net();
  const int T = 128;
  size_t want = (n + T - 1) / T;
  size_t cap = (size_t)m->sm_count * 8;
  int grid = (int)(want < cap ? want : cap);
  const size_t per_thread = (size_t)hess_slots(net.M) * net.H * sizeof(float);
  if (ensure_ws(m, (size_t)grid * T * per_thread)) return 1;
  sdf_hess_kernel<<<grid, T, 0, st>>>(net, x, y, sbar, n, hxx, hxy, hyy, m->d_ws);
  NLO_CHECK_LAUNCH();
  return 0;
}
