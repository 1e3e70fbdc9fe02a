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
#include <cstdlib>

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

// ---- Hessian (K1b): forward-over-reverse, the structure of jac_adj1_nn_sdf.pt ---------------------------
// One warp per point; lanes own output neurons (lane, lane+32, ...).  Every dense product - forward, its two
// tangents, reverse and its two tangents - is "loop over the input index, broadcast the input from shared
// memory, read one coalesced row of weights": the forward passes read the transposed copies W_l^T (net.wt),
// the reverse passes read W_l itself.  Per-warp shared-memory slots of H floats, layers l = 0..M:
//   5l+0 h_l   5l+1 phi'(a_l)   5l+2 phi''(a_l)   5l+3 adot_l^x   5l+4 adot_l^y
//   then g, gdot^x, gdot^y (ping) and (pong)
__host__ __device__ inline int hess_slots(int M) { return 5 * (M + 1) + 6; }

__global__ void __launch_bounds__(256) sdf_hess_kernel(SdfNetDev net, const float* __restrict__ wt, const float* __restrict__ x,
                                                       const float* __restrict__ y, const float* __restrict__ sbar, size_t n,
                                                       float* __restrict__ hxx, float* __restrict__ hxy, float* __restrict__ hyy) {
  extern __shared__ __align__(16) float smem[];
  const int H = net.H, M = net.M;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  float* S = smem + (size_t)warp * hess_slots(M) * H;
  auto at = [&](int slot, int k) -> float& { return S[slot * H + k]; };
  const float* wts = net.w;
  const float* W0 = wts + net.off_W0();
  const float* b0 = wts + net.off_b0();
  const float* wout = wts + net.off_wout();
  const int GP = 5 * (M + 1);
  for (size_t i = (size_t)blockIdx.x * nwarp + warp; i < n; i += (size_t)gridDim.x * nwarp) {
    const float px = x[i], py = y[i];
    const float seed = sbar ? sbar[i] : 1.f;
    for (int k = lane; k < H; k += 32) {
      const float a = fmaf(W0[2 * k], px, fmaf(W0[2 * k + 1], py, b0[k]));
      float d, d2; nlo_phi_d2(a, net.act0, net.p0, d, d2);
      at(0, k) = nlo_phi(a, net.act0, net.p0); at(1, k) = d; at(2, k) = d2;
      at(3, k) = W0[2 * k]; at(4, k) = W0[2 * k + 1];
    }
    __syncwarp();
    for (int l = 1; l <= M; ++l) {
      const float* WT = wt + (size_t)(l - 1) * H * H;           // WT[k][j] = W_l[j][k]
      const float* b = wts + net.off_b(l);
      const int pi = 5 * (l - 1), po = 5 * l;
      for (int j0 = 0; j0 < H; j0 += 128) {                      // up to 4 outputs per lane at a time
        float a[4], ax[4], ay[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { const int j = j0 + lane + 32 * q; a[q] = j < H ? b[j] : 0.f; ax[q] = 0.f; ay[q] = 0.f; }
        for (int k = 0; k < H; ++k) {
          const float h = at(pi, k), d1 = at(pi + 1, k);
          const float tx = d1 * at(pi + 3, k), ty = d1 * at(pi + 4, k);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int j = j0 + lane + 32 * q;
            if (j < H) { const float w = WT[(size_t)k * H + j]; a[q] = fmaf(w, h, a[q]); ax[q] = fmaf(w, tx, ax[q]); ay[q] = fmaf(w, ty, ay[q]); }
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = j0 + lane + 32 * q;
          if (j < H) {
            float d, d2; nlo_phi_d2(a[q], net.act, net.p, d, d2);
            at(po, j) = nlo_phi(a[q], net.act, net.p); at(po + 1, j) = d; at(po + 2, j) = d2; at(po + 3, j) = ax[q]; at(po + 4, j) = ay[q];
          }
        }
      }
      __syncwarp();
    }
    int cur = GP, nxt = GP + 3;
    for (int j = lane; j < H; j += 32) {
      const int po = 5 * M;
      const float sw = seed * wout[j];
      at(cur, j) = sw * at(po + 1, j);
      at(cur + 1, j) = sw * at(po + 2, j) * at(po + 3, j);
      at(cur + 2, j) = sw * at(po + 2, j) * at(po + 4, j);
    }
    __syncwarp();
    for (int l = M; l >= 1; --l) {
      const float* W = wts + net.off_W(l);                       // W[j][k]: row j is contiguous in k
      const int pi = 5 * (l - 1);
      for (int k0 = 0; k0 < H; k0 += 128) {
        float back[4], bx[4], by[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { back[q] = 0.f; bx[q] = 0.f; by[q] = 0.f; }
        for (int j = 0; j < H; ++j) {
          const float g = at(cur, j), gx = at(cur + 1, j), gy = at(cur + 2, j);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int k = k0 + lane + 32 * q;
            if (k < H) { const float w = W[(size_t)j * H + k]; back[q] = fmaf(w, g, back[q]); bx[q] = fmaf(w, gx, bx[q]); by[q] = fmaf(w, gy, by[q]); }
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int k = k0 + lane + 32 * q;
          if (k < H) {
            const float d1 = at(pi + 1, k), d2 = at(pi + 2, k);
            at(nxt, k) = back[q] * d1;
            at(nxt + 1, k) = fmaf(bx[q], d1, back[q] * d2 * at(pi + 3, k));
            at(nxt + 2, k) = fmaf(by[q], d1, back[q] * d2 * at(pi + 4, k));
          }
        }
      }
      __syncwarp();
      const int tmp = cur; cur = nxt; nxt = tmp;
    }
    float vxx = 0.f, vxy = 0.f, vyy = 0.f;
    for (int k = lane; k < H; k += 32) {
      const float gx = at(cur + 1, k), gy = at(cur + 2, k);
      vxx = fmaf(gx, W0[2 * k], vxx);        // d(adj_x)/dx
      vxy = fmaf(gy, W0[2 * k], vxy);        // d(adj_x)/dy  (== d(adj_y)/dx)
      vyy = fmaf(gy, W0[2 * k + 1], vyy);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      vxx += __shfl_xor_sync(0xffffffffu, vxx, o); vxy += __shfl_xor_sync(0xffffffffu, vxy, o); vyy += __shfl_xor_sync(0xffffffffu, vyy, o);
    }
    if (lane == 0) { if (hxx) hxx[i] = vxx; if (hxy) hxy[i] = vxy; if (hyy) hyy[i] = vyy; }
    __syncwarp();
  }
}

// ---- K1 general path, register-tiled (H = 16*TN in {64, 128, 256}, any depth, any activation) -----------------
// A CTA of 256 threads owns a tile of 128 points and runs every layer as a small SGEMM on the FP32 FMA pipe:
//   * the layer input (activations going forward, adjoints going back) is a [H][128] tile in shared memory
//     (row stride 132 floats: conflict-free for both the row-wise 128-bit stores and the broadcast loads);
//   * weights stream from L2 in slabs of 16 rows through a cp.async double buffer - W_l^T rows going forward,
//     W_l rows going back - so the width is not limited by shared memory (H = 256: 256 KB of weights);
//   * thread (tx, ty) accumulates 8 points (8*ty..) x TN neurons (4 consecutive ones per 64-wide group: 64g + 4tx + 0..3) in
//     registers: per k it issues 2 broadcast LDS.128 for the points + TN/4 LDS.128 for the weights for 8*TN FMAs.
// Pre-activations of intermediate layers (M >= 2) go to a tile-private global scratch.
constexpr int GT_P = 128, GT_S = 132, GT_KS = 16, GT_THREADS = 256;
// column (neuron) owned by accumulator slot c of thread tx
__device__ __forceinline__ int gt_col(int tx, int c) { return 64 * (c >> 2) + 4 * tx + (c & 3); }

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_all;" ::: "memory"); }

// acc[i][c] = sum_k As[k][p0+i] * Wg[k][tx + 16c]     (Wg: global, row-major [H][H])
template <int TN>
__device__ __forceinline__ void gemm_tile(const float* __restrict__ Wg, const float* __restrict__ As, float* __restrict__ Ws,
                                          float (&acc)[8][TN], int t, int tx, int p0) {
  constexpr int H = 16 * TN;
  constexpr int SLAB_F4 = GT_KS * H / 4;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < TN; ++c) acc[i][c] = 0.f;
  for (int q = t; q < SLAB_F4; q += GT_THREADS) cp_async16(Ws + 4 * q, Wg + 4 * q);
  for (int s = 0; s < H / GT_KS; ++s) {
    cp_async_commit_wait_all();
    __syncthreads();                                   // slab s landed; everyone finished slab s-1
    const float* Wcur = Ws + (s & 1) * GT_KS * H;
    if (s + 1 < H / GT_KS) {
      float* Wn = Ws + ((s + 1) & 1) * GT_KS * H;
      const float* src = Wg + (size_t)(s + 1) * GT_KS * H;
      for (int q = t; q < SLAB_F4; q += GT_THREADS) cp_async16(Wn + 4 * q, src + 4 * q);
    }
#pragma unroll 4
    for (int kk = 0; kk < GT_KS; ++kk) {
      const float* arow = As + (size_t)(s * GT_KS + kk) * GT_S + p0;
      const float4 a0 = *reinterpret_cast<const float4*>(arow);
      const float4 a1 = *reinterpret_cast<const float4*>(arow + 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[TN];
#pragma unroll
      for (int g = 0; g < TN / 4; ++g) {
        const float4 w4 = *reinterpret_cast<const float4*>(Wcur + kk * H + 64 * g + 4 * tx);
        bv[4 * g + 0] = w4.x; bv[4 * g + 1] = w4.y; bv[4 * g + 2] = w4.z; bv[4 * g + 3] = w4.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < TN; ++c) acc[i][c] = fmaf(av[i], bv[c], acc[i][c]);
    }
  }
  __syncthreads();                                     // all reads of As done: the caller may overwrite the tile
}

template <int TN>
__global__ void __launch_bounds__(GT_THREADS, (TN <= 8 ? 2 : 1))
sdf_gemm_kernel(SdfNetDev net, const float* __restrict__ wt, const float* __restrict__ x, const float* __restrict__ y,
                const float* __restrict__ sbar, size_t n, float* __restrict__ s_out, float* __restrict__ jx_out,
                float* __restrict__ jy_out, float* __restrict__ ws) {
  constexpr int H = 16 * TN;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                              // [H][GT_S]
  float* Ws = As + H * GT_S;                     // [2][GT_KS][H]
  float* Px = Ws + 2 * GT_KS * H;                // [GT_P] x, y, seed of the tile
  float* Py = Px + GT_P;
  float* Sd = Py + GT_P;
  const int M = net.M;
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4, p0 = ty * 8;
  const float* __restrict__ gw = net.w;
  const float* W0 = gw + net.off_W0();
  const float* b0 = gw + net.off_b0();
  const float* wout = gw + net.off_wout();
  const float bout = gw[net.off_bout()];
  const bool want_jac = (jx_out != nullptr) || (jy_out != nullptr);
  float* tile_ws = ws ? ws + (size_t)blockIdx.x * (M > 1 ? M - 1 : 0) * H * GT_P : nullptr;
  const size_t n_tiles = (n + GT_P - 1) / GT_P;
  for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const size_t base = tile * GT_P;
    if (t < GT_P) {
      const size_t i = base + t < n ? base + t : n - 1;
      Px[t] = x[i]; Py[t] = y[i]; Sd[t] = sbar ? sbar[i] : 1.f;
    }
    __syncthreads();
    // layer 0: h0 tile, filled row-wise (coalesced along the points)
    for (int q = t; q < H * (GT_P / 4); q += GT_THREADS) {
      const int r = q / (GT_P / 4), c4 = (q % (GT_P / 4)) * 4;
      const float wx = W0[2 * r], wy = W0[2 * r + 1], bb = b0[r];
      float4 v;
      v.x = nlo_phi(fmaf(wx, Px[c4 + 0], fmaf(wy, Py[c4 + 0], bb)), net.act0, net.p0);
      v.y = nlo_phi(fmaf(wx, Px[c4 + 1], fmaf(wy, Py[c4 + 1], bb)), net.act0, net.p0);
      v.z = nlo_phi(fmaf(wx, Px[c4 + 2], fmaf(wy, Py[c4 + 2], bb)), net.act0, net.p0);
      v.w = nlo_phi(fmaf(wx, Px[c4 + 3], fmaf(wy, Py[c4 + 3], bb)), net.act0, net.p0);
      *reinterpret_cast<float4*>(As + (size_t)r * GT_S + c4) = v;
    }
    __syncthreads();
    float acc[8][TN];
    float sp[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) sp[i] = 0.f;
    // forward through the hidden layers
    for (int l = 1; l <= M; ++l) {
      gemm_tile<TN>(wt + (size_t)(l - 1) * H * H, As, Ws, acc, t, tx, p0);
      const float* bl = gw + net.off_b(l);
#pragma unroll
      for (int c = 0; c < TN; ++c) {
        const int nn = gt_col(tx, c);
        const float bb = bl[nn];
        float o[8];
        if (l < M) {
          float* zrow = tile_ws + ((size_t)(l - 1) * H + nn) * GT_P + p0;
#pragma unroll
          for (int i = 0; i < 8; ++i) { const float a = acc[i][c] + bb; zrow[i] = a; o[i] = nlo_phi(a, net.act, net.p); }
        } else {
          const float w2 = wout[nn];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float v, d; nlo_phi_d(acc[i][c] + bb, net.act, net.p, v, d);
            sp[i] = fmaf(w2, v, sp[i]);
            o[i] = Sd[p0 + i] * w2 * d;                                  // g_M
          }
        }
        float* orow = As + (size_t)nn * GT_S + p0;
        *reinterpret_cast<float4*>(orow) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(orow + 4) = make_float4(o[4], o[5], o[6], o[7]);
      }
      __syncthreads();
    }
    // value: reduce the per-thread partial dot products over the 16 threads that share these 8 points
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = sp[i];
      v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8);
      sp[i] = v + bout;
    }
    if (tx == 0 && s_out) {
#pragma unroll
      for (int i = 0; i < 8; ++i) if (base + p0 + i < n) s_out[base + p0 + i] = sp[i];
    }
    if (want_jac) {
      for (int l = M; l >= 1; --l) {
        gemm_tile<TN>(gw + net.off_W(l), As, Ws, acc, t, tx, p0);       // acc[i][c] = sum_j g_l[j][p] W_l[j][k]
        if (l > 1) {
#pragma unroll
          for (int c = 0; c < TN; ++c) {
            const int kk = gt_col(tx, c);
            const float* zrow = tile_ws + ((size_t)(l - 2) * H + kk) * GT_P + p0;
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { float d, d2; nlo_phi_d2(zrow[i], net.act, net.p, d, d2); o[i] = acc[i][c] * d; }
            float* orow = As + (size_t)kk * GT_S + p0;
            *reinterpret_cast<float4*>(orow) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(orow + 4) = make_float4(o[4], o[5], o[6], o[7]);
          }
          __syncthreads();
        } else {
          float jx[8], jy[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { jx[i] = 0.f; jy[i] = 0.f; }
#pragma unroll
          for (int c = 0; c < TN; ++c) {
            const int kk = gt_col(tx, c);
            const float wx = W0[2 * kk], wy = W0[2 * kk + 1], bb = b0[kk];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float d, d2; nlo_phi_d2(fmaf(wx, Px[p0 + i], fmaf(wy, Py[p0 + i], bb)), net.act0, net.p0, d, d2);
              const float g0 = acc[i][c] * d;
              jx[i] = fmaf(g0, wx, jx[i]); jy[i] = fmaf(g0, wy, jy[i]);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float a = jx[i], b = jy[i];
            a += __shfl_xor_sync(0xffffffffu, a, 1); b += __shfl_xor_sync(0xffffffffu, b, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2); b += __shfl_xor_sync(0xffffffffu, b, 2);
            a += __shfl_xor_sync(0xffffffffu, a, 4); b += __shfl_xor_sync(0xffffffffu, b, 4);
            a += __shfl_xor_sync(0xffffffffu, a, 8); b += __shfl_xor_sync(0xffffffffu, b, 8);
            if (tx == 0 && base + p0 + i < n) { if (jx_out) jx_out[base + p0 + i] = a; if (jy_out) jy_out[base + p0 + i] = b; }
          }
        }
      }
    }
    __syncthreads();                                   // Px/Py/As are rewritten by the next tile
  }
}

template <int TN>
int launch_gemm(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n, float* s, float* jx, float* jy,
                cudaStream_t st);

// Activation workspace of the stream that launches: two streams evaluating one model (the lanes of nlo_nlp_eval_host) must not
// share it - the tail CTAs of one launch and the first CTAs of the next would write the same blocks.
int ensure_ws(nlo_sdf_model* m, cudaStream_t st, size_t bytes, float** out) {
  const int slot = nlo_model_stream_slot(m, st);
  if (slot < 0) return 1;
  if (m->ws_cap[slot] < bytes) {
    if (m->d_ws[slot]) cudaFree(m->d_ws[slot]);          // (cudaFree waits for the device: no launch still reads the old block)
    m->d_ws[slot] = nullptr; m->ws_cap[slot] = 0;
    NLO_CUDA(cudaMalloc(&m->d_ws[slot], bytes));
    m->ws_cap[slot] = bytes;
  }
  *out = m->d_ws[slot];
  return 0;
}

template <int TN>
int launch_gemm(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n, float* s, float* jx, float* jy,
                cudaStream_t st) {
  constexpr int H = 16 * TN;
  SdfNetDev net = m->net();
  const size_t smem = (size_t)(H * GT_S + 2 * GT_KS * H + 3 * GT_P) * sizeof(float);
  auto kfn = sdf_gemm_kernel<TN>;
  static bool attr_set[64] = {false};                    // per device: the attribute call is not free on a 20 us path
  if (!attr_set[m->device & 63]) {
    NLO_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[m->device & 63] = true;
  }
  const int per_sm = TN <= 8 ? 2 : 1;
  const size_t tiles = (n + GT_P - 1) / GT_P;
  const size_t cap = (size_t)m->sm_count * per_sm;
  const int grid = (int)(tiles < cap ? tiles : cap);
  float* ws = nullptr;
  if (net.M > 1) {
    if (ensure_ws(m, st, (size_t)grid * (net.M - 1) * H * GT_P * sizeof(float), &ws)) return 1;
  }
  kfn<<<grid, GT_THREADS, smem, st>>>(net, m->d_wt, x, y, sbar, n, s, jx, jy, ws);
  NLO_CHECK_LAUNCH();
  return 0;
}

}  // namespace

int nlo_sdf_simt_launch(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n,
                        float* s, float* jx, float* jy, cudaStream_t st) {
  if (n == 0) return 0;
  SdfNetDev net = m->net();
  const int H = net.H, M = net.M;
  if (H % 4 != 0) return nlo_fail("hidden width must be a multiple of 4 (got %d)", H);
  if (M >= 1 && (H == 64 || H == 128 || H == 256) && !getenv("NLO_B200_SIMT_LEGACY")) {
    if (H == 64) return launch_gemm<4>(m, x, y, sbar, n, s, jx, jy, st);
    if (H == 128) return launch_gemm<8>(m, x, y, sbar, n, s, jx, jy, st);
    return launch_gemm<16>(m, x, y, sbar, n, s, jx, jy, st);
  }
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
    if (ensure_ws(m, st, (size_t)grid * T * per_thread, &ws)) return 1;
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
  const size_t per_warp = (size_t)hess_slots(net.M) * net.H * sizeof(float);
  int warps = (int)((200 * 1024) / per_warp);
  if (warps < 1) return nlo_fail("network too large for the Hessian kernel (H=%d, M=%d)", net.H, net.M);
  if (warps > 8) warps = 8;
  const size_t smem = per_warp * warps;
  NLO_CUDA(cudaFuncSetAttribute(sdf_hess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
  size_t want = (n + warps - 1) / warps;
  int per_sm = (int)((227 * 1024) / (smem + 1024)); if (per_sm < 1) per_sm = 1; if (per_sm > 4) per_sm = 4;
  size_t cap = (size_t)m->sm_count * per_sm;
  int grid = (int)(want < cap ? want : cap);
  sdf_hess_kernel<<<grid, warps * 32, smem, st>>>(net, m->d_wt, x, y, sbar, n, hxx, hxy, hyy);
  NLO_CHECK_LAUNCH();
  return 0;
}
