// K1 on tensor cores for networks with TWO or THREE H x H matrices (H in {64, 128}): the reference's deeper defaults -
// core/config.py:203-212 (num_hidden_layers 3 -> two H x H matrices for `mlp`, three for `fourier` / `siren`),
// scripts/run_benchmark.py:64-83 with the YAMLs' num_hidden_layers 2 for `fourier` / `siren` (two H x H matrices),
// core/nn_architectures.py:42-100.  Same arithmetic as sdf_tc.cu (error-compensated split-fp16, three kind::f16 passes per
// contraction, FP32 accumulation in tensor memory, exact power-of-two scaling), chained through the layers:
//
//   layer 0 (SIMT)        h0 = phi0(W0 p + b0)                                -> A (TMEM, row-scaled fp16 hi | lo)
//   for l = 1 .. M        Z_l = H_{l-1} W_l^T        (tcgen05, W_l image K-major)
//     epilogue, l < M     h_l = phi(z_l + b_l) -> A;  phi'(z_l) kept for the way back: one bit per neuron in registers for
//                         ReLU / leaky ReLU, an FP32 column of tensor memory per neuron otherwise
//     epilogue, l = M     s = w_out . phi(z_M + b_M) + b_out;  g_M = w_out * phi'(z_M) -> A
//   for l = M .. 1        Y = G_l W_l                (tcgen05, the SAME image read through an MN-major descriptor)
//     epilogue, l > 1     g_{l-1} = Y * phi'(z_{l-1}) -> A
//     epilogue, l = 1     g_0 = Y * phi0'(a0);  J = sbar * g_0 . W0
//
// Tensor memory per tile group: A (H columns) + D (H columns) [+ (M - 1) H columns of saved derivatives for smooth hidden
// activations]; groups per SM = 512 / that.  All M images (hi | lo each) stay resident in shared memory: 128 KB (M = 2) or
// 192 KB (M = 3) at H = 128.  Everything else - thread = point = TMEM lane (two threads per point at H = 128), small vectors as
// constant-bank operands in fully unrolled bodies, one elected lane issuing the MMAs behind a CTA-wide lock, tiles handed out by
// a global counter - is the design of sdf_tc.cu.
#pragma once
#include "nlo_common.cuh"
#include "tc_ptx.cuh"
#include <cuda_fp16.h>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace {

constexpr int DTILE = 128;         // points per tile == TMEM lanes
constexpr int DEEP_MAXM = 3;

struct TcDeepParams {              // built by nlo_sdf_tc_deep_prepare
  float inv_sw[DEEP_MAXM];         // 1 / (power-of-two scale of the image of W_l), index l - 1
  float c_fwd[DEEP_MAXM];          // max row sum of |W_l|: |z_l| <= c_fwd * max|h_{l-1}| + max|b_l|
  float maxb[DEEP_MAXM];
  float sg[DEEP_MAXM];             // power-of-two scale of the reverse seed g_l (A operand of the reverse GEMM through W_l)
  float inv_sg[DEEP_MAXM];
  float max_w0x, max_w0y, max_b0;
};
struct TcDeepConst {               // small vectors, __constant__ memory: constant-bank operands of the unrolled bodies
  float w0x[128], w0y[128], b0[128], b[DEEP_MAXM][128], wout[128], wouts[128];
  float bout;
};
__constant__ TcDeepConst cd;

template <int H, int M, bool PWL>
struct DeepCfg {
  static constexpr int IMG_HALFS = H * H;
  static constexpr uint32_t GROUP_COLS = 2 * H + (PWL ? 0 : (M - 1) * H);
  static constexpr int NGROUPS = 512 / GROUP_COLS;                  // tensor memory decides how many tiles an SM holds
  static constexpr int SPLIT = 4 / NGROUPS > H / 32 ? H / 32 : 4 / NGROUPS;   // threads per point: 512 threads per CTA whenever the width allows
  static constexpr int GROUP_THREADS = DTILE * SPLIT;
  static constexpr int THREADS = NGROUPS * GROUP_THREADS;
  static constexpr size_t bytes() {
    return (size_t)M * 2 * IMG_HALFS * 2 + (size_t)NGROUPS * SPLIT * DTILE * 4 * 4 + NGROUPS * 8 + 16 + NGROUPS * 4;
  }
};

struct DeepCtx {
  uint32_t tmem_base, lane_base, sB, mbar_addr, bar_id;      // sB: shared-memory address of image 1 hi (images: l hi | l lo, l = 1..M)
  int* lock;
  unsigned int* ctr;
  uint32_t* next_slot;
  uint32_t ctr_bias;
  float prm0, prm;
  TcDeepParams p;
};

// 3-pass split-fp16 GEMM of the group's A operand with image l: forward (K-major: D = A . W_l^T) or reverse (MN-major: D = A . W_l)
template <int H>
__device__ __forceinline__ void deep_issue_mmas(const DeepCtx& c, int l, bool fwd) {
  constexpr uint32_t COL_AHI = 0, COL_ALO = H / 2, COL_D = H;
  constexpr uint32_t IDESC_K = umma_idesc_f16(DTILE, H, 0), IDESC_MN = umma_idesc_f16(DTILE, H, 1);
  const uint32_t b_hi = c.sB + (uint32_t)(l - 1) * 2u * (uint32_t)(H * H * 2), b_lo = b_hi + (uint32_t)(H * H * 2);
  const uint32_t lbo = fwd ? 16u * H : 128u, sbo = fwd ? 128u : 16u * H;
  const uint32_t kstep_bytes = fwd ? 32u * H : 256u;
  const uint32_t idesc = fwd ? IDESC_K : IDESC_MN;
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {                 // smallest terms first: lo.hi, hi.lo, hi.hi
    const uint32_t a_col = (pass == 0) ? COL_ALO : COL_AHI;
    const uint32_t b_base = (pass == 1) ? b_lo : b_hi;
#pragma unroll
    for (int ks = 0; ks < H / 16; ++ks)
      tc_mma_f16_ts(c.tmem_base + COL_D, c.tmem_base + a_col + ks * 8, umma_desc(b_base + ks * kstep_bytes, lbo, sbo), idesc, (pass | ks) != 0);
  }
  *reinterpret_cast<volatile int*>(c.lock) = 0;          // the lock goes back right behind the last MMA, ahead of the commit (sdf_tc.cu)
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(c.mbar_addr) : "memory");
}
template <int H>
__device__ __forceinline__ uint32_t deep_issue_gemm(const DeepCtx& c, int l, bool fwd, bool fetch_next) {
  tc_fence_after();
  while (atomicCAS(c.lock, 0, 1) != 0) { }
  if (elect_one(1u)) deep_issue_mmas<H>(c, l, fwd);
  // next tile of this group: requested here, stored to next_slot at the end of the tile (sdf_tc.cu: the issuing thread must not
  // sit out the counter's round trip in front of its epilogue)
  return (fetch_next && c.ctr) ? atomicAdd(c.ctr, 1u) : 0u;      // (the bias is added where the value is consumed)
}

// One tile of one group.  SPLIT threads own a point: thread HALF (0 .. SPLIT-1) the neurons / D columns [HALF * H / SPLIT, ...).
// Every thread handles HH = 64 or 32 neurons per layer.
template <int H, int ACT0, int ACT, int M, bool PWL, int SPLIT, int HALF>
__device__ __noinline__ uint32_t sdf_deep_tile(DeepCtx c, uint32_t phase, float px, float py, float seed, bool want_jac,
                                               float* __restrict__ part, float* __restrict__ s_ptr, float* __restrict__ j_ptr, ptrdiff_t jy_off) {
  constexpr int HH = H / SPLIT;
  static_assert(HH == 64 || HH == 32, "every thread owns 64 or 32 neurons");
  constexpr int C0 = HALF * HH;
  constexpr int NCH = HH / 32;
  constexpr int NT = DTILE * SPLIT;
  constexpr uint32_t COL_AHI = HALF * (HH / 2), COL_ALO = H / 2 + HALF * (HH / 2), COL_D = H + C0;
  const int tg = threadIdx.x % NT;
  const int pt = tg & (DTILE - 1);
  uint32_t next_tile_idx = 0u;
  // ---- layer 0 -> A ---------------------------------------------------------------------------------------------------------
  float bound = act_bound(ACT0, c.prm0, fmaf(fabsf(px), c.p.max_w0x, fmaf(fabsf(py), c.p.max_w0y, c.p.max_b0))) + 1e-30f;
  float sc, inv;
  row_scale(bound, sc, inv);
#pragma unroll
  for (int cc = 0; cc < NCH; ++cc) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      float v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = C0 + cc * 32 + 2 * q + e;
        v[e] = nlo_phi_tc(fmaf(cd.w0x[k], px, fmaf(cd.w0y[k], py, cd.b0[k])), ACT0, c.prm0) * sc;
      }
      split_pack_f16(v[0], v[1], hi[q], lo[q]);
    }
    TmemIO<16>::st(c.lane_base + COL_AHI + cc * 16, hi);
    TmemIO<16>::st(c.lane_base + COL_ALO + cc * 16, lo);
  }
  uint32_t mask[M > 1 ? M - 1 : 1][2];                      // PWL: bit j of mask[l-1] = (z_l[C0 + j] > 0)
  float s = HALF == 0 ? cd.bout : 0.f;
  // ---- forward through the H x H layers ------------------------------------------------------------------------------------------
#pragma unroll
  for (int l = 1; l <= M; ++l) {
    tc_wait_st();
    tc_fence_before();
    group_bar<NT>(c.bar_id);
    if (tg == 0) { const uint32_t nx = deep_issue_gemm<H>(c, l, true, l == 1); if (l == 1) next_tile_idx = nx; }
    mbar_wait_addr(c.mbar_addr, phase); phase ^= 1;
    tc_fence_after();
    const float unscale = inv * c.p.inv_sw[l - 1];
    float sc_n = 1.f, inv_n = 1.f;
    if (l < M) {
      bound = act_bound(ACT, c.prm, fmaf(c.p.c_fwd[l - 1], bound, c.p.maxb[l - 1])) + 1e-30f;
      row_scale(bound, sc_n, inv_n);
      if (PWL) { mask[l - 1][0] = 0u; mask[l - 1][1] = 0u; }
    }
#pragma unroll
    for (int cc = 0; cc < NCH; ++cc) {
      uint32_t z[32];
      tmem_ld32(c.lane_base + COL_D + cc * 32, z);
      tc_wait_ld();
      uint32_t hi[16], lo[16], dsave[32];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        float o[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = C0 + cc * 32 + 2 * q + e;
          const float zz = fmaf(__uint_as_float(z[2 * q + e]), unscale, cd.b[l - 1][j]);
          float v, d;
          act_vd<ACT>(zz, ACT, c.prm, v, d);
          if (l < M) {
            o[e] = v * sc_n;
            if (PWL) { if (zz > 0.f) mask[l - 1][cc] |= 1u << (2 * q + e); }
            else dsave[2 * q + e] = __float_as_uint(d);
          } else {
            s = fmaf(cd.wout[j], v, s);
            o[e] = cd.wouts[j] * d;
          }
        }
        split_pack_f16(o[0], o[1], hi[q], lo[q]);
      }
      if (l < M || want_jac) {
        TmemIO<16>::st(c.lane_base + COL_AHI + cc * 16, hi);
        TmemIO<16>::st(c.lane_base + COL_ALO + cc * 16, lo);
      }
      if (l < M && !PWL) tmem_st32(c.lane_base + 2 * H + (l - 1) * H + C0 + cc * 32, dsave);
    }
    sc = sc_n; inv = inv_n;
  }
  float jx = 0.f, jy = 0.f;
  if (want_jac) {
    // ---- reverse through the H x H layers -------------------------------------------------------------------------------------------
#pragma unroll
    for (int l = M; l >= 1; --l) {
      tc_wait_st();
      tc_fence_before();
      group_bar<NT>(c.bar_id);
      if (tg == 0) deep_issue_gemm<H>(c, l, false, false);
      mbar_wait_addr(c.mbar_addr, phase); phase ^= 1;
      tc_fence_after();
      const float unscale = c.p.inv_sg[l - 1] * c.p.inv_sw[l - 1];
      if (l > 1) {
        const float rescale = unscale * c.p.sg[l - 2];        // straight to the scale of g_{l-1}
#pragma unroll
        for (int cc = 0; cc < NCH; ++cc) {
          uint32_t y[32], dsave[32];
          tmem_ld32(c.lane_base + COL_D + cc * 32, y);
          if (!PWL) tmem_ld32(c.lane_base + 2 * H + (l - 2) * H + C0 + cc * 32, dsave);
          tc_wait_ld();
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            float g[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float yy = __uint_as_float(y[2 * q + e]) * rescale;
              if (PWL) {
                const bool pos = (mask[l - 2][cc] >> (2 * q + e)) & 1u;
                g[e] = ACT == NLO_ACT_RELU ? (pos ? yy : 0.f) : (pos ? yy : yy * NLO_LEAKY_SLOPE);
              } else {
                g[e] = yy * __uint_as_float(dsave[2 * q + e]);
              }
            }
            split_pack_f16(g[0], g[1], hi[q], lo[q]);
          }
          TmemIO<16>::st(c.lane_base + COL_AHI + cc * 16, hi);
          TmemIO<16>::st(c.lane_base + COL_ALO + cc * 16, lo);
        }
      } else {
        // through layer 0 to the Jacobian
#pragma unroll
        for (int cc = 0; cc < NCH; ++cc) {
          uint32_t gz[32];
          tmem_ld32(c.lane_base + COL_D + cc * 32, gz);
          tc_wait_ld();
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const int k = C0 + cc * 32 + q;
            const float a = fmaf(cd.w0x[k], px, fmaf(cd.w0y[k], py, cd.b0[k]));
            if (ACT0 == NLO_ACT_RELU) {
              if (a > 0.f) { jx = fmaf(__uint_as_float(gz[q]), cd.w0x[k], jx); jy = fmaf(__uint_as_float(gz[q]), cd.w0y[k], jy); }
              continue;
            }
            float v, d;
            act_vd<ACT0>(a, ACT0, c.prm0, v, d);
            const float g0 = __uint_as_float(gz[q]) * d;
            jx = fmaf(g0, cd.w0x[k], jx);
            jy = fmaf(g0, cd.w0y[k], jy);
          }
        }
        const float u = seed * unscale;
        jx *= u; jy *= u;
      }
    }
  }
  if (SPLIT == 1) {
    if (s_ptr) *s_ptr = s;
    if (j_ptr) { j_ptr[0] = jx; j_ptr[jy_off] = jy; }
    if (HALF == 0 && tg == 0 && c.ctr) *c.next_slot = next_tile_idx + c.ctr_bias;
    tc_fence_before();
    group_bar<NT>(c.bar_id);                    // the next tile's tcgen05.st / MMA must not overtake this tile's TMEM reads
    tc_fence_after();
    return phase;
  }
  // partial sums of the SPLIT owners of a point meet in shared memory; owner 0 adds them up and stores
  float* mine = part + (HALF * DTILE + pt) * 4;
  mine[0] = s; mine[1] = jx; mine[2] = jy;
  if (HALF == 0 && tg == 0 && c.ctr) *c.next_slot = next_tile_idx + c.ctr_bias;
  tc_fence_before();
  group_bar<NT>(c.bar_id);                      // (also keeps the next tile's tcgen05.st / MMA behind this tile's TMEM reads)
  tc_fence_after();
  if (HALF == 0) {
#pragma unroll
    for (int h = 1; h < SPLIT; ++h) {
      const float* other = part + (h * DTILE + pt) * 4;
      s += other[0]; jx += other[1]; jy += other[2];
    }
    if (s_ptr) *s_ptr = s;
    if (j_ptr) { j_ptr[0] = jx; j_ptr[jy_off] = jy; }
  }
  // `part` is rewritten by the next tile only after its first group barrier, which owner 0 reaches after these reads
  return phase;
}

template <int H, int ACT0, int ACT, int M>
__global__ void __launch_bounds__((DeepCfg<H, M, (ACT == NLO_ACT_RELU || ACT == NLO_ACT_LEAKY_RELU)>::THREADS), 1)
sdf_tc_deep_kernel(SdfNetDev net, TcDeepParams prm, const __half* __restrict__ bimg, const float* __restrict__ x, const float* __restrict__ y,
                   const float* __restrict__ sbar, size_t n, float* __restrict__ s_out, float* __restrict__ jx_out, float* __restrict__ jy_out,
                   unsigned int* __restrict__ tile_ctr) {
  constexpr bool PWL = (ACT == NLO_ACT_RELU || ACT == NLO_ACT_LEAKY_RELU);
  using Cfg = DeepCfg<H, M, PWL>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __half* sB = reinterpret_cast<__half*>(smem_raw);
  float* sPart = reinterpret_cast<float*>(smem_raw + (size_t)M * 2 * Cfg::IMG_HALFS * 2);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sPart + Cfg::NGROUPS * Cfg::SPLIT * DTILE * 4);
  int* lock = reinterpret_cast<int*>(mbar + Cfg::NGROUPS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lock + 1);
  uint32_t* next_tile = tmem_slot + 1;
  const int t = threadIdx.x, warp = t >> 5;
  const int grp = t / Cfg::GROUP_THREADS, tg = t % Cfg::GROUP_THREADS;
  const int half = (tg >> 5) >> 2;               // which of the SPLIT owners of the point this thread is
  const int pt = tg & (DTILE - 1);
  {
    const uint4* src = reinterpret_cast<const uint4*>(bimg);
    uint4* dst = reinterpret_cast<uint4*>(sB);
    for (int i = t; i < M * 2 * Cfg::IMG_HALFS / 8; i += Cfg::THREADS) dst[i] = src[i];
    if (t == 0) {
      for (int g = 0; g < Cfg::NGROUPS; ++g) mbar_init(mbar + g, 1);
      *lock = 0;
    }
    fence_async_smem();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  DeepCtx c;
  const uint32_t tmem_all = *tmem_slot;
  c.tmem_base = tmem_all + (uint32_t)grp * Cfg::GROUP_COLS;
  c.lane_base = c.tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  c.sB = smem_u32(sB);
  c.mbar_addr = smem_u32(mbar + grp);
  c.bar_id = 1 + grp;
  c.lock = lock;
  c.ctr = tile_ctr; c.next_slot = next_tile + grp; c.ctr_bias = 2 * gridDim.x * Cfg::NGROUPS;
  c.prm0 = net.p0; c.prm = net.p; c.p = prm;
  const bool want_jac = (jx_out != nullptr) || (jy_out != nullptr);
  float* part = sPart + grp * (Cfg::SPLIT * DTILE * 4);
  uint32_t phase = 0;
  const size_t n_tiles = (n + DTILE - 1) / DTILE;
  const size_t stride = (size_t)gridDim.x * Cfg::NGROUPS;
  size_t tile_next = (size_t)blockIdx.x * Cfg::NGROUPS + grp + stride;
  for (size_t tile = (size_t)blockIdx.x * Cfg::NGROUPS + grp; tile < n_tiles;) {
    if (tile_next < n_tiles && half == 0) {
      const size_t in = min(tile_next * DTILE + pt, n - 1);
      asm volatile("prefetch.global.L1 [%0];" :: "l"(x + in));
      asm volatile("prefetch.global.L1 [%0];" :: "l"(y + in));
      if (sbar) asm volatile("prefetch.global.L1 [%0];" :: "l"(sbar + in));
    }
    const size_t i = tile * DTILE + pt;
    const bool valid = i < n;
    const size_t ic = valid ? i : n - 1;
    const float px = x[ic], py = y[ic], seed = sbar ? sbar[ic] : 1.f;
    float* sp = (valid && s_out) ? s_out + i : nullptr;
    float* jp = (valid && jx_out) ? jx_out + i : nullptr;
    const ptrdiff_t jo = jy_out - jx_out;
#define NLO_DEEP_TILE(Q) phase = sdf_deep_tile<H, ACT0, ACT, M, PWL, Cfg::SPLIT, Q>(c, phase, px, py, seed, want_jac, part, sp, jp, jo)
    if constexpr (Cfg::SPLIT == 1) { NLO_DEEP_TILE(0); }
    else if constexpr (Cfg::SPLIT == 2) { if (half == 0) NLO_DEEP_TILE(0); else NLO_DEEP_TILE(1); }
    else { if (half == 0) NLO_DEEP_TILE(0); else if (half == 1) NLO_DEEP_TILE(1); else if (half == 2) NLO_DEEP_TILE(2); else NLO_DEEP_TILE(3); }
#undef NLO_DEEP_TILE
    tile = tile_next;
    tile_next = tile_ctr ? (size_t)next_tile[grp] : tile_next + stride;
  }
  tc_fence_before();
  __syncthreads();
  if (tile_ctr && t == 0) {                      // the last CTA resets the counter pair for the next launch on this stream
    __threadfence();
    if (atomicAdd(tile_ctr + 1, 1u) == gridDim.x - 1) { tile_ctr[0] = 0u; tile_ctr[1] = 0u; __threadfence(); }
  }
  if (warp == 0) tmem_dealloc(tmem_all, 512);
}

unsigned long long g_deep_owner[64] = {0};    // per device: uid of the model whose vectors sit in `cd` (per translation unit)
std::mutex g_deep_mu;

template <int H, int ACT0, int ACT, int M>
int launch_deep(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n, float* s, float* jx, float* jy, cudaStream_t st) {
  constexpr bool PWL = (ACT == NLO_ACT_RELU || ACT == NLO_ACT_LEAKY_RELU);
  using Cfg = DeepCfg<H, M, PWL>;
  auto kfn = sdf_tc_deep_kernel<H, ACT0, ACT, M>;
  const size_t smem = Cfg::bytes();
  static bool attr_set[64] = {false};
  if (!attr_set[m->device & 63]) {
    NLO_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[m->device & 63] = true;
  }
  const size_t tiles = (n + DTILE - 1) / DTILE;
  const size_t want = (tiles + Cfg::NGROUPS - 1) / Cfg::NGROUPS;
  const int grid = (int)(want < (size_t)m->sm_count ? want : (size_t)m->sm_count);
  TcDeepParams prm;
  memcpy(&prm, m->h_deep, sizeof(prm));
  std::lock_guard<std::mutex> lk(g_deep_mu);
  if (g_deep_owner[m->device] != m->uid) {
    NLO_CUDA(cudaDeviceSynchronize());
    NLO_CUDA(cudaMemcpyToSymbol(cd, m->h_deep + 32, sizeof(TcDeepConst), 0, cudaMemcpyHostToDevice));
    g_deep_owner[m->device] = m->uid;
  }
  unsigned int* ctr = nullptr;
  if (tiles > (size_t)grid * Cfg::NGROUPS) {
    const int slot = nlo_model_stream_slot(m, st);
    if (slot < 0) return 1;
    ctr = reinterpret_cast<unsigned int*>(static_cast<char*>(m->d_tc) + m->tc_bytes) + 2 * slot;
  }
  kfn<<<grid, Cfg::THREADS, smem, st>>>(m->net(), prm, reinterpret_cast<const __half*>(m->d_tc), x, y, sbar, n, s, jx, jy, ctr);
  NLO_CHECK_LAUNCH();
  return 0;
}

// the activation pairs of the layer zoo (core/nn_architectures.py:42-100, l4casadi's naive MLP) for one (H, M)
template <int H, int M>
int dispatch_deep(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n, float* s, float* jx, float* jy, cudaStream_t st) {
  const int a0 = (int)m->desc.act0, a = (int)m->desc.act;
#define NLO_DEEP_ONE(A0, A1) if (a0 == A0 && a == A1) return launch_deep<H, A0, A1, M>(m, x, y, sbar, n, s, jx, jy, st)
  NLO_DEEP_ONE(NLO_ACT_RELU, NLO_ACT_RELU); NLO_DEEP_ONE(NLO_ACT_COS_SCALE, NLO_ACT_RELU);
  NLO_DEEP_ONE(NLO_ACT_TANH, NLO_ACT_TANH); NLO_DEEP_ONE(NLO_ACT_SIGMOID, NLO_ACT_SIGMOID);
  NLO_DEEP_ONE(NLO_ACT_LEAKY_RELU, NLO_ACT_LEAKY_RELU); NLO_DEEP_ONE(NLO_ACT_COS_SCALE, NLO_ACT_TANH);
  NLO_DEEP_ONE(NLO_ACT_COS_SCALE, NLO_ACT_SIGMOID); NLO_DEEP_ONE(NLO_ACT_COS_SCALE, NLO_ACT_LEAKY_RELU);
  NLO_DEEP_ONE(NLO_ACT_SIN, NLO_ACT_SIN);
#undef NLO_DEEP_ONE
  return nlo_fail("tensor-tile path (deep): activation pair (%d, %d) has no compiled tile body", a0, a);
}

}  // namespace
