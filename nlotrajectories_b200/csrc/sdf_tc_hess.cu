// K1b on tensor cores: value, Jacobian and HESSIAN (sbar * d2s/dp2, the jac_adj1_nn_sdf of _l4c_generated/nn_sdf.cpp:88-104) of
// networks 2 -> H -> H -> 1 (H in {64, 128}) whose hidden activation is smooth (tanh / sigmoid / sine), in one launch.
//
// With two inputs the second derivative needs no reverse-mode tangents.  For directions d in {x, y}
//     hdot_d = phi0'(a0) * W0[:, d]            (tangent of the first layer, SIMT)
//     zdot_d = W1 hdot_d                       (two more FORWARD GEMMs with the same operand image as z1 = W1 h0 + b1)
// and then, with g1 = w2 * phi'(z1) and G0 = W1^T g1 (the reverse GEMM that the Jacobian needs anyway),
//     H = sum_j w2_j phi''(z1_j) zdot_j zdot_j^T  +  sum_k G0_k phi0''(a0_k) W0_k W0_k^T,      J = sum_k G0_k phi0'(a0_k) W0_k.
// So: three forward GEMMs into three accumulators (z1, zdot_x, zdot_y: 3 H tensor-memory columns), one epilogue over all three, one
// reverse GEMM, one final epilogue - 12 fp16 MMA passes per tile against 6 for value + Jacobian, instead of the FP32 warp-per-point
// kernel (sdf_simt.cu: sdf_hess_kernel, 0.046 G points/s).  Tensor memory per tile group: A (H columns: one operand at a time) + 3 D
// (3 H columns) = 4 H: one group per SM at H = 128 (four threads per point), two at H = 64 (two threads per point); every thread
// handles 32 neurons.  The operand images are the first two images of the M = 1 kernel (W1 hi | lo, sdf_tc.cu).
#include "nlo_common.cuh"
#include "tc_ptx.cuh"
#include <cuda_fp16.h>
#include <cmath>
#include <cstring>
#include <mutex>

namespace {

constexpr int HTILE = 128;
__constant__ TcConst csth;

template <int H>
struct HessCfg {
  static constexpr int IMG_HALFS = H * H;
  static constexpr uint32_t GROUP_COLS = 4 * H;
  static constexpr int NGROUPS = 512 / GROUP_COLS;             // 1 (H = 128) or 2 (H = 64)
  static constexpr int SPLIT = H / 32;                         // 4 or 2 threads per point, 32 neurons each
  static constexpr int GROUP_THREADS = HTILE * SPLIT;
  static constexpr int THREADS = NGROUPS * GROUP_THREADS;      // 512
  static constexpr size_t bytes() { return (size_t)2 * IMG_HALFS * 2 + (size_t)NGROUPS * SPLIT * HTILE * 8 * 4 + NGROUPS * 8 + 16 + NGROUPS * 4; }
};

struct HessCtx {
  uint32_t tmem_base, lane_base, sB, mbar_addr, bar_id;
  int* lock;
  float prm0, prm, inv_sw, inv_sc1, st, inv_st, max_w0x, max_w0y, max_b0;
};

// 3-pass split-fp16 GEMM of the group's A operand with the W1 image into the accumulator at column d_col
template <int H>
__device__ __forceinline__ void hess_issue_gemm(const HessCtx& c, uint32_t d_col, bool fwd) {
  constexpr uint32_t COL_AHI = 0, COL_ALO = H / 2;
  constexpr uint32_t IDESC_K = umma_idesc_f16(HTILE, H, 0), IDESC_MN = umma_idesc_f16(HTILE, H, 1);
  tc_fence_after();
  while (atomicCAS(c.lock, 0, 1) != 0) { }
  if (elect_one(1u)) {
    const uint32_t b_hi = c.sB, b_lo = c.sB + (uint32_t)(H * H * 2);
    const uint32_t lbo = fwd ? 16u * H : 128u, sbo = fwd ? 128u : 16u * H;
    const uint32_t kstep_bytes = fwd ? 32u * H : 256u;
    const uint32_t idesc = fwd ? IDESC_K : IDESC_MN;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
      const uint32_t a_col = (pass == 0) ? COL_ALO : COL_AHI;
      const uint32_t b_base = (pass == 1) ? b_lo : b_hi;
#pragma unroll
      for (int ks = 0; ks < H / 16; ++ks)
        tc_mma_f16_ts(c.tmem_base + d_col, c.tmem_base + a_col + ks * 8, umma_desc(b_base + ks * kstep_bytes, lbo, sbo), idesc, (pass | ks) != 0);
    }
    *reinterpret_cast<volatile int*>(c.lock) = 0;        // the lock goes back right behind the last MMA, ahead of the commit (sdf_tc.cu)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(c.mbar_addr) : "memory");
  }
}

// value, first and second derivative of the compile-time activation
template <int A>
__device__ __forceinline__ void act_vdd(float a, float prm, float& v, float& d, float& d2) {
  nlo_phi_d_tc(a, A, prm, v, d);
  d2 = nlo_phi_d2_from_vd(A, prm, v, d);
}

template <int H, int ACT0, int ACT, int SPLIT, int PART>
__device__ __noinline__ uint32_t sdf_hess_tile(HessCtx c, uint32_t phase, float px, float py, float seed, float* __restrict__ part,
                                               float* __restrict__ s_ptr, float* __restrict__ j_ptr, ptrdiff_t jy_off, float* __restrict__ h_ptr,
                                               ptrdiff_t hxy_off, ptrdiff_t hyy_off) {
  constexpr int HH = 32;
  constexpr int C0 = PART * HH;
  constexpr int NT = HTILE * SPLIT;
  constexpr uint32_t COL_AHI = PART * (HH / 2), COL_ALO = H / 2 + PART * (HH / 2);
  const int tg = threadIdx.x % NT;
  const int pt = tg & (HTILE - 1);
  float sc0, inv0;
  row_scale(act_bound(ACT0, c.prm0, fmaf(fabsf(px), c.max_w0x, fmaf(fabsf(py), c.max_w0y, c.max_b0))) + 1e-30f, sc0, inv0);
  // ---- three forward GEMMs: A = h0, then the two tangents of the first layer, each into its own accumulator -------------------------
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      float v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = C0 + 2 * q + e;
        float val, d;
        act_vd<ACT0>(fmaf(csth.w0x[k], px, fmaf(csth.w0y[k], py, csth.b0[k])), ACT0, c.prm0, val, d);
        v[e] = t == 0 ? val * sc0 : d * (t == 1 ? csth.w0x[k] : csth.w0y[k]) * c.st;
      }
      split_pack_f16(v[0], v[1], hi[q], lo[q]);
    }
    if (t > 0) {                                   // the previous GEMM must be done with the operand before it is overwritten
      mbar_wait_addr(c.mbar_addr, phase); phase ^= 1;
      tc_fence_after();
    }
    TmemIO<16>::st(c.lane_base + COL_AHI, hi);
    TmemIO<16>::st(c.lane_base + COL_ALO, lo);
    tc_wait_st();
    tc_fence_before();
    group_bar<NT>(c.bar_id);
    if (tg == 0) hess_issue_gemm<H>(c, (uint32_t)((1 + t) * H), true);
  }
  mbar_wait_addr(c.mbar_addr, phase); phase ^= 1;
  tc_fence_after();
  // ---- epilogue over z1, zdot_x, zdot_y: value, reverse seed, the hidden layer's share of the Hessian ----------------------------------
  float s = PART == 0 ? csth.bout : 0.f;
  float hxx = 0.f, hxy = 0.f, hyy = 0.f;
  {
    uint32_t z[32], zx[32], zy[32];
    tmem_ld32(c.lane_base + 1 * H + C0, z);
    tmem_ld32(c.lane_base + 2 * H + C0, zx);
    tmem_ld32(c.lane_base + 3 * H + C0, zy);
    tc_wait_ld();
    const float u0 = inv0 * c.inv_sw, ut = c.inv_st * c.inv_sw;
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      float g[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = C0 + 2 * q + e;
        float v, d, d2;
        act_vdd<ACT>(fmaf(__uint_as_float(z[2 * q + e]), u0, csth.b1[j]), c.prm, v, d, d2);
        s = fmaf(csth.w2[j], v, s);
        g[e] = csth.w2s[j] * d;
        const float tx = __uint_as_float(zx[2 * q + e]) * ut, ty = __uint_as_float(zy[2 * q + e]) * ut, wd = csth.w2[j] * d2;
        hxx = fmaf(wd * tx, tx, hxx); hxy = fmaf(wd * tx, ty, hxy); hyy = fmaf(wd * ty, ty, hyy);
      }
      split_pack_f16(g[0], g[1], hi[q], lo[q]);
    }
    TmemIO<16>::st(c.lane_base + COL_AHI, hi);
    TmemIO<16>::st(c.lane_base + COL_ALO, lo);
  }
  tc_wait_st();
  tc_fence_before();
  group_bar<NT>(c.bar_id);
  if (tg == 0) hess_issue_gemm<H>(c, (uint32_t)H, false);
  mbar_wait_addr(c.mbar_addr, phase); phase ^= 1;
  tc_fence_after();
  // ---- final epilogue: through layer 0 to the Jacobian and the first layer's share of the Hessian --------------------------------------
  float jx = 0.f, jy = 0.f, h1xx = 0.f, h1xy = 0.f, h1yy = 0.f;
  {
    uint32_t gz[32];
    tmem_ld32(c.lane_base + 1 * H + C0, gz);
    tc_wait_ld();
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      const int k = C0 + q;
      float v, d, d2;
      act_vdd<ACT0>(fmaf(csth.w0x[k], px, fmaf(csth.w0y[k], py, csth.b0[k])), c.prm0, v, d, d2);
      const float G = __uint_as_float(gz[q]);
      const float g0 = G * d, tt = G * d2;
      jx = fmaf(g0, csth.w0x[k], jx); jy = fmaf(g0, csth.w0y[k], jy);
      const float tx = tt * csth.w0x[k], ty = tt * csth.w0y[k];
      h1xx = fmaf(tx, csth.w0x[k], h1xx); h1xy = fmaf(tx, csth.w0y[k], h1xy); h1yy = fmaf(ty, csth.w0y[k], h1yy);
    }
  }
  const float ur = c.inv_sc1 * c.inv_sw;           // the reverse GEMM's accumulator carries the scales of g1 and of the image
  jx *= ur; jy *= ur;
  hxx = fmaf(h1xx, ur, hxx); hxy = fmaf(h1xy, ur, hxy); hyy = fmaf(h1yy, ur, hyy);
  // partial sums of the SPLIT owners of a point meet in shared memory; owner 0 adds them up, applies the adjoint seed and stores
  float* mine = part + (PART * HTILE + pt) * 8;
  mine[0] = s; mine[1] = jx; mine[2] = jy; mine[3] = hxx; mine[4] = hxy; mine[5] = hyy;
  tc_fence_before();
  group_bar<NT>(c.bar_id);                         // (also keeps the next tile's tcgen05.st / MMA behind this tile's TMEM reads)
  tc_fence_after();
  if (PART == 0) {
#pragma unroll
    for (int h = 1; h < SPLIT; ++h) {
      const float* o = part + (h * HTILE + pt) * 8;
      s += o[0]; jx += o[1]; jy += o[2]; hxx += o[3]; hxy += o[4]; hyy += o[5];
    }
    if (s_ptr) *s_ptr = s;
    if (j_ptr) { j_ptr[0] = seed * jx; j_ptr[jy_off] = seed * jy; }
    if (h_ptr) { h_ptr[0] = seed * hxx; h_ptr[hxy_off] = seed * hxy; h_ptr[hyy_off] = seed * hyy; }
  }
  return phase;
}

template <int H, int ACT0, int ACT>
__global__ void __launch_bounds__(HessCfg<H>::THREADS, 1)
sdf_tc_hess_kernel(SdfNetDev net, TcParams prm, float st, const __half* __restrict__ bimg, const float* __restrict__ x, const float* __restrict__ y,
                   const float* __restrict__ sbar, size_t n, float* __restrict__ s_out, float* __restrict__ jx_out, float* __restrict__ jy_out,
                   float* __restrict__ hxx_out, float* __restrict__ hxy_out, float* __restrict__ hyy_out) {
  using Cfg = HessCfg<H>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __half* sB = reinterpret_cast<__half*>(smem_raw);
  float* sPart = reinterpret_cast<float*>(smem_raw + (size_t)2 * Cfg::IMG_HALFS * 2);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sPart + Cfg::NGROUPS * Cfg::SPLIT * HTILE * 8);
  int* lock = reinterpret_cast<int*>(mbar + Cfg::NGROUPS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lock + 1);
  const int t = threadIdx.x, warp = t >> 5;
  const int grp = t / Cfg::GROUP_THREADS, tg = t % Cfg::GROUP_THREADS;
  const int owner = (tg >> 5) >> 2;
  const int pt = tg & (HTILE - 1);
  {
    const uint4* src = reinterpret_cast<const uint4*>(bimg);
    uint4* dst = reinterpret_cast<uint4*>(sB);
    for (int i = t; i < 2 * Cfg::IMG_HALFS / 8; i += Cfg::THREADS) dst[i] = src[i];
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
  HessCtx c;
  const uint32_t tmem_all = *tmem_slot;
  c.tmem_base = tmem_all + (uint32_t)grp * Cfg::GROUP_COLS;
  c.lane_base = c.tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  c.sB = smem_u32(sB);
  c.mbar_addr = smem_u32(mbar + grp);
  c.bar_id = 1 + grp;
  c.lock = lock;
  c.prm0 = net.p0; c.prm = net.p; c.inv_sw = prm.inv_sw; c.inv_sc1 = prm.inv_sc1; c.st = st; c.inv_st = 1.f / st;
  c.max_w0x = prm.max_w0x; c.max_w0y = prm.max_w0y; c.max_b0 = prm.max_b0;
  float* part = sPart + grp * (Cfg::SPLIT * HTILE * 8);
  uint32_t phase = 0;
  const size_t n_tiles = (n + HTILE - 1) / HTILE;
  const size_t stride = (size_t)gridDim.x * Cfg::NGROUPS;
  for (size_t tile = (size_t)blockIdx.x * Cfg::NGROUPS + grp; tile < n_tiles; tile += stride) {
    const size_t i = tile * HTILE + pt;
    const bool valid = i < n;
    const size_t ic = valid ? i : n - 1;
    const float px = x[ic], py = y[ic], seed = sbar ? sbar[ic] : 1.f;
    float* sp = (valid && s_out) ? s_out + i : nullptr;
    float* jp = (valid && jx_out) ? jx_out + i : nullptr;
    float* hp = valid ? hxx_out + i : nullptr;
    const ptrdiff_t jo = jy_out - jx_out, hxyo = hxy_out - hxx_out, hyyo = hyy_out - hxx_out;
#define NLO_HESS_TILE(Q) phase = sdf_hess_tile<H, ACT0, ACT, Cfg::SPLIT, Q>(c, phase, px, py, seed, part, sp, jp, jo, hp, hxyo, hyyo)
    if constexpr (Cfg::SPLIT == 2) { if (owner == 0) NLO_HESS_TILE(0); else NLO_HESS_TILE(1); }
    else { if (owner == 0) NLO_HESS_TILE(0); else if (owner == 1) NLO_HESS_TILE(1); else if (owner == 2) NLO_HESS_TILE(2); else NLO_HESS_TILE(3); }
#undef NLO_HESS_TILE
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_all, 512);
}

unsigned long long g_hess_owner[64] = {0};
std::mutex g_hess_mu;

template <int H, int ACT0, int ACT>
int launch_hess(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n, float* s, float* jx, float* jy, float* hxx,
                float* hxy, float* hyy, cudaStream_t st) {
  using Cfg = HessCfg<H>;
  auto kfn = sdf_tc_hess_kernel<H, ACT0, ACT>;
  const size_t smem = Cfg::bytes();
  static bool attr_set[64] = {false};
  if (!attr_set[m->device & 63]) {
    NLO_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[m->device & 63] = true;
  }
  const size_t tiles = (n + HTILE - 1) / HTILE;
  const size_t want = (tiles + Cfg::NGROUPS - 1) / Cfg::NGROUPS;
  const int grid = (int)(want < (size_t)m->sm_count ? want : (size_t)m->sm_count);
  TcParams prm;
  memcpy(&prm, m->tc_params, sizeof(prm));
  const TcConst* hc = reinterpret_cast<const TcConst*>(m->tc_const);
  // tangents of the first layer: |phi0'(a0) W0[k, d]| <= sup|phi0'| max|W0|; one power-of-two scale for all rows
  const int a0 = (int)m->desc.act0;
  const float dmax0 = a0 == NLO_ACT_SIGMOID ? 0.25f : (a0 == NLO_ACT_SIN || a0 == NLO_ACT_COS_SCALE) ? fabsf(m->desc.p0) : 1.f;
  const float tb = dmax0 * fmaxf(prm.max_w0x, prm.max_w0y);
  int ex = 0;
  if (tb > 0.f) frexpf(tb, &ex);
  const float tscale = ldexpf(1.f, 14 - ex);
  std::lock_guard<std::mutex> lk(g_hess_mu);
  if (g_hess_owner[m->device] != m->uid) {
    NLO_CUDA(cudaDeviceSynchronize());
    NLO_CUDA(cudaMemcpyToSymbol(csth, hc, sizeof(TcConst), 0, cudaMemcpyHostToDevice));
    g_hess_owner[m->device] = m->uid;
  }
  kfn<<<grid, Cfg::THREADS, smem, st>>>(m->net(), prm, tscale, reinterpret_cast<const __half*>(m->d_tc), x, y, sbar, n, s, jx, jy, hxx, hxy, hyy);
  NLO_CHECK_LAUNCH();
  return 0;
}

}  // namespace

// networks with one H x H matrix, H in {64, 128}, smooth hidden activation, and a compiled (first layer, hidden) pair
bool nlo_sdf_tc_hess_gemm_supported(const nlo_sdf_model* m) {
  if (!m->d_tc || m->desc.n_hidden_mats != 1 || !(m->desc.hidden == 64 || m->desc.hidden == 128)) return false;
  const uint32_t a0 = m->desc.act0, a = m->desc.act;
  const bool same = a0 == a && (a == NLO_ACT_TANH || a == NLO_ACT_SIGMOID || a == NLO_ACT_SIN);
  const bool fourier = a0 == NLO_ACT_COS_SCALE && (a == NLO_ACT_TANH || a == NLO_ACT_SIGMOID);
  return same || fourier;
}

int nlo_sdf_tc_hess_gemm_launch(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n, float* s, float* jx, float* jy,
                                float* hxx, float* hxy, float* hyy, cudaStream_t st) {
  if (n == 0) return 0;
  if (!nlo_sdf_tc_hess_gemm_supported(m)) return nlo_fail("tensor-tile Hessian (GEMM form): unsupported network");
  if (!hxx || !hxy || !hyy) return nlo_fail("tensor-tile Hessian: all three outputs are required");
  if ((jx == nullptr) != (jy == nullptr)) return nlo_fail("jx and jy must be requested together");
  const int H = (int)m->desc.hidden, a0 = (int)m->desc.act0, a = (int)m->desc.act;
#define NLO_HESS_ONE(HH, A0, A1) if (H == HH && a0 == A0 && a == A1) return launch_hess<HH, A0, A1>(m, x, y, sbar, n, s, jx, jy, hxx, hxy, hyy, st)
#define NLO_HESS_ALL(HH)                                                                                          \
  NLO_HESS_ONE(HH, NLO_ACT_TANH, NLO_ACT_TANH); NLO_HESS_ONE(HH, NLO_ACT_SIGMOID, NLO_ACT_SIGMOID);               \
  NLO_HESS_ONE(HH, NLO_ACT_SIN, NLO_ACT_SIN); NLO_HESS_ONE(HH, NLO_ACT_COS_SCALE, NLO_ACT_TANH);                  \
  NLO_HESS_ONE(HH, NLO_ACT_COS_SCALE, NLO_ACT_SIGMOID)
  NLO_HESS_ALL(128);
  NLO_HESS_ALL(64);
#undef NLO_HESS_ALL
#undef NLO_HESS_ONE
  return nlo_fail("tensor-tile Hessian (GEMM form): no compiled tile body for this network");
}
