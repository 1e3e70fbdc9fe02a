// K1 (tcgen05 form): fused learned-SDF value + Jacobian / adjoint with the two H x H contractions on
// 5th-generation tensor cores: error-compensated split-fp16 (hi + lo, three products) with FP32
// accumulation in tensor memory, i.e. FP32-grade results at three kind::f16 MMA passes.
//
// Replaces the per-point TorchScript calls behind _l4c_generated/nn_sdf.cpp:57-83 (nn_sdf, jac_nn_sdf,
// adj1_nn_sdf) for networks 2 -> H -> H -> 1 with H in {64, 128}: the benchmark model
// (scripts/run_benchmark.py:65 with benchmarks/*.yaml model: hidden_dim 128, num_hidden_layers 2) and
// the shipped FourierMLP-128 (_l4c_generated/nn_sdf.pt).
//
// One persistent CTA of 512 threads per SM; its threads form independent tile groups (H = 128: two groups of 256
// threads, two threads per point; H = 64: four groups of 128) that share one shared-memory copy of the operand images
// and each own 2H tensor-memory columns, an mbarrier and a named barrier.  Tile = 128 points == the 128 TMEM lanes;
// tiles are handed out by a global counter (SMs do not all run at the same speed).  While one group's MMAs run, the
// other groups' SIMT phases fill the SM; an MMA-issue lock keeps the groups' GEMMs first-come-first-served.
// The small vectors (W0, b0, b1, w2) sit in __constant__ memory and every loop is fully unrolled, so they enter the
// arithmetic as constant-bank operands and shared memory is left to the tensor core's B-operand fetches:
//   layer 0 (SIMT)   h0 = phi0(W0 p + b0)  -> row-scaled, split into fp16 hi/lo -> tcgen05.st -> A (TMEM)
//   GEMM 1 (tcgen05) Z1[128 x H] = H0 . W1^T   3 passes (lo.hi, hi.lo, hi.hi), B = W1 K-major in smem
//   epilogue 1       tcgen05.ld Z1; s = w2.phi(z1+b1)+b2; g1 = w2*phi'(z1+b1) -> A (TMEM)
//   GEMM 2 (tcgen05) G0[128 x H] = G1 . W1
//                      ReLU hidden layer: G1 = mask . diag(w2), so A = the 0/1 mask (exact in fp16, hi only) and
//                      B = V = diag(w2) W1 (its own hi/lo images): 2 passes instead of 3 and no split in epilogue 1;
//                      other activations: 3 passes over the W1 image read through an MN-major descriptor
//   epilogue 2       tcgen05.ld G0; g0 = G0 * phi0'(a0); J = sbar * g0 . W0   (the adjoint seed is applied last)
// fp16 has tf32's 11 significant bits at twice the MMA rate and half the bytes; its narrow exponent is
// handled by exact power-of-two scaling: W1 and V by one global factor each (host), every A row (= point) by its own
// factor from a cheap bound on the row, all undone in the epilogues.
// W1*S is split once on the host into fp16 hi + fp16 lo, each stored in UMMA core-matrix order (no swizzle):
// element (n,k) at ((k/8)*(H/8) + n/8)*128 + (n%8)*16 + (k%8)*2 bytes, which is simultaneously the
// canonical K-major layout of B(n,k) = W1[n][k] (LBO = 16H, SBO = 128) and the canonical MN-major layout
// of B'(i,j) = W1[j][i] (LBO = 128, SBO = 16H).  (tf32 cannot do this: its MN-major form exists only in the
// 128B_BASE32B swizzle, which has no K-major twin, and two tf32 images of W1 do not fit in shared memory.)
#include "nlo_common.cuh"
#include "tc_ptx.cuh"
#include <cuda_fp16.h>
#include <vector>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace {

constexpr int TILE = 128;          // points per tile == TMEM lanes

__constant__ TcConst cst;

// One persistent CTA of 512 threads per SM.  Its threads form NGROUPS independent tile groups (H = 128: two groups of
// 256 threads, two threads per point; H = 64: four groups of 128 threads) - each group owns 2H tensor-memory columns,
// one mbarrier and one named barrier, and walks its own tiles; the groups meet only at the MMA-issue lock.  While one
// group's MMAs run, the other groups' SIMT phases fill the SM.  All groups share one copy of the operand images.
template <int H>
struct TcCfg {
  static constexpr int IMG_HALFS = H * H;                           // one fp16 image
  static constexpr uint32_t TMEM_COLS = 2 * H;                      // per group: A hi H/2 | A lo H/2 | D H
  static constexpr int SPLIT = (H >= 128) ? 2 : 1;                  // threads per point
  static constexpr int GROUP_THREADS = TILE * SPLIT;
  static constexpr int NGROUPS = 512 / TMEM_COLS;
  static constexpr int THREADS = NGROUPS * GROUP_THREADS;           // 512
  static constexpr size_t bytes(bool relu_hidden) {
    return (size_t)(relu_hidden ? 4 : 2) * IMG_HALFS * 2 + (size_t)NGROUPS * 2 * TILE * 4 * 4 + NGROUPS * 8 + 16 + NGROUPS * 4 +
           (size_t)NGROUPS * 2 * 3 * TILE * 4;      // + the coordinate staging of the next tile, double-buffered per group
  }
};

struct TileCtx {
  uint32_t tmem_base, lane_base, sB_hi, sB_lo, sV_hi, sV_lo, mbar_addr, bar_id;
  int* lock;
  unsigned int* ctr;       // global tile counter (dynamic scheduling) or nullptr (static striding)
  uint32_t* next_slot;     // the group's next tile index, fetched while GEMM 1 runs
  uint32_t ctr_bias;
  ptrdiff_t hxy_off, hyy_off;   // hxy / hyy outputs relative to hxx (HESS variants)
  int act0, act;
  float prm0, prm, inv_sw, unscale2, max_w0x, max_w0y, max_b0;
  long long* dbg;          // optional phase timeline (NLO_B200_TC_TIMELINE): clock stamps per tile of CTA 0, groups 0 / 1
};

template <int H, bool RELU_BWD>
__device__ __forceinline__ void issue_mmas(const TileCtx& c, bool fwd, long long* dbg = nullptr) {
  constexpr uint32_t COL_AHI = 0, COL_ALO = H / 2, COL_D = H;
  constexpr uint32_t IDESC_K = umma_idesc_f16(TILE, H, 0), IDESC_MN = umma_idesc_f16(TILE, H, 1);
  if (!fwd && RELU_BWD) {
    if (dbg) dbg[30] = clock64();
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {               // smaller term first: mask.lo, mask.hi
      const uint32_t b_base = (pass == 0) ? c.sV_lo : c.sV_hi;
#pragma unroll
      for (int ks = 0; ks < H / 16; ++ks)
        tc_mma_f16_ts(c.tmem_base + COL_D, c.tmem_base + COL_AHI + ks * 8, umma_desc(b_base + ks * 32u * H, 16u * H, 128u), IDESC_K,
                      (pass | ks) != 0);
    }
    if (dbg) dbg[31] = clock64();
  } else {
    const uint32_t lbo = fwd ? 16u * H : 128u, sbo = fwd ? 128u : 16u * H;
    const uint32_t kstep_bytes = fwd ? 32u * H : 256u;     // 16 k: two K-adjacent core matrices
    const uint32_t idesc = fwd ? IDESC_K : IDESC_MN;
    if (dbg && fwd) dbg[13] = clock64();
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {               // smallest terms first: lo.hi, hi.lo, hi.hi
      const uint32_t a_col = (pass == 0) ? COL_ALO : COL_AHI;
      const uint32_t b_base = (pass == 1) ? c.sB_lo : c.sB_hi;
#pragma unroll
      for (int ks = 0; ks < H / 16; ++ks)
        tc_mma_f16_ts(c.tmem_base + COL_D, c.tmem_base + a_col + ks * 8, umma_desc(b_base + ks * kstep_bytes, lbo, sbo), idesc,
                      (pass | ks) != 0);
      if (dbg && fwd) dbg[pass == 0 ? 14 : pass == 1 ? 15 : 29] = clock64();
    }
  }
  // The lock goes back right behind the last MMA, ahead of the commit, with a plain store: it only shapes the order in which the groups'
  // GEMMs enter the tensor queue.  (With commit + __threadfence_block() + atomicExch in front of the release the lock stayed taken for
  // 300 - 600 cycles after the last MMA - NLO_B200_TC_TIMELINE - while the other group's issuer waited for it with the queue draining.)
  *reinterpret_cast<volatile int*>(c.lock) = 0;
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(c.mbar_addr) : "memory");
}
// Lane 0 of the group issues the MMAs of a GEMM and commits them to the group's mbarrier.  The CTA-wide lock makes the
// groups' GEMMs run first-come-first-served instead of interleaved instruction by instruction (tcgen05.mma issue
// blocks while the tensor queue is full, so two concurrent issuers stretch both GEMMs to the sum of the two).
// The MMAs sit under elect.sync (single-lane mask): ptxas then emits them on the uniform datapath, 2 instructions per
// MMA, instead of wrapping each one in a ~16-instruction R2UR waterfall loop.  Measured on B200 (G points/s, ReLU
// H=128 / H=64): this form 6.39 / 15.97; plain single lane 6.02 / 12.54; whole warp converged + elect 5.54 / 15.15.
template <int H, bool RELU_BWD>
__device__ __forceinline__ uint32_t issue_gemm(const TileCtx& c, bool fwd, long long* dbg) {
  tc_fence_after();
  while (atomicCAS(c.lock, 0, 1) != 0) { }
  if (dbg) dbg[fwd ? 10 : 11] = clock64();
  if (elect_one(1u)) issue_mmas<H, RELU_BWD>(c, fwd, dbg);
  // next tile of this group: requested here, consumed (stored to next_slot) at the end of the tile - the round trip to the one hot
  // counter is longer than the GEMM, and with the store right here the issuing thread sat it out while its group waited for it
  return (fwd && c.ctr) ? atomicAdd(c.ctr, 1u) : 0u;      // (the bias is added where the value is consumed, not here)
}
// One tile (128 points) of one group.  SPLIT = 2: two threads own a point - HALF 0 the low half of the neurons / D
// columns, HALF 1 the high half; they meet only to add three partial sums per point through shared memory.
// Deliberately NOT inlined into the persistent loop: the constants are loop-invariant, and an inlined body makes
// the compiler hoist hundreds of them into registers (and spill) instead of feeding them to the arithmetic from
// the constant bank.
// HESS (ReLU hidden layer only): also sbar * Hessian.  With phi'' = 0 in the hidden layer the forward-over-reverse
// tangents need no further GEMM: H = sum_k G0[k] phi0''(a0[k]) W0[k,:] W0[k,:]^T, and G0 is what epilogue 2 already holds.
template <int H, int ACT0, int ACT, bool FULL, int SPLIT, int HALF, bool HESS>
__device__ __noinline__ uint32_t sdf_tc_tile(TileCtx c, uint32_t phase, float px, float py, float seed, bool want_jac,
                                             float* __restrict__ part, float* __restrict__ s_ptr, float* __restrict__ j_ptr,
                                             ptrdiff_t jy_off, float* __restrict__ h_ptr) {
  constexpr int HH = H / SPLIT;                    // neurons per thread
  constexpr int C0 = HALF * HH;                    // first neuron / D column of this thread
  constexpr int NCH = HH / 32;                     // 32-neuron chunks per thread
  constexpr int NT = TILE * SPLIT;                 // threads of the group
  constexpr bool RH = (ACT == NLO_ACT_RELU);       // ReLU hidden layer: mask x V reverse GEMM
  constexpr uint32_t COL_AHI = HALF * (HH / 2), COL_ALO = H / 2 + HALF * (HH / 2), COL_D = H + C0;
  const int tg = threadIdx.x % NT;
  const int pt = tg & (TILE - 1);
  long long* dbg = (c.dbg && pt == 0) ? c.dbg + HALF * 16 : nullptr;
#define TC_STAMP(i) do { if (dbg) dbg[i] = clock64(); } while (0)
  TC_STAMP(0);
  if (dbg) { unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); dbg[12] = (long long)gt; }
  uint32_t next_tile_idx = 0u;
  // ---- layer 0 -> A operand (row-scaled fp16 hi/lo) ------------------------------------------------------------
  float sc0, inv0;
  row_scale(act_bound(c.act0, c.prm0, fmaf(fabsf(px), c.max_w0x, fmaf(fabsf(py), c.max_w0y, c.max_b0))) + 1e-30f, sc0, inv0);
  auto layer0 = [&](int cc) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      float v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = C0 + cc * 32 + 2 * q + e;
        const float a = fmaf(cst.w0x[k], px, fmaf(cst.w0y[k], py, cst.b0[k]));
        if (ACT0 >= 0) v[e] = nlo_phi_tc(a, ACT0, c.prm0) * sc0; else v[e] = nlo_phi_tc(a, c.act0, c.prm0) * sc0;
      }
      split_pack_f16(v[0], v[1], hi[q], lo[q]);
    }
    TmemIO<16>::st(c.lane_base + COL_AHI + cc * 16, hi);
    TmemIO<16>::st(c.lane_base + COL_ALO + cc * 16, lo);
  };
  if (FULL) {
#pragma unroll
    for (int cc = 0; cc < NCH; ++cc) layer0(cc);
  } else {
#pragma unroll 1
    for (int cc = 0; cc < NCH; ++cc) layer0(cc);
  }
  TC_STAMP(1);
  tc_wait_st();
  tc_fence_before();
  group_bar<NT>(c.bar_id);
  TC_STAMP(8);
  if (tg == 0) next_tile_idx = issue_gemm<H, RH>(c, true, dbg);
  TC_STAMP(2);
  mbar_wait_addr(c.mbar_addr, phase); phase ^= 1;
  tc_fence_after();
  TC_STAMP(3);

  // ---- epilogue 1: value, and the reverse seed g1 = w2 * phi'(z1) -> A operand (the adjoint seed sbar is linear ----
  // ---- in everything downstream and multiplies the Jacobian at the very end)                                   ----
  float s = HALF == 0 ? cst.bout : 0.f;
  const float unscale1 = inv0 * c.inv_sw;
  auto epi1 = [&](int cc, const uint32_t* z) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      float g[2];
      uint32_t mask = 0u;                                  // RH: the pair of fp16 0 / 1 values, built from the two predicates
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = C0 + cc * 32 + 2 * q + e;
        const float zz = fmaf(__uint_as_float(z[2 * q + e]), unscale1, cst.b1[j]);
        if (RH) {
          // one predicate serves the value (predicated FFMA instead of max + FFMA) and the mask, which is exact in fp16:
          // the w2 factor of the reverse seed lives in V = diag(w2) W1
          if (zz > 0.f) { s = fmaf(cst.w2[j], zz, s); mask |= (e == 0 ? 0x3C00u : 0x3C000000u); }
        } else {
          float v, d;
          act_vd<ACT>(zz, c.act, c.prm, v, d);
          s = fmaf(cst.w2[j], v, s);
          g[e] = cst.w2s[j] * d;
        }
      }
      if (RH) hi[q] = mask; else split_pack_f16(g[0], g[1], hi[q], lo[q]);
    }
    if (want_jac) {
      TmemIO<16>::st(c.lane_base + COL_AHI + cc * 16, hi);
      if (!RH) TmemIO<16>::st(c.lane_base + COL_ALO + cc * 16, lo);
    }
  };
  // All of this thread's D columns in flight at once: one exposed tcgen05.ld latency per epilogue instead of one per chunk.
  // Measured: +1 % for H = 128 with a cheap first layer, -1 % for H = 64 and for the Fourier first layer (register pressure).
  constexpr bool LD_ALL = FULL && H == 128 && ACT0 != NLO_ACT_COS_SCALE;
  if (LD_ALL) {
    uint32_t z[NCH][32];
#pragma unroll
    for (int cc = 0; cc < NCH; ++cc) tmem_ld32(c.lane_base + COL_D + cc * 32, z[cc]);
    tc_wait_ld();
#pragma unroll
    for (int cc = 0; cc < NCH; ++cc) epi1(cc, z[cc]);
  } else if (FULL) {
#pragma unroll
    for (int cc = 0; cc < NCH; ++cc) {
      uint32_t z[32];
      tmem_ld32(c.lane_base + COL_D + cc * 32, z);
      tc_wait_ld();
      epi1(cc, z);
    }
  } else {
#pragma unroll 1
    for (int cc = 0; cc < NCH; ++cc) {
      uint32_t z[32];
      tmem_ld32(c.lane_base + COL_D + cc * 32, z);
      tc_wait_ld();
      epi1(cc, z);
    }
  }
  float jx = 0.f, jy = 0.f, hxx = 0.f, hxy = 0.f, hyy = 0.f;
  if (want_jac) {
    TC_STAMP(4);
    tc_wait_st();
    tc_fence_before();
    group_bar<NT>(c.bar_id);
    TC_STAMP(9);
    if (tg == 0) issue_gemm<H, RH>(c, false, dbg);
    TC_STAMP(5);
    mbar_wait_addr(c.mbar_addr, phase); phase ^= 1;
    tc_fence_after();
    TC_STAMP(6);
    // ---- epilogue 2: through layer 0 to the Jacobian ---------------------------------------------------------------
    auto epi2 = [&](int cc, const uint32_t* gz) {
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const int k = C0 + cc * 32 + q;
        const float a = fmaf(cst.w0x[k], px, fmaf(cst.w0y[k], py, cst.b0[k]));
        if (ACT0 == NLO_ACT_RELU && !HESS) {
          // phi0' is 0 / 1: two predicated FFMAs instead of a compare-to-float, a multiply and two FFMAs
          if (a > 0.f) { jx = fmaf(__uint_as_float(gz[q]), cst.w0x[k], jx); jy = fmaf(__uint_as_float(gz[q]), cst.w0y[k], jy); }
          continue;
        }
        float v, d;
        act_vd<ACT0>(a, c.act0, c.prm0, v, d);
        const float g0 = __uint_as_float(gz[q]) * d;
        jx = fmaf(g0, cst.w0x[k], jx);
        jy = fmaf(g0, cst.w0y[k], jy);
        if (HESS) {
          const float tt = __uint_as_float(gz[q]) * nlo_phi_d2_from_vd(ACT0 >= 0 ? ACT0 : c.act0, c.prm0, v, d);
          const float tx = tt * cst.w0x[k], ty = tt * cst.w0y[k];
          hxx = fmaf(tx, cst.w0x[k], hxx); hxy = fmaf(tx, cst.w0y[k], hxy); hyy = fmaf(ty, cst.w0y[k], hyy);
        }
      }
    };
    if (LD_ALL) {
      uint32_t gz[NCH][32];
#pragma unroll
      for (int cc = 0; cc < NCH; ++cc) tmem_ld32(c.lane_base + COL_D + cc * 32, gz[cc]);
      tc_wait_ld();
#pragma unroll
      for (int cc = 0; cc < NCH; ++cc) epi2(cc, gz[cc]);
    } else if (FULL) {
#pragma unroll
      for (int cc = 0; cc < NCH; ++cc) {
        uint32_t gz[32];
        tmem_ld32(c.lane_base + COL_D + cc * 32, gz);
        tc_wait_ld();
        epi2(cc, gz);
      }
    } else {
#pragma unroll 1
      for (int cc = 0; cc < NCH; ++cc) {
        uint32_t gz[32];
        tmem_ld32(c.lane_base + COL_D + cc * 32, gz);
        tc_wait_ld();
        epi2(cc, gz);
      }
    }
    const float unscale2 = seed * c.unscale2;
    jx *= unscale2; jy *= unscale2;
    if (HESS) { hxx *= unscale2; hxy *= unscale2; hyy *= unscale2; }
  }
  TC_STAMP(7);
  if (SPLIT == 1) {
    if (s_ptr) *s_ptr = s;
    if (j_ptr) { j_ptr[0] = jx; j_ptr[jy_off] = jy; }
    if (HESS && h_ptr) { h_ptr[0] = hxx; h_ptr[c.hxy_off] = hxy; h_ptr[c.hyy_off] = hyy; }
    if (HALF == 0 && tg == 0 && c.ctr) *c.next_slot = next_tile_idx + c.ctr_bias;
    asm volatile("cp.async.wait_all;" ::: "memory");       // the next tile's coordinates (staged by the persistent loop) have landed
    // the next tile's tcgen05.st / MMA must not overtake this tile's TMEM reads
    tc_fence_before();
    group_bar<NT>(c.bar_id);
    tc_fence_after();
    return phase;
  }
  // ---- hand the other half what it stores: HALF 0 writes s, jx (hxx, hxy), HALF 1 writes jy (hyy) ------------------
  float* mine = part + (HALF * TILE + pt) * 4;
  if (HALF == 0) { mine[0] = jy; if (HESS) mine[1] = hyy; }
  else { mine[0] = s; mine[1] = jx; if (HESS) { mine[2] = hxx; mine[3] = hxy; } }
  if (HALF == 0 && tg == 0 && c.ctr) *c.next_slot = next_tile_idx + c.ctr_bias;
  asm volatile("cp.async.wait_all;" ::: "memory");         // the next tile's coordinates (staged by the persistent loop) have landed
  // (this barrier also keeps the next tile's tcgen05.st / MMA from overtaking this tile's TMEM reads)
  tc_fence_before();
  group_bar<NT>(c.bar_id);
  tc_fence_after();
  const float* other = part + ((1 - HALF) * TILE + pt) * 4;
  if (HALF == 0) {
    if (s_ptr) *s_ptr = s + other[0];
    if (j_ptr) *j_ptr = jx + other[1];
    if (HESS && h_ptr) { h_ptr[0] = hxx + other[2]; h_ptr[c.hxy_off] = hxy + other[3]; }
  } else {
    if (j_ptr) *j_ptr = jy + other[0];
    if (HESS && h_ptr) h_ptr[c.hyy_off] = hyy + other[1];
  }
  // `part` is rewritten by the next tile only after its first group barrier, which every thread reaches after these reads
  if (HALF == 0) TC_STAMP(27);
  return phase;
}

template <int H, int ACT0, int ACT, bool HESS>
__global__ void __launch_bounds__(TcCfg<H>::THREADS, 1)
sdf_tc_kernel(SdfNetDev net, TcParams prm_tc, const __half* __restrict__ bimg, const float* __restrict__ x,
              const float* __restrict__ y, const float* __restrict__ sbar, size_t n, float* __restrict__ s_out,
              float* __restrict__ jx_out, float* __restrict__ jy_out, float* __restrict__ hxx_out, float* __restrict__ hxy_out,
              float* __restrict__ hyy_out, unsigned int* __restrict__ tile_ctr, long long* __restrict__ dbg) {
  using Cfg = TcCfg<H>;
  constexpr bool RH = (ACT == NLO_ACT_RELU);
  constexpr int NIMG = RH ? 4 : 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __half* sB = reinterpret_cast<__half*>(smem_raw);                  // W1 hi | W1 lo | (V hi | V lo)
  float* sPart = reinterpret_cast<float*>(smem_raw + (size_t)NIMG * Cfg::IMG_HALFS * 2);   // [NGROUPS][2][TILE][4]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sPart + Cfg::NGROUPS * 2 * TILE * 4);       // [NGROUPS]
  int* lock = reinterpret_cast<int*>(mbar + Cfg::NGROUPS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lock + 1);
  uint32_t* next_tile = tmem_slot + 1;                                                      // [NGROUPS]
  float* sCoord = reinterpret_cast<float*>(next_tile + Cfg::NGROUPS);                       // [NGROUPS][2][3][TILE]: x | y | seed of the next tile
  const int t = threadIdx.x, warp = t >> 5;
  const int grp = t / Cfg::GROUP_THREADS, tg = t % Cfg::GROUP_THREADS;
  const int half = (tg >> 5) >> 2;
  const int pt = tg & (TILE - 1);

  // ---- one-time setup: operand images to smem, barriers, lock, tensor memory ---------------------------------
  {
    const uint4* src = reinterpret_cast<const uint4*>(bimg);
    uint4* dst = reinterpret_cast<uint4*>(sB);
    for (int i = t; i < NIMG * Cfg::IMG_HALFS / 8; i += Cfg::THREADS) dst[i] = src[i];
    if (t == 0) {
      for (int g = 0; g < Cfg::NGROUPS; ++g) mbar_init(mbar + g, 1);
      *lock = 0;
    }
    fence_async_smem();                          // generic-proxy smem writes -> visible to the tensor-core (async) proxy
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  TileCtx c;
  const uint32_t tmem_all = *tmem_slot;
  c.tmem_base = tmem_all + (uint32_t)grp * Cfg::TMEM_COLS;
  c.lane_base = c.tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  c.sB_hi = smem_u32(sB); c.sB_lo = smem_u32(sB + Cfg::IMG_HALFS);
  c.sV_hi = smem_u32(sB + 2 * Cfg::IMG_HALFS); c.sV_lo = smem_u32(sB + 3 * Cfg::IMG_HALFS);
  c.mbar_addr = smem_u32(mbar + grp);
  c.bar_id = 1 + grp;
  c.lock = lock;
  c.ctr = tile_ctr; c.next_slot = next_tile + grp; c.ctr_bias = 2 * gridDim.x * Cfg::NGROUPS;
  c.act0 = ACT0 >= 0 ? ACT0 : net.act0; c.act = ACT >= 0 ? ACT : net.act;
  c.prm0 = net.p0; c.prm = net.p; c.inv_sw = prm_tc.inv_sw;
  c.unscale2 = RH ? prm_tc.inv_sv : prm_tc.inv_sc1 * prm_tc.inv_sw;
  c.max_w0x = prm_tc.max_w0x; c.max_w0y = prm_tc.max_w0y; c.max_b0 = prm_tc.max_b0;
  const bool want_jac = (jx_out != nullptr) || (jy_out != nullptr) || HESS;
  constexpr bool FULL = (ACT0 >= 0 && ACT >= 0);   // compile-time activations: unrolled, vectors as constant-bank operands
  float* part = sPart + grp * (2 * TILE * 4);
  c.hxy_off = HESS ? hxy_out - hxx_out : 0; c.hyy_off = HESS ? hyy_out - hxx_out : 0;
  uint32_t phase = 0;
  if (dbg && t == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); dbg[2 * 64 * 32 + blockIdx.x * 4 + 0] = (long long)gt; dbg[2 * 64 * 32 + blockIdx.x * 4 + 2] = clock64(); }
  const size_t n_tiles = (n + TILE - 1) / TILE;
  int it = 0;
  // tiles: the first one by position, the rest from the global counter (SMs do not all run at the same speed; static
  // striding left the slowest 25 % behind) - or by striding when no counter is given (small launches)
  // The counter runs two tiles ahead: the index of the tile after this one is known when this one starts, so its coordinates
  // can be pulled towards the SM while this tile computes (the loads at the top of a tile are otherwise an exposed DRAM latency).
  const size_t stride = (size_t)gridDim.x * Cfg::NGROUPS;
  size_t tile_next = (size_t)blockIdx.x * Cfg::NGROUPS + grp + stride;
  // The coordinates of the group's NEXT tile travel global -> shared memory with cp.async while this tile computes (its index is known
  // a tile ahead, see above), double-buffered per group; the tile body waits for the copies just before its final barrier.  Loading them
  // at the top of their own tile left 400 - 800 cycles of L2 / DRAM latency exposed in front of every tile's layer 0 (a prefetch.global.L1
  // a tile ahead did not survive in the 28 KB of L1 that the operand images leave).
  float* coord = sCoord + grp * (2 * 3 * TILE);
  for (size_t tile = (size_t)blockIdx.x * Cfg::NGROUPS + grp; tile < n_tiles; ++it) {
    if (tile_next < n_tiles && half == 0) {
      const size_t in = min(tile_next * TILE + pt, n - 1);
      float* dst = coord + ((it + 1) & 1) * (3 * TILE) + pt;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(dst)), "l"(x + in) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(dst + TILE)), "l"(y + in) : "memory");
      if (sbar) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(dst + 2 * TILE)), "l"(sbar + in) : "memory");
    }
    // timeline: groups 0 and 1 of CTA 0, first 64 tiles
    c.dbg = (dbg && (it & 15) == 0 && (it >> 4) < 64 && blockIdx.x == 0 && grp < 2) ? dbg + (grp * 64 + (it >> 4)) * 32 : nullptr;
    if (c.dbg && tg == 0) c.dbg[26] = clock64();
    const size_t i = tile * TILE + pt;
    const bool valid = i < n;
    float px, py, seed;
    if (it == 0) {
      const size_t ic = valid ? i : n - 1;
      px = x[ic]; py = y[ic]; seed = sbar ? sbar[ic] : 1.f;
    } else {
      const float* src = coord + (it & 1) * (3 * TILE) + pt;
      px = src[0]; py = src[TILE]; seed = sbar ? src[2 * TILE] : 1.f;
    }
    if (Cfg::SPLIT == 1) {
      // jx and jy are both written by the one owner of the point (both non-null whenever a Jacobian is requested through this path)
      phase = sdf_tc_tile<H, ACT0, ACT, FULL, 1, 0, HESS>(c, phase, px, py, seed, want_jac, part, (valid && s_out) ? s_out + i : nullptr,
                                                          (valid && jx_out) ? jx_out + i : nullptr, jy_out - jx_out,
                                                          (HESS && valid) ? hxx_out + i : nullptr);
    } else if (half == 0) {
      phase = sdf_tc_tile<H, ACT0, ACT, FULL, Cfg::SPLIT, 0, HESS>(c, phase, px, py, seed, want_jac, part,
                                                                   (valid && s_out) ? s_out + i : nullptr,
                                                                   (valid && jx_out) ? jx_out + i : nullptr, 0,
                                                                   (HESS && valid) ? hxx_out + i : nullptr);
    } else {
      phase = sdf_tc_tile<H, ACT0, ACT, FULL, Cfg::SPLIT, Cfg::SPLIT - 1, HESS>(c, phase, px, py, seed, want_jac, part, nullptr,
                                                                                (valid && jy_out) ? jy_out + i : nullptr, 0,
                                                                                (HESS && valid) ? hxx_out + i : nullptr);
    }
    tile = tile_next;
    tile_next = tile_ctr ? (size_t)next_tile[grp] : tile_next + stride;
  }
  tc_fence_before();
  __syncthreads();
  // The last CTA to finish puts the tile counter back to zero for the next launch on this stream (tile_ctr[1] counts finished
  // CTAs): no memset per launch.  Every group's atomicAdd on tile_ctr[0] has returned before the barrier above.
  if (tile_ctr && t == 0) {
    __threadfence();
    if (atomicAdd(tile_ctr + 1, 1u) == gridDim.x - 1) { tile_ctr[0] = 0u; tile_ctr[1] = 0u; __threadfence(); }
  }
  if (dbg && t == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); dbg[2 * 64 * 32 + blockIdx.x * 4 + 1] = (long long)gt; dbg[2 * 64 * 32 + blockIdx.x * 4 + 3] = clock64(); }
  if (warp == 0) tmem_dealloc(tmem_all, 512);
}

unsigned long long g_const_owner[64] = {0};   // per device: uid of the model whose vectors sit in `cst`
std::mutex g_const_mu;                        // check owner -> (sync + upload) -> launch is one critical section per process

template <int H, int ACT0, int ACT, bool HESS = false>
int launch_tc(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n, float* s, float* jx, float* jy, cudaStream_t st,
              float* hxx = nullptr, float* hxy = nullptr, float* hyy = nullptr) {
  using Cfg = TcCfg<H>;
  auto kfn = sdf_tc_kernel<H, ACT0, ACT, HESS>;
  const size_t smem = Cfg::bytes(ACT == NLO_ACT_RELU);
  static bool attr_set[64] = {false};                    // per device; the call costs microseconds on the single-problem path
  if (!attr_set[m->device & 63]) {
    NLO_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[m->device & 63] = true;
  }
  const size_t tiles = (n + TILE - 1) / TILE;
  const size_t want = (tiles + Cfg::NGROUPS - 1) / Cfg::NGROUPS;
  const int grid = (int)(want < (size_t)m->sm_count ? want : (size_t)m->sm_count);   // one persistent CTA per SM
  TcParams prm;
  memcpy(&prm, m->tc_params, sizeof(prm));
  std::lock_guard<std::mutex> lk(g_const_mu);   // another host thread must not swap the constants between this check and the launch
  if (g_const_owner[m->device] != m->uid) {
    // A different model's vectors sit in constant memory: drain whatever may still read them, then upload ours
    // synchronously so that launches on any stream see them.  Rare (model switch), so the device-wide sync is fine.
    NLO_CUDA(cudaDeviceSynchronize());
    NLO_CUDA(cudaMemcpyToSymbol(cst, m->tc_const, sizeof(TcConst), 0, cudaMemcpyHostToDevice));
    g_const_owner[m->device] = m->uid;
  }
  long long* dbg = nullptr;
  if (getenv("NLO_B200_TC_TIMELINE")) {             // debugging aid: dump phase clocks of two groups after the launch
    NLO_CUDA(cudaMalloc(&dbg, (2 * 64 * 32 + 4 * 256) * sizeof(long long)));
    NLO_CUDA(cudaMemsetAsync(dbg, 0, (2 * 64 * 32 + 4 * 256) * sizeof(long long), st));
  }
  unsigned int* ctr = nullptr;
  if (tiles > (size_t)grid * Cfg::NGROUPS) {         // more than one tile per group: balance dynamically
    // the counter pair of the launching stream: launches on one stream are ordered, and the kernel leaves its pair at zero
    const int slot = nlo_model_stream_slot(m, st);
    if (slot < 0) return 1;
    ctr = reinterpret_cast<unsigned int*>(static_cast<char*>(m->d_tc) + m->tc_bytes) + 2 * slot;
  }
  kfn<<<grid, Cfg::THREADS, smem, st>>>(m->net(), prm, reinterpret_cast<const __half*>(m->d_tc), x, y, sbar, n, s, jx, jy, hxx, hxy, hyy, ctr, dbg);
  if (dbg) {
    std::vector<long long> h(2 * 64 * 32 + 4 * 256);
    NLO_CUDA(cudaMemcpyAsync(h.data(), dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
    NLO_CUDA(cudaStreamSynchronize(st));
    cudaFree(dbg);
    {
      const long long* q = h.data() + 2 * 64 * 32;
      long long t0 = q[0], t1 = q[1], dmin = 1LL << 60, dmax = 0; double mhz_min = 1e9, mhz_max = 0;
      for (int b = 0; b < grid && b < 256; ++b) {
        t0 = q[4 * b] < t0 ? q[4 * b] : t0; t1 = q[4 * b + 1] > t1 ? q[4 * b + 1] : t1;
        const long long d = q[4 * b + 1] - q[4 * b]; dmin = d < dmin ? d : dmin; dmax = d > dmax ? d : dmax;
        const double mhz = 1e3 * (double)(q[4 * b + 3] - q[4 * b + 2]) / (double)d; mhz_min = mhz < mhz_min ? mhz : mhz_min; mhz_max = mhz > mhz_max ? mhz : mhz_max;
      }
      fprintf(stderr, "[tc timeline] kernel %lld ns first-start to last-end; per-CTA duration %lld..%lld ns; per-CTA clock %.0f..%.0f MHz; CTA0 %lld ns\n",
              t1 - t0, dmin, dmax, mhz_min, mhz_max, q[1] - q[0]);
    }
    for (int g = 0; g < 2; ++g)
      for (int it = 8; it < 14; ++it) {
        const long long* r = h.data() + (g * 64 + it) * 32;
        if (it == 8) {
          for (int q = 0; q + 8 < 64; q += 8) {
            const long long* r0 = h.data() + (g * 64 + q) * 32; const long long* r1 = h.data() + (g * 64 + q + 8) * 32;
            if (r1[0] == 0) break;
            fprintf(stderr, "[tc timeline] group%d tiles %4d..%4d: %lld cycles in %lld ns -> %.0f MHz, %.0f cycles/tile\n", g, q * 16, (q + 8) * 16,
                    r1[0] - r0[0], r1[12] - r0[12], 1e3 * (double)(r1[0] - r0[0]) / (double)(r1[12] - r0[12]), (double)(r1[0] - r0[0]) / 128.0);
          }
        }
        fprintf(stderr, "[tc timeline] group%d tile%2d half0: L0 %5lld bar %5lld lock %5lld issue %5lld mma1 %5lld E1 %5lld bar %5lld lock %5lld issue %5lld mma2 %5lld E2 %5lld | start-to-start %6lld | half1 L0 %5lld E1 %5lld E2 %5lld\n",
                g, it, r[1] - r[0], r[8] - r[1], r[10] - r[8], r[2] - r[10], r[3] - r[2], r[4] - r[3], r[9] - r[4], r[11] - r[9], r[5] - r[11], r[6] - r[5], r[7] - r[6],
                0LL, r[17] - r[16], r[20] - r[19], r[23] - r[22]);
        fprintf(stderr, "[tc timeline]   loop top -> tile body %5lld | E2 end -> tile end (exchange, barrier, stores) %5lld | whole tile incl. loads %6lld\n",
                r[0] - r[26], r[27] - r[7], r[27] - r[26]);
        fprintf(stderr, "[tc timeline]   GEMM 1 issue: lock -> first MMA %5lld, pass 0 %5lld, pass 1 %5lld, pass 2 %5lld, -> returned %5lld | GEMM 2: lock -> first %5lld, MMAs %5lld, -> returned %5lld\n",
                r[13] - r[10], r[14] - r[13], r[15] - r[14], r[29] - r[15], r[2] - r[29], r[30] - r[11], r[31] - r[30], r[5] - r[31]);
      }
  }
  NLO_CHECK_LAUNCH();
  return 0;
}


// =====================================================================================================================
// ReLU first layer + ReLU hidden layer, H = 128, value + Jacobian (the benchmark shape, scripts/run_benchmark.py:65):
// LAYER 0 ON THE TENSOR CORE.  The tile bodies above are bound by their SIMT instructions (ncu: issue slots 61 %, ALU pipe 40 %,
// tensor pipe 51 %), and a third of those instructions evaluate a0 = W0 p + b0 - twice: in layer 0 and again in epilogue 2 for
// the sign of a0.  Here a0 is one more MMA per tile:
//   D0[128 points x 128 neurons] = P[128 x 16] . B0[16 x 128]
// where a row of P holds the point in fp16 pieces (px = h + m + l, py likewise, and a power of two for the bias) and B0 the
// matching pieces of W0 and b0 (two pieces of W0, three of b0), arranged so that the 13 used K slots are exactly the products that
// matter: h.H, h.L, m.H, m.L, l.H per coordinate and c.b0{H,M,L}.  Every product is exact in the FP32 accumulator; the sum carries a0
// to 2^-22 relative (the two-piece W0), like the GEMMs behind it.  The row scale of the A operand of GEMM 1 is folded into P
// (P = p . sc0 / s_W0), so D0 = sc0 . a0 arrives ready to split:
//   layer 0   tcgen05.ld D0;  hi = truncate(D0) to fp16's 11 bits (one LOP3), lo = D0 - hi >= 0 (one FADD); ReLU is the .relu of the
//             two packed conversions (a negative D0 has a negative hi AND a negative remainder)          3 instructions / neuron (was 9)
//   epilogue 1  as above; the 0/1 mask goes to the A-lo columns (dead after GEMM 1), so that h0's hi image SURVIVES in the A-hi columns
//   epilogue 2  phi0'(a0) = [hi(h0) != 0] read back from the A-hi columns: a bit test instead of re-evaluating a0
//                                                                                                          3.5 instructions / neuron (was 6)
// The layer-0 MMA of the group's NEXT tile is issued from epilogue 2, as soon as every thread holds its G0 columns (the D columns
// are free from then on, and the 8 A columns it reads sit in the dead A-lo range): its latency hides behind epilogue 2 and the
// loop top, and the tile body starts by waiting on its mbarrier.  41 MMAs per tile instead of 40.
// =====================================================================================================================
constexpr int RR_H = 128;
constexpr int RR_B0_HALFS = 16 * RR_H;      // layer-0 operand image: 16 K slots x 128 neurons, K-major core-matrix order

// What a tile body needs: the hot words by value (registers), the launch-wide pointers behind one shared-memory pointer.
struct RrGlobal {
  const float* x; const float* y; const float* sbar;
  float* s_out; float* jx_out; float* jy_out;
  unsigned int* ctr;
  long long* dbg;
  uint32_t n, n_tiles, ctr_bias, pad;
  long long ret_clock[2];  // timeline only: when thread 0 of each group last returned from a tile body
  // rows mode (K3 fused, nlo_sdf_tc_rows_launch): the points are the footprint points of an NLP batch, formed in the kernel from the
  // poses in w, and the results go straight to the SDF rows of g and their Jacobian entries (core/geometry.py:78-83, 107-117)
  const float* w; float* g_rows; float* jac; const int* nz;
  size_t ld;
  uint32_t P, nx, nb, nz_per_knot;
  float bx[4], by[4];
};
struct RrCtx {
  uint32_t tmem_base, lane_base, sB_hi, sB_lo, sV_hi, sV_lo, sB0, mbar_addr, mbar2_addr, bar_id;
  int* lock;
  const RrGlobal* g;       // shared memory
  uint32_t* next_slot;     // the group's next-but-one tile index
  float* coord;            // the group's coordinate staging: [2][RR_COORD_SLOTS][TILE]  (x | y | seed | rows mode: (k, b) | dpx | dpy)
  float* part;             // the group's exchange buffer [2][TILE][4]
  float* rows;             // rows mode: the group's row bookkeeping [3][3][TILE]
  float inv_sw, unscale2, max_w0x, max_w0y, max_b0, spx_mul, spy_mul, cb_mul;
  uint32_t n_tiles, flags;   // copies of what every tile asks (flags: 1 = dynamic tile counter, 2 = timeline, 4 = adjoint seed given)
};

__device__ __forceinline__ float rr_bound(const RrCtx& c, float px, float py) {
  return fmaf(fabsf(px), c.max_w0x, fmaf(fabsf(py), c.max_w0y, c.max_b0)) + 1e-30f;
}
// v = h + m + l with every part an fp16 value (33 significant bits together; the parts are returned as floats)
__device__ __forceinline__ void rr_split3(float v, float& h, float& m, float& l) {
  h = __half2float(__float2half_rn(v));
  const float r = v - h;
  m = __half2float(__float2half_rn(r));
  l = __half2float(__float2half_rn(r - m));
}
// One row of the layer-0 A operand (16 fp16 values = 8 tensor-memory columns):
//   slot  0   1   2   3   4   5   6   7   8   9   10  11  12  13..15
//   A     xh  xh  xm  xm  xl  yh  yh  ym  ym  yl  c   c   c   0          x = px sc0 / s_W0x, y = py sc0 / s_W0y, c = sc0 / s_b0
//   B0    XH  XL  XH  XL  XH  YH  YL  YH  YL  YH  bH  bM  bL  0          X = W0[:,0] s_W0x = XH + XL, b = b0 s_b0 = bH + bM + bL
// |x| <= 2^14 / (max|W0x| s_W0x) <= 2^10 and c <= 4 by the choice of the scales (nlo_sdf_tc_prepare); the clamps only keep
// degenerate rows (a zero weight column, bound underflow) free of inf . 0.
__device__ __forceinline__ void rr_point_row(const RrCtx& c, float px, float py, uint32_t (&w)[8]) {
  float sc0, inv0;
  row_scale(rr_bound(c, px, py), sc0, inv0);
  const float lim = 60000.f;
  float xh, xm, xl, yh, ym, yl;
  rr_split3(fminf(fmaxf(px * (sc0 * c.spx_mul), -lim), lim), xh, xm, xl);
  rr_split3(fminf(fmaxf(py * (sc0 * c.spy_mul), -lim), lim), yh, ym, yl);
  const float cb = fminf(sc0 * c.cb_mul, 32768.f);
  w[0] = pack_f16(xh, xh); w[1] = pack_f16(xm, xm); w[2] = pack_f16(xl, yh); w[3] = pack_f16(yh, ym);
  w[4] = pack_f16(ym, yl); w[5] = pack_f16(cb, cb); w[6] = pack_f16(cb, 0.f); w[7] = 0u;
}

// The thread that issues a group's MMAs sits in the HALF-1 warps: the issuing warp is blocked for most of a GEMM (the tensor queue takes
// ~9 MMAs), and the HALF-0 warps carry the work that runs in the shadow of the GEMMs (coordinate staging, the next tile's A row)
constexpr int RR_ISSUER = TILE;
// Rows mode: point q of the batch is footprint point b of knot k of problem p, q = (k nb + b) P + p.  Its pose (x, y, heading) was
// staged into slots 0..2 of the tile's coordinate buffer; this forms the point, leaves it in slots 0, 1 (both threads of the point read
// it for the row scale) and what the deferred store needs in the tile's row-info entry: (k, b), d point / d heading.
// The row bookkeeping lives in its own three-deep rotation (a tile's entry is written while the tile before it computes and read while the
// tile after it computes, and in the value-only form no group barrier separates that read from the write two tiles later).
constexpr int RR_COORD_SLOTS = 3;
__device__ __forceinline__ void rr_rows_point(const RrGlobal* g, uint32_t q, float* slot, float* rinfo, float x, float y, float th, float& px, float& py) {
  const uint32_t r = q / g->P, k = r / g->nb, b = r - k * g->nb;
  float sn, cs;
  nlo_sincos_fast(th, sn, cs);
  const float bx = g->bx[b], by = g->by[b];
  px = x + cs * bx - sn * by;
  py = y + sn * bx + cs * by;
  slot[0] = px; slot[TILE] = py;
  rinfo[0] = __uint_as_float(k * 4u + b);
  rinfo[TILE] = -sn * bx - cs * by;
  rinfo[2 * TILE] = cs * bx - sn * by;
}
constexpr uint32_t RR_COL_AHI = 0, RR_COL_ALO = RR_H / 2, RR_COL_D = RR_H, RR_COL_P = RR_H / 2;   // P: the first 8 A-lo columns

// the layer-0 MMA of one tile: D0 = P . B0.  No lock: one MMA slipping into the other group's GEMM is harmless (other D columns).
__device__ __forceinline__ void rr_issue_l0(const RrCtx& c) {
  tc_fence_after();
  if (elect_one(1u)) {
    tc_mma_f16_ts(c.tmem_base + RR_COL_D, c.tmem_base + RR_COL_P, umma_desc(c.sB0, 16u * RR_H, 128u), umma_idesc_f16(TILE, RR_H, 0), 0u);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(c.mbar2_addr) : "memory");
  }
}
__device__ __forceinline__ void rr_issue_gemm(const RrCtx& c, bool fwd) {
  constexpr uint32_t H = RR_H, IDESC_K = umma_idesc_f16(TILE, RR_H, 0);
  tc_fence_after();
  while (atomicCAS(c.lock, 0, 1) != 0) { }
  if (elect_one(1u)) {
    if (!fwd) {
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {               // smaller term first: mask.lo, mask.hi; the mask sits in the A-lo columns
        const uint32_t b_base = (pass == 0) ? c.sV_lo : c.sV_hi;
#pragma unroll
        for (int ks = 0; ks < (int)H / 16; ++ks)
          tc_mma_f16_ts(c.tmem_base + RR_COL_D, c.tmem_base + RR_COL_ALO + ks * 8, umma_desc(b_base + ks * 32u * H, 16u * H, 128u), IDESC_K,
                        (pass | ks) != 0);
      }
    } else {
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {               // smallest terms first: lo.hi, hi.lo, hi.hi
        const uint32_t a_col = (pass == 0) ? RR_COL_ALO : RR_COL_AHI;
        const uint32_t b_base = (pass == 1) ? c.sB_lo : c.sB_hi;
#pragma unroll
        for (int ks = 0; ks < (int)H / 16; ++ks)
          tc_mma_f16_ts(c.tmem_base + RR_COL_D, c.tmem_base + a_col + ks * 8, umma_desc(b_base + ks * 32u * H, 16u * H, 128u), IDESC_K,
                        (pass | ks) != 0);
      }
    }
    *reinterpret_cast<volatile int*>(c.lock) = 0;          // right behind the last MMA (see issue_mmas)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(c.mbar_addr) : "memory");
  }
}

// One tile of one group; two threads per point (HALF 0: neurons / D columns 0..63, stores s and jx; HALF 1: 64..127, stores jy).
// `state`: bit 0 = parity of the GEMM mbarrier, bit 1 = parity of the layer-0 mbarrier, bit 2 = which half of the coordinate
// staging holds THIS tile.  Everything the persistent loop used to do around the call - staging the next tile's coordinates,
// fetching this tile's, forming the output addresses - happens in here, in the shadow of the GEMM waits: with it in the loop
// (64-bit index arithmetic, S2R, a dozen pointer arguments) 800 of a tile's 8,900 cycles went by between two calls.
// JAC = false (value only): no mask, no reverse GEMM; the next tile's layer-0 MMA goes out from epilogue 1 instead.  The value is
// computed by the same instruction sequence in both forms, so a value-only evaluation (a line-search trial) reproduces the value
// of a full one bit for bit.
struct RrTile {            // what a tile carries from its front half (layer 0, GEMM 1 issued) to its back half
  uint32_t ph, ph2, buf, tile, tile_next, next_tile_idx, prev_tile, have_prev, stride, rb;
  float inv0;
  long long* dbg;
};
#define RR_STAMP(i) do { if (t.dbg) t.dbg[i] = clock64(); } while (0)
#define RR_TILE_CONSTS                                                                                                     \
  constexpr int HH = RR_H / 2, C0 = HALF * HH, NCH = HH / 32, NT = 2 * TILE;                                               \
  constexpr uint32_t COL_AHI = RR_COL_AHI + HALF * (HH / 2), COL_ALO = RR_COL_ALO + HALF * (HH / 2), COL_D = RR_COL_D + C0; \
  const int tg = threadIdx.x % NT;                                                                                         \
  const int pt = tg & (TILE - 1);                                                                                          \
  const float* cur = c.coord + t.buf * (RR_COORD_SLOTS * TILE) + pt; \
  float* nxt = c.coord + (t.buf ^ 1u) * (RR_COORD_SLOTS * TILE) + pt; \
  (void)COL_AHI; (void)COL_ALO; (void)COL_D; (void)C0; (void)NCH; (void)cur; (void)nxt; (void)tg

// The previous tile's results: each half left its partial sums {s, jx, jy} in the exchange buffer; HALF 0 stores s and jx, HALF 1 jy.
template <int HALF, bool JAC, bool ROWS>
__device__ __forceinline__ void rr_store_prev(const RrCtx& c, const RrTile& t) {
  const int pt = threadIdx.x & (TILE - 1);
  const RrGlobal* g = c.g;
  const uint32_t i = t.prev_tile * (uint32_t)TILE + (uint32_t)pt;
  if (i < g->n) {
    const float* pa = c.part + pt * 4;
    const float* pb = c.part + (TILE + pt) * 4;
    if (ROWS) {
      // row (k, b) of problem p: g row value, d/dx, d/dy (HALF 0: value and d/dx) and d/dheading through the footprint transform (HALF 1)
      const float* rinfo = c.rows + ((t.rb + 2u) % 3u) * (3 * TILE) + pt;              // the previous tile's entry
      const uint32_t kb = __float_as_uint(rinfo[0]), k = kb >> 2, b = kb & 3u;
      const uint32_t r = k * g->nb + b, p = i - r * g->P;
      const size_t ld = g->ld;
      const int* nz = g->nz + k * g->nz_per_knot + 3u * b;
      if (HALF == 0) {
        if (g->g_rows) g->g_rows[(size_t)r * ld + p] = pa[0] + pb[0];
        if (JAC) g->jac[(size_t)nz[0] * ld + p] = pa[1] + pb[1];
      } else if (JAC) {
        const float gx = pa[1] + pb[1], gy = pb[2] + pa[2];
        g->jac[(size_t)nz[1] * ld + p] = gy;
        g->jac[(size_t)nz[2] * ld + p] = gx * rinfo[TILE] + gy * rinfo[2 * TILE];
      }
    } else if (HALF == 0) {
      if (g->s_out) g->s_out[i] = pa[0] + pb[0];
      if (JAC) g->jx_out[i] = pa[1] + pb[1];
    } else {
      if (JAC) g->jy_out[i] = pb[2] + pa[2];
    }
  }
}

// front half of a tile: layer 0 from the tensor core's D0, GEMM 1 issued, the next tile's coordinates on their way
template <int HALF, bool JAC, bool ROWS>
__device__ __forceinline__ void rr_front(const RrCtx& c, RrTile& t) {
  RR_TILE_CONSTS;
  t.dbg = (tg == RR_ISSUER && (c.flags & 2u) && blockIdx.x == 0 && t.tile >= 32u * 2u * gridDim.x && t.tile < 40u * 2u * gridDim.x)
              ? c.g->dbg + ((threadIdx.x / NT) * 8 + (t.tile / (2u * gridDim.x) - 32u)) * 16 : nullptr;
  RR_STAMP(0);
  if (t.dbg) t.dbg[14] = c.g->ret_clock[threadIdx.x / NT];
  float sc0;
  row_scale(rr_bound(c, cur[0], cur[TILE]), sc0, t.inv0);
  // ---- layer 0: D0 = sc0 . a0 from the tensor core -> ReLU, fp16 hi / lo -> A operand ----------------------------------
  mbar_wait_addr(c.mbar2_addr, t.ph2); t.ph2 ^= 1u;
  tc_fence_after();
  RR_STAMP(1);
  {
    uint32_t d[NCH][32];
#pragma unroll
    for (int cc = 0; cc < NCH; ++cc) tmem_ld32(c.lane_base + COL_D + cc * 32, d[cc]);
    tc_wait_ld();
#pragma unroll
    for (int cc = 0; cc < NCH; ++cc) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float v0 = __uint_as_float(d[cc][2 * q]), v1 = __uint_as_float(d[cc][2 * q + 1]);
        // hi = fp16 of max(v, 0) rounded TOWARD ZERO (so that the remainder of a positive v is >= 0), unpacked again on the FMA pipe;
        // the ALU pipe - the busiest of this kernel - sees two packed conversions per pair instead of two LOP3 and two conversions
        hi[q] = pack_relu_rz_f16(v0, v1);
        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi[q]));
        lo[q] = pack_relu_f16(v0 - hf.x, v1 - hf.y);
      }
      TmemIO<16>::st(c.lane_base + COL_AHI + cc * 16, hi);
      TmemIO<16>::st(c.lane_base + COL_ALO + cc * 16, lo);
    }
  }
  RR_STAMP(2);
  tc_wait_st();
  tc_fence_before();
  group_bar<NT>(c.bar_id);
  RR_STAMP(3);
  t.next_tile_idx = 0u;
  if (tg == RR_ISSUER) {
    rr_issue_gemm(c, true);
    // the tile after the next one: requested here, stored at the end of the tile (the round trip to the one hot counter is long)
    if (c.flags & 1u) t.next_tile_idx = atomicAdd(c.g->ctr, 1u);
  }
  RR_STAMP(4);
  // ---- in the shadow of GEMM 1 (everything below used to sit between two tiles, behind a group barrier of its own) ----------
  // (the barrier in front of GEMM 1 is what orders these shared-memory reads behind the previous tile's writes)
  if (t.have_prev) {
    // which tile follows the next one: thread 0 stored it at the end of the previous tile
    t.tile_next = (c.flags & 1u) ? *c.next_slot : t.tile + t.stride;
    rr_store_prev<HALF, JAC, ROWS>(c, t);                     // the previous tile's results: both halves' partial sums -> global memory
  }
  const bool has_next = t.tile_next < c.n_tiles;
  // the next tile's coordinates start their way global -> shared memory (consumed in epilogue 2 and by the next tile)
  if (HALF == 0 && has_next) {
    const RrGlobal* g = c.g;
    const uint32_t in = min(t.tile_next * (uint32_t)TILE + (uint32_t)pt, g->n - 1u);
    if (ROWS) {                        // the pose of the point's knot: x, y, heading
      const uint32_t r = in / g->P, k = r / g->nb;
      const float* src = g->w + (size_t)(k * g->nx) * g->ld + (in - r * g->P);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(nxt)), "l"(src) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(nxt + TILE)), "l"(src + g->ld) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(nxt + 2 * TILE)), "l"(src + 2 * g->ld) : "memory");
    } else {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(nxt)), "l"(g->x + in) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(nxt + TILE)), "l"(g->y + in) : "memory");
      if (g->sbar) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(nxt + 2 * TILE)), "l"(g->sbar + in) : "memory");
    }
  }
}

// back half of a tile: GEMM 1 done -> epilogue 1 -> (GEMM 2 -> epilogue 2) -> stores.  The loop's backward branch sits between the
// front half and this wait: a far branch costs ~700 cycles of instruction fetch (measured with the timeline stamps - whether the tile body
// was a function called per tile or a loop), and here that latency runs under GEMM 1 instead of between two tiles.
template <int HALF, bool JAC, bool ROWS>
__device__ __forceinline__ void rr_back(const RrCtx& c, RrTile& t) {
  RR_TILE_CONSTS;
  const bool has_next = t.tile_next < c.n_tiles;
  mbar_wait_addr(c.mbar_addr, t.ph); t.ph ^= 1u;
  tc_fence_after();
  RR_STAMP(5);
  // ---- epilogue 1: value; the 0 / 1 mask of the hidden layer -> A-lo columns --------------------------------------------
  float s = HALF == 0 ? cst.bout : 0.f;
  uint32_t wrow[8];
  {
    const float unscale1 = t.inv0 * c.inv_sw;
    uint32_t z[NCH][32];
#pragma unroll
    for (int cc = 0; cc < NCH; ++cc) tmem_ld32(c.lane_base + COL_D + cc * 32, z[cc]);
    if (!JAC) {                       // D is free once every thread holds its columns: the next tile's layer-0 MMA
      if (HALF == 0 && has_next) {
        asm volatile("cp.async.wait_all;" ::: "memory");     // this thread staged the point itself
        if (ROWS) {
          float qx, qy;
          rr_rows_point(c.g, min(t.tile_next * (uint32_t)TILE + (uint32_t)pt, c.g->n - 1u), nxt, c.rows + ((t.rb + 1u) % 3u) * (3 * TILE) + pt, nxt[0], nxt[TILE],
                        nxt[2 * TILE], qx, qy);
          rr_point_row(c, qx, qy, wrow);
        } else {
          rr_point_row(c, nxt[0], nxt[TILE], wrow);
        }
        TmemIO<8>::st(c.lane_base + RR_COL_P, wrow);
        tc_wait_st();
      }
      tc_wait_ld();
      tc_fence_before();
      group_bar<NT>(c.bar_id);
      if (tg == RR_ISSUER && has_next) rr_issue_l0(c);
    } else {
      tc_wait_ld();
    }
#pragma unroll
    for (int cc = 0; cc < NCH; ++cc) {
      uint32_t mk[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        uint32_t mask = 0u;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = C0 + cc * 32 + 2 * q + e;
          const float zz = fmaf(__uint_as_float(z[cc][2 * q + e]), unscale1, cst.b1[j]);
          if (zz > 0.f) { s = fmaf(cst.w2[j], zz, s); mask |= (e == 0 ? 0x3C00u : 0x3C000000u); }
        }
        mk[q] = mask;
      }
      if (JAC) TmemIO<16>::st(c.lane_base + COL_ALO + cc * 16, mk);
    }
  }
  float jx = 0.f, jy = 0.f;
  if (JAC) {
    RR_STAMP(6);
    tc_wait_st();
    tc_fence_before();
    group_bar<NT>(c.bar_id);
    RR_STAMP(7);
    if (tg == RR_ISSUER) rr_issue_gemm(c, false);
    RR_STAMP(8);
    // while GEMM 2 runs: the next tile's layer-0 A row (its coordinates were staged during GEMM 1)
    if (HALF == 0 && has_next) {
      asm volatile("cp.async.wait_all;" ::: "memory");       // this thread staged the point itself
      if (ROWS) {
        float qx, qy;
        rr_rows_point(c.g, min(t.tile_next * (uint32_t)TILE + (uint32_t)pt, c.g->n - 1u), nxt, c.rows + ((t.rb + 1u) % 3u) * (3 * TILE) + pt, nxt[0], nxt[TILE],
                        nxt[2 * TILE], qx, qy);
        rr_point_row(c, qx, qy, wrow);
      } else {
        rr_point_row(c, nxt[0], nxt[TILE], wrow);
      }
    }
    mbar_wait_addr(c.mbar_addr, t.ph); t.ph ^= 1u;
    tc_fence_after();
    RR_STAMP(9);
    // ---- epilogue 2: through layer 0 to the Jacobian; the next tile's layer-0 MMA goes out as soon as D is free ---------
    uint32_t gz[NCH][32];
#pragma unroll
    for (int cc = 0; cc < NCH; ++cc) tmem_ld32(c.lane_base + COL_D + cc * 32, gz[cc]);
    if (HALF == 0 && has_next) {
      TmemIO<8>::st(c.lane_base + RR_COL_P, wrow);           // the A-lo columns are dead once GEMM 2 has completed
      tc_wait_st();
    }
    tc_wait_ld();
    tc_fence_before();
    group_bar<NT>(c.bar_id);
    RR_STAMP(10);
    if (tg == RR_ISSUER && has_next) rr_issue_l0(c);
    RR_STAMP(11);
#pragma unroll
    for (int cc = 0; cc < NCH; ++cc) {
      uint32_t hw[16];
      TmemIO<16>::ld(c.lane_base + COL_AHI + cc * 16, hw);   // hi(h0) of THIS tile: nonzero exactly where a0 > 0
      tc_wait_ld();
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        // select, then two unconditional FFMAs: under a per-thread predicate the constants would arrive one LDC each
        // instead of four per LDCU.128.  One packed half compare yields both predicates of a pair (hi(h0) >= +0).
        float gsel[2];
        asm("{\n\t.reg .pred p, q;\n\tsetp.gt.f16x2 p|q, %2, %3;\n\tselp.f32 %0, %4, 0f00000000, p;\n\tselp.f32 %1, %5, 0f00000000, q;\n\t}"
            : "=f"(gsel[0]), "=f"(gsel[1]) : "r"(hw[q]), "r"(0u), "f"(__uint_as_float(gz[cc][2 * q])), "f"(__uint_as_float(gz[cc][2 * q + 1])));
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int k = C0 + cc * 32 + 2 * q + e;
          jx = fmaf(gsel[e], cst.w0x[k], jx);
          jy = fmaf(gsel[e], cst.w0y[k], jy);
        }
      }
    }
    const float unscale2 = ((!ROWS && (c.flags & 4u)) ? cur[2 * TILE] : 1.f) * c.unscale2;
    jx *= unscale2; jy *= unscale2;
  }
  RR_STAMP(12);
  // ---- leave the partial sums for the deferred store (rr_store_prev, in the shadow of the next tile's GEMM 1).  No barrier here: a warp
  // ---- touches only its own tensor-memory lanes and columns between MMAs, and every MMA is issued behind a group barrier ------------
  {
    float* mine = c.part + (HALF * TILE + pt) * 4;
    mine[0] = s; mine[1] = jx; mine[2] = jy;
  }
  if (tg == RR_ISSUER && (c.flags & 1u)) *c.next_slot = t.next_tile_idx + c.g->ctr_bias;
  t.prev_tile = t.tile; t.have_prev = 1u;
  RR_STAMP(13);
  t.buf ^= 1u;
  t.rb = (t.rb + 1u) % 3u;
}
#undef RR_STAMP
#undef RR_TILE_CONSTS

// All tiles of one group half.  The loop is rotated: its backward branch follows the issue of GEMM 1 (see rr_back).
template <int HALF, bool JAC, bool ROWS>
__device__ __noinline__ void sdf_rr_group(RrCtx c, uint32_t tile, uint32_t tile_next, uint32_t stride) {
  RrTile t;
  t.ph = t.ph2 = t.buf = 0u; t.tile = tile; t.tile_next = tile_next; t.next_tile_idx = 0u; t.inv0 = 1.f; t.dbg = nullptr;
  t.prev_tile = 0u; t.have_prev = 0u; t.stride = stride; t.rb = 0u;
  if (t.tile >= c.n_tiles) return;
  rr_front<HALF, JAC, ROWS>(c, t);
#pragma unroll 1
  for (;;) {
    rr_back<HALF, JAC, ROWS>(c, t);
    if ((c.flags & 2u) && threadIdx.x % (2 * TILE) == RR_ISSUER) const_cast<RrGlobal*>(c.g)->ret_clock[threadIdx.x / (2 * TILE)] = clock64();
    t.tile = t.tile_next;                 // (its successor is read in the front half, behind the barrier that orders it)
    if (t.tile >= c.n_tiles) break;
    rr_front<HALF, JAC, ROWS>(c, t);
  }
  group_bar<2 * TILE>(c.bar_id);
  rr_store_prev<HALF, JAC, ROWS>(c, t);
}

constexpr size_t rr_smem_bytes() {
  return (size_t)4 * RR_H * RR_H * 2 + (size_t)RR_B0_HALFS * 2 + (size_t)2 * 2 * TILE * 4 * 4 + 4 * 8 + sizeof(RrGlobal) + 32 +
         (size_t)2 * 2 * RR_COORD_SLOTS * TILE * 4 + (size_t)2 * 3 * 3 * TILE * 4;
}

template <bool JAC, bool ROWS>
__global__ void __launch_bounds__(512, 1)
sdf_tc_rr_kernel(TcParams prm_tc, const __half* __restrict__ bimg, RrGlobal gl) {
  constexpr int H = RR_H, NG = 2, GT = 2 * TILE, IMG = H * H;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __half* sB = reinterpret_cast<__half*>(smem_raw);                       // W1 hi | W1 lo | V hi | V lo | B0
  float* sPart = reinterpret_cast<float*>(smem_raw + ((size_t)4 * IMG + RR_B0_HALFS) * 2);   // [NG][2][TILE][4]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sPart + NG * 2 * TILE * 4);                  // [NG] GEMMs | [NG] layer 0
  RrGlobal* sG = reinterpret_cast<RrGlobal*>(mbar + 2 * NG);
  int* lock = reinterpret_cast<int*>(sG + 1);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lock + 1);
  uint32_t* next_tile = tmem_slot + 1;                                                       // [NG]
  float* sCoord = reinterpret_cast<float*>(next_tile + NG);                                  // [NG][2][RR_COORD_SLOTS][TILE]
  const int t = threadIdx.x, warp = t >> 5;
  const int grp = t / GT, tg = t % GT;
  const int half = (tg >> 5) >> 2;
  const int pt = tg & (TILE - 1);
  {
    const uint4* src = reinterpret_cast<const uint4*>(bimg);
    uint4* dst = reinterpret_cast<uint4*>(sB);
    for (int i = t; i < (4 * IMG + RR_B0_HALFS) / 8; i += 512) dst[i] = src[i];
    if (t == 0) {
      for (int g = 0; g < 2 * NG; ++g) mbar_init(mbar + g, 1);
      *lock = 0;
      gl.ctr_bias = 2 * gridDim.x * NG;
      *sG = gl;
    }
    fence_async_smem();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  RrCtx c;
  const uint32_t tmem_all = *tmem_slot;
  c.tmem_base = tmem_all + (uint32_t)grp * (2 * H);
  c.lane_base = c.tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  c.sB_hi = smem_u32(sB); c.sB_lo = smem_u32(sB + IMG);
  c.sV_hi = smem_u32(sB + 2 * IMG); c.sV_lo = smem_u32(sB + 3 * IMG);
  c.sB0 = smem_u32(sB + 4 * IMG);
  c.mbar_addr = smem_u32(mbar + grp); c.mbar2_addr = smem_u32(mbar + NG + grp);
  c.bar_id = 1 + grp;
  c.lock = lock;
  c.g = sG;
  c.next_slot = next_tile + grp;
  c.coord = sCoord + grp * (2 * RR_COORD_SLOTS * TILE);
  c.rows = sCoord + 2 * (2 * RR_COORD_SLOTS * TILE) + grp * (3 * 3 * TILE);
  c.part = sPart + grp * (2 * TILE * 4);
  c.inv_sw = prm_tc.inv_sw; c.unscale2 = prm_tc.inv_sv;
  c.max_w0x = prm_tc.max_w0x; c.max_w0y = prm_tc.max_w0y; c.max_b0 = prm_tc.max_b0;
  c.spx_mul = prm_tc.spx_mul; c.spy_mul = prm_tc.spy_mul; c.cb_mul = prm_tc.cb_mul;
  c.n_tiles = gl.n_tiles; c.flags = (gl.ctr ? 1u : 0u) | (gl.dbg ? 2u : 0u) | (gl.sbar ? 4u : 0u);
  const uint32_t n_tiles = gl.n_tiles;
  const uint32_t stride = gridDim.x * NG;
  uint32_t tile = blockIdx.x * NG + grp;
  uint32_t tile_next = tile + stride;
  if (tile < n_tiles) {                       // (group-uniform) the first tile: coordinates and layer-0 MMA up front
    if (half == 0) {
      const uint32_t ic = min(tile * (uint32_t)TILE + (uint32_t)pt, gl.n - 1u);
      float px, py;
      if (ROWS) {
        const uint32_t r = ic / gl.P, k = r / gl.nb;
        const float* src = gl.w + (size_t)(k * gl.nx) * gl.ld + (ic - r * gl.P);
        rr_rows_point(sG, ic, c.coord + pt, c.rows + pt, src[0], src[gl.ld], src[2 * gl.ld], px, py);
      } else {
        px = gl.x[ic]; py = gl.y[ic];
        c.coord[pt] = px; c.coord[TILE + pt] = py; c.coord[2 * TILE + pt] = gl.sbar ? gl.sbar[ic] : 1.f;
      }
      uint32_t w[8];
      rr_point_row(c, px, py, w);
      TmemIO<8>::st(c.lane_base + RR_COL_P, w);
      tc_wait_st();
    }
    tc_fence_before();
    group_bar<GT>(c.bar_id);
    if (tg == 0) rr_issue_l0(c);
  }
  // tiles: the first two by position, the rest from the global counter (the SMs do not all run at the same speed), which the tile
  // body polls two tiles ahead - or by striding when no counter is given (small launches)
  if (half == 0) sdf_rr_group<0, JAC, ROWS>(c, tile, tile_next, stride); else sdf_rr_group<1, JAC, ROWS>(c, tile, tile_next, stride);
  tc_fence_before();
  __syncthreads();
  // The last CTA to finish puts the tile counter back to zero for the next launch on this stream
  if (gl.ctr && t == 0) {
    __threadfence();
    if (atomicAdd(gl.ctr + 1, 1u) == gridDim.x - 1) { gl.ctr[0] = 0u; gl.ctr[1] = 0u; __threadfence(); }
  }
  if (warp == 0) tmem_dealloc(tmem_all, 512);
}

// NLO_B200_TC_L0=0 keeps the SIMT layer 0 (the A / B switch of this form)
bool rr_enabled() {
  static const int on = [] { const char* e = getenv("NLO_B200_TC_L0"); return (e && e[0] == '0') ? 0 : 1; }();
  return on != 0;
}

template <bool JAC, bool ROWS>
int launch_tc_rr_gl(nlo_sdf_model* m, RrGlobal gl, size_t n, cudaStream_t st) {
  auto kfn = sdf_tc_rr_kernel<JAC, ROWS>;
  const size_t smem = rr_smem_bytes();
  static bool attr_set[64] = {false};
  if (!attr_set[m->device & 63]) {
    NLO_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[m->device & 63] = true;
  }
  const size_t tiles = (n + TILE - 1) / TILE;
  const size_t want = (tiles + 1) / 2;
  const int grid = (int)(want < (size_t)m->sm_count ? want : (size_t)m->sm_count);
  TcParams prm;
  memcpy(&prm, m->tc_params, sizeof(prm));
  std::lock_guard<std::mutex> lk(g_const_mu);
  if (g_const_owner[m->device] != m->uid) {
    NLO_CUDA(cudaDeviceSynchronize());
    NLO_CUDA(cudaMemcpyToSymbol(cst, m->tc_const, sizeof(TcConst), 0, cudaMemcpyHostToDevice));
    g_const_owner[m->device] = m->uid;
  }
  gl.ctr = nullptr; gl.dbg = nullptr; gl.n = (uint32_t)n; gl.n_tiles = (uint32_t)tiles; gl.ctr_bias = 0; gl.pad = 0; gl.ret_clock[0] = gl.ret_clock[1] = 0;
  if (tiles > (size_t)grid * 2) {
    const int slot = nlo_model_stream_slot(m, st);
    if (slot < 0) return 1;
    gl.ctr = reinterpret_cast<unsigned int*>(static_cast<char*>(m->d_tc) + m->tc_bytes) + 2 * slot;
  }
  if (getenv("NLO_B200_TC_TIMELINE")) {             // debugging aid: phase clocks of thread 0 of both groups of CTA 0, its tiles 32..39
    NLO_CUDA(cudaMalloc(&gl.dbg, 2 * 8 * 16 * sizeof(long long)));
    NLO_CUDA(cudaMemsetAsync(gl.dbg, 0, 2 * 8 * 16 * sizeof(long long), st));
  }
  kfn<<<grid, 512, smem, st>>>(prm, reinterpret_cast<const __half*>(m->d_tc), gl);
  if (gl.dbg) {
    long long h[2 * 8 * 16];
    NLO_CUDA(cudaMemcpyAsync(h, gl.dbg, sizeof(h), cudaMemcpyDeviceToHost, st));
    NLO_CUDA(cudaStreamSynchronize(st));
    cudaFree(gl.dbg);
    for (int g = 0; g < 2; ++g)
      for (int it = 0; it < 8; ++it) {
        const long long* r = h + (g * 8 + it) * 16;
        if (r[0] == 0) continue;
        fprintf(stderr, "[rr timeline] group%d slot%d: wait L0-MMA %4lld L0 %4lld bar %4lld issue1 %5lld mma1 %4lld | E1 %4lld bar %4lld issue2 %5lld mma2 %4lld | "
                        "E2: ld+row+bar %4lld issueL0 %4lld compute %4lld | end %4lld | tile body %5lld | return -> next body %5lld\n",
                g, it, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[6] - r[5], r[7] - r[6], r[8] - r[7], r[9] - r[8],
                r[10] - r[9], r[11] - r[10], r[12] - r[11], r[13] - r[12], r[13] - r[0], r[0] - r[14]);
      }
  }
  NLO_CHECK_LAUNCH();
  return 0;
}

template <bool JAC>
int launch_tc_rr(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n, float* s, float* jx, float* jy, cudaStream_t st) {
  RrGlobal gl;
  memset(&gl, 0, sizeof(gl));
  gl.x = x; gl.y = y; gl.sbar = sbar; gl.s_out = s; gl.jx_out = jx; gl.jy_out = jy;
  return launch_tc_rr_gl<JAC, false>(m, gl, n, st);
}

}  // namespace

bool nlo_sdf_tc_supported(const nlo_sdf_desc* d) {
  return (d->n_hidden_mats == 1 && (d->hidden == 64 || d->hidden == 128)) || nlo_sdf_tc256_supported(d) || nlo_sdf_tc_deep_supported(d);
}

// power-of-two scale that puts mx into [2^13, 2^14)
static float tc_pow2_scale(float mx) {
  int ex = 0;
  if (mx > 0.f) frexpf(mx, &ex);                        // mx = f * 2^ex, f in [0.5, 1)
  return ldexpf(1.f, 14 - ex);
}
// element (n, k) of a K-major operand image in UMMA core-matrix order (no swizzle), in halfs
static size_t tc_img_off(int n, int k, int H) { return ((size_t)(k / 8) * (H / 8) + n / 8) * 64 + (n % 8) * 8 + (k % 8); }

// Scale W1 (and V = diag(w2) W1) by a power of two into fp16's range, split into fp16 hi + lo, store in UMMA
// core-matrix order: images W1 hi | W1 lo | V hi | V lo.
int nlo_sdf_tc_prepare(nlo_sdf_model* m, const float* w) {
  if (m->desc.hidden == 256) return nlo_sdf_tc256_prepare(m, w);
  if (m->desc.n_hidden_mats >= 2) return nlo_sdf_tc_deep_prepare(m, w);
  const int H = (int)m->desc.hidden;
  const float* W0 = w;
  const float* b0 = w + 2 * H;
  const float* W1 = w + 3 * H;                          // blob order: W0[H][2], b0[H], W1[H][H], b1[H], w_out[H], b_out
  const float* b1 = W1 + (size_t)H * H;
  const float* w2 = b1 + H;
  float mx = 0.f, mv = 0.f, mw2 = 0.f;
  for (int j = 0; j < H; ++j) {
    mw2 = fmaxf(mw2, fabsf(w2[j]));
    for (int k = 0; k < H; ++k) {
      mx = fmaxf(mx, fabsf(W1[(size_t)j * H + k]));
      mv = fmaxf(mv, (float)fabs((double)w2[j] * (double)W1[(size_t)j * H + k]));
    }
  }
  const float sw = tc_pow2_scale(mx), sv = tc_pow2_scale(mv);
  const int act = (int)m->desc.act;
  const float dbound = act == NLO_ACT_SIGMOID ? 0.25f : (act == NLO_ACT_SIN || act == NLO_ACT_COS_SCALE) ? fabsf(m->desc.p) : 1.f;
  const float sc1 = tc_pow2_scale(mw2 * dbound);
  TcParams prm;
  prm.inv_sw = 1.f / sw; prm.inv_sv = 1.f / sv; prm.inv_sc1 = 1.f / sc1;
  prm.max_w0x = prm.max_w0y = prm.max_b0 = 0.f;
  for (int k = 0; k < H; ++k) {
    prm.max_w0x = fmaxf(prm.max_w0x, fabsf(W0[2 * k])); prm.max_w0y = fmaxf(prm.max_w0y, fabsf(W0[2 * k + 1]));
    prm.max_b0 = fmaxf(prm.max_b0, fabsf(b0[k]));
  }
  // layer 0 on the tensor core (ReLU / ReLU form): W0 columns scaled into [2^4, 2^5), b0 into [2^12, 2^13) - see rr_point_row
  float mb0 = 0.f;
  for (int k = 0; k < H; ++k) mb0 = fmaxf(mb0, fabsf(b0[k]));
  const float s_w0x = tc_pow2_scale(prm.max_w0x) / 512.f, s_w0y = tc_pow2_scale(prm.max_w0y) / 512.f, s_b0 = tc_pow2_scale(mb0) / 2.f;
  prm.spx_mul = 1.f / s_w0x; prm.spy_mul = 1.f / s_w0y; prm.cb_mul = 1.f / s_b0;
  static_assert(sizeof(TcParams) <= sizeof(m->tc_params), "tc_params too small");
  memcpy(m->tc_params, &prm, sizeof(prm));
  static_assert(sizeof(TcConst) <= sizeof(m->tc_const), "tc_const too small");
  TcConst* cst = reinterpret_cast<TcConst*>(m->tc_const);
  memset(cst, 0, sizeof(TcConst));
  for (int k = 0; k < H; ++k) {
    cst->w0x[k] = W0[2 * k]; cst->w0y[k] = W0[2 * k + 1]; cst->b0[k] = b0[k]; cst->b1[k] = b1[k]; cst->w2[k] = w2[k];
    cst->w2s[k] = w2[k] * sc1;
  }
  cst->bout = w2[H];
  const size_t HH = (size_t)H * H;
  std::vector<__half> img(4 * HH + (H == RR_H ? RR_B0_HALFS : 0), __float2half_rn(0.f));
  if (H == RR_H) {
    // B0: the layer-0 operand, 16 K slots x H neurons, same core-matrix order as one k-step of the W1 image (slots: rr_point_row)
    __half* b0img = img.data() + 4 * HH;
    for (int nn = 0; nn < H; ++nn) {
      const float xv = W0[2 * nn] * s_w0x, yv = W0[2 * nn + 1] * s_w0y, bv = b0[nn] * s_b0;
      const __half xH = __float2half_rn(xv), xL = __float2half_rn(xv - __half2float(xH));
      const __half yH = __float2half_rn(yv), yL = __float2half_rn(yv - __half2float(yH));
      const __half bH = __float2half_rn(bv);
      const float br = bv - __half2float(bH);
      const __half bM = __float2half_rn(br), bL = __float2half_rn(br - __half2float(bM));
      const __half rows[16] = {xH, xL, xH, xL, xH, yH, yL, yH, yL, yH, bH, bM, bL, __float2half_rn(0.f), __float2half_rn(0.f), __float2half_rn(0.f)};
      for (int k = 0; k < 16; ++k) b0img[tc_img_off(nn, k, H)] = rows[k];
    }
  }
  for (int nn = 0; nn < H; ++nn)
    for (int k = 0; k < H; ++k) {
      // forward / generic reverse operand: B(n, k) = W1[n][k]
      const float v = W1[(size_t)nn * H + k] * sw;
      const __half hi = __float2half_rn(v);
      const size_t off = tc_img_off(nn, k, H);
      img[off] = hi;
      img[HH + off] = __float2half_rn(v - __half2float(hi));
      // ReLU reverse operand: B'(n' = nn, k' = k) = V[k][nn] = w2[k] W1[k][nn]   (n' = input neuron, k' = hidden neuron)
      const double vv = (double)w2[k] * (double)W1[(size_t)k * H + nn] * (double)sv;
      const __half vhi = __float2half_rn((float)vv);
      img[2 * HH + off] = vhi;
      img[3 * HH + off] = __float2half_rn((float)(vv - (double)__half2float(vhi)));
    }
  if (m->d_tc) cudaFree(m->d_tc);
  m->d_tc = nullptr;
  NLO_CUDA(cudaMalloc(&m->d_tc, img.size() * sizeof(__half) + 2 * NLO_STREAM_SLOTS * sizeof(unsigned int)));
  NLO_CUDA(cudaMemcpy(m->d_tc, img.data(), img.size() * sizeof(__half), cudaMemcpyHostToDevice));
  m->tc_bytes = img.size() * sizeof(__half);
  NLO_CUDA(cudaMemset(static_cast<char*>(m->d_tc) + m->tc_bytes, 0, 2 * NLO_STREAM_SLOTS * sizeof(unsigned int)));
  return 0;
}

// Fused value + Jacobian + Hessian on the tensor path: networks whose hidden layer is ReLU and whose first layer is the
// cosine feature map (the shipped FourierMLP shape).  jx/jy must both be given or both be NULL.
bool nlo_sdf_tc_hess_supported(const nlo_sdf_model* m) {
  return m->d_tc && m->desc.n_hidden_mats == 1 && (m->desc.hidden == 64 || m->desc.hidden == 128) &&
         m->desc.act0 == NLO_ACT_COS_SCALE && m->desc.act == NLO_ACT_RELU;
}
int nlo_sdf_tc_hess_launch(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n,
                           float* s, float* jx, float* jy, float* hxx, float* hxy, float* hyy, cudaStream_t st) {
  if (n == 0) return 0;
  if (!nlo_sdf_tc_hess_supported(m)) return nlo_fail("tensor-tile Hessian: unsupported network");
  if (!hxx || !hxy || !hyy) return nlo_fail("tensor-tile Hessian: all three outputs are required");
  if (m->desc.hidden == 128) return launch_tc<128, NLO_ACT_COS_SCALE, NLO_ACT_RELU, true>(m, x, y, sbar, n, s, jx, jy, st, hxx, hxy, hyy);
  return launch_tc<64, NLO_ACT_COS_SCALE, NLO_ACT_RELU, true>(m, x, y, sbar, n, s, jx, jy, st, hxx, hxy, hyy);
}

int nlo_sdf_tc_launch(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n,
                      float* s, float* jx, float* jy, cudaStream_t st) {
  if (n == 0) return 0;
  if (m->desc.hidden == 256) return nlo_sdf_tc256_launch(m, x, y, sbar, n, s, jx, jy, st);
  if (m->desc.n_hidden_mats >= 2) return nlo_sdf_tc_deep_launch(m, x, y, sbar, n, s, jx, jy, st);
  if (!m->d_tc) return nlo_fail("tensor-tile operands were not prepared");
  const int H = (int)m->desc.hidden, a0 = (int)m->desc.act0, a = (int)m->desc.act;
  if (H == RR_H && a0 == NLO_ACT_RELU && a == NLO_ACT_RELU && n <= 0x7FFFFF00u && rr_enabled()) {
    if (jx && jy) return launch_tc_rr<true>(m, x, y, sbar, n, s, jx, jy, st);
    if (!jx && !jy) return launch_tc_rr<false>(m, x, y, sbar, n, s, nullptr, nullptr, st);
  }
#define NLO_TC_ONE(HH, A0, A1) if (a0 == A0 && a == A1) return launch_tc<HH, A0, A1>(m, x, y, sbar, n, s, jx, jy, st)
  // compile-time activation pairs of the layer zoo (core/nn_architectures.py:42-100, l4casadi's naive MLP): unrolled tile bodies with
  // constant-bank operands; anything else runs the generic instantiation (run-time activation switch, ~10x slower)
#define NLO_TC(HH)                                                                                                 \
  do {                                                                                                             \
    NLO_TC_ONE(HH, NLO_ACT_RELU, NLO_ACT_RELU); NLO_TC_ONE(HH, NLO_ACT_COS_SCALE, NLO_ACT_RELU);                    \
    NLO_TC_ONE(HH, NLO_ACT_TANH, NLO_ACT_TANH); NLO_TC_ONE(HH, NLO_ACT_SIGMOID, NLO_ACT_SIGMOID);                   \
    NLO_TC_ONE(HH, NLO_ACT_LEAKY_RELU, NLO_ACT_LEAKY_RELU); NLO_TC_ONE(HH, NLO_ACT_COS_SCALE, NLO_ACT_TANH);        \
    NLO_TC_ONE(HH, NLO_ACT_COS_SCALE, NLO_ACT_SIGMOID); NLO_TC_ONE(HH, NLO_ACT_COS_SCALE, NLO_ACT_LEAKY_RELU);      \
    NLO_TC_ONE(HH, NLO_ACT_SIN, NLO_ACT_SIN);                                                                       \
    return launch_tc<HH, -1, -1>(m, x, y, sbar, n, s, jx, jy, st);                                                 \
  } while (0)
  if (H == 128) NLO_TC(128);
  if (H == 64) NLO_TC(64);
#undef NLO_TC
#undef NLO_TC_ONE
  return nlo_fail("tensor-tile path: unsupported width %d", H);
}

// K3 fused into K1 (hard SDF rows of a footprint with a heading: benchmark_6): footprint points from the poses in w, results straight into
// the rows of g (g_rows = g + g_off_sdf * ld; may be NULL) and their Jacobian entries (jac NULL: values only).  nz: the CCS slots of the
// rows' entries, nz_per_knot apart per knot, three per footprint point (d/dx, d/dy, d/dheading).
bool nlo_sdf_tc_rows_supported(const nlo_sdf_model* m) {
  static const int on = [] { const char* e = getenv("NLO_B200_FUSED_ROWS"); return (e && e[0] == '0') ? 0 : 1; }();
  return on && m->prec == NLO_PREC_TC_3XF16 && m->d_tc && m->desc.n_hidden_mats == 1 && m->desc.hidden == RR_H &&
         m->desc.act0 == NLO_ACT_RELU && m->desc.act == NLO_ACT_RELU && rr_enabled();
}
int nlo_sdf_tc_rows_launch(nlo_sdf_model* m, const float* w, size_t P, size_t ld, int n_knots, int nx, int nb, const float* bx, const float* by,
                           float* g_rows, float* jac, const int* nz, int nz_per_knot, cudaStream_t st) {
  const size_t n = (size_t)n_knots * nb * P;
  if (n == 0) return 0;
  if (n > 0x7FFFFF00u || nb > 4 || nb < 1) return nlo_fail("fused SDF rows: batch too large or unsupported footprint");
  RrGlobal gl;
  memset(&gl, 0, sizeof(gl));
  gl.w = w; gl.g_rows = g_rows; gl.jac = jac; gl.nz = nz; gl.ld = ld;
  gl.P = (uint32_t)P; gl.nx = (uint32_t)nx; gl.nb = (uint32_t)nb; gl.nz_per_knot = (uint32_t)nz_per_knot;
  for (int b = 0; b < nb; ++b) { gl.bx[b] = bx[b]; gl.by[b] = by[b]; }
  return jac ? launch_tc_rr_gl<true, true>(m, gl, n, st) : launch_tc_rr_gl<false, true>(m, gl, n, st);
}
