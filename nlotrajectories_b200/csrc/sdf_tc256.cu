// K1 (tcgen05 form, H = 256): fused learned-SDF value + Jacobian / adjoint for networks 2 -> 256 -> 256 -> 1 with a
// ReLU hidden layer (the "width 256" rows of the synthetic sweep, BASELINE.json configs[4]; same math as sdf_tc.cu and
// the same reference: _l4c_generated/nn_sdf.cpp:57-83, scripts/run_benchmark.py:64-83).
//
// At H = 256 the operand images no longer fit in shared memory (W1 hi+lo and V = diag(w2) W1 hi+lo: 512 KB), and one
// tile fills the whole tensor memory (A hi 128 | A lo 128 | D 256 columns), so the structure differs from sdf_tc.cu:
//   * the B operands stream from L2: the host lays all four images out as 64 chunks of 8 KB in exactly the order the
//     MMAs consume them (forward k-step ks: W1hi[ks], W1lo[ks]; reverse: Vlo[ks], Vhi[ks]); a dedicated producer warp
//     moves them with cp.async.bulk into a 24-slot shared-memory ring (full / empty mbarriers, tcgen05.commit frees
//     a slot), running ahead of the MMAs across GEMMs and tiles;
//   * one persistent CTA per SM, warp-specialised: 16 compute warps (four threads per point, 64 neurons each), the
//     producer warp and an MMA-issuer warp (M = 128, N = 256, K = 16).  The k-steps are consumed in an order that lets
//     half of each GEMM start as soon as every thread has written the first 32-neuron chunk of its A operand, so the
//     second half of layer 0 / epilogue 1 overlaps the tensor core; the epilogues read all 64 of a thread's D columns
//     up front, which frees D for the next GEMM (per-warp mbarrier arrivals, no CTA-wide barrier on that path);
//   * the small vectors sit in shared memory (read as broadcast LDS.128), one code path for all four neuron quarters.
// Numerics are those of sdf_tc.cu: split-fp16 hi/lo with FP32 accumulation, exact power-of-two scaling, the reverse
// GEMM as exact 0/1 mask x V (two passes).
#include "nlo_common.cuh"
#include "tc_ptx.cuh"
#include <cuda_fp16.h>
#include <vector>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <cstdio>

namespace {

constexpr int H = 256;
constexpr int TILE = 128;                 // points per tile == TMEM lanes
constexpr int NQ = 4;                     // threads per point (neuron quarters)
constexpr int NCOMPUTE = TILE * NQ;       // 512 compute threads
constexpr int NCW = NCOMPUTE / 32;          // compute warps
constexpr int THREADS = NCOMPUTE + 64;    // + producer warp + MMA warp
constexpr int CHUNK = 8192;               // one K = 16 step of a 256-row operand image
constexpr int NSLOT = 24;
constexpr int CHUNKS_FWD = 32, CHUNKS_BWD = 32;
constexpr uint32_t COL_AHI = 0, COL_ALO = 128, COL_D = 256;

struct Tc256Params { float inv_sw, inv_sv, max_w0x, max_w0y, max_b0; };

struct Smem {
  alignas(1024) uint8_t ring[NSLOT][CHUNK];
  alignas(16) float w0x[H], w0y[H], b0[H], b1[H], w2[H];
  float part[NQ][TILE][3];
  alignas(8) uint64_t full[NSLOT], empty[NSLOT], mma_bar[2], a_bar[4], dfree_bar[2];
  uint32_t tmem_slot;
};

__device__ __forceinline__ void mbar_wait256(uint32_t a, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void tc_commit_addr(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_lane0() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 1;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory"); }
// one arrival per compute warp, after every lane's tensor-memory traffic is ordered before the barrier
__device__ __forceinline__ void warp_arrive(uint32_t a, int lane) {
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(a);
}
__device__ __forceinline__ void compute_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NCOMPUTE) : "memory"); }
__device__ __forceinline__ void st16(uint32_t a, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};"
               ::"r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                 "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(a) : "memory");
}
__device__ __forceinline__ uint32_t pack2_f16(float v0, float v1) {
  const __half2 p = __floats2half2_rn(v0, v1);
  return *reinterpret_cast<const uint32_t*>(&p);
}

// k-steps in consumption order: a thread's first 32-neuron chunk (cc = 0) covers k-steps {4q, 4q+1}, its second {4q+2, 4q+3}
// (q = neuron quarter), so half of a GEMM can run as soon as every thread has written its first chunk of the A operand.
__host__ __device__ constexpr int ks_of(int i) { return ((i & 7) >> 1) * 4 + (i >> 3) * 2 + (i & 1); }   // i = 0..15

// MMA warp, elected lane: consume the ring and issue the MMAs of k-steps order[i0 .. i0+8) of one GEMM.  `c` = chunks consumed so far.
template <int CL>
__device__ __forceinline__ void issue_half256(Smem* sm, uint32_t tmem, bool fwd, int i0, uint32_t& c) {
  constexpr uint16_t MASK = (uint16_t)((1u << CL) - 1u);
  constexpr uint32_t IDESC = umma_idesc_f16(TILE, H, 0);
  const uint32_t ring = smem_u32(sm->ring), full = smem_u32(sm->full), empty = smem_u32(sm->empty);
  for (int i = i0; i < i0 + 8; ++i) {
    const int ks = ks_of(i);
    // first chunk of the k-step: forward W1hi[ks] (A lo and A hi both multiply it); reverse Vlo[ks]
    uint32_t slot = c % NSLOT;
    mbar_wait256(full + slot * 8, (c / NSLOT) & 1);
    tc_fence_after();
    uint64_t desc = umma_desc(ring + slot * CHUNK, 16u * H, 128u);
    if (fwd) {
      tc_mma_f16_ts(tmem + COL_D, tmem + COL_ALO + ks * 8, desc, IDESC, i != 0);
      tc_mma_f16_ts(tmem + COL_D, tmem + COL_AHI + ks * 8, desc, IDESC, 1);
    } else {
      tc_mma_f16_ts(tmem + COL_D, tmem + COL_AHI + ks * 8, desc, IDESC, i != 0);
    }
    if (CL > 1) tc_commit_mc(empty + slot * 8, MASK); else tc_commit_addr(empty + slot * 8);
    ++c;
    // second chunk: forward W1lo[ks]; reverse Vhi[ks]
    slot = c % NSLOT;
    mbar_wait256(full + slot * 8, (c / NSLOT) & 1);
    tc_fence_after();
    desc = umma_desc(ring + slot * CHUNK, 16u * H, 128u);
    tc_mma_f16_ts(tmem + COL_D, tmem + COL_AHI + ks * 8, desc, IDESC, 1);
    if (CL > 1) tc_commit_mc(empty + slot * 8, MASK); else tc_commit_addr(empty + slot * 8);
    ++c;
  }
}

// CL = thread-block cluster size (1, 2 or 4).  With CL > 1 every CTA of a cluster fetches 1/CL of each chunk and
// multicasts it into the ring of all CL CTAs (same slot, same offset), which divides the L2 -> SM traffic - the
// bound of the CL = 1 form - by CL; a slot is reused once the MMAs of all CL CTAs have read it (empty barrier of
// count CL, tcgen05.commit multicast), so the CTAs of a cluster walk their tiles in step.
template <int ACT0, int CL>
__global__ void __launch_bounds__(THREADS, 1)
sdf_tc256_kernel(SdfNetDev net, Tc256Params prm, const uint8_t* __restrict__ img, const float* __restrict__ x,
                 const float* __restrict__ y, const float* __restrict__ sbar, size_t n, float* __restrict__ s_out,
                 float* __restrict__ jx_out, float* __restrict__ jy_out, long long* __restrict__ dbg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Smem* sm = reinterpret_cast<Smem*>(smem_raw);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const bool want_jac = (jx_out != nullptr) || (jy_out != nullptr);
  const size_t n_tiles = (n + TILE - 1) / TILE;
  // every CTA runs the same number of tiles (the ring is shared cluster-wide); tiles past the end compute on clamped
  // points and store nothing
  const size_t my_tiles = (n_tiles + gridDim.x - 1) / gridDim.x;
  const uint32_t rank = CL > 1 ? cluster_rank() : 0u;
  const int act0 = ACT0 >= 0 ? ACT0 : net.act0;

  // ---- one-time setup -------------------------------------------------------------------------------------------
  {
    const float* W0 = net.w + net.off_W0();
    const float* b0 = net.w + net.off_b0();
    const float* b1 = net.w + net.off_b(1);
    const float* w2 = net.w + net.off_wout();
    for (int k = t; k < H; k += THREADS) {
      sm->w0x[k] = W0[2 * k]; sm->w0y[k] = W0[2 * k + 1]; sm->b0[k] = b0[k]; sm->b1[k] = b1[k]; sm->w2[k] = w2[k];
    }
    if (t == 0) {
      for (int i = 0; i < NSLOT; ++i) { mbar_init(sm->full + i, 1); mbar_init(sm->empty + i, CL); }
      mbar_init(sm->mma_bar + 0, 1); mbar_init(sm->mma_bar + 1, 1);
      for (int i = 0; i < 4; ++i) mbar_init(sm->a_bar + i, NCW);
      mbar_init(sm->dfree_bar + 0, NCW); mbar_init(sm->dfree_bar + 1, NCW);
    }
    fence_async_smem();
  }
  if (warp == 0) tmem_alloc(&sm->tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();                // the peers' barriers exist before anything is multicast into them
  tc_fence_after();
  const uint32_t tmem = sm->tmem_slot;

  if (warp == NCW) {
    // ---- producer: stream the operand chunks in consumption order, as far ahead as the ring allows ------------------
    if (lane == 0) {
      const uint32_t ring = smem_u32(sm->ring), full = smem_u32(sm->full), empty = smem_u32(sm->empty);
      const int per_tile = want_jac ? CHUNKS_FWD + CHUNKS_BWD : CHUNKS_FWD;
      uint32_t c = 0;
      for (size_t tl = 0; tl < my_tiles; ++tl)
        for (int ch = 0; ch < per_tile; ++ch, ++c) {
          const uint32_t slot = c % NSLOT;
          if (c >= NSLOT) mbar_wait256(empty + slot * 8, ((c / NSLOT) - 1) & 1);
          mbar_expect_tx(full + slot * 8, CHUNK);
          if (CL > 1) {
            constexpr uint32_t PART = CHUNK / CL;
            bulk_g2s_mc(ring + slot * CHUNK + rank * PART, img + (size_t)ch * CHUNK + rank * PART, PART, full + slot * 8, (uint16_t)((1u << CL) - 1u));
          } else {
            bulk_g2s(ring + slot * CHUNK, img + (size_t)ch * CHUNK, CHUNK, full + slot * 8);
          }
        }
    }
  } else if (warp == NCW + 1) {
    // ---- MMA issuer: half a GEMM as soon as the compute warps have written the matching half of the A operand -------
    if (lane == 0 && elect_lane0()) {
      const uint32_t a_bar = smem_u32(sm->a_bar), dfree = smem_u32(sm->dfree_bar), mma_bar = smem_u32(sm->mma_bar);
      uint32_t consumed = 0;
      for (size_t it = 0; it < my_tiles; ++it) {
        const uint32_t ph = (uint32_t)(it & 1);
        // D of the previous tile must have been read: by epilogue 2, or by epilogue 1 when no Jacobian is asked for
        if (it > 0) mbar_wait256(dfree + (want_jac ? 8 : 0), ph ^ 1);
        mbar_wait256(a_bar + 0, ph); tc_fence_after();
        issue_half256<CL>(sm, tmem, true, 0, consumed);
        mbar_wait256(a_bar + 8, ph); tc_fence_after();
        issue_half256<CL>(sm, tmem, true, 8, consumed);
        tc_commit_addr(mma_bar + 0);
        if (want_jac) {
          mbar_wait256(dfree + 0, ph);                       // every z of GEMM 1 is in registers: D may be overwritten
          mbar_wait256(a_bar + 16, ph); tc_fence_after();
          issue_half256<CL>(sm, tmem, false, 0, consumed);
          mbar_wait256(a_bar + 24, ph); tc_fence_after();
          issue_half256<CL>(sm, tmem, false, 8, consumed);
          tc_commit_addr(mma_bar + 8);
        }
      }
    }
  } else {
    // ---- compute warps ---------------------------------------------------------------------------------------------
    const int q = warp >> 2;                                   // neuron quarter: neurons [64 q, 64 q + 64)
    const int pt = (warp & 3) * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t a_bar = smem_u32(sm->a_bar), dfree = smem_u32(sm->dfree_bar), mma_bar = smem_u32(sm->mma_bar);
    for (size_t it = 0; it < my_tiles; ++it) {
      const uint32_t ph = (uint32_t)(it & 1);
      const size_t tile = blockIdx.x + it * gridDim.x;
      const size_t i = tile * TILE + pt;
      const bool valid = i < n;
      const size_t ic = valid ? i : n - 1;
      const float px = x[ic], py = y[ic], seed = sbar ? sbar[ic] : 1.f;
      long long* dg = (dbg && t == 0 && blockIdx.x == 0 && it < 16) ? dbg + it * 8 : nullptr;
#define T256_STAMP(i) do { if (dg) dg[i] = clock64(); } while (0)
      T256_STAMP(0);
      // ---- layer 0 -> A operand (row-scaled fp16 hi/lo); GEMM 1 starts on the first half while the second is computed ----
      float sc0, inv0;
      row_scale(act_bound(act0, net.p0, fmaf(fabsf(px), prm.max_w0x, fmaf(fabsf(py), prm.max_w0y, prm.max_b0))) + 1e-30f, sc0, inv0);
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int base = q * 64 + cc * 32;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 wx = *reinterpret_cast<const float4*>(sm->w0x + base + 4 * j4);
          const float4 wy = *reinterpret_cast<const float4*>(sm->w0y + base + 4 * j4);
          const float4 bb = *reinterpret_cast<const float4*>(sm->b0 + base + 4 * j4);
          const float a0 = fmaf(wx.x, px, fmaf(wy.x, py, bb.x)), a1 = fmaf(wx.y, px, fmaf(wy.y, py, bb.y));
          const float a2 = fmaf(wx.z, px, fmaf(wy.z, py, bb.z)), a3 = fmaf(wx.w, px, fmaf(wy.w, py, bb.w));
          split_pack_f16(nlo_phi_tc(a0, act0, net.p0) * sc0, nlo_phi_tc(a1, act0, net.p0) * sc0, hi[2 * j4], lo[2 * j4]);
          split_pack_f16(nlo_phi_tc(a2, act0, net.p0) * sc0, nlo_phi_tc(a3, act0, net.p0) * sc0, hi[2 * j4 + 1], lo[2 * j4 + 1]);
        }
        st16(lane_base + COL_AHI + q * 32 + cc * 16, hi);
        st16(lane_base + COL_ALO + q * 32 + cc * 16, lo);
        tc_wait_st();
        warp_arrive(a_bar + cc * 8, lane);
      }
      T256_STAMP(1);
      mbar_wait256(mma_bar + 0, ph);
      tc_fence_after();
      T256_STAMP(3);
      // ---- epilogue 1: value, and the 0/1 mask of the ReLU layer -> A operand.  All 64 pre-activations of this thread are read
      // ---- first, so that GEMM 2 may overwrite D while the second half of the mask is still being computed -------------------
      float s = q == 0 ? net.w[net.off_bout()] : 0.f;
      const float unscale1 = inv0 * prm.inv_sw;
      {
        uint32_t z[2][32];
        tmem_ld32(lane_base + COL_D + q * 64, z[0]);
        tmem_ld32(lane_base + COL_D + q * 64 + 32, z[1]);
        tc_wait_ld();
        warp_arrive(dfree + 0, lane);
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int base = q * 64 + cc * 32;
          uint32_t hi[16];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 bb = *reinterpret_cast<const float4*>(sm->b1 + base + 4 * j4);
            const float4 ww = *reinterpret_cast<const float4*>(sm->w2 + base + 4 * j4);
            const float z0 = fmaf(__uint_as_float(z[cc][4 * j4 + 0]), unscale1, bb.x), z1 = fmaf(__uint_as_float(z[cc][4 * j4 + 1]), unscale1, bb.y);
            const float z2 = fmaf(__uint_as_float(z[cc][4 * j4 + 2]), unscale1, bb.z), z3 = fmaf(__uint_as_float(z[cc][4 * j4 + 3]), unscale1, bb.w);
            // one predicate per neuron serves the value (predicated FFMA) and the fp16 0 / 1 mask
            uint32_t m0 = 0u, m1 = 0u;
            if (z0 > 0.f) { s = fmaf(ww.x, z0, s); m0 |= 0x3C00u; }
            if (z1 > 0.f) { s = fmaf(ww.y, z1, s); m0 |= 0x3C000000u; }
            if (z2 > 0.f) { s = fmaf(ww.z, z2, s); m1 |= 0x3C00u; }
            if (z3 > 0.f) { s = fmaf(ww.w, z3, s); m1 |= 0x3C000000u; }
            hi[2 * j4] = m0; hi[2 * j4 + 1] = m1;
          }
          if (want_jac) {
            st16(lane_base + COL_AHI + q * 32 + cc * 16, hi);
            tc_wait_st();
            warp_arrive(a_bar + (2 + cc) * 8, lane);
          }
        }
      }
      T256_STAMP(4);
      float jx = 0.f, jy = 0.f;
      if (want_jac) {
        mbar_wait256(mma_bar + 8, ph);
        tc_fence_after();
        T256_STAMP(6);
        // ---- epilogue 2: through layer 0 to the Jacobian (D is read up front: the next tile's GEMM 1 may then start) ---------
        uint32_t gz[2][32];
        tmem_ld32(lane_base + COL_D + q * 64, gz[0]);
        tmem_ld32(lane_base + COL_D + q * 64 + 32, gz[1]);
        tc_wait_ld();
        warp_arrive(dfree + 8, lane);
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int base = q * 64 + cc * 32;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 wx = *reinterpret_cast<const float4*>(sm->w0x + base + 4 * j4);
            const float4 wy = *reinterpret_cast<const float4*>(sm->w0y + base + 4 * j4);
            const float4 bb = *reinterpret_cast<const float4*>(sm->b0 + base + 4 * j4);
            const float wxs[4] = {wx.x, wx.y, wx.z, wx.w}, wys[4] = {wy.x, wy.y, wy.z, wy.w}, bs[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (ACT0 == NLO_ACT_RELU) {          // phi0' is 0 / 1: predicated FFMAs
                if (fmaf(wxs[e], px, fmaf(wys[e], py, bs[e])) > 0.f) {
                  jx = fmaf(__uint_as_float(gz[cc][4 * j4 + e]), wxs[e], jx); jy = fmaf(__uint_as_float(gz[cc][4 * j4 + e]), wys[e], jy);
                }
                continue;
              }
              float v, d;
              nlo_phi_d_tc(fmaf(wxs[e], px, fmaf(wys[e], py, bs[e])), act0, net.p0, v, d);
              const float g0 = __uint_as_float(gz[cc][4 * j4 + e]) * d;
              jx = fmaf(g0, wxs[e], jx); jy = fmaf(g0, wys[e], jy);
            }
          }
        }
        const float unscale2 = seed * prm.inv_sv;
        jx *= unscale2; jy *= unscale2;
      }
      T256_STAMP(7);
      // ---- add the four quarters of a point -----------------------------------------------------------------------------
      sm->part[q][pt][0] = s; sm->part[q][pt][1] = jx; sm->part[q][pt][2] = jy;
      compute_bar();
      if (q == 0 && valid) {
        float r[3];
#pragma unroll
        for (int e = 0; e < 3; ++e) r[e] = sm->part[0][pt][e] + sm->part[1][pt][e] + sm->part[2][pt][e] + sm->part[3][pt][e];
        if (s_out) s_out[i] = r[0];
        if (jx_out) jx_out[i] = r[1];
        if (jy_out) jy_out[i] = r[2];
      }
      compute_bar();                                   // `part` is rewritten by the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();                // no CTA leaves while a peer may still multicast into it
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int ACT0, int CL>
int launch_tc256_cl(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n, float* s, float* jx, float* jy, cudaStream_t st) {
  auto kfn = sdf_tc256_kernel<ACT0, CL>;
  const size_t smem = sizeof(Smem) + 1024;
  NLO_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const size_t tiles = (n + TILE - 1) / TILE;
  size_t grid = tiles < (size_t)m->sm_count ? tiles : (size_t)m->sm_count;
  grid = (grid + CL - 1) / CL * CL;
  if (grid > (size_t)m->sm_count) grid = (size_t)m->sm_count / CL * CL;
  Tc256Params prm;
  memcpy(&prm, m->tc_params, sizeof(prm));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  long long* dbg = nullptr;
  if (getenv("NLO_B200_TC_TIMELINE")) { NLO_CUDA(cudaMalloc(&dbg, 16 * 8 * sizeof(long long))); NLO_CUDA(cudaMemsetAsync(dbg, 0, 16 * 8 * sizeof(long long), st)); }
  NLO_CUDA(cudaLaunchKernelEx(&cfg, kfn, m->net(), prm, static_cast<const uint8_t*>(m->d_tc), x, y, sbar, n, s, jx, jy, dbg));
  if (dbg) {
    long long h[16 * 8];
    NLO_CUDA(cudaMemcpyAsync(h, dbg, sizeof(h), cudaMemcpyDeviceToHost, st));
    NLO_CUDA(cudaStreamSynchronize(st));
    cudaFree(dbg);
    for (int it = 4; it < 8; ++it) {
      const long long* r = h + it * 8;
      fprintf(stderr, "[tc256 timeline] tile%2d: L0 %5lld wait-GEMM1 %5lld E1 %5lld wait-GEMM2 %5lld E2 %5lld | period %6lld\n", it,
              r[1] - r[0], r[3] - r[1], r[4] - r[3], r[6] - r[4], r[7] - r[6], (h + (it + 1) * 8)[0] - r[0]);
    }
  }
  NLO_CHECK_LAUNCH();
  return 0;
}

template <int ACT0>
int launch_tc256(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n, float* s, float* jx, float* jy, cudaStream_t st) {
  static const int cl = [] { const char* e = getenv("NLO_B200_TC256_CLUSTER"); const int v = e ? atoi(e) : 2; return (v == 1 || v == 2 || v == 4) ? v : 2; }();
  if (cl == 4) return launch_tc256_cl<ACT0, 4>(m, x, y, sbar, n, s, jx, jy, st);
  if (cl == 2) return launch_tc256_cl<ACT0, 2>(m, x, y, sbar, n, s, jx, jy, st);
  return launch_tc256_cl<ACT0, 1>(m, x, y, sbar, n, s, jx, jy, st);
}

float pow2_scale(float mx) {
  int ex = 0;
  if (mx > 0.f) frexpf(mx, &ex);
  return ldexpf(1.f, 14 - ex);
}

}  // namespace

bool nlo_sdf_tc256_supported(const nlo_sdf_desc* d) {
  return d->n_hidden_mats == 1 && d->hidden == 256 && d->act == NLO_ACT_RELU;
}

// Images in consumption order: position i holds k-step ks_of(i); chunk 2i = W1hi[ks], 2i+1 = W1lo[ks] (forward), 32+2i = Vlo[ks],
// 32+2i+1 = Vhi[ks] (reverse); inside a chunk element (n, kk) of the K = 16 step sits at ((kk/8)*32 + n/8)*128 + (n%8)*16 + (kk%8)*2 bytes
// (UMMA K-major core-matrix order, no swizzle: LBO = 4096, SBO = 128).
int nlo_sdf_tc256_prepare(nlo_sdf_model* m, const float* w) {
  const float* W0 = w;
  const float* b0 = w + 2 * H;
  const float* W1 = w + 3 * H;
  const float* b1 = W1 + (size_t)H * H;
  const float* w2 = b1 + H;
  float mx = 0.f, mv = 0.f;
  for (int j = 0; j < H; ++j)
    for (int k = 0; k < H; ++k) {
      mx = fmaxf(mx, fabsf(W1[(size_t)j * H + k]));
      mv = fmaxf(mv, (float)fabs((double)w2[j] * (double)W1[(size_t)j * H + k]));
    }
  const float sw = pow2_scale(mx), sv = pow2_scale(mv);
  Tc256Params prm;
  prm.inv_sw = 1.f / sw; prm.inv_sv = 1.f / sv;
  prm.max_w0x = prm.max_w0y = prm.max_b0 = 0.f;
  for (int k = 0; k < H; ++k) {
    prm.max_w0x = fmaxf(prm.max_w0x, fabsf(W0[2 * k])); prm.max_w0y = fmaxf(prm.max_w0y, fabsf(W0[2 * k + 1]));
    prm.max_b0 = fmaxf(prm.max_b0, fabsf(b0[k]));
  }
  static_assert(sizeof(Tc256Params) <= sizeof(m->tc_params), "tc_params too small");
  memcpy(m->tc_params, &prm, sizeof(prm));
  std::vector<__half> img((size_t)(CHUNKS_FWD + CHUNKS_BWD) * CHUNK / 2);
  auto at = [&](int chunk, int nn, int kk) -> __half& {
    return img[(size_t)chunk * (CHUNK / 2) + ((size_t)(kk / 8) * (H / 8) + nn / 8) * 64 + (nn % 8) * 8 + (kk % 8)];
  };
  for (int i = 0; i < H / 16; ++i)                       // consumption position; ks_of(i) = the k-step it holds
    for (int nn = 0; nn < H; ++nn)
      for (int kk = 0; kk < 16; ++kk) {
        const int ks = ks_of(i);
        const int k = ks * 16 + kk;
        // forward: B(n, k) = W1[n][k]
        const float v = W1[(size_t)nn * H + k] * sw;
        const __half hi = __float2half_rn(v);
        at(2 * i, nn, kk) = hi;
        at(2 * i + 1, nn, kk) = __float2half_rn(v - __half2float(hi));
        // reverse: B'(n' = nn, k' = k) = V[k][nn] = w2[k] W1[k][nn]
        const double vv = (double)w2[k] * (double)W1[(size_t)k * H + nn] * (double)sv;
        const __half vhi = __float2half_rn((float)vv);
        at(CHUNKS_FWD + 2 * i + 1, nn, kk) = vhi;
        at(CHUNKS_FWD + 2 * i, nn, kk) = __float2half_rn((float)(vv - (double)__half2float(vhi)));
      }
  if (m->d_tc) cudaFree(m->d_tc);
  m->d_tc = nullptr;
  NLO_CUDA(cudaMalloc(&m->d_tc, img.size() * sizeof(__half)));
  NLO_CUDA(cudaMemcpy(m->d_tc, img.data(), img.size() * sizeof(__half), cudaMemcpyHostToDevice));
  m->tc_bytes = img.size() * sizeof(__half);
  return 0;
}

int nlo_sdf_tc256_launch(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n,
                         float* s, float* jx, float* jy, cudaStream_t st) {
  if (n == 0) return 0;
  if (!m->d_tc) return nlo_fail("tensor-tile operands were not prepared");
  const int a0 = (int)m->desc.act0;
  if (a0 == NLO_ACT_RELU) return launch_tc256<NLO_ACT_RELU>(m, x, y, sbar, n, s, jx, jy, st);
  if (a0 == NLO_ACT_COS_SCALE) return launch_tc256<NLO_ACT_COS_SCALE>(m, x, y, sbar, n, s, jx, jy, st);
  return launch_tc256<-1>(m, x, y, sbar, n, s, jx, jy, st);
}
