// K1 (tcgen05 form): fused learned-SDF value + Jacobian / adjoint with the two H x H contractions on
// 5th-generation tensor cores, error-compensated 3xTF32 with FP32 accumulation in tensor memory.
//
// Replaces the per-point TorchScript calls behind _l4c_generated/nn_sdf.cpp:57-83 (nn_sdf, jac_nn_sdf,
// adj1_nn_sdf) for networks 2 -> H -> H -> 1 with H in {64, 128}: the benchmark model
// (scripts/run_benchmark.py:65 with benchmarks/*.yaml model: hidden_dim 128, num_hidden_layers 2) and
// the shipped FourierMLP-128 (_l4c_generated/nn_sdf.pt).
//
// One persistent CTA per SM, 128 threads, tile = 128 points; thread t owns point t == TMEM lane t, so
// there is no cross-thread traffic at all:
//   layer 0 (SIMT)   h0 = phi0(W0 p + b0)                       -> split hi/lo -> tcgen05.st -> A (TMEM)
//   GEMM 1 (tcgen05) Z1[128 x H] = H0 . W1^T   3 passes (lo.hi, hi.lo, hi.hi), B = W1 K-major in smem
//   epilogue 1       tcgen05.ld Z1; s = w2.phi(z1+b1)+b2; g1 = sbar*w2*phi'(z1+b1) -> hi/lo -> A (TMEM)
//   GEMM 2 (tcgen05) G0[128 x H] = G1 . W1     same smem bytes read through an MN-major descriptor
//   epilogue 2       tcgen05.ld G0; g0 = G0 * phi0'(a0); J = g0 . W0
// W1 is split once on the host into tf32 hi + tf32 lo and stored in UMMA core-matrix order (no swizzle):
// element (n,k) at ((k/4)*(H/8) + n/8)*128 + (n%8)*16 + (k%4)*4 bytes, which is simultaneously the
// canonical K-major layout of B(n,k) = W1[n][k] (LBO = 16H, SBO = 128) and the canonical MN-major layout
// of B'(i,j) = W1[j][i] (LBO = 128, SBO = 16H).
#include "nlo_common.cuh"
#include <vector>
#include <cstring>

namespace {

// ---- PTX wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem descriptor], kind::tf32, issued by one thread
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  return d;                                     // base_offset 0, lbo_mode 0, layout_type 0 = no swizzle
}
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, int b_mn_major) {
  return (1u << 4)                              // D format: F32
         | (2u << 7) | (2u << 10)               // A, B format: TF32
         | ((uint32_t)b_mn_major << 16)         // B major: 0 = K, 1 = MN
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

#define TM_R32(v) "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),            \
  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),          \
  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
#define TM_W32(v) "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),          \
  "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),                    \
  "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),                  \
  "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])

// 32 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
      "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : TM_R32(v) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
      "%24,%25,%26,%27,%28,%29,%30,%31};"
      ::TM_W32(v), "r"(taddr) : "memory");
}

// tf32 split of a runtime value: hi = top 19 bits (truncation), lo = exact remainder, truncated the same way
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(v) & 0xffffe000u;
  lo = __float_as_uint(v - __uint_as_float(hi)) & 0xffffe000u;
}

template <int A>
__device__ __forceinline__ void act_vd(float a, int rt, float prm, float& v, float& d) {
  if (A >= 0) nlo_phi_d(a, A, prm, v, d); else nlo_phi_d(a, rt, prm, v, d);
}

constexpr int TILE = 128;

template <int H>
struct TcSmem {
  static constexpr int B_FLOATS = 2 * H * H;                        // hi | lo images
  static constexpr int VEC_FLOATS = 2 * H + 3 * H + 4;              // W0 | b0 | b1 | w2 | b_out,pad
  static constexpr size_t BYTES = (size_t)(B_FLOATS + VEC_FLOATS) * 4 + 16;
  static constexpr uint32_t TMEM_COLS = (3 * H <= 256) ? 256 : 512;
};

template <int H, int ACT0, int ACT>
__global__ void __launch_bounds__(TILE, 1) sdf_tc_kernel(SdfNetDev net, const float* __restrict__ bimg, const float* __restrict__ x,
                                                         const float* __restrict__ y, const float* __restrict__ sbar, size_t n,
                                                         float* __restrict__ s_out, float* __restrict__ jx_out, float* __restrict__ jy_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  float* sB = reinterpret_cast<float*>(smem_raw);
  float* sW0 = sB + TcSmem<H>::B_FLOATS;
  float* sb0 = sW0 + 2 * H;
  float* sb1 = sb0 + H;
  float* sw2 = sb1 + H;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sw2 + H + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int t = threadIdx.x, warp = t >> 5;
  constexpr int NCH = H / 32;

  // ---- one-time setup: operand images to smem, barrier, tensor memory ---------------------------------------
  {
    const float4* src = reinterpret_cast<const float4*>(bimg);
    float4* dst = reinterpret_cast<float4*>(sB);
    for (int i = t; i < TcSmem<H>::B_FLOATS / 4; i += TILE) dst[i] = src[i];
    for (int i = t; i < 2 * H; i += TILE) sW0[i] = net.w[net.off_W0() + i];
    for (int i = t; i < H; i += TILE) { sb0[i] = net.w[net.off_b0() + i]; sb1[i] = net.w[net.off_b(1) + i]; sw2[i] = net.w[net.off_wout() + i]; }
    if (t == 0) { sw2[H] = net.w[net.off_bout()]; mbar_init(mbar, 1); }
    fence_async_smem();                          // generic-proxy smem writes -> visible to the tensor-core (async) proxy
  }
  if (warp == 0) tmem_alloc(tmem_slot, TcSmem<H>::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
  const uint32_t COL_AHI = 0, COL_ALO = H, COL_D = 2 * H;
  const float bout = sw2[H];
  const bool want_jac = (jx_out != nullptr) || (jy_out != nullptr);
  const float prm0 = net.p0, prm = net.p;
  const uint32_t sB_hi = smem_u32(sB), sB_lo = smem_u32(sB + H * H);
  constexpr uint32_t IDESC_FWD = umma_idesc_tf32(TILE, H, 0), IDESC_BWD = umma_idesc_tf32(TILE, H, 1);
  uint32_t phase = 0;

  // one thread issues the 3 x (H/8) MMAs of a GEMM and commits them to the mbarrier
  auto issue_gemm = [&](bool fwd) {
    if (t == 0) {
      tc_fence_after();
      const uint32_t lbo = fwd ? 16u * H : 128u, sbo = fwd ? 128u : 16u * H;
      const uint32_t kstep_bytes = fwd ? 32u * H : 128u;
      const uint32_t idesc = fwd ? IDESC_FWD : IDESC_BWD;
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {               // smallest terms first: lo.hi, hi.lo, hi.hi
        const uint32_t a_col = (pass == 0) ? COL_ALO : COL_AHI;
        const uint32_t b_base = (pass == 1) ? sB_lo : sB_hi;
#pragma unroll 4
        for (int ks = 0; ks < H / 8; ++ks) {
          tc_mma_tf32_ts(tmem_base + COL_D, tmem_base + a_col + ks * 8, umma_desc(b_base + ks * kstep_bytes, lbo, sbo), idesc,
                         (pass | ks) != 0);
        }
      }
      tc_commit(mbar);
    }
  };

  const size_t n_tiles = (n + TILE - 1) / TILE;
  for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const size_t i = tile * TILE + t;
    const bool valid = i < n;
    const size_t ic = valid ? i : n - 1;
    const float px = x[ic], py = y[ic];
    const float seed = sbar ? sbar[ic] : 1.f;

    // ---- layer 0 -> A operand ----------------------------------------------------------------------------------
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const int k = c * 32 + q;
        const float2 w0 = *reinterpret_cast<const float2*>(sW0 + 2 * k);
        const float a = fmaf(w0.x, px, fmaf(w0.y, py, sb0[k]));
        float v;
        if (ACT0 >= 0) v = nlo_phi(a, ACT0, prm0); else v = nlo_phi(a, net.act0, prm0);
        split_tf32(v, hi[q], lo[q]);
      }
      tmem_st32(lane_base + COL_AHI + c * 32, hi);
      tmem_st32(lane_base + COL_ALO + c * 32, lo);
    }
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    issue_gemm(true);
    mbar_wait(mbar, phase); phase ^= 1;
    tc_fence_after();

    // ---- epilogue 1: value, and g1 -> A operand ------------------------------------------------------------------
    float s = bout;
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      uint32_t z[32];
      tmem_ld32(lane_base + COL_D + c * 32, z);
      tc_wait_ld();
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const int j = c * 32 + q;
        float v, d;
        act_vd<ACT>(__uint_as_float(z[q]) + sb1[j], net.act, prm, v, d);
        const float w2 = sw2[j];
        s = fmaf(w2, v, s);
        split_tf32(seed * w2 * d, hi[q], lo[q]);
      }
      if (want_jac) {
        tmem_st32(lane_base + COL_AHI + c * 32, hi);
        tmem_st32(lane_base + COL_ALO + c * 32, lo);
      }
    }
    if (valid && s_out) s_out[i] = s;
    if (want_jac) {
      tc_wait_st();
      tc_fence_before();
      __syncthreads();
      issue_gemm(false);
      mbar_wait(mbar, phase); phase ^= 1;
      tc_fence_after();
      // ---- epilogue 2: through layer 0 to the Jacobian ---------------------------------------------------------------
      float jx = 0.f, jy = 0.f;
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        uint32_t gz[32];
        tmem_ld32(lane_base + COL_D + c * 32, gz);
        tc_wait_ld();
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const int k = c * 32 + q;
          const float2 w0 = *reinterpret_cast<const float2*>(sW0 + 2 * k);
          const float a = fmaf(w0.x, px, fmaf(w0.y, py, sb0[k]));
          float v, d;
          act_vd<ACT0>(a, net.act0, prm0, v, d);
          const float g0 = __uint_as_float(gz[q]) * d;
          jx = fmaf(g0, w0.x, jx);
          jy = fmaf(g0, w0.y, jy);
        }
      }
      if (valid) { if (jx_out) jx_out[i] = jx; if (jy_out) jy_out[i] = jy; }
    }
    // the next tile's tcgen05.st / MMA must not overtake this tile's TMEM reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TcSmem<H>::TMEM_COLS);
}

inline uint32_t rn_tf32_bits(float v) {
  uint32_t b; memcpy(&b, &v, 4);
  if ((b & 0x7f800000u) == 0x7f800000u) return b & 0xffffe000u;
  b += 0x00000fffu + ((b >> 13) & 1u);
  return b & 0xffffe000u;
}

template <int H, int ACT0, int ACT>
int launch_tc(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n, float* s, float* jx, float* jy, cudaStream_t st) {
  auto kfn = sdf_tc_kernel<H, ACT0, ACT>;
  // H = 64 allocates 256 TMEM columns: at most two CTAs may share an SM, so pad the request past a third of the SM
  size_t smem = TcSmem<H>::BYTES;
  if (H == 64 && smem < 80 * 1024) smem = 80 * 1024;
  NLO_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const size_t tiles = (n + TILE - 1) / TILE;
  const size_t cap = (size_t)m->sm_count * (H == 64 ? 2 : 1);
  const int grid = (int)(tiles < cap ? tiles : cap);
  kfn<<<grid, TILE, smem, st>>>(m->net(), m->d_tc, x, y, sbar, n, s, jx, jy);
  NLO_CHECK_LAUNCH();
  return 0;
}

}  // namespace

bool nlo_sdf_tc_supported(const nlo_sdf_desc* d) {
  return d->n_hidden_mats == 1 && (d->hidden == 64 || d->hidden == 128);
}

// Split W1 into tf32 hi + lo (round-to-nearest) and store both in UMMA core-matrix order.
int nlo_sdf_tc_prepare(nlo_sdf_model* m, const float* w) {
  const int H = (int)m->desc.hidden;
  const float* W1 = w + 3 * H;                          // blob order: W0[H][2], b0[H], W1[H][H], ...
  std::vector<float> img((size_t)2 * H * H);
  for (int nn = 0; nn < H; ++nn)
    for (int k = 0; k < H; ++k) {
      const float v = W1[(size_t)nn * H + k];
      const uint32_t hb = rn_tf32_bits(v);
      float hi; memcpy(&hi, &hb, 4);
      const uint32_t lb = rn_tf32_bits(v - hi);
      float lo; memcpy(&lo, &lb, 4);
      const size_t off = ((size_t)(k / 4) * (H / 8) + nn / 8) * 32 + (nn % 8) * 4 + (k % 4);   // in floats
      img[off] = hi;
      img[(size_t)H * H + off] = lo;
    }
  if (m->d_tc) cudaFree(m->d_tc);
  m->d_tc = nullptr;
  NLO_CUDA(cudaMalloc(&m->d_tc, img.size() * sizeof(float)));
  NLO_CUDA(cudaMemcpy(m->d_tc, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice));
  m->tc_bytes = img.size() * sizeof(float);
  return 0;
}

int nlo_sdf_tc_launch(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n,
                      float* s, float* jx, float* jy, cudaStream_t st) {
  if (n == 0) return 0;
  if (!m->d_tc) return nlo_fail("tensor-tile operands were not prepared");
  const int H = (int)m->desc.hidden, a0 = (int)m->desc.act0, a = (int)m->desc.act;
#define NLO_TC(HH)                                                                                                 \
  do {                                                                                                             \
    if (a0 == NLO_ACT_RELU && a == NLO_ACT_RELU) return launch_tc<HH, NLO_ACT_RELU, NLO_ACT_RELU>(m, x, y, sbar, n, s, jx, jy, st);          \
    if (a0 == NLO_ACT_COS_SCALE && a == NLO_ACT_RELU) return launch_tc<HH, NLO_ACT_COS_SCALE, NLO_ACT_RELU>(m, x, y, sbar, n, s, jx, jy, st); \
    return launch_tc<HH, -1, -1>(m, x, y, sbar, n, s, jx, jy, st);                                                 \
  } while (0)
  if (H == 128) NLO_TC(128);
  if (H == 64) NLO_TC(64);
#undef NLO_TC
  return nlo_fail("tensor-tile path: unsupported width %d", H);
}
