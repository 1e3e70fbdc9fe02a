// tcgen05 / TMEM / mbarrier PTX wrappers and the split-fp16 helpers shared by the tensor-path kernels
// (sdf_tc.cu: H = 64 / 128 with resident operand images; sdf_tc256.cu: H = 256 with streamed operand images).
#pragma once
#include "nlo_common.cuh"
#include <cuda_fp16.h>

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
// D[tmem] (+)= A[tmem] . B[smem descriptor], kind::f16, issued by one thread
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
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
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N, int b_mn_major) {
  return (1u << 4)                              // D format: F32
         | (0u << 7) | (0u << 10)               // A, B format: F16
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

// Split two row-scaled values into fp16 hi + fp16 lo and pack the pair the way a 16-bit A operand sits in
// tensor memory (element 2c in the low half of column c, element 2c+1 in the high half).
// hi = round-to-nearest fp16 of the value, lo = round-to-nearest fp16 of the exact remainder (22 significant bits
// together).  One packed conversion each way; the unpack + subtract run on the FMA pipe, keeping the ALU pipe free.
__device__ __forceinline__ void split_pack_f16(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  const __half2 ph = __floats2half2_rn(v0, v1);
  const float2 hf = __half22float2(ph);
  const __half2 pl = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&ph);
  lo = *reinterpret_cast<const uint32_t*>(&pl);
}
// 2^e such that bound * 2^e lies in [2^13, 2^15): exact scaling into fp16's comfortable range
__device__ __forceinline__ void row_scale(float bound, float& sc, float& inv) {
  const int ex = (int)((__float_as_uint(bound) >> 23) & 0xffu) - 127;      // floor(log2(bound)) for normal bound
  int e = 13 - ex;
  e = e < -60 ? -60 : (e > 60 ? 60 : e);
  sc = __uint_as_float((uint32_t)(127 + e) << 23);
  inv = __uint_as_float((uint32_t)(127 - e) << 23);
}

struct TcParams {          // built by nlo_sdf_tc_prepare
  float inv_sw;            // 1 / (power-of-two scale applied to W1 in its fp16 images)
  float inv_sv;            // 1 / (power-of-two scale applied to V = diag(w2) W1 in its fp16 images)
  float inv_sc1;           // 1 / (power-of-two scale folded into w2s, the generic reverse-pass seed vector)
  float max_w0x, max_w0y, max_b0;
  // layer 0 on the tensor core (sdf_tc.cu, ReLU / ReLU form): 1 / (power-of-two scales of the W0 columns and of b0 in their fp16 pieces)
  float spx_mul, spy_mul, cb_mul;
};
// Small vectors of the network live in __constant__ memory: with fully unrolled loops every use is an FFMA/FMUL
// with a constant-bank operand (c[3][imm]) - no load instruction and no shared-memory bandwidth (which the tensor
// core needs for its B-operand fetches).  One copy per device context; nlo_sdf_tc_launch re-uploads it (after a
// device-wide sync) whenever a different model is evaluated.
struct TcConst {
  float w0x[128], w0y[128], b0[128], b1[128], w2[128], w2s[128];
  float bout;
};
template <int N> struct TmemIO;
template <> struct TmemIO<16> {
  __device__ static __forceinline__ void ld(uint32_t a, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(a) : "memory");
  }
  __device__ static __forceinline__ void st(uint32_t a, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};"
                 ::"r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                   "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(a) : "memory");
  }
};

template <> struct TmemIO<8> {
  __device__ static __forceinline__ void st(uint32_t a, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"
                 ::"r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(a) : "memory");
  }
};
// {fp16(max(v0, 0)) in the low half, fp16(max(v1, 0)) in the high half}: ReLU folded into the packed conversion
__device__ __forceinline__ uint32_t pack_relu_f16(float v0, float v1) {
  uint32_t d;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(v1), "f"(v0));
  return d;
}

// the same with round-toward-zero: the result never exceeds the value, so value - result >= 0 for a positive value
__device__ __forceinline__ uint32_t pack_relu_rz_f16(float v0, float v1) {
  uint32_t d;
  asm("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(v1), "f"(v0));
  return d;
}

// named barrier of one tile group
template <int NT>
__device__ __forceinline__ void group_bar(uint32_t id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NT) : "memory"); }

__device__ __forceinline__ bool elect_one(uint32_t mask) {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, %1;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred) : "r"(mask));
  return pred != 0;
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t a, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ uint32_t pack_f16(float v0, float v1) {
  const __half2 p = __floats2half2_rn(v0, v1);
  return *reinterpret_cast<const uint32_t*>(&p);
}

template <int A>
__device__ __forceinline__ void act_vd(float a, int rt, float prm, float& v, float& d) {
  if (A >= 0) nlo_phi_d_tc(a, A, prm, v, d); else nlo_phi_d_tc(a, rt, prm, v, d);
}
// upper bound of |phi(a)| given |a| <= ba, and of |phi'|
__device__ __forceinline__ float act_bound(int act, float prm, float ba) {
  switch (act) {
    case NLO_ACT_TANH: case NLO_ACT_SIGMOID: case NLO_ACT_SIN: return 1.f;
    case NLO_ACT_COS_SCALE: return fabsf(prm);
    default: return ba;
  }
}

}  // namespace
