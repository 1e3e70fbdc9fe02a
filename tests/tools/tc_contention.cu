// What stretches the H = 128 tile kernel's MMAs beyond their floor?  One CTA of 512 threads per SM: lane 0 of warp 0 issues
// back-to-back "GEMMs" of 24 tcgen05.mma (M128 N128 K16, A in tensor memory, B in shared memory - the forward GEMM of
// sdf_tc_kernel<128>) while warps 8..15 run one kind of background work, as the other tile group's SIMT phases would.
// Prints one JSON line per background mode: cycles per GEMM (floor: 24 x 64 = 1536).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I nlotrajectories_b200/csrc tests/tools/tc_contention.cu -o tests/tools/tc_contention
#include "tc_ptx.cuh"
#include <cstdio>
#include <vector>

namespace {
constexpr int H = 128, TILE = 128;
__constant__ float kc[256];

template <int n_mma>
__global__ void __launch_bounds__(512, 1) contention_kernel(const __half* __restrict__ bimg, int mode, int n_gemm, int bg_warps, int sched_mask,
                                                            long long* __restrict__ cycles, float* __restrict__ sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __half* sB = reinterpret_cast<__half*>(smem_raw);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + 2 * H * H * 2);
  volatile int* done = reinterpret_cast<volatile int*>(mbar + 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(const_cast<int*>(done) + 1);
  const int t = threadIdx.x, warp = t >> 5;
  {
    const uint4* src = reinterpret_cast<const uint4*>(bimg);
    uint4* dst = reinterpret_cast<uint4*>(sB);
    for (int i = t; i < 2 * H * H / 8; i += 512) dst[i] = src[i];
    if (t == 0) { mbar_init(mbar, 1); mbar_init(mbar + 1, 1); *done = 0; }
    fence_async_smem();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  {                                                         // A operand: something non-trivial in columns 0..127 of every lane
    uint32_t v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = 0x3C003800u + (uint32_t)(t * 16 + q) * 0x00010001u % 0x03ff03ffu;
    if (warp < 4)
      for (int cc = 0; cc < 8; ++cc) TmemIO<16>::st(lane_base + cc * 16, v);
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t sB_hi = smem_u32(sB), sB_lo = smem_u32(sB + H * H), mb = smem_u32(mbar);
  constexpr uint32_t IDESC = umma_idesc_f16(TILE, H, 0);
  if (warp == 0) {
    long long t0 = clock64();
    uint32_t phase = 0;
    for (int g = 0; g < n_gemm; ++g) {
      if (t == 0 && elect_one(1u)) {
#pragma unroll
        for (int i = 0; i < n_mma; ++i) {
          const int pass = i >> 3, ks = i & 7;
          const uint32_t a_col = (pass % 3 == 0) ? 64u : 0u;
          const uint32_t b_base = (pass % 3 == 1) ? sB_lo : sB_hi;
          tc_mma_f16_ts(tmem + 128, tmem + a_col + ks * 8, umma_desc(b_base + ks * 32u * H, 16u * H, 128u), IDESC, i != 0);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mb) : "memory");
      }
      __syncwarp();
      mbar_wait_addr(mb, phase); phase ^= 1;
      tc_fence_after();
    }
    long long t1 = clock64();
    if (t == 0) { cycles[blockIdx.x] = t1 - t0; *done = 1; }
  } else if (warp >= 16 - bg_warps && ((sched_mask >> (warp & 3)) & 1)) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = (float)(t + i);
    long long iters = 0;
    while (!*done) {
      ++iters;
      if (mode == 1) {                                       // FFMA, register operands
#pragma unroll
        for (int r = 0; r < 32; ++r)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fmaf(acc[i], 1.0001f, 0.5f);
      } else if (mode == 2) {                                // tcgen05.ld x32 + wait
        uint32_t z[32];
        tmem_ld32(lane_base + 256 + (iters & 3) * 32, z);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += __uint_as_float(z[i] ^ z[i + 8] ^ z[i + 16] ^ z[i + 24]);
      } else if (mode == 3) {                                // tcgen05.st x16 + wait
        uint32_t v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = __float_as_uint(acc[q & 7]) + q;
        TmemIO<16>::st(lane_base + 256 + (iters & 7) * 16, v);
        tc_wait_st();
        acc[0] += 1.f;
      } else if (mode == 4) {                                // FFMA, constant-bank operands
#pragma unroll
        for (int r = 0; r < 32; ++r)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fmaf(acc[i], kc[r * 8 + i], kc[(r * 8 + i + 1) & 255]);
      } else if (mode == 5) {                                // an epilogue-1-like mix: ld 32, ~5 FFMA-class per column, st 16
        uint32_t z[32], v[16];
        tmem_ld32(lane_base + 256 + (iters & 3) * 32, z);
        tc_wait_ld();
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float a = fmaf(__uint_as_float(z[2 * q]), kc[q], kc[q + 32]), b = fmaf(__uint_as_float(z[2 * q + 1]), kc[q + 64], kc[q + 96]);
          if (a > 0.f) acc[q & 7] = fmaf(kc[q + 128], a, acc[q & 7]);
          if (b > 0.f) acc[(q + 1) & 7] = fmaf(kc[q + 160], b, acc[(q + 1) & 7]);
          v[q] = (a > 0.f ? 0x3C00u : 0u) | (b > 0.f ? 0x3C000000u : 0u);
        }
        TmemIO<16>::st(lane_base + 384 + (iters & 7) * 16, v);
        tc_wait_st();
      } else if (mode == 6) {                                // layer-0-like: FFMA + cvt + st hi / lo
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float a = fmaxf(fmaf(kc[q], acc[0], fmaf(kc[q + 32], acc[1], kc[q + 64])), 0.f) * acc[2];
          const float b = fmaxf(fmaf(kc[q + 96], acc[0], fmaf(kc[q + 128], acc[1], kc[q + 160])), 0.f) * acc[2];
          split_pack_f16(a, b, hi[q], lo[q]);
        }
        TmemIO<16>::st(lane_base + 256 + (iters & 3) * 16, hi);
        TmemIO<16>::st(lane_base + 320 + (iters & 3) * 16, lo);
        tc_wait_st();
        acc[0] += 1e-3f;
      } else if (mode == 7 || mode == 8) {                   // epilogue-2-like, dense: per neuron 2 FFMA + compare + 2 predicated FFMA, constants
        float px = acc[0], py = acc[1];                      // as uniform-register operands, no waits inside (the D columns are loaded once)
#pragma unroll
        for (int q = 0; q < 64; ++q) {
          const float a = fmaf(kc[q], px, fmaf(kc[q + 64], py, kc[q + 128]));
          if (a > 0.f) { acc[2 + (q & 1)] = fmaf(acc[4 + (q & 3)], kc[q], acc[2 + (q & 1)]); acc[6 + (q & 1)] = fmaf(acc[4 + (q & 3)], kc[q + 64], acc[6 + (q & 1)]); }
          if (mode == 8 && (q & 15) == 15) __nanosleep(0);
        }
        acc[0] += 1e-3f;
      } else {
        __nanosleep(200);
      }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 12345.678f) sink[t] = s + (float)iters;
    if ((t & 31) == 0) atomicAdd(reinterpret_cast<unsigned long long*>(cycles + 256 + blockIdx.x), (unsigned long long)iters);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
}  // namespace

int main(int argc, char** argv) {
  const int n_gemm = argc > 1 ? atoi(argv[1]) : 400;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  std::vector<__half> img(2 * H * H);
  for (size_t i = 0; i < img.size(); ++i) img[i] = __float2half((float)((i * 2654435761u) % 2001) / 1000.f - 1.f);
  std::vector<float> kh(256);
  for (int i = 0; i < 256; ++i) kh[i] = 0.5f + (float)i / 512.f;
  cudaMemcpyToSymbol(kc, kh.data(), sizeof(float) * 256);
  __half* d_img; long long* d_cyc; float* d_sink;
  cudaMalloc(&d_img, img.size() * 2); cudaMalloc(&d_cyc, 512 * 8); cudaMalloc(&d_sink, 512 * 4);
  cudaMemcpy(d_img, img.data(), img.size() * 2, cudaMemcpyHostToDevice);
  const size_t smem = 2 * H * H * 2 + 64;
  cudaFuncSetAttribute(contention_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(contention_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const char* names[] = {"idle", "ffma_reg", "tcgen05_ld", "tcgen05_st", "ffma_const", "epilogue1_mix", "layer0_mix", "epilogue2_dense", "epilogue2_dense_yield"};
  // sched: which of the four warp schedulers (warp % 4) may host background warps; the issuing warp 0 sits on scheduler 0
  for (int n_mma : {24, 16})
    for (int bg : {8, 12})
     for (int sched : {0xF, 0xE, 0x1})
      for (int mode = 0; mode < 9; ++mode) {
        if (mode == 0 && (bg != 8 || sched != 0xF)) continue;
        if (sched != 0xF && mode != 1 && mode < 5) continue;
        for (int rep = 0; rep < 2; ++rep) {
          cudaMemset(d_cyc, 0, 512 * 8);
          if (n_mma == 24) contention_kernel<24><<<sms, 512, smem>>>(d_img, mode, n_gemm, bg, sched, d_cyc, d_sink);
          else contention_kernel<16><<<sms, 512, smem>>>(d_img, mode, n_gemm, bg, sched, d_cyc, d_sink);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
        }
        std::vector<long long> c(512);
        cudaMemcpy(c.data(), d_cyc, 512 * 8, cudaMemcpyDeviceToHost);
        double mean = 0, mx = 0, it = 0;
        for (int i = 0; i < sms; ++i) { mean += (double)c[i]; mx = c[i] > mx ? (double)c[i] : mx; it += (double)c[256 + i]; }
        mean /= sms;
        printf("{\"background\": \"%s\", \"background_warps\": %d, \"schedulers\": %d, \"mmas_per_gemm\": %d, \"cycles_per_gemm_mean\": %.1f, \"cycles_per_gemm_max\": %.1f, "
               "\"cycles_per_mma\": %.2f, \"background_iters_per_warp_per_gemm\": %.2f}\n",
               names[mode], bg, sched, n_mma, mean / n_gemm, mx / n_gemm, mean / n_gemm / n_mma, it / sms / bg / n_gemm);
        fflush(stdout);
      }
  return 0;
}
