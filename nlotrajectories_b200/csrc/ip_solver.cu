// Batched interior-point solver on the device: the CUDA kernels around the bodies of ip_core.cuh, the
// backend that ip_solve_loop drives, and the C ABI (nlo_ip_*).  It is the caller on both sides of the evaluation hot path -
// every iteration is one nlo_nlp_eval, one nlo_nlp_hess, a few value-only evaluations for the line search, and the
// block-tridiagonal factorisation of the condensed KKT matrix - and replaces what the reference delegates to IPOPT
// (core/runner.py:112-133).  Nothing leaves the device between the initial upload of the starts and the final download.
#include "nlo_common.cuh"
#include "nlp_internal.cuh"
#include "ip_core.cuh"
#include "ip_tables.hpp"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace {

constexpr int IP_TEAM = 8;      // warps per team: a thread block is 32 problems (lanes) x IP_TEAM workers (warps)
constexpr int BT_TPB = 64;      // threads per block of the factorisation kernels (shared-memory scratch per thread)

// The team reducer of ip_core.cuh on the device: worker = warp, reductions per lane (= per problem) through shared memory.
// Lanes beyond the batch take no part in any loop (part() is out of every range) and store nothing, but reach every barrier.
struct IpBlockRed {
  double* sm;                    // [IP_TEAM][32]
  int w, lane;
  bool act;
  __device__ __forceinline__ int part() const { return act ? w : (1 << 30); }
  __device__ __forceinline__ int nparts() const { return IP_TEAM; }
  template <class Op>
  __device__ __forceinline__ double reduce(double v, Op op) {
    sm[w * 32 + lane] = v;
    __syncthreads();
    double r = sm[lane];
#pragma unroll
    for (int i = 1; i < IP_TEAM; ++i) r = op(r, sm[i * 32 + lane]);
    __syncthreads();
    return r;
  }
  __device__ __forceinline__ double sum(double v) { return reduce(v, [](double a, double b) { return a + b; }); }
  __device__ __forceinline__ double max(double v) { return reduce(v, [](double a, double b) { return a > b ? a : b; }); }
  __device__ __forceinline__ double min(double v) { return reduce(v, [](double a, double b) { return a < b ? a : b; }); }
  __device__ __forceinline__ bool any(bool v) { return reduce(v ? 1.0 : 0.0, [](double a, double b) { return a > b ? a : b; }) != 0.0; }
  __device__ __forceinline__ void sync() { __syncthreads(); }
};
#define IP_TEAM_SETUP(n_)                                                      \
  __shared__ double red_sm[IP_TEAM * 32];                                      \
  const size_t q_ = (size_t)blockIdx.x * 32 + threadIdx.x;                     \
  IpBlockRed red{red_sm, (int)threadIdx.y, (int)threadIdx.x, q_ < (n_)};       \
  const size_t qc_ = red.act ? q_ : (n_) - 1        /* inactive lanes read (only) a valid column */

__global__ void __launch_bounds__(32 * IP_TEAM) ip_init_kernel(IpTables T, IpState S, IpWork W, size_t P, double mu0, int max_iter) {
  IP_TEAM_SETUP(P);
  ip_init_body(T, S, W, qc_, mu0, max_iter, red);
}
__global__ void __launch_bounds__(32 * IP_TEAM) ip_residual_kernel(IpTables T, IpState S, IpWork W, size_t P, int it, double tol) {
  IP_TEAM_SETUP(P);
  const int done = ip_residual_body(T, S, W, qc_, it, tol, red);
  if (threadIdx.y == 0) {
    const unsigned m = __ballot_sync(0xffffffffu, done && red.act);
    if (threadIdx.x == 0 && m) atomicAdd(W.counters, __popc(m));
  }
}
__global__ void __launch_bounds__(32 * IP_TEAM) ip_step_kernel(IpTables T, IpState S, IpWork W, size_t P) {
  IP_TEAM_SETUP(P);
  ip_step_body(T, S, W, qc_, red);
}
// trial points of the line search, fp32 for the evaluation kernels: column q of the trial batch is problem list[q] (or q itself)
__global__ void __launch_bounds__(256) ip_trial_kernel(IpState S, IpWork W, size_t n, int n_w, const int* __restrict__ list) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const size_t p = list ? (size_t)list[q] : q;
  const double a = W.alpha[p];
  for (int c = blockIdx.y; c < n_w; c += gridDim.y)
    W.wt32[(size_t)c * S.ld + q] = (float)(S.w[(size_t)c * S.ld + p] + a * W.dw[(size_t)c * S.ld + p]);
}
// problems whose trial step is refused are appended to `next` (their order does not matter: every problem is independent)
__global__ void __launch_bounds__(32 * IP_TEAM) ip_merit_kernel(IpTables T, IpState S, IpWork W, size_t n, const int* __restrict__ list,
                                                                int* __restrict__ next) {
  IP_TEAM_SETUP(n);
  const size_t p = list ? (size_t)list[qc_] : qc_;
  const int rej = ip_merit_body(T, S, W, p, qc_, red);
  if (threadIdx.y == 0 && red.act && rej) next[atomicAdd(W.counters + 1, 1)] = (int)p;
}
// ---- the remaining trials of the line search side by side: for every refused problem, nj of the 13 step lengths a, a/2, ... a/2^12
// (a = the halved step the sequential search would try next) are nj columns of one trial batch; ip_ls_select_kernel takes the first
// accepted one.  Two stages: the three largest step lengths first, the other ten only for the problems that refuse all three.
#define IP_LS_LADDER (IP_LS_TRIALS - 1)
__global__ void __launch_bounds__(256) ip_trial_ladder_kernel(IpState S, IpWork W, size_t n, int n_w, const int* __restrict__ list, int j0, int nj) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n * nj) return;
  const size_t p = (size_t)list[q / nj];
  const double a = ldexp(W.alpha[p], -(j0 + (int)(q % nj)));
  for (int c = blockIdx.y; c < n_w; c += gridDim.y)
    W.wt32[(size_t)c * S.ld + q] = (float)(S.w[(size_t)c * S.ld + p] + a * W.dw[(size_t)c * S.ld + p]);
}
__global__ void __launch_bounds__(32 * IP_TEAM) ip_merit_ladder_kernel(IpTables T, IpState S, IpWork W, size_t n, const int* __restrict__ list,
                                                                       int* __restrict__ flags, int j0, int nj) {
  IP_TEAM_SETUP(n * nj);
  const size_t p = (size_t)list[qc_ / nj];
  const double a = ldexp(W.alpha[p], -(j0 + (int)(qc_ % nj)));
  const bool ok = ip_merit_ok(T, S, W, p, qc_, a, false, red);
  if (threadIdx.y == 0 && red.act) flags[qc_] = ok ? 1 : 0;
}
__global__ void __launch_bounds__(256) ip_ls_select_kernel(IpWork W, size_t n, const int* __restrict__ list, const int* __restrict__ flags, int j0,
                                                           int nj, int* __restrict__ next, int* __restrict__ next_count) {
  const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const size_t p = (size_t)list[r];
  int j = 0;
  while (j < nj && !flags[r * nj + j]) ++j;
  if (j < nj) { W.accepted[p] = 1; W.alpha[p] = ldexp(W.alpha[p], -(j0 + j)); }
  else if (next) next[atomicAdd(next_count, 1)] = (int)p;
  else W.alpha[p] = ldexp(W.alpha[p], -IP_LS_LADDER);      // refused to the end: the (tiny) last step is taken, as in the sequential search
}
__global__ void __launch_bounds__(32 * IP_TEAM) ip_update_kernel(IpTables T, IpState S, IpWork W, size_t P) {
  IP_TEAM_SETUP(P);
  ip_update_body(T, S, W, qc_, red);
}
__global__ void __launch_bounds__(32 * IP_TEAM) ip_lsq_prep_kernel(IpTables T, IpState S, IpWork W, size_t P) {
  IP_TEAM_SETUP(P);
  ip_lsq_prep_body(T, S, W, qc_, red);
}
__global__ void __launch_bounds__(32 * IP_TEAM) ip_lsq_choose_kernel(IpTables T, IpState S, IpWork W, size_t P) {
  IP_TEAM_SETUP(P);
  ip_lsq_choose_body(T, S, W, qc_, red);
}
__global__ void __launch_bounds__(32 * IP_TEAM) ip_flush_kernel(IpTables T, IpState S, IpWork W, IpOut O, size_t P, int only_done) {
  IP_TEAM_SETUP(P);
  if (only_done && !S.done[qc_]) red.act = false;      // (per lane; the lane still reaches the team's barriers)
  ip_flush_body(T, S, W, O, qc_, red);
}
__global__ void __launch_bounds__(256) ip_w32_kernel(IpState S, IpWork W, size_t P, int n_w) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  for (int c = blockIdx.y; c < n_w; c += gridDim.y) W.w32[(size_t)c * S.ld + p] = (float)S.w[(size_t)c * S.ld + p];
}
// working-set compaction: column keep[q] of every row -> column q
template <typename V>
__global__ void __launch_bounds__(256) ip_gather_kernel(const V* __restrict__ src, V* __restrict__ dst, int rows, size_t ld,
                                                        const int* __restrict__ keep, size_t n_keep) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_keep) return;
  const size_t from = (size_t)keep[q];
  for (int r = blockIdx.y; r < rows; r += gridDim.y) dst[(size_t)r * ld + q] = src[(size_t)r * ld + from];
}
// [n_in_rows][n_in_cols] (row stride ld_in) -> [n_in_cols][ld_out], with a type conversion
template <typename A, typename B>
__global__ void __launch_bounds__(256) ip_transpose_kernel(const A* __restrict__ in, B* __restrict__ out, size_t n_in_rows, size_t n_in_cols,
                                                           size_t ld_in, size_t ld_out) {
  __shared__ B tile[32][33];
  const size_t tiles_c = (n_in_cols + 31) / 32, tiles_r = (n_in_rows + 31) / 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (size_t tidx = blockIdx.x; tidx < tiles_c * tiles_r; tidx += gridDim.x) {
    const size_t tr = tidx / tiles_c, tc = tidx - tr * tiles_c;
    for (int j = ty; j < 32; j += 8) {
      const size_t r = tr * 32 + j, c = tc * 32 + tx;
      tile[j][tx] = (r < n_in_rows && c < n_in_cols) ? (B)in[r * ld_in + c] : (B)0;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      const size_t c = tc * 32 + j, r = tr * 32 + tx;
      if (c < n_in_cols && r < n_in_rows) out[c * ld_out + r] = tile[tx][j];
    }
    __syncthreads();
  }
}

// ---- block-tridiagonal systems ----------------------------------------------------------------------------------------------
// one thread per (slot, problem): the slot's terms are summed and written once - no atomics, no scratch
__global__ void __launch_bounds__(256) bt_assemble_kernel(BtTables B, const float* __restrict__ jac, const float* __restrict__ hess,
                                                          const double* __restrict__ omega, double* __restrict__ K, size_t P, size_t ld) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int n_slots = B.nb * B.SLK;
  for (int s = blockIdx.y; s < n_slots; s += gridDim.y) K[(size_t)s * ld + p] = bt_slot_value(B, s, jac, hess, omega, ld, p);
}

template <int NS, int NXR>
__global__ void __launch_bounds__(BT_TPB) bt_kkt_kernel(BtTables B, const double* __restrict__ K, double* __restrict__ Lf,
                                                        const double* __restrict__ rhs, double* __restrict__ dw, double* __restrict__ dw_alt,
                                                        size_t P, size_t ld, int n_unknown, const double* __restrict__ delta_in,
                                                        const int* __restrict__ skip, double* __restrict__ dwt_out) {
  extern __shared__ double bt_smem[];
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double dwt;
  bt_kkt_body<NS, NXR>(B, K, Lf, rhs, dw, dw_alt, ld, p, n_unknown, delta_in[p], skip ? skip[p] : 0, &dwt, bt_smem + threadIdx.x, BT_TPB);
  dwt_out[p] = dwt;
}
// plain SPD solve with a per-problem diagonal shift (least-squares multipliers); a failed factorisation keeps `fallback`
template <int NS, int NXR>
__global__ void __launch_bounds__(BT_TPB) bt_spd_kernel(BtTables B, const double* __restrict__ K, double* __restrict__ Lf,
                                                        const double* __restrict__ rhs, double* __restrict__ x, size_t P, size_t ld, int n_unknown,
                                                        const double* __restrict__ shift, const double* __restrict__ fallback) {
  extern __shared__ double bt_smem[];
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  if (!bt_solve_attempt<NS, NXR>(B, K, Lf, rhs, x, ld, p, p, shift[p], HUGE_VAL, bt_smem + threadIdx.x, BT_TPB))
    for (int e = 0; e < n_unknown; ++e) x[(size_t)e * ld + p] = fallback[(size_t)e * ld + p];
}

template <int NS, int NXR>
int launch_bt_kkt(const BtTables& B, const double* K, double* Lf, const double* rhs, double* dw, double* dw_alt, size_t P, size_t ld,
                  int n_unknown, const double* delta_in, const int* skip, double* dwt_out, int device, cudaStream_t st) {
  constexpr size_t smem = (size_t)(2 * (NXR * NS + NS * (NS + 1) / 2) + NS) * BT_TPB * sizeof(double);
  static bool attr[64] = {false};
  if (!attr[device & 63]) {
    NLO_CUDA(cudaFuncSetAttribute(bt_kkt_kernel<NS, NXR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr[device & 63] = true;
  }
  bt_kkt_kernel<NS, NXR><<<(unsigned)((P + BT_TPB - 1) / BT_TPB), BT_TPB, smem, st>>>(B, K, Lf, rhs, dw, dw_alt, P, ld, n_unknown, delta_in, skip, dwt_out);
  NLO_CHECK_LAUNCH();
  return 0;
}
template <int NS, int NXR>
int launch_bt_spd(const BtTables& B, const double* K, double* Lf, const double* rhs, double* x, size_t P, size_t ld, int n_unknown,
                  const double* shift, const double* fallback, int device, cudaStream_t st) {
  constexpr size_t smem = (size_t)(2 * (NXR * NS + NS * (NS + 1) / 2) + NS) * BT_TPB * sizeof(double);
  static bool attr[64] = {false};
  if (!attr[device & 63]) {
    NLO_CUDA(cudaFuncSetAttribute(bt_spd_kernel<NS, NXR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr[device & 63] = true;
  }
  bt_spd_kernel<NS, NXR><<<(unsigned)((P + BT_TPB - 1) / BT_TPB), BT_TPB, smem, st>>>(B, K, Lf, rhs, x, P, ld, n_unknown, shift, fallback);
  NLO_CHECK_LAUNCH();
  return 0;
}

// ---- the same factor + solve with a TEAM per 32 problems ------------------------------------------------------------------------
// lanes = 32 consecutive problems (coalesced global access), warps = the NS rows of a block + one warp for the right-hand side.
// The blocks live in shared memory ([slot][lane]); a Cholesky column costs two block barriers, the triangular solves of the
// sub-diagonal rows and of the right-hand side are row-local, and the Schur update of the next diagonal block is one row per warp.
// One thread per problem (bt_kkt_kernel above) walks the 81 stages alone - ~1 ms per attempt however few problems there are; a team
// shares each stage among NS + 1 warps.
template <int NS, int NXR>
struct BtTeam {
  static constexpr int ND = NS * (NS + 1) / 2, NO = NXR * NS, SLK = ND + NO, SLL = ND + NO + NS;
  static constexpr int S = 0, LD = S + ND, LO = LD + NS, Y = LO + NO, TOTAL = Y + NS;      // shared-memory slots
  static constexpr int NW = NS + 1;
  static constexpr size_t smem_bytes = (size_t)TOTAL * 32 * sizeof(double);
};

// One attempt for the 32 problems of the block with `delta` on the diagonal; lanes with store == false run along without writing.
// Returns (to every thread of a lane) whether the factorisation succeeded and the solution is finite with max |x| < xmax.
template <int NS, int NXR>
__device__ __forceinline__ bool bt_team_attempt(const BtTables& B, const double* __restrict__ K, double* __restrict__ Lf,
                                                const double* __restrict__ rhs, double* __restrict__ x, size_t ld, size_t p, double delta,
                                                double xmax, bool store, double* sm, int* s_flag) {
  using T = BtTeam<NS, NXR>;
  constexpr int ND = T::ND, NO = T::NO, SLK = T::SLK, SLL = T::SLL;
  const int lane = threadIdx.x, w = threadIdx.y;
#define SM(slot) sm[(slot) * 32 + lane]
  const int nb = B.nb;
  bool ok = true;
  // Three register rows per thread, named by the role of its warp (a union: the row warps never touch the right-hand side's values):
  //   row warp w < NS:        r1 = next diagonal block's row (dn)   r2 = sub-diagonal block's row (on)   r3 = its row of L_{k+1,k} (lo)
  //   right-hand-side warp:   r1 = next block's right-hand side (bn) r2 = current right-hand side (b)     r3 = y_k
  double dinv[NS], r1[NS], r2[NS], r3[NS];
  const bool rhs_warp = (w == NS);
#pragma unroll
  for (int i = 0; i < NS; ++i) { r1[i] = 0.0; r2[i] = 0.0; r3[i] = 0.0; }
  __syncthreads();                                           // the previous attempt's readers are done with shared memory
  if (!rhs_warp) {
    const bool real = B.var[w] >= 0;
#pragma unroll
    for (int j = 0; j < NS; ++j)
      if (j <= w) SM(T::S + BT_LI(w, j)) = K[(size_t)BT_LI(w, j) * ld + p] + ((j == w && real) ? delta : 0.0);
  } else {
#pragma unroll
    for (int i = 0; i < NS; ++i) { const int u = B.var[i]; r2[i] = u >= 0 ? rhs[(size_t)u * ld + p] : 0.0; }
  }
  for (int k = 0; k < nb; ++k) {
    // rows of the next stage travel while this block is factorised
    if (k < nb - 1) {
      const double* Kn = K + (size_t)(k + 1) * SLK * ld + p;
      const double* Ok = K + ((size_t)k * SLK + ND) * ld + p;
      if (!rhs_warp) {
#pragma unroll
        for (int j = 0; j < NS; ++j) if (j <= w) r1[j] = Kn[(size_t)BT_LI(w, j) * ld];
        if (w < NXR) {
#pragma unroll
          for (int j = 0; j < NS; ++j) r2[j] = Ok[(size_t)(w * NS + j) * ld];
        }
      } else {
#pragma unroll
        for (int i = 0; i < NS; ++i) { const int u = B.var[(k + 1) * NS + i]; r1[i] = u >= 0 ? rhs[(size_t)u * ld + p] : 0.0; }
      }
    }
    // Cholesky, column by column: every thread of the lane follows the pivots (so `ok` and 1 / pivot are known to all rows)
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      __syncthreads();
      double d = SM(T::S + BT_LI(j, j));
      if (!(d > 0.0) || !ip_finite(d)) { ok = false; d = 1.0; }
      dinv[j] = rsqrt(d);                                  // one dependent sequence instead of sqrt + divide
      double lwj = 0.0;
      if (w == j) SM(T::LD + j) = d * dinv[j];
      if (w > j && w < NS) { lwj = SM(T::S + BT_LI(w, j)) * dinv[j]; SM(T::S + BT_LI(w, j)) = lwj; }
      __syncthreads();
      if (w > j && w < NS) {
#pragma unroll
        for (int m = j + 1; m < NS; ++m) if (m <= w) SM(T::S + BT_LI(w, m)) -= lwj * SM(T::S + BT_LI(m, j));
      }
    }
    __syncthreads();
    const bool st = store && ok;
    double* Lk = Lf + (size_t)k * SLL * ld + p;
    if (!rhs_warp) {
      if (k < nb - 1 && w < NXR) {                           // row w of L_{k+1,k} = O_k L_kk^-T
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          double v = r2[j];
#pragma unroll
          for (int m = 0; m < NS; ++m) if (m < j) v -= r3[m] * SM(T::S + BT_LI(j, m));
          r3[j] = v * dinv[j];
        }
#pragma unroll
        for (int j = 0; j < NS; ++j) { SM(T::LO + w * NS + j) = r3[j]; if (st) Lk[(size_t)(ND + w * NS + j) * ld] = r3[j]; }
      }
      if (st) {                                              // row w of L_kk
#pragma unroll
        for (int j = 0; j < NS; ++j) if (j < w) Lk[(size_t)BT_LI(w, j) * ld] = SM(T::S + BT_LI(w, j));
        Lk[(size_t)BT_LI(w, w) * ld] = SM(T::LD + w);
      }
    } else {                                                 // y = L^-1 b
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        double v = r2[i];
#pragma unroll
        for (int m = 0; m < NS; ++m) if (m < i) v -= SM(T::S + BT_LI(i, m)) * r3[m];
        r3[i] = v * dinv[i];
      }
#pragma unroll
      for (int i = 0; i < NS; ++i) { SM(T::Y + i) = r3[i]; if (st) Lk[(size_t)(ND + NO + i) * ld] = r3[i]; }
    }
    __syncthreads();
    if (k == nb - 1) break;
    if (!rhs_warp) {                                         // row w of the next diagonal block's Schur complement
      const bool real = B.var[(k + 1) * NS + w] >= 0;
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        if (j > w) continue;
        double v = r1[j];
        if (w < NXR) {
#pragma unroll
          for (int m = 0; m < NS; ++m) v -= r3[m] * SM(T::LO + j * NS + m);
        }
        if (j == w && real) v += delta;
        SM(T::S + BT_LI(w, j)) = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        double v = r1[i];
        if (i < NXR) {
#pragma unroll
          for (int m = 0; m < NS; ++m) v -= SM(T::LO + i * NS + m) * r3[m];
        }
        r2[i] = v;
      }
    }
  }
  // backward pass: the factor blocks of stage k come back into shared memory row by row, the right-hand-side warp substitutes
  // (r1 = y_k - L_{k+1,k}^T x_{k+1}, r2 = x_{k+1} then x_k)
  double amax = 0.0;
  bool fin = true;
#pragma unroll
  for (int i = 0; i < NS; ++i) r2[i] = 0.0;
  for (int k = nb - 1; k >= 0; --k) {
    const double* Lk = Lf + (size_t)k * SLL * ld + p;
    if (!rhs_warp) {
#pragma unroll
      for (int j = 0; j < NS; ++j) if (j < w) SM(T::S + BT_LI(w, j)) = Lk[(size_t)BT_LI(w, j) * ld];
      SM(T::LD + w) = Lk[(size_t)BT_LI(w, w) * ld];
      if (k < nb - 1 && w < NXR) {
#pragma unroll
        for (int j = 0; j < NS; ++j) SM(T::LO + w * NS + j) = Lk[(size_t)(ND + w * NS + j) * ld];
      }
    } else {
#pragma unroll
      for (int i = 0; i < NS; ++i) r1[i] = Lk[(size_t)(ND + NO + i) * ld];
    }
    __syncthreads();
    if (rhs_warp) {
      if (k < nb - 1) {
#pragma unroll
        for (int i = 0; i < NXR; ++i)
#pragma unroll
          for (int j = 0; j < NS; ++j) r1[j] -= SM(T::LO + i * NS + j) * r2[i];
      }
#pragma unroll
      for (int i = NS - 1; i >= 0; --i) {
        double v = r1[i];
#pragma unroll
        for (int m = 0; m < NS; ++m) if (m > i) v -= SM(T::S + BT_LI(m, i)) * r2[m];
        r2[i] = v / SM(T::LD + i);
      }
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        const int u = B.var[k * NS + i];
        if (u >= 0) {
          if (store && ok) x[(size_t)u * ld + p] = r2[i];
          if (!ip_finite(r2[i])) fin = false;
          amax = ip_max(amax, fabs(r2[i]));
        }
      }
    }
    __syncthreads();
  }
  if (w == NS) s_flag[lane] = (ok && fin && amax < xmax) ? 1 : 0;
  __syncthreads();
  return s_flag[lane] != 0;
#undef SM
}

// KKT mode: max_attempts 16, bump 1, xmax 1e3, fallback NULL (zero step on failure), delta_out written.
// SPD mode (least-squares multipliers): one attempt with the given shift, no bump, fallback = the vector kept on failure.
template <int NS, int NXR>
__global__ void __launch_bounds__(32 * (NS + 1), 2) bt_team_kernel(BtTables B, const double* __restrict__ K, double* __restrict__ Lf,
                                                                const double* __restrict__ rhs, double* __restrict__ x, double* __restrict__ x_alt,
                                                                size_t P, size_t ld, int n_unknown, const double* __restrict__ delta_in,
                                                                const int* __restrict__ skip, double* __restrict__ delta_out, int max_attempts,
                                                                int bump, double xmax, const double* __restrict__ fallback,
                                                                int* __restrict__ fail_list, int* __restrict__ fail_count) {
  using T = BtTeam<NS, NXR>;
  extern __shared__ double bt_team_smem[];
  __shared__ int s_flag[32];
  const int lane = threadIdx.x, w = threadIdx.y;
  const size_t p_raw = (size_t)blockIdx.x * 32 + lane;
  const bool in_range = p_raw < P;
  const size_t p = in_range ? p_raw : P - 1;                 // lanes beyond the batch shadow the last problem and store nothing
  const bool skipped = in_range && skip && skip[p];
  const double d0 = delta_in[p];
  double dwt = d0;
  int state = (in_range && !skipped) ? 0 : 2;                // 0 to solve, 1 solved, 2 not taking part
  if (skipped)
    for (int c = w; c < n_unknown; c += T::NW) x[(size_t)c * ld + p] = 0.0;
  for (int attempt = 0; attempt < max_attempts; ++attempt) {
    const bool active = state == 0;
    if (!__syncthreads_or(active)) break;
    const bool good = bt_team_attempt<NS, NXR>(B, K, Lf, rhs, x_alt, ld, p, dwt, xmax, active, bt_team_smem, s_flag);
    if (active) {
      if (good) { state = 1; for (int c = w; c < n_unknown; c += T::NW) x[(size_t)c * ld + p] = x_alt[(size_t)c * ld + p]; }
      else dwt = ip_min(ip_max(dwt * 8.0, 1e-4), 1e8);
    }
  }
  if (state == 0) {                                          // every attempt failed
    for (int c = w; c < n_unknown; c += T::NW) x[(size_t)c * ld + p] = fallback ? fallback[(size_t)c * ld + p] : 0.0;
    if (fail_list && w == 0) fail_list[atomicAdd(fail_count, 1)] = (int)p;
  }
  if (bump) {
    // a problem that needed more regularisation than last time is solved once more with twice the value that first passed
    const bool again = in_range && !skipped && dwt > d0;
    if (__syncthreads_or(again)) {
      const bool good = bt_team_attempt<NS, NXR>(B, K, Lf, rhs, x_alt, ld, p, 2.0 * dwt, xmax, again, bt_team_smem, s_flag);
      if (again && good) {
        for (int c = w; c < n_unknown; c += T::NW) x[(size_t)c * ld + p] = x_alt[(size_t)c * ld + p];
        dwt = 2.0 * dwt;
      }
    }
  }
  if (delta_out && w == 0 && in_range) delta_out[p] = dwt;
}

template <int NS, int NXR>
int launch_bt_team(const BtTables& B, const double* K, double* Lf, const double* rhs, double* x, double* x_alt, size_t P, size_t ld,
                   int n_unknown, const double* delta_in, const int* skip, double* delta_out, int max_attempts, int bump, double xmax,
                   const double* fallback, int device, cudaStream_t st, int* fail_list = nullptr, int* fail_count = nullptr) {
  using T = BtTeam<NS, NXR>;
  static bool attr[64] = {false};
  if (!attr[device & 63]) {
    NLO_CUDA(cudaFuncSetAttribute(bt_team_kernel<NS, NXR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::smem_bytes));
    attr[device & 63] = true;
  }
  bt_team_kernel<NS, NXR><<<(unsigned)((P + 31) / 32), dim3(32, T::NW), T::smem_bytes, st>>>(B, K, Lf, rhs, x, x_alt, P, ld, n_unknown, delta_in, skip,
                                                                                          delta_out, max_attempts, bump, xmax, fallback, fail_list, fail_count);
  NLO_CHECK_LAUNCH();
  return 0;
}

// ---- the regularisation search for the problems whose first factorisation failed: all remaining attempts AT ONCE -------------------
// The search of bt_kkt_body is sequential per problem (delta x8 until the Cholesky succeeds, then once more with twice that value),
// and in a batch of thousands some problem needs ten or more attempts in almost every iteration: with the attempts inside one kernel
// the whole batch waited ~1 ms per attempt of its slowest member (72 % of a benchmark_6 x 8,192 solve).  Here every failed problem gets
// 32 lanes: lane a - 1 (a = 1..15) factorises with the delta of attempt a, lane 16 + a - 1 with twice that; bt_select_kernel then takes
// the first success in the order of the sequential search, so the result is the same and the cost is one more attempt, not fifteen.
// A stage of the search covers attempts a0 .. a0 + n_att - 1: 2^lanes_log2 lanes per problem, the first half factorises with the
// delta of those attempts, the second half with twice that.  Most searches end at the first step of the ladder (benchmark_6 x
// 65,536: 76 % at attempt 1, 18 % at 2..5), so stage A tries attempts 1 and 2 with 4 lanes per problem and only what is left gets
// the 32 lanes of stage B (attempts 3..15).
template <int NS, int NXR>
__global__ void __launch_bounds__(BT_TPB) bt_ladder_kernel(BtTables B, const double* __restrict__ K, double* __restrict__ Lf,
                                                           const double* __restrict__ rhs, double* __restrict__ xbuf, size_t ld,
                                                           const int* __restrict__ list, size_t n_list, const double* __restrict__ delta_in,
                                                           int* __restrict__ flags, int lanes_log2, int a0, int n_att) {
  extern __shared__ double bt_smem[];
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (n_list << lanes_log2)) return;
  const size_t prob = (size_t)list[q >> lanes_log2];
  const int half = 1 << (lanes_log2 - 1), i = (int)(q & ((1u << lanes_log2) - 1)), k = i & (half - 1);
  if (k >= n_att) { flags[q] = 0; return; }
  const double d = bt_ladder_delta(delta_in[prob], a0 + k) * (i >= half ? 2.0 : 1.0);
  flags[q] = bt_solve_attempt<NS, NXR>(B, K, Lf, rhs, xbuf, ld, prob, q, d, 1e3, bt_smem + threadIdx.x, BT_TPB) ? 1 : 0;
}
// one warp per problem of the stage: the first successful attempt (and its doubled twin when that succeeded too) -> dw, delta;
// problems without a success go to `next` (the following stage) or, at the last stage, get a zero step
__global__ void __launch_bounds__(256) bt_select_kernel(const int* __restrict__ list, size_t n_list, const int* __restrict__ flags,
                                                        const double* __restrict__ xbuf, double* __restrict__ dw, size_t ld, int n_unknown,
                                                        const double* __restrict__ delta_in, double* __restrict__ delta_out,
                                                        unsigned long long* __restrict__ hist, int lanes_log2, int a0,
                                                        int* __restrict__ next, int* __restrict__ next_count) {
  const size_t wq = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wq >= n_list) return;
  const int LP = 1 << lanes_log2, half = LP >> 1;
  const size_t prob = (size_t)list[wq], q0 = wq << lanes_log2;
  const unsigned m = __ballot_sync(0xffffffffu, lane < LP && flags[q0 + lane] != 0);
  const unsigned first = m & ((1u << half) - 1u);
  const double d0 = delta_in[prob];
  if (!first) {
    if (next) { if (lane == 0) next[atomicAdd(next_count, 1)] = (int)prob; return; }
    // all sixteen attempts failed: no step, delta as the sequential search leaves it
    for (int c = lane; c < n_unknown; c += 32) dw[(size_t)c * ld + prob] = 0.0;
    if (lane == 0) { delta_out[prob] = bt_ladder_delta(d0, 16); atomicAdd(hist + 15, 1ull); }
    return;
  }
  const int i = __ffs(first) - 1;
  if (lane == 0) atomicAdd(hist + (a0 - 1 + i), 1ull);
  const bool twice = (m >> (half + i)) & 1u;
  const size_t src = q0 + (twice ? half + i : i);
  for (int c = lane; c < n_unknown; c += 32) dw[(size_t)c * ld + prob] = xbuf[(size_t)c * ld + src];
  if (lane == 0) delta_out[prob] = bt_ladder_delta(d0, a0 + i) * (twice ? 2.0 : 1.0);
}
template <int NS, int NXR>
int launch_bt_ladder(const BtTables& B, const double* K, double* Lf, const double* rhs, double* xbuf, size_t ld, const int* list, size_t n_list,
                     const double* delta_in, int* flags, int lanes_log2, int a0, int n_att, int device, cudaStream_t st) {
  constexpr size_t smem = (size_t)(2 * (NXR * NS + NS * (NS + 1) / 2) + NS) * BT_TPB * sizeof(double);
  static bool attr[64] = {false};
  if (!attr[device & 63]) {
    NLO_CUDA(cudaFuncSetAttribute(bt_ladder_kernel<NS, NXR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr[device & 63] = true;
  }
  bt_ladder_kernel<NS, NXR><<<(unsigned)(((n_list << lanes_log2) + BT_TPB - 1) / BT_TPB), BT_TPB, smem, st>>>(B, K, Lf, rhs, xbuf, ld, list, n_list,
                                                                                                         delta_in, flags, lanes_log2, a0, n_att);
  NLO_CHECK_LAUNCH();
  return 0;
}
inline bool ip_thread_kkt() { static const bool v = [] { const char* e = getenv("NLO_B200_IP_THREAD_KKT"); return e && *e && *e != '0'; }(); return v; }   // A / B switch: one thread per problem

// block sizes of the six dynamics models (core/dynamics.py:151-158) with and without slack: NS = nx + nu + slack, NXR = nx
#define IP_KKT_SIZES(X) X(5, 3) X(6, 3) X(6, 4) X(7, 4) X(7, 5) X(8, 5) X(9, 7) X(10, 7)
#define IP_LSQ_SIZES(X) X(3, 3) X(4, 4) X(5, 5) X(7, 7)

}  // namespace

namespace {
// Where a solve spends its time: a CUDA event at the start of every phase, read back at the loop's own synchronisation points
// (no extra synchronisation); the interval between two marks is charged to the phase of the first.
enum { PH_EVAL = 0, PH_RESIDUAL, PH_HESSIAN, PH_KKT, PH_STEP, PH_LINESEARCH, PH_UPDATE, PH_MULTIPLIERS, PH_COMPACT, PH_COUNT };
struct PhaseTimer {
  static constexpr int CAP = 96;
  cudaEvent_t ev[CAP]; int id[CAP]; int n = 0; bool ready = false;
  double ms[PH_COUNT] = {0};
  int init() { for (auto& e : ev) NLO_CUDA(cudaEventCreate(&e)); ready = true; return 0; }
  void destroy() { if (ready) for (auto& e : ev) cudaEventDestroy(e); ready = false; }
  void reset() { n = 0; for (double& m : ms) m = 0.0; }
  void mark(int phase, cudaStream_t st) { if (ready && n < CAP) { cudaEventRecord(ev[n], st); id[n++] = phase; } }
  void collect() {                                         // call right after a stream synchronisation that follows the last mark
    if (!ready || n < 2) return;
    for (int i = 0; i + 1 < n; ++i) { float t = 0.f; if (cudaEventElapsedTime(&t, ev[i], ev[i + 1]) == cudaSuccess) ms[id[i]] += t; }
    std::swap(ev[0], ev[n - 1]); id[0] = id[n - 1]; n = 1;
  }
};

}  // namespace

// ---- the solver object ----------------------------------------------------------------------------------------------------
struct nlo_ip {
  nlo_nlp* nlp;
  int device;
  size_t cap;                        // problems the buffers hold == leading dimension of every array
  cudaStream_t st;
  IpHostTables HT;
  BtHost kkt_h, lsq_h;
  int* d_itab; double* d_dtab;
  IpTables T; BtTables KB, LB;
  size_t state_rows;
  double* d_state[2]; int* d_istate[2]; int cur;
  float* d_f32; double* d_f64; int* d_i32;
  double *d_K, *d_L;
  double* d_out; int* d_iout;
  int* d_keep;
  int* h_pin;                        // pinned: counters[2] + done flags
  IpState S; IpWork W; IpOut O;
  IpOptions opt;
  PhaseTimer timer;
  unsigned long long* d_hist;        // [16] how many regularisation searches ended at attempt 1..15 (slot 15: none succeeded)
  long long kkt_problems, kkt_retries;
};

namespace {

void ip_bind_state(nlo_ip* s, int which) {
  const IpTables& T = s->T;
  const size_t ld = s->cap;
  double* b = s->d_state[which];
  IpState& S = s->S;
  S.ld = ld;
  S.w = b; b += (size_t)T.n_w * ld;
  S.s = b; b += (size_t)T.nI * ld;
  S.zl = b; b += (size_t)T.nI * ld;
  S.zu = b; b += (size_t)T.nI * ld;
  S.lamE = b; b += (size_t)T.nE * ld;
  S.mu = b; b += ld; S.nu = b; b += ld; S.delta_w = b; b += ld; S.err0 = b; b += ld; S.f_mark = b;
  int* ib = s->d_istate[which];
  S.iters = ib; S.stalled = ib + ld; S.orig = ib + 2 * ld; S.done = ib + 3 * ld;
  s->cur = which;
}

template <typename V>
int upload(V** dst, const std::vector<V>& v) {
  NLO_CUDA(cudaMalloc(dst, std::max<size_t>(v.size(), 1) * sizeof(V)));
  if (!v.empty()) NLO_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(V), cudaMemcpyHostToDevice));
  return 0;
}

// Regularised Newton step for P problems whose matrix blocks are assembled in s->d_K: one attempt for everybody (team kernel), then
// every remaining attempt of the failed ones side by side (ladder + select kernels).  NLO_B200_IP_THREAD_KKT=1 selects the sequential
// search inside one kernel (one thread per problem) instead - same results, for A / B timing.
template <int NS, int NXR>
int ip_kkt_device(nlo_ip* s, const double* rhs, double* dw, size_t P, size_t ld, const double* delta_in, const int* skip, double* delta_out,
                  cudaStream_t st) {
  const BtTables& B = s->KB;
  if (ip_thread_kkt() || ld < 32)
    return launch_bt_kkt<NS, NXR>(B, s->d_K, s->d_L, rhs, dw, s->W.dw_alt, P, ld, s->T.n_w, delta_in, skip, delta_out, s->device, st);
  int* fail_count = s->W.counters + 2;
  int* fail_list = s->W.ls_list[0];
  int* flags = s->W.ls_list[1];
  NLO_CUDA(cudaMemsetAsync(fail_count, 0, sizeof(int), st));
  if (launch_bt_team<NS, NXR>(B, s->d_K, s->d_L, rhs, dw, s->W.dw_alt, P, ld, s->T.n_w, delta_in, skip, delta_out, 1, 0, 1e3, nullptr, s->device, st,
                              fail_list, fail_count)) return 1;
  NLO_CUDA(cudaMemcpyAsync(s->h_pin, s->W.counters, 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
  NLO_CUDA(cudaStreamSynchronize(st));
  const size_t n_fail = (size_t)s->h_pin[2];
  s->kkt_problems += (long long)P; s->kkt_retries += (long long)n_fail;
  if (n_fail == 0) return 0;
  // stage A: attempts 1 and 2 (+ their doubled twins), 4 lanes per problem; what is left goes to the list of stage B
  int* list_b = s->d_keep;                                   // (the compaction's index buffer is free during an iteration)
  int* count_b = s->W.counters + 3;
  NLO_CUDA(cudaMemsetAsync(count_b, 0, sizeof(int), st));
  for (size_t off = 0, chunk = ld / 4; off < n_fail; off += chunk) {
    const size_t n = std::min(chunk, n_fail - off);
    if (launch_bt_ladder<NS, NXR>(B, s->d_K, s->d_L, rhs, s->W.dw_alt, ld, fail_list + off, n, delta_in, flags, 2, 1, 2, s->device, st)) return 1;
    bt_select_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(fail_list + off, n, flags, s->W.dw_alt, dw, ld, s->T.n_w, delta_in, delta_out,
                                                                       s->d_hist, 2, 1, list_b, count_b);
    NLO_CHECK_LAUNCH();
  }
  NLO_CUDA(cudaMemcpyAsync(s->h_pin, s->W.counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
  NLO_CUDA(cudaStreamSynchronize(st));
  const size_t n_b = (size_t)s->h_pin[3];
  // stage B: attempts 3..15, 32 lanes per problem
  for (size_t off = 0, chunk = ld / 32; off < n_b; off += chunk) {
    const size_t n = std::min(chunk, n_b - off);
    if (launch_bt_ladder<NS, NXR>(B, s->d_K, s->d_L, rhs, s->W.dw_alt, ld, list_b + off, n, delta_in, flags, 5, 3, 13, s->device, st)) return 1;
    bt_select_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(list_b + off, n, flags, s->W.dw_alt, dw, ld, s->T.n_w, delta_in, delta_out,
                                                                       s->d_hist, 5, 3, nullptr, nullptr);
    NLO_CHECK_LAUNCH();
  }
  return 0;
}

// the backend ip_solve_loop drives: every method enqueues kernels on the solver's stream; the few that return a count synchronise
struct GpuBackend {
  nlo_ip* s;
  cudaStream_t st;
  PhaseTimer* tm;
  dim3 g1(size_t P) const { return dim3((unsigned)((P + 31) / 32)); }
  dim3 b1() const { return dim3(32, IP_TEAM); }

  int read_counter(int which, size_t* out) {
    NLO_CUDA(cudaMemcpyAsync(s->h_pin, s->W.counters, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    NLO_CUDA(cudaStreamSynchronize(st));
    *out = (size_t)s->h_pin[which];
    return 0;
  }
  int eval_full(size_t P) {
    tm->mark(PH_EVAL, st);
    return nlo_nlp_eval(s->nlp, s->W.w32, P, s->cap, s->W.g, s->W.jac, s->W.f, s->W.grad, st);
  }
  int init(size_t P, double mu0, int max_iter) {
    ip_init_kernel<<<g1(P), b1(), 0, st>>>(s->T, s->S, s->W, P, mu0, max_iter);
    NLO_CHECK_LAUNCH();
    return 0;
  }
  int residual(size_t P, int it, double tol, size_t* n_done) {
    tm->mark(PH_RESIDUAL, st);
    NLO_CUDA(cudaMemsetAsync(s->W.counters, 0, 2 * sizeof(int), st));
    ip_residual_kernel<<<g1(P), b1(), 0, st>>>(s->T, s->S, s->W, P, it, tol);
    NLO_CHECK_LAUNCH();
    tm->mark(PH_HESSIAN, st);                              // (closes the residual interval; the Hessian is what follows)
    if (read_counter(0, n_done)) return 1;
    tm->collect();
    return 0;
  }
  int hessian(size_t P) { tm->mark(PH_HESSIAN, st); return nlo_nlp_hess(s->nlp, s->W.w32, nullptr, s->W.lam32, P, s->cap, s->W.hess, st); }
  int assemble(const BtTables& B, size_t P) {
    const int n_slots = B.nb * B.SLK;
    bt_assemble_kernel<<<dim3((unsigned)((P + 255) / 256), (unsigned)std::min(n_slots, 65535)), 256, 0, st>>>(B, s->W.jac, s->W.hess, s->W.omega, s->d_K, P, s->cap);
    NLO_CHECK_LAUNCH();
    return 0;
  }
  int kkt_solve(size_t P) {
    tm->mark(PH_KKT, st);
    if (assemble(s->KB, P)) return 1;
#define IP_CASE(NS_, NXR_)                                                                                                      \
    if (s->KB.NS == NS_ && s->KB.NXR == NXR_)                                                                                   \
      return ip_kkt_device<NS_, NXR_>(s, s->W.rhs, s->W.dw, P, s->cap, s->S.delta_w, s->S.done, s->W.dwt, st);
    IP_KKT_SIZES(IP_CASE)
#undef IP_CASE
    return nlo_fail("interior point: no factorisation kernel for stage blocks of %d unknowns (%d states)", s->KB.NS, s->KB.NXR);
  }
  int step(size_t P) {
    tm->mark(PH_STEP, st);
    ip_step_kernel<<<g1(P), b1(), 0, st>>>(s->T, s->S, s->W, P);
    NLO_CHECK_LAUNCH();
    return 0;
  }
  int trial(size_t n, int ls, size_t* rejected) {
    tm->mark(PH_LINESEARCH, st);
    const int* list = ls > 0 ? s->W.ls_list[(ls + 1) & 1] : nullptr;
    ip_trial_kernel<<<dim3((unsigned)((n + 255) / 256), 32), 256, 0, st>>>(s->S, s->W, n, s->T.n_w, list);
    NLO_CHECK_LAUNCH();
    if (nlo_nlp_eval(s->nlp, s->W.wt32, n, s->cap, s->W.gt, nullptr, s->W.ft, nullptr, st)) return 1;
    NLO_CUDA(cudaMemsetAsync(s->W.counters + 1, 0, sizeof(int), st));
    ip_merit_kernel<<<g1(n), b1(), 0, st>>>(s->T, s->S, s->W, n, list, s->W.ls_list[ls & 1]);
    NLO_CHECK_LAUNCH();
    return read_counter(1, rejected);
  }
  int line_search(size_t P, IpStats* stats) {
    static const bool sequential = [] { const char* e = getenv("NLO_B200_IP_SEQ_LS"); return e && *e && *e != '0'; }();
    if (sequential || s->cap < 2 * IP_LS_LADDER) return ip_line_search_sequential(*this, P, stats);
    size_t rejected = 0;
    if (trial(P, 0, &rejected)) return 1;                    // refused problems: ls_list[0], step length already halved
    ++stats->trials; stats->trial_problems += (long long)P;
    if (rejected == 0) return 0;
    int* flags = s->W.ls_list[1];
    int* count_b = s->W.counters + 3;
    const int* list = s->W.ls_list[0];
    size_t n_list = rejected;
    const int stage_j0[2] = {0, 3}, stage_nj[2] = {3, IP_LS_LADDER - 3};
    for (int stage = 0; stage < 2 && n_list > 0; ++stage) {
      const int j0 = stage_j0[stage], nj = stage_nj[stage];
      int* next = stage == 0 ? s->d_keep : nullptr;           // (the compaction's index buffer is free during an iteration)
      if (next) NLO_CUDA(cudaMemsetAsync(count_b, 0, sizeof(int), st));
      const size_t chunk = s->cap / nj;
      for (size_t off = 0; off < n_list; off += chunk) {
        const size_t n = std::min(chunk, n_list - off), cols = n * nj;
        tm->mark(PH_LINESEARCH, st);
        ip_trial_ladder_kernel<<<dim3((unsigned)((cols + 255) / 256), 32), 256, 0, st>>>(s->S, s->W, n, s->T.n_w, list + off, j0, nj);
        NLO_CHECK_LAUNCH();
        if (nlo_nlp_eval(s->nlp, s->W.wt32, cols, s->cap, s->W.gt, nullptr, s->W.ft, nullptr, st)) return 1;
        ip_merit_ladder_kernel<<<g1(cols), b1(), 0, st>>>(s->T, s->S, s->W, n, list + off, flags, j0, nj);
        NLO_CHECK_LAUNCH();
        ip_ls_select_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s->W, n, list + off, flags, j0, nj, next, count_b);
        NLO_CHECK_LAUNCH();
        ++stats->trials; stats->trial_problems += (long long)cols;
      }
      if (!next) break;
      NLO_CUDA(cudaMemcpyAsync(s->h_pin, s->W.counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
      NLO_CUDA(cudaStreamSynchronize(st));
      n_list = (size_t)s->h_pin[3];
      list = s->d_keep;
    }
    return 0;
  }
  int update(size_t P) {
    tm->mark(PH_UPDATE, st);
    ip_update_kernel<<<g1(P), b1(), 0, st>>>(s->T, s->S, s->W, P);
    NLO_CHECK_LAUNCH();
    return 0;
  }
  int lsq_multipliers(size_t P) {
    if (s->T.nE == 0) return 0;
    tm->mark(PH_MULTIPLIERS, st);
    ip_lsq_prep_kernel<<<g1(P), b1(), 0, st>>>(s->T, s->S, s->W, P);
    NLO_CHECK_LAUNCH();
    if (assemble(s->LB, P)) return 1;
    int rc = -1;
#define IP_CASE(NS_, NXR_)                                                                                                      \
    if (rc < 0 && s->LB.NS == NS_)                                                                                              \
      rc = ip_thread_kkt() ? launch_bt_spd<NS_, NXR_>(s->LB, s->d_K, s->d_L, s->W.v, s->W.lam_ls, P, s->cap, s->T.nE, s->W.eps_ls, s->S.lamE, s->device, st) \
                           : launch_bt_team<NS_, NXR_>(s->LB, s->d_K, s->d_L, s->W.v, s->W.lam_ls, s->W.dw_alt, P, s->cap, s->T.nE, s->W.eps_ls, nullptr,  \
                                                       nullptr, 1, 0, HUGE_VAL, s->S.lamE, s->device, st);
    IP_LSQ_SIZES(IP_CASE)
#undef IP_CASE
    if (rc < 0) return nlo_fail("interior point: no multiplier kernel for %d states", s->LB.NS);
    if (rc) return rc;
    ip_lsq_choose_kernel<<<g1(P), b1(), 0, st>>>(s->T, s->S, s->W, P);
    NLO_CHECK_LAUNCH();
    return 0;
  }
  int flush(size_t P, int only_done) {
    ip_flush_kernel<<<g1(P), b1(), 0, st>>>(s->T, s->S, s->W, s->O, P, only_done);
    NLO_CHECK_LAUNCH();
    return 0;
  }
  int flush_all(size_t P) { return flush(P, 0); }
  int compact(size_t P, size_t* newP) {
    tm->mark(PH_COMPACT, st);
    if (flush(P, 1)) return 1;
    int* h_done = s->h_pin + 2;
    NLO_CUDA(cudaMemcpyAsync(h_done, s->S.done, P * sizeof(int), cudaMemcpyDeviceToHost, st));
    NLO_CUDA(cudaStreamSynchronize(st));
    std::vector<int> keep;
    keep.reserve(P);
    for (size_t p = 0; p < P; ++p) if (!h_done[p]) keep.push_back((int)p);
    *newP = keep.size();
    if (keep.empty()) return 0;
    NLO_CUDA(cudaMemcpyAsync(s->d_keep, keep.data(), keep.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    const int other = 1 - s->cur;
    const unsigned gx = (unsigned)((keep.size() + 255) / 256);
    ip_gather_kernel<double><<<dim3(gx, 64), 256, 0, st>>>(s->d_state[s->cur], s->d_state[other], (int)s->state_rows, s->cap, s->d_keep, keep.size());
    NLO_CHECK_LAUNCH();
    ip_gather_kernel<int><<<dim3(gx, 4), 256, 0, st>>>(s->d_istate[s->cur], s->d_istate[other], 4, s->cap, s->d_keep, keep.size());
    NLO_CHECK_LAUNCH();
    NLO_CUDA(cudaStreamSynchronize(st));          // `keep` (pageable) must outlive the copy
    ip_bind_state(s, other);
    ip_w32_kernel<<<dim3(gx, 32), 256, 0, st>>>(s->S, s->W, keep.size(), s->T.n_w);
    NLO_CHECK_LAUNCH();
    return 0;
  }
  void report(size_t P, int it, size_t n_done) { fprintf(stderr, "[nlo_ip] it %3d  active %zu  done %zu\n", it, P, n_done); }
};

}  // namespace

extern "C" {

void nlo_ip_destroy(nlo_ip* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  void* bufs[] = {s->d_itab, s->d_dtab, s->d_state[0], s->d_state[1], s->d_istate[0], s->d_istate[1], s->d_f32, s->d_f64, s->d_i32,
                  s->d_K, s->d_L, s->d_out, s->d_iout, s->d_keep, s->d_hist};
  for (void* b : bufs) if (b) cudaFree(b);
  if (s->h_pin) cudaFreeHost(s->h_pin);
  s->timer.destroy();
  if (s->st) cudaStreamDestroy(s->st);
  delete s;
}

int nlo_ip_create(nlo_nlp* p, const double* lbg, const double* ubg, size_t max_problems, nlo_ip** out) {
  if (!out) return nlo_fail("null out");
  *out = nullptr;
  if (!p || !lbg || !ubg) return nlo_fail("null argument");
  if (max_problems == 0) return nlo_fail("max_problems must be positive");
  NLO_CUDA(cudaSetDevice(p->device));
  nlo_ip* s = new (std::nothrow) nlo_ip();
  if (!s) return nlo_fail("out of host memory");
  s->nlp = p; s->device = p->device; s->cap = (max_problems + 31) / 32 * 32;
  s->d_itab = nullptr; s->d_dtab = nullptr; s->d_state[0] = s->d_state[1] = nullptr; s->d_istate[0] = s->d_istate[1] = nullptr;
  s->d_f32 = nullptr; s->d_f64 = nullptr; s->d_i32 = nullptr; s->d_K = s->d_L = nullptr; s->d_out = nullptr; s->d_iout = nullptr;
  s->d_keep = nullptr; s->h_pin = nullptr; s->st = nullptr; s->d_hist = nullptr; s->kkt_problems = s->kkt_retries = 0;
  const NlpDev& L = p->L;
  {
    std::vector<int> jc(L.n_w + 1, 0), hc(L.n_w + 1, 0);
    for (int z = 0; z < L.nnz; ++z) jc[p->cols_ccs[z] + 1]++;
    for (size_t z = 0; z < p->hcols_ccs.size(); ++z) hc[p->hcols_ccs[z] + 1]++;
    for (int c = 0; c < L.n_w; ++c) { jc[c + 1] += jc[c]; hc[c + 1] += hc[c]; }
    ip_build_tables(L.n_w, L.n_g, jc.data(), p->rows_ccs.data(), (int)p->hrows_ccs.size(), hc.data(), p->hrows_ccs.data(), lbg, ubg, &s->HT);
  }
  IpStages stg = {L.N, L.nx, L.nu, L.use_slack, L.n_term, L.g_off_dyn};
  if (!ip_build_kkt_system(s->HT, stg, &s->kkt_h)) { std::string e = s->kkt_h.error; nlo_ip_destroy(s); return nlo_fail("interior point: %s", e.c_str()); }
  if (s->HT.nE > 0 && !ip_build_lsq_system(s->HT, stg, &s->lsq_h)) { std::string e = s->lsq_h.error; nlo_ip_destroy(s); return nlo_fail("interior point: %s", e.c_str()); }
  // ---- device tables: one int blob, one double blob ----
  const IpHostTables& H = s->HT;
  std::vector<int> it;
  auto push = [&](const std::vector<int>& v) { const size_t off = it.size(); it.insert(it.end(), v.begin(), v.end()); return off; };
  const size_t o_rkind = push(H.rkind), o_ridx = push(H.ridx), o_colind = push(H.colind), o_row = push(H.row), o_rptr = push(H.rptr),
               o_rnz = push(H.rnz), o_rcol = push(H.rcol), o_hcolind = push(H.hcolind), o_hrow = push(H.hrow),
               o_kvar = push(s->kkt_h.var), o_kptr = push(s->kkt_h.term_ptr), o_kterms = push(s->kkt_h.terms),
               o_lvar = push(s->lsq_h.var), o_lptr = push(s->lsq_h.term_ptr), o_lterms = push(s->lsq_h.terms);
  std::vector<double> dt(H.lb);
  dt.insert(dt.end(), H.ub.begin(), H.ub.end());
  if (upload(&s->d_itab, it) || upload(&s->d_dtab, dt)) { nlo_ip_destroy(s); return 1; }
  IpTables& T = s->T;
  T.n_w = H.n_w; T.n_g = H.n_g; T.nnz = H.nnz; T.nnzh = H.nnzh; T.nE = H.nE; T.nI = H.nI;
  T.rkind = s->d_itab + o_rkind; T.ridx = s->d_itab + o_ridx; T.colind = s->d_itab + o_colind; T.row = s->d_itab + o_row;
  T.rptr = s->d_itab + o_rptr; T.rnz = s->d_itab + o_rnz; T.rcol = s->d_itab + o_rcol; T.hcolind = s->d_itab + o_hcolind; T.hrow = s->d_itab + o_hrow;
  T.lb = s->d_dtab; T.ub = s->d_dtab + H.n_g;
  auto bind_bt = [&](BtTables& B, const BtHost& h, size_t ov, size_t op, size_t ot) {
    B.nb = h.nb; B.NS = h.NS; B.NXR = h.NXR; B.SLK = h.SLK; B.SLL = h.SLL;
    B.var = s->d_itab + ov; B.term_ptr = s->d_itab + op; B.terms = s->d_itab + ot;
  };
  bind_bt(s->KB, s->kkt_h, o_kvar, o_kptr, o_kterms);
  bind_bt(s->LB, s->lsq_h, o_lvar, o_lptr, o_lterms);
  // ---- buffers ----
  const size_t ld = s->cap;
  s->state_rows = (size_t)T.n_w + 3 * (size_t)T.nI + T.nE + 5;
  const size_t f32_rows = 1 + (size_t)T.n_w + T.n_g + T.nnz + T.nnzh + T.n_w + T.n_g + T.n_w + 1 + T.n_g;
  const size_t f64_rows = (size_t)T.n_g * 2 + T.n_w * 3 + T.nI * 3 + T.nE * 2 + 6;
  const size_t k_slots = std::max((size_t)s->kkt_h.nb * s->kkt_h.SLK, (size_t)s->lsq_h.nb * s->lsq_h.SLK);
  const size_t l_slots = std::max((size_t)s->kkt_h.nb * s->kkt_h.SLL, (size_t)s->lsq_h.nb * s->lsq_h.SLL);
  bool ok = true;
  for (int b = 0; b < 2 && ok; ++b)
    ok = cudaMalloc(&s->d_state[b], s->state_rows * ld * sizeof(double)) == cudaSuccess && cudaMalloc(&s->d_istate[b], 4 * ld * sizeof(int)) == cudaSuccess;
  ok = ok && cudaMalloc(&s->d_f32, f32_rows * ld * sizeof(float)) == cudaSuccess && cudaMalloc(&s->d_f64, f64_rows * ld * sizeof(double)) == cudaSuccess &&
       cudaMalloc(&s->d_i32, (3 * ld + 8) * sizeof(int)) == cudaSuccess && cudaMalloc(&s->d_K, k_slots * ld * sizeof(double)) == cudaSuccess &&
       cudaMalloc(&s->d_L, l_slots * ld * sizeof(double)) == cudaSuccess &&
       cudaMalloc(&s->d_out, ((size_t)T.n_w + T.n_g + 3) * ld * sizeof(double)) == cudaSuccess && cudaMalloc(&s->d_iout, 2 * ld * sizeof(int)) == cudaSuccess &&
       cudaMalloc(&s->d_keep, ld * sizeof(int)) == cudaSuccess && cudaMalloc(&s->d_hist, 16 * sizeof(unsigned long long)) == cudaSuccess && cudaHostAlloc(&s->h_pin, (ld + 2) * sizeof(int), cudaHostAllocDefault) == cudaSuccess &&
       cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking) == cudaSuccess;
  if (!ok) {
    const double gb = ((2 * s->state_rows + f64_rows + k_slots + l_slots + T.n_w + T.n_g + 3) * 8.0 + f32_rows * 4.0) * ld / 1e9;
    nlo_ip_destroy(s);
    return nlo_fail("interior point: device allocation failed (%.1f GB for %zu problems)", gb, max_problems);
  }
  IpWork& W = s->W;
  W.ld = ld;
  float* f = s->d_f32;
  W.f = f; f += ld; W.grad = f; f += (size_t)T.n_w * ld; W.g = f; f += (size_t)T.n_g * ld; W.jac = f; f += (size_t)T.nnz * ld;
  W.hess = f; f += (size_t)T.nnzh * ld; W.w32 = f; f += (size_t)T.n_w * ld; W.lam32 = f; f += (size_t)T.n_g * ld;
  W.wt32 = f; f += (size_t)T.n_w * ld; W.ft = f; f += ld; W.gt = f;
  double* d = s->d_f64;
  W.omega = d; d += (size_t)T.n_g * ld; W.v = d; d += (size_t)T.n_g * ld; W.rhs = d; d += (size_t)T.n_w * ld; W.dw = d; d += (size_t)T.n_w * ld;
  W.dw_alt = d; d += (size_t)T.n_w * ld; W.ds = d; d += (size_t)T.nI * ld; W.dzl = d; d += (size_t)T.nI * ld; W.dzu = d; d += (size_t)T.nI * ld;
  W.dlamE = d; d += (size_t)T.nE * ld; W.lam_ls = d; d += (size_t)T.nE * ld;
  W.dwt = d; d += ld; W.alpha = d; d += ld; W.alpha_d = d; d += ld; W.phi0 = d; d += ld; W.dphi = d; d += ld; W.eps_ls = d; W.viol = nullptr;
  W.accepted = s->d_i32; W.ls_list[0] = s->d_i32 + ld; W.ls_list[1] = s->d_i32 + 2 * ld; W.counters = s->d_i32 + 3 * ld;
  IpOut& O = s->O;
  O.ld = ld; O.w = s->d_out; O.lam = O.w + (size_t)T.n_w * ld; O.f = O.lam + (size_t)T.n_g * ld; O.viol = O.f + ld; O.err = O.viol + ld;
  O.iters = s->d_iout; O.status = s->d_iout + ld;
  ip_bind_state(s, 0);
  if (nlo_nlp_reserve(p, s->cap)) { nlo_ip_destroy(s); return 1; }
  *out = s;
  return 0;
}

size_t nlo_ip_capacity(const nlo_ip* s) { return s ? s->cap : 0; }

int nlo_ip_solve(nlo_ip* s, const double* w0_host, size_t P, const nlo_ip_options* opt_in, double* w_host, double* f_host, double* viol_host,
                 double* kkt_err_host, int* iters_host, int* status_host, double* lam_host, nlo_ip_stats* stats_out) {
  if (!s || !w0_host) return nlo_fail("null argument");
  if (P == 0) return 0;
  if (P > s->cap) return nlo_fail("interior point: %zu problems exceed the capacity %zu given to nlo_ip_create", P, s->cap);
  NLO_CUDA(cudaSetDevice(s->device));
  IpOptions opt = {1e-4, 300, 0.1, 1, 1, 0};
  if (opt_in) { opt.tol = opt_in->tol; opt.max_iter = opt_in->max_iter; opt.mu0 = opt_in->mu0; opt.ls_multipliers = opt_in->ls_multipliers;
                opt.compact = opt_in->compact; opt.verbose = opt_in->verbose; }
  if (!(opt.tol > 0.0) || opt.max_iter < 1 || !(opt.mu0 > 0.0)) return nlo_fail("interior point: bad options");
  const IpTables& T = s->T;
  const size_t ld = s->cap;
  cudaStream_t st = s->st;
  ip_bind_state(s, 0);
  // starts: problem-major doubles on the host -> variable-major doubles + floats on the device (staged through the K buffer)
  NLO_CUDA(cudaMemcpyAsync(s->d_K, w0_host, P * (size_t)T.n_w * sizeof(double), cudaMemcpyHostToDevice, st));
  {
    const size_t tiles = ((size_t)(T.n_w + 31) / 32) * ((P + 31) / 32);
    ip_transpose_kernel<double, double><<<(unsigned)std::min<size_t>(tiles, 148 * 16), 256, 0, st>>>(s->d_K, s->S.w, P, (size_t)T.n_w, (size_t)T.n_w, ld);
    NLO_CHECK_LAUNCH();
    ip_w32_kernel<<<dim3((unsigned)((P + 255) / 256), 32), 256, 0, st>>>(s->S, s->W, P, T.n_w);
    NLO_CHECK_LAUNCH();
    std::vector<int> orig(P);
    for (size_t p = 0; p < P; ++p) orig[p] = (int)p;
    NLO_CUDA(cudaMemcpyAsync(s->S.orig, orig.data(), P * sizeof(int), cudaMemcpyHostToDevice, st));
    NLO_CUDA(cudaStreamSynchronize(st));
  }
  if (!s->timer.ready && s->timer.init()) return 1;
  s->timer.reset();
  s->kkt_problems = s->kkt_retries = 0;
  NLO_CUDA(cudaMemsetAsync(s->d_hist, 0, 16 * sizeof(unsigned long long), st));
  GpuBackend x{s, st, &s->timer};
  IpStats stats;
  if (ip_solve_loop(x, P, opt, &stats)) return 1;
  s->timer.mark(PH_COMPACT, st);
  NLO_CUDA(cudaStreamSynchronize(st));
  s->timer.collect();
  // results: variable-major on the device -> problem-major on the host (staged through the K buffer)
  auto download = [&](const double* soa, int rows, double* host) -> int {
    if (!host) return 0;
    const size_t tiles = ((size_t)(rows + 31) / 32) * ((P + 31) / 32);
    ip_transpose_kernel<double, double><<<(unsigned)std::min<size_t>(tiles, 148 * 16), 256, 0, st>>>(soa, s->d_K, (size_t)rows, P, ld, (size_t)rows);
    NLO_CHECK_LAUNCH();
    NLO_CUDA(cudaMemcpyAsync(host, s->d_K, P * (size_t)rows * sizeof(double), cudaMemcpyDeviceToHost, st));
    NLO_CUDA(cudaStreamSynchronize(st));
    return 0;
  };
  if (download(s->O.w, T.n_w, w_host) || download(s->O.lam, T.n_g, lam_host)) return 1;
  if (f_host) NLO_CUDA(cudaMemcpyAsync(f_host, s->O.f, P * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (viol_host) NLO_CUDA(cudaMemcpyAsync(viol_host, s->O.viol, P * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (kkt_err_host) NLO_CUDA(cudaMemcpyAsync(kkt_err_host, s->O.err, P * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (iters_host) NLO_CUDA(cudaMemcpyAsync(iters_host, s->O.iters, P * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (status_host) NLO_CUDA(cudaMemcpyAsync(status_host, s->O.status, P * sizeof(int), cudaMemcpyDeviceToHost, st));
  NLO_CUDA(cudaStreamSynchronize(st));
  if (stats_out) {
    stats_out->iterations = stats.iterations; stats_out->evaluations = stats.evaluations; stats_out->hessians = stats.hessians;
    stats_out->trials = stats.trials; stats_out->compactions = stats.compactions; stats_out->trial_problems = stats.trial_problems;
    for (int i = 0; i < PH_COUNT; ++i) stats_out->phase_ms[i] = s->timer.ms[i];
    stats_out->kkt_problems = s->kkt_problems; stats_out->kkt_retries = s->kkt_retries;
    unsigned long long hist[16];
    NLO_CUDA(cudaMemcpy(hist, s->d_hist, sizeof(hist), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 16; ++i) stats_out->kkt_retry_hist[i] = (long long)hist[i];
  }
  return 0;
}

// One regularised Newton step of the condensed KKT system for inputs that already live on the device (all variable-major, leading
// dimension ld <= capacity):  (H + J^T diag(omega) J + delta I) dw = rhs  per problem, delta chosen by the inertia test.
int nlo_ip_kkt_step(nlo_ip* s, const float* jac, const float* hess, const double* omega, const double* rhs, const double* delta_in, size_t P,
                    size_t ld, double* dw, double* delta_out, void* stream) {
  if (!s || !jac || !hess || !omega || !rhs || !delta_in || !dw || !delta_out) return nlo_fail("null argument");
  if (P == 0) return 0;
  if (ld < P || ld > s->cap) return nlo_fail("interior point: ld (%zu) must lie in [P, capacity %zu]", ld, s->cap);
  NLO_CUDA(cudaSetDevice(s->device));
  cudaStream_t st = (cudaStream_t)stream;
  const BtTables& B = s->KB;
  const int n_slots = B.nb * B.SLK;
  bt_assemble_kernel<<<dim3((unsigned)((P + 255) / 256), (unsigned)std::min(n_slots, 65535)), 256, 0, st>>>(B, jac, hess, omega, s->d_K, P, ld);
  NLO_CHECK_LAUNCH();
#define IP_CASE(NS_, NXR_)                                                                                                      \
  if (B.NS == NS_ && B.NXR == NXR_) return ip_kkt_device<NS_, NXR_>(s, rhs, dw, P, ld, delta_in, nullptr, delta_out, st);
  IP_KKT_SIZES(IP_CASE)
#undef IP_CASE
  return nlo_fail("interior point: no factorisation kernel for stage blocks of %d unknowns (%d states)", B.NS, B.NXR);
}

}  // extern "C"
