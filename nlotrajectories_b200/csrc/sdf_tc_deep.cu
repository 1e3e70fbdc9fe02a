// Host side of the deep tensor-tile path (sdf_tc_deep.cuh): operand images and scales of networks with two or three H x H
// matrices, and the routing to the per-shape translation units.
#include "sdf_tc_deep.cuh"

int nlo_sdf_tc_deep_launch_h128m2(nlo_sdf_model*, const float*, const float*, const float*, size_t, float*, float*, float*, cudaStream_t);
int nlo_sdf_tc_deep_launch_h128m3(nlo_sdf_model*, const float*, const float*, const float*, size_t, float*, float*, float*, cudaStream_t);
int nlo_sdf_tc_deep_launch_h64m2(nlo_sdf_model*, const float*, const float*, const float*, size_t, float*, float*, float*, cudaStream_t);
int nlo_sdf_tc_deep_launch_h64m3(nlo_sdf_model*, const float*, const float*, const float*, size_t, float*, float*, float*, cudaStream_t);

bool nlo_sdf_tc_deep_supported(const nlo_sdf_desc* d) {
  if (!(d->hidden == 64 || d->hidden == 128) || d->n_hidden_mats < 2 || d->n_hidden_mats > DEEP_MAXM) return false;
  const uint32_t a0 = d->act0, a = d->act;
  const bool same = a0 == a && (a == NLO_ACT_RELU || a == NLO_ACT_TANH || a == NLO_ACT_SIGMOID || a == NLO_ACT_LEAKY_RELU || a == NLO_ACT_SIN);
  const bool fourier = a0 == NLO_ACT_COS_SCALE && (a == NLO_ACT_RELU || a == NLO_ACT_TANH || a == NLO_ACT_SIGMOID || a == NLO_ACT_LEAKY_RELU);
  return same || fourier;                     // the activation pairs with a compiled tile body (dispatch_deep)
}

static float deep_pow2_scale(float mx) {     // power of two that puts mx into [2^13, 2^14)
  int ex = 0;
  if (mx > 0.f) frexpf(mx, &ex);
  return ldexpf(1.f, 14 - ex);
}
static size_t deep_img_off(int n, int k, int H) { return ((size_t)(k / 8) * (H / 8) + n / 8) * 64 + (n % 8) * 8 + (k % 8); }

int nlo_sdf_tc_deep_prepare(nlo_sdf_model* m, const float* w) {
  const int H = (int)m->desc.hidden, M = (int)m->desc.n_hidden_mats;
  const float* W0 = w;
  const float* b0 = w + 2 * H;
  auto Wl = [&](int l) { return w + 3 * H + (size_t)(l - 1) * ((size_t)H * H + H); };      // l = 1..M
  auto bl = [&](int l) { return Wl(l) + (size_t)H * H; };
  const float* wout = w + 3 * H + (size_t)M * ((size_t)H * H + H);
  static_assert(sizeof(TcDeepParams) <= 32 * sizeof(float), "parameter block is 32 floats");
  if (m->h_deep) free(m->h_deep);
  m->h_deep = static_cast<float*>(calloc(32 + sizeof(TcDeepConst) / sizeof(float) + 1, sizeof(float)));
  if (!m->h_deep) return nlo_fail("out of host memory");
  TcDeepParams prm;
  memset(&prm, 0, sizeof(prm));
  TcDeepConst* cst = reinterpret_cast<TcDeepConst*>(m->h_deep + 32);
  const size_t HH = (size_t)H * H;
  std::vector<__half> img((size_t)M * 2 * HH);
  const int act = (int)m->desc.act;
  const float dmax = act == NLO_ACT_SIGMOID ? 0.25f : act == NLO_ACT_SIN ? fabsf(m->desc.p) : 1.f;
  std::vector<float> colsum(M + 1, 0.f);
  for (int l = 1; l <= M; ++l) {
    const float* W = Wl(l);
    float mx = 0.f, rmax = 0.f, cmax = 0.f;
    std::vector<float> cs(H, 0.f);
    for (int j = 0; j < H; ++j) {
      float rs = 0.f;
      for (int k = 0; k < H; ++k) { const float a = fabsf(W[(size_t)j * H + k]); mx = fmaxf(mx, a); rs += a; cs[k] += a; }
      rmax = fmaxf(rmax, rs);
      prm.maxb[l - 1] = fmaxf(prm.maxb[l - 1], fabsf(bl(l)[j]));
      cst->b[l - 1][j] = bl(l)[j];
    }
    for (int k = 0; k < H; ++k) cmax = fmaxf(cmax, cs[k]);
    colsum[l] = cmax;
    prm.c_fwd[l - 1] = rmax;
    const float sw = deep_pow2_scale(mx);
    prm.inv_sw[l - 1] = 1.f / sw;
    __half* hi = img.data() + (size_t)(l - 1) * 2 * HH;
    __half* lo = hi + HH;
    for (int nn = 0; nn < H; ++nn)
      for (int k = 0; k < H; ++k) {
        const float v = W[(size_t)nn * H + k] * sw;
        const __half h = __float2half_rn(v);
        const size_t off = deep_img_off(nn, k, H);
        hi[off] = h;
        lo[off] = __float2half_rn(v - __half2float(h));
      }
  }
  // reverse seeds: |g_M| <= max|w_out| sup|phi'|,  |g_{l-1}| <= (max column sum of |W_l|) |g_l| sup|phi'|
  float mw = 0.f;
  for (int j = 0; j < H; ++j) mw = fmaxf(mw, fabsf(wout[j]));
  float gb = mw * dmax;
  for (int l = M; l >= 1; --l) {
    prm.sg[l - 1] = deep_pow2_scale(gb);
    prm.inv_sg[l - 1] = 1.f / prm.sg[l - 1];
    gb = colsum[l] * gb * dmax;
  }
  for (int k = 0; k < H; ++k) {
    cst->w0x[k] = W0[2 * k]; cst->w0y[k] = W0[2 * k + 1]; cst->b0[k] = b0[k]; cst->wout[k] = wout[k]; cst->wouts[k] = wout[k] * prm.sg[M - 1];
    prm.max_w0x = fmaxf(prm.max_w0x, fabsf(W0[2 * k])); prm.max_w0y = fmaxf(prm.max_w0y, fabsf(W0[2 * k + 1]));
    prm.max_b0 = fmaxf(prm.max_b0, fabsf(b0[k]));
  }
  cst->bout = wout[H];
  memcpy(m->h_deep, &prm, sizeof(prm));
  if (m->d_tc) cudaFree(m->d_tc);
  m->d_tc = nullptr;
  NLO_CUDA(cudaMalloc(&m->d_tc, img.size() * sizeof(__half) + 2 * NLO_STREAM_SLOTS * sizeof(unsigned int)));
  NLO_CUDA(cudaMemcpy(m->d_tc, img.data(), img.size() * sizeof(__half), cudaMemcpyHostToDevice));
  m->tc_bytes = img.size() * sizeof(__half);
  NLO_CUDA(cudaMemset(static_cast<char*>(m->d_tc) + m->tc_bytes, 0, 2 * NLO_STREAM_SLOTS * sizeof(unsigned int)));
  return 0;
}

int nlo_sdf_tc_deep_launch(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n, float* s, float* jx, float* jy,
                           cudaStream_t st) {
  if (n == 0) return 0;
  if (!m->d_tc || !m->h_deep) return nlo_fail("tensor-tile operands were not prepared");
  const int H = (int)m->desc.hidden, M = (int)m->desc.n_hidden_mats;
  if (H == 128 && M == 2) return nlo_sdf_tc_deep_launch_h128m2(m, x, y, sbar, n, s, jx, jy, st);
  if (H == 128 && M == 3) return nlo_sdf_tc_deep_launch_h128m3(m, x, y, sbar, n, s, jx, jy, st);
  if (H == 64 && M == 2) return nlo_sdf_tc_deep_launch_h64m2(m, x, y, sbar, n, s, jx, jy, st);
  if (H == 64 && M == 3) return nlo_sdf_tc_deep_launch_h64m3(m, x, y, sbar, n, s, jx, jy, st);
  return nlo_fail("tensor-tile path (deep): unsupported shape H=%d M=%d", H, M);
}
