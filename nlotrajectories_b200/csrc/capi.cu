// C ABI of libnlo_b200.so (declared in include/nlo_b200.h): handles, batched entry points,
// host-buffer entry points and the CasADi external ABI that replaces _l4c_generated/nn_sdf.cpp.
#include "nlo_common.cuh"
#include "nlp_internal.cuh"
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

// ---- errors / counters ---------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static thread_local unsigned long long g_launches = 0;

int nlo_fail(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  if (getenv("NLO_B200_VERBOSE")) fprintf(stderr, "[nlo_b200] error: %s\n", g_err);
  return 1;
}
void nlo_count_launch(unsigned n) { g_launches += n; }

int nlo_model_stream_slot(nlo_sdf_model* m, cudaStream_t st) {
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  for (int i = 0; i < m->n_slots; ++i) if (m->slot_stream[i] == st) return i;
  if (m->n_slots == NLO_STREAM_SLOTS) {
    // every slot belongs to some other stream: drain the device, then nothing in flight owns a slot any more
    if (cudaDeviceSynchronize() != cudaSuccess) { nlo_fail("device synchronisation failed while recycling stream slots"); return -1; }
    m->n_slots = 0;
  }
  m->slot_stream[m->n_slots] = st;
  return m->n_slots++;
}

extern "C" {

int nlo_version(void) { return 100; }
const char* nlo_last_error(void) { return g_err; }
unsigned long long nlo_launch_count(void) { return g_launches; }

int nlo_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { nlo_fail("cudaGetDeviceCount failed: no usable CUDA device"); return -1; }
  return n;
}
int nlo_device_sm_count(int device) {
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) { nlo_fail("cannot query device %d", device); return -1; }
  return n;
}

// ---- SDF model --------------------------------------------------------------------------------------
size_t nlo_sdf_weight_count(const nlo_sdf_desc* d) {
  const size_t H = d->hidden, M = d->n_hidden_mats;
  return 3 * H + M * (H * H + H) + H + 1;
}

static int validate_desc(const nlo_sdf_desc* d) {
  if (!d) return nlo_fail("null descriptor");
  if (d->hidden < 4 || d->hidden > 1024 || d->hidden % 4) return nlo_fail("hidden width %u unsupported (multiple of 4 in [4,1024])", d->hidden);
  if (d->n_hidden_mats > 16) return nlo_fail("too many hidden layers (%u)", d->n_hidden_mats);
  if (d->act0 > NLO_ACT_IDENTITY || d->act > NLO_ACT_IDENTITY) return nlo_fail("unknown activation id");
  return 0;
}

int nlo_sdf_create(const nlo_sdf_desc* desc, const float* weights, size_t n_weights, int device, nlo_sdf_model** out) {
  if (!out) return nlo_fail("null out");
  *out = nullptr;
  if (validate_desc(desc)) return 1;
  if (!weights) return nlo_fail("null weights");
  if (n_weights != nlo_sdf_weight_count(desc))
    return nlo_fail("weight count %zu does not match descriptor (%zu expected)", n_weights, nlo_sdf_weight_count(desc));
  int ndev = nlo_device_count();
  if (ndev <= 0) return nlo_fail("no CUDA device available: libnlo_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return nlo_fail("device %d out of range (have %d)", device, ndev);
  NLO_CUDA(cudaSetDevice(device));
  nlo_sdf_model* m = new (std::nothrow) nlo_sdf_model();
  if (!m) return nlo_fail("out of host memory");
  memset(m, 0, sizeof(*m));
  m->desc = *desc; m->device = device; m->n_w = n_weights;
  { static std::mutex mu; static unsigned long long next_uid = 1; std::lock_guard<std::mutex> lk(mu); m->uid = next_uid++; }
  NLO_CUDA(cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, device));
  NLO_CUDA(cudaMalloc(&m->d_w, n_weights * sizeof(float)));
  NLO_CUDA(cudaMemcpy(m->d_w, weights, n_weights * sizeof(float), cudaMemcpyHostToDevice));
  if (desc->n_hidden_mats > 0) {
    const size_t H = desc->hidden, M = desc->n_hidden_mats;
    std::vector<float> wt(M * H * H);
    for (size_t l = 0; l < M; ++l) {
      const float* W = weights + 3 * H + l * (H * H + H);
      for (size_t j = 0; j < H; ++j) for (size_t k = 0; k < H; ++k) wt[l * H * H + k * H + j] = W[j * H + k];
    }
    NLO_CUDA(cudaMalloc(&m->d_wt, wt.size() * sizeof(float)));
    NLO_CUDA(cudaMemcpy(m->d_wt, wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  NLO_CUDA(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  m->prec = NLO_PREC_FP32_SIMT;
  if (nlo_sdf_tc_supported(desc)) {
    if (nlo_sdf_tc_prepare(m, weights)) { nlo_sdf_destroy(m); return 1; }
    m->prec = NLO_PREC_TC_3XF16;     // AUTO default: tensor tiles when the shape supports them
  }
  const char* env = getenv("NLO_B200_PRECISION");
  if (env && !strcmp(env, "fp32")) m->prec = NLO_PREC_FP32_SIMT;
  *out = m;
  return 0;
}

struct NlowHeader { char magic[4]; uint32_t version; nlo_sdf_desc desc; uint32_t pad0; uint64_t n_weights; uint8_t pad[16]; };
static_assert(sizeof(nlo_sdf_desc) == 28, "descriptor layout is part of the file format");
static_assert(sizeof(NlowHeader) == 64, "header must be 64 bytes");

int nlo_sdf_save(const char* path, const nlo_sdf_desc* desc, const float* weights, size_t n_weights) {
  if (validate_desc(desc)) return 1;
  if (n_weights != nlo_sdf_weight_count(desc)) return nlo_fail("weight count mismatch");
  FILE* fp = fopen(path, "wb");
  if (!fp) return nlo_fail("cannot open %s for writing", path);
  NlowHeader h; memset(&h, 0, sizeof(h));
  memcpy(h.magic, "NLOW", 4); h.version = 1; h.desc = *desc; h.n_weights = n_weights;
  bool ok = fwrite(&h, sizeof(h), 1, fp) == 1 && fwrite(weights, sizeof(float), n_weights, fp) == n_weights;
  fclose(fp);
  return ok ? 0 : nlo_fail("short write to %s", path);
}

int nlo_sdf_load(const char* path, int device, nlo_sdf_model** out) {
  if (!path) return nlo_fail("null path");
  FILE* fp = fopen(path, "rb");
  if (!fp) return nlo_fail("cannot open weight file %s", path);
  NlowHeader h;
  if (fread(&h, sizeof(h), 1, fp) != 1 || memcmp(h.magic, "NLOW", 4) || h.version != 1) { fclose(fp); return nlo_fail("%s is not a .nlow v1 file", path); }
  if (validate_desc(&h.desc) || h.n_weights != nlo_sdf_weight_count(&h.desc)) { fclose(fp); return nlo_fail("%s: inconsistent header", path); }
  std::vector<float> w(h.n_weights);
  bool ok = fread(w.data(), sizeof(float), w.size(), fp) == w.size();
  fclose(fp);
  if (!ok) return nlo_fail("%s: truncated", path);
  return nlo_sdf_create(&h.desc, w.data(), w.size(), device, out);
}

void nlo_sdf_destroy(nlo_sdf_model* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  if (m->d_w) cudaFree(m->d_w);
  if (m->d_wt) cudaFree(m->d_wt);
  if (m->d_tc) cudaFree(m->d_tc);
  if (m->d_io) cudaFree(m->d_io);
  for (float* b : m->d_ws) if (b) cudaFree(b);
  if (m->h_io) cudaFreeHost(m->h_io);
  if (m->h_deep) free(m->h_deep);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

int nlo_sdf_set_precision(nlo_sdf_model* m, int prec) {
  if (!m) return nlo_fail("null model");
  if (prec == NLO_PREC_AUTO) prec = nlo_sdf_tc_supported(&m->desc) ? NLO_PREC_TC_3XF16 : NLO_PREC_FP32_SIMT;
  if (prec == NLO_PREC_TC_3XF16 && !nlo_sdf_tc_supported(&m->desc))
    return nlo_fail("tensor-tile path supports H in {64,128} with 1-3 hidden matrices (compiled activation pairs) and H=256 ReLU; this model has H=%u M=%u",
                    m->desc.hidden, m->desc.n_hidden_mats);
  if (prec != NLO_PREC_TC_3XF16 && prec != NLO_PREC_FP32_SIMT) return nlo_fail("unknown precision %d", prec);
  m->prec = prec;
  return 0;
}
int nlo_sdf_get_precision(const nlo_sdf_model* m) { return m ? m->prec : -1; }
int nlo_sdf_describe(const nlo_sdf_model* m, nlo_sdf_desc* out) { if (!m || !out) return nlo_fail("null argument"); *out = m->desc; return 0; }

int nlo_sdf_eval(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n,
                 float* s, float* jx, float* jy, void* stream) {
  if (!m) return nlo_fail("null model");
  if (n == 0) return 0;
  if (!x || !y) return nlo_fail("null coordinate array");
  NLO_CUDA(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  if ((jx == nullptr) != (jy == nullptr)) return nlo_fail("jx and jy must be requested together");
  if (m->prec == NLO_PREC_TC_3XF16) return nlo_sdf_tc_launch(m, x, y, sbar, n, s, jx, jy, st);
  return nlo_sdf_simt_launch(m, x, y, sbar, n, s, jx, jy, st);
}

// second derivatives of piecewise-linear networks vanish identically (what sdf_hess_kernel would compute, exactly)
static bool sdf_is_piecewise_linear(const nlo_sdf_model* m) {
  auto lin = [](uint32_t a) { return a == NLO_ACT_RELU || a == NLO_ACT_LEAKY_RELU; };
  return lin(m->desc.act0) && (m->desc.n_hidden_mats == 0 || lin(m->desc.act));
}

int nlo_sdf_hess(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n,
                 float* hxx, float* hxy, float* hyy, void* stream) {
  if (!m) return nlo_fail("null model");
  if (n == 0) return 0;
  if (!x || !y) return nlo_fail("null coordinate array");
  NLO_CUDA(cudaSetDevice(m->device));
  if (sdf_is_piecewise_linear(m)) {
    for (float* b : {hxx, hxy, hyy}) if (b) NLO_CUDA(cudaMemsetAsync(b, 0, n * sizeof(float), (cudaStream_t)stream));
    return 0;
  }
  if (m->prec == NLO_PREC_TC_3XF16 && hxx && hxy && hyy && nlo_sdf_tc_hess_supported(m))
    return nlo_sdf_tc_hess_launch(m, x, y, sbar, n, nullptr, nullptr, nullptr, hxx, hxy, hyy, (cudaStream_t)stream);
  if (m->prec == NLO_PREC_TC_3XF16 && hxx && hxy && hyy && nlo_sdf_tc_hess_gemm_supported(m))
    return nlo_sdf_tc_hess_gemm_launch(m, x, y, sbar, n, nullptr, nullptr, nullptr, hxx, hxy, hyy, (cudaStream_t)stream);
  return nlo_sdf_simt_hess_launch(m, x, y, sbar, n, hxx, hxy, hyy, (cudaStream_t)stream);
}

static int ensure_io(nlo_sdf_model* m, size_t floats) {
  if (m->io_cap < floats) {
    if (m->d_io) cudaFree(m->d_io);
    m->d_io = nullptr; m->io_cap = 0;
    NLO_CUDA(cudaMalloc(&m->d_io, floats * sizeof(float)));
    m->io_cap = floats;
  }
  return 0;
}

// Host-buffer evaluation on the model's private stream.  Pinned caller buffers move at full PCIe
// speed; pageable ones are staged by the driver.
int nlo_sdf_eval_host(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n,
                      float* s, float* jx, float* jy) {
  if (!m) return nlo_fail("null model");
  if (n == 0) return 0;
  if (!x || !y) return nlo_fail("null coordinate array");
  NLO_CUDA(cudaSetDevice(m->device));
  const size_t n_in = sbar ? 3 : 2;
  if (ensure_io(m, (n_in + 3) * n)) return 1;
  float* d = m->d_io;
  cudaStream_t st = m->stream;
  NLO_CUDA(cudaMemcpyAsync(d, x, n * sizeof(float), cudaMemcpyHostToDevice, st));
  NLO_CUDA(cudaMemcpyAsync(d + n, y, n * sizeof(float), cudaMemcpyHostToDevice, st));
  if (sbar) NLO_CUDA(cudaMemcpyAsync(d + 2 * n, sbar, n * sizeof(float), cudaMemcpyHostToDevice, st));
  float* ds = d + n_in * n; float* djx = ds + n; float* djy = djx + n;
  int rc = (m->prec == NLO_PREC_TC_3XF16)
               ? nlo_sdf_tc_launch(m, d, d + n, sbar ? d + 2 * n : nullptr, n, s ? ds : nullptr, jx ? djx : nullptr, jy ? djy : nullptr, st)
               : nlo_sdf_simt_launch(m, d, d + n, sbar ? d + 2 * n : nullptr, n, s ? ds : nullptr, jx ? djx : nullptr, jy ? djy : nullptr, st);
  if (rc) return rc;
  if (s) NLO_CUDA(cudaMemcpyAsync(s, ds, n * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (jx) NLO_CUDA(cudaMemcpyAsync(jx, djx, n * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (jy) NLO_CUDA(cudaMemcpyAsync(jy, djy, n * sizeof(float), cudaMemcpyDeviceToHost, st));
  NLO_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int nlo_sdf_hess_host(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n,
                      float* hxx, float* hxy, float* hyy) {
  if (!m) return nlo_fail("null model");
  if (n == 0) return 0;
  NLO_CUDA(cudaSetDevice(m->device));
  if (ensure_io(m, 6 * n)) return 1;
  float* d = m->d_io; cudaStream_t st = m->stream;
  NLO_CUDA(cudaMemcpyAsync(d, x, n * sizeof(float), cudaMemcpyHostToDevice, st));
  NLO_CUDA(cudaMemcpyAsync(d + n, y, n * sizeof(float), cudaMemcpyHostToDevice, st));
  if (sbar) NLO_CUDA(cudaMemcpyAsync(d + 2 * n, sbar, n * sizeof(float), cudaMemcpyHostToDevice, st));
  if (nlo_sdf_hess(m, d, d + n, sbar ? d + 2 * n : nullptr, n, d + 3 * n, d + 4 * n, d + 5 * n, st)) return 1;
  if (hxx) NLO_CUDA(cudaMemcpyAsync(hxx, d + 3 * n, n * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (hxy) NLO_CUDA(cudaMemcpyAsync(hxy, d + 4 * n, n * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (hyy) NLO_CUDA(cudaMemcpyAsync(hyy, d + 5 * n, n * sizeof(float), cudaMemcpyDeviceToHost, st));
  NLO_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// ---- CasADi external ABI (replaces _l4c_generated/nn_sdf.cpp) ----------------------------------------
static std::mutex g_mu;
static nlo_sdf_model* g_model = nullptr;
static bool g_model_owned = false;
static long long g_batch = 0;
static std::vector<long long> g_sp[8];

int nlo_casadi_bind(nlo_sdf_model* m) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_model_owned && g_model) nlo_sdf_destroy(g_model);
  g_model = m; g_model_owned = false;
  return 0;
}

static long long batch_size() {
  if (g_batch <= 0) { const char* e = getenv("NLO_B200_BATCH"); g_batch = e ? atoll(e) : 1; if (g_batch <= 0) g_batch = 1; }
  return g_batch;
}
int nlo_casadi_set_batch(long long n) {
  if (n <= 0) return nlo_fail("batch must be positive");
  std::lock_guard<std::mutex> lk(g_mu);
  g_batch = n;
  for (auto& v : g_sp) v.clear();
  return 0;
}

static nlo_sdf_model* casadi_model() {
  if (g_model) return g_model;
  const char* path = getenv("NLO_B200_WEIGHTS");
  if (!path) { nlo_fail("no model bound: call nlo_casadi_bind() or set NLO_B200_WEIGHTS=<file.nlow>"); return nullptr; }
  const char* dev = getenv("NLO_B200_DEVICE");
  nlo_sdf_model* m = nullptr;
  if (nlo_sdf_load(path, dev ? atoi(dev) : 0, &m)) return nullptr;
  g_model = m; g_model_owned = true;
  return m;
}

// Staging for the CasADi externals: pinned, device-mapped host memory.  Small calls (the reference's pattern is ONE 1x2 point per
// call, nn_sdf.cpp:57-104) skip every memcpy: the kernel reads its inputs from and writes its results to this buffer directly
// (zero-copy over PCIe), so a call costs one launch and one stream synchronisation.
static float* g_pin = nullptr;
static size_t g_pin_cap = 0;
static int ensure_pin(size_t floats) {
  if (g_pin_cap >= floats) return 0;
  if (g_pin) cudaFreeHost(g_pin);
  g_pin = nullptr; g_pin_cap = 0;
  const size_t want = floats < 4096 ? 4096 : floats;
  NLO_CUDA(cudaHostAlloc(&g_pin, want * sizeof(float), cudaHostAllocMapped | cudaHostAllocPortable));
  g_pin_cap = want;
  return 0;
}
constexpr size_t NLO_ZERO_COPY_MAX = 16384;      // points per call up to which the zero-copy form is used

// mode 0: value, 1: jac, 2: adj1, 3: jac_adj1.  in: x[P], y[P] doubles (column-major P x 2).
static int casadi_eval(int mode, long long P, const double* p, const double* seed, double* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  nlo_sdf_model* m = casadi_model();
  if (!m) { fprintf(stderr, "[nlo_b200] %s\n", g_err); return 1; }
  const size_t n = (size_t)P;
  if (cudaSetDevice(m->device) != cudaSuccess || ensure_pin(6 * n)) { fprintf(stderr, "[nlo_b200] %s\n", g_err); return 1; }
  float* hx = g_pin; float* hy = hx + n; float* hs = hy + n; float* o0 = hs + n; float* o1 = o0 + n; float* o2 = o1 + n;
  for (size_t i = 0; i < n; ++i) { hx[i] = p ? (float)p[i] : 0.f; hy[i] = p ? (float)p[n + i] : 0.f; hs[i] = seed ? (float)seed[i] : 0.f; }
  const bool use_seed = (mode == 2 || mode == 3);   // NULL seed == zeros (CasADi convention)
  int rc;
  if (n <= NLO_ZERO_COPY_MAX) {
    // unified addressing: the pinned, mapped buffer is valid as a device pointer
    if (mode == 0) rc = nlo_sdf_eval(m, hx, hy, nullptr, n, o0, nullptr, nullptr, m->stream);
    else if (mode == 1 || mode == 2) rc = nlo_sdf_eval(m, hx, hy, use_seed ? hs : nullptr, n, nullptr, o0, o1, m->stream);
    else rc = nlo_sdf_hess(m, hx, hy, hs, n, o0, o1, o2, m->stream);
    if (!rc && cudaStreamSynchronize(m->stream) != cudaSuccess) rc = nlo_fail("stream synchronisation failed");
  } else if (mode == 0) rc = nlo_sdf_eval_host(m, hx, hy, nullptr, n, o0, nullptr, nullptr);
  else if (mode == 1 || mode == 2) rc = nlo_sdf_eval_host(m, hx, hy, use_seed ? hs : nullptr, n, nullptr, o0, o1);
  else rc = nlo_sdf_hess_host(m, hx, hy, hs, n, o0, o1, o2);
  if (rc) { fprintf(stderr, "[nlo_b200] %s\n", g_err); return 1; }
  if (mode == 0) for (size_t i = 0; i < n; ++i) out[i] = o0[i];
  else if (mode == 1 || mode == 2) for (size_t i = 0; i < n; ++i) { out[i] = o0[i]; out[n + i] = o1[i]; }
  else for (size_t i = 0; i < n; ++i) {                  // column-major over the 4-per-point pattern
    out[2 * i] = o0[i]; out[2 * i + 1] = o1[i]; out[2 * n + 2 * i] = o1[i]; out[2 * n + 2 * i + 1] = o2[i];
  }
  return 0;
}

static const long long s_in0[3] = {1, 2, 1};
static const long long s_out0[3] = {1, 1, 1};
long long nn_sdf_n_in(void) { return 1; }
long long nn_sdf_n_out(void) { return 1; }
const long long* nn_sdf_sparsity_in(long long i) { return i == 0 ? s_in0 : nullptr; }
const long long* nn_sdf_sparsity_out(long long i) { return i == 0 ? s_out0 : nullptr; }
void nn_sdf_incref(void) {}
void nn_sdf_decref(void) {}
int nn_sdf(const double** arg, double** res, long long*, double*, int) {
  if (!res || !res[0]) return 0;
  return casadi_eval(0, 1, arg ? arg[0] : nullptr, nullptr, res[0]);
}
long long jac_nn_sdf_n_in(void) { return 2; }
long long jac_nn_sdf_n_out(void) { return 1; }
int jac_nn_sdf(const double** arg, double** res, long long*, double*, int) {
  if (!res || !res[0]) return 0;
  return casadi_eval(1, 1, arg ? arg[0] : nullptr, nullptr, res[0]);
}
long long adj1_nn_sdf_n_in(void) { return 3; }
long long adj1_nn_sdf_n_out(void) { return 1; }
int adj1_nn_sdf(const double** arg, double** res, long long*, double*, int) {
  // adj1 [i0, out_o0, adj_o0] -> [out_adj_i0]   (nn_sdf.cpp:80)
  if (!res || !res[0]) return 0;
  return casadi_eval(2, 1, arg ? arg[0] : nullptr, arg ? arg[2] : nullptr, res[0]);
}
long long jac_adj1_nn_sdf_n_in(void) { return 4; }
long long jac_adj1_nn_sdf_n_out(void) { return 3; }
static int jac_adj1_common(long long P, const double** arg, double** res) {
  // jac_adj1 [i0, out_o0, adj_o0, out_adj_i0] -> [jac_adj_i0_i0, jac_adj_i0_out_o0, jac_adj_i0_adj_o0]  (nn_sdf.cpp:92)
  if (!res) return 0;
  if (res[1] != nullptr) { nlo_fail("jac_adj_i0_out_o0 is not provided"); fprintf(stderr, "[nlo_b200] %s\n", g_err); return 1; }
  if (res[2] != nullptr) { nlo_fail("jac_adj_i0_adj_o0 is not provided"); fprintf(stderr, "[nlo_b200] %s\n", g_err); return 1; }
  if (res[0] == nullptr) { nlo_fail("only jac_adj_i0_i0 can be provided"); fprintf(stderr, "[nlo_b200] %s\n", g_err); return 1; }
  return casadi_eval(3, P, arg ? arg[0] : nullptr, arg ? arg[2] : nullptr, res[0]);
}
int jac_adj1_nn_sdf(const double** arg, double** res, long long*, double*, int) {
  if (!res || !res[0]) return jac_adj1_common(1, arg, res);
  // single point: dense 2x2 column-major {hxx, hxy, hxy, hyy}
  double tmp[4];
  double* r2[3] = {tmp, res[1], res[2]};
  int rc = jac_adj1_common(1, arg, r2);
  if (!rc) { res[0][0] = tmp[0]; res[0][1] = tmp[1]; res[0][2] = tmp[2]; res[0][3] = tmp[3]; }
  return rc;
}

// batched forms ------------------------------------------------------------------------------------------
static const long long* dense_sp(int slot, long long r, long long c) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto& v = g_sp[slot];
  if (v.empty()) { v = {r, c, 1}; }
  return v.data();
}
long long nn_sdf_batch_n_in(void) { return 1; }
long long nn_sdf_batch_n_out(void) { return 1; }
const long long* nn_sdf_batch_sparsity_in(long long i) { return i == 0 ? dense_sp(0, batch_size(), 2) : nullptr; }
const long long* nn_sdf_batch_sparsity_out(long long i) { return i == 0 ? dense_sp(1, batch_size(), 1) : nullptr; }
int nn_sdf_batch(const double** arg, double** res, long long*, double*, int) {
  if (!res || !res[0]) return 0;
  return casadi_eval(0, batch_size(), arg ? arg[0] : nullptr, nullptr, res[0]);
}
long long jac_nn_sdf_batch_n_in(void) { return 2; }
long long jac_nn_sdf_batch_n_out(void) { return 1; }
const long long* jac_nn_sdf_batch_sparsity_in(long long i) {
  return i == 0 ? dense_sp(0, batch_size(), 2) : i == 1 ? dense_sp(1, batch_size(), 1) : nullptr;
}
const long long* jac_nn_sdf_batch_sparsity_out(long long i) {
  if (i != 0) return nullptr;
  const long long P = batch_size();
  std::lock_guard<std::mutex> lk(g_mu);
  auto& v = g_sp[2];
  if (v.empty()) {           // P x 2P, column c holds row c % P
    v.reserve(2 + 2 * P + 1 + 2 * P);
    v.push_back(P); v.push_back(2 * P);
    for (long long c = 0; c <= 2 * P; ++c) v.push_back(c);
    for (long long c = 0; c < 2 * P; ++c) v.push_back(c % P);
  }
  return v.data();
}
int jac_nn_sdf_batch(const double** arg, double** res, long long*, double*, int) {
  if (!res || !res[0]) return 0;
  return casadi_eval(1, batch_size(), arg ? arg[0] : nullptr, nullptr, res[0]);
}
long long adj1_nn_sdf_batch_n_in(void) { return 3; }
long long adj1_nn_sdf_batch_n_out(void) { return 1; }
const long long* adj1_nn_sdf_batch_sparsity_in(long long i) {
  return i == 0 ? dense_sp(0, batch_size(), 2) : (i == 1 || i == 2) ? dense_sp(1, batch_size(), 1) : nullptr;
}
const long long* adj1_nn_sdf_batch_sparsity_out(long long i) { return i == 0 ? dense_sp(0, batch_size(), 2) : nullptr; }
int adj1_nn_sdf_batch(const double** arg, double** res, long long*, double*, int) {
  if (!res || !res[0]) return 0;
  return casadi_eval(2, batch_size(), arg ? arg[0] : nullptr, arg ? arg[2] : nullptr, res[0]);
}
long long jac_adj1_nn_sdf_batch_n_in(void) { return 4; }
long long jac_adj1_nn_sdf_batch_n_out(void) { return 3; }
const long long* jac_adj1_nn_sdf_batch_sparsity_in(long long i) {
  return (i == 0 || i == 3) ? dense_sp(0, batch_size(), 2) : (i == 1 || i == 2) ? dense_sp(1, batch_size(), 1) : nullptr;
}
const long long* jac_adj1_nn_sdf_batch_sparsity_out(long long i) {
  const long long P = batch_size();
  if (i == 0) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto& v = g_sp[3];
    if (v.empty()) {         // 2P x 2P, column c holds rows {c % P, c % P + P}
      v.push_back(2 * P); v.push_back(2 * P);
      for (long long c = 0; c <= 2 * P; ++c) v.push_back(2 * c);
      for (long long c = 0; c < 2 * P; ++c) { v.push_back(c % P); v.push_back(c % P + P); }
    }
    return v.data();
  }
  if (i == 1 || i == 2) {    // structurally empty 2P x P blocks
    std::lock_guard<std::mutex> lk(g_mu);
    auto& v = g_sp[4];
    if (v.empty()) { v.push_back(2 * P); v.push_back(P); for (long long c = 0; c <= P; ++c) v.push_back(0); }
    return v.data();
  }
  return nullptr;
}
int jac_adj1_nn_sdf_batch(const double** arg, double** res, long long*, double*, int) {
  if (!res) return 0;
  // outputs 1 and 2 are structurally empty here, so CasADi may pass non-NULL zero-length buffers: ignore them
  if (!res[0]) return 0;
  return casadi_eval(3, batch_size(), arg ? arg[0] : nullptr, arg ? arg[2] : nullptr, res[0]);
}

// ---- NLP -------------------------------------------------------------------------------------------------
int nlo_nlp_create(const nlo_nlp_desc* desc, nlo_sdf_model* model, int device, nlo_nlp** out) {
  if (!out) return nlo_fail("null out");
  *out = nullptr;
  if (!desc) return nlo_fail("null descriptor");
  if (desc->sdf_mode == NLO_SDF_LEARNED && !model) return nlo_fail("learned-SDF mode needs a model");
  if (desc->sdf_mode > NLO_SDF_CIRCLES) return nlo_fail("unknown sdf_mode");
  int ndev = nlo_device_count();
  if (ndev <= 0) return nlo_fail("no CUDA device available: libnlo_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return nlo_fail("device %d out of range", device);
  if (model && model->device != device) return nlo_fail("model lives on device %d, nlp requested on %d", model->device, device);
  NLO_CUDA(cudaSetDevice(device));
  nlo_nlp* p = new (std::nothrow) nlo_nlp();
  if (!p) return nlo_fail("out of host memory");
  p->desc = *desc; p->model = model; p->device = device;
  p->d_tables = nullptr; p->d_hmap = nullptr; p->hs_cap_P = 0; p->zc = nullptr; p->zc_cap = 0;
  p->d_hs[0] = p->d_hs[1] = p->d_hs[2] = nullptr;
  memset(&p->scratch, 0, sizeof(p->scratch));
  memset(p->lane, 0, sizeof(p->lane));
  std::vector<int> nzmap, copy_row, copy_var, copy_emit;
  std::vector<float> const_ccs;
  p->d_compact = nullptr;
  if (nlo_nlp_build_layout(desc, &p->L, &p->rows_ccs, &p->cols_ccs, &nzmap, &copy_row, &copy_var, &copy_emit, &const_ccs)) { delete p; return 1; }
  {
    // what varies with w and what does not (compact host form)
    const NlpDev& L = p->L;
    std::vector<char> is_copy(L.n_g, 0);
    for (size_t i = 0; i < copy_row.size(); ++i) { is_copy[copy_row[i]] = 1; p->cg_copy_row.push_back(copy_row[i]); p->cg_copy_var.push_back(copy_var[i]); }
    for (int r = 0; r < L.n_g; ++r) if (!is_copy[r]) p->cg_rows.push_back(r);
    for (int z = 0; z < L.nnz; ++z) {
      if (std::isnan(const_ccs[z])) p->cj_nz.push_back(z);
      else { p->cj_const_nz.push_back(z); p->cj_const_val.push_back(const_ccs[z]); }
    }
    for (int k = 0; k <= L.N; ++k) { p->cgr_idx.push_back(k * L.nx); p->cgr_idx.push_back(k * L.nx + 1); }
    if (L.use_smooth) for (int k = 0; k < L.N - 1; ++k) for (int i = 0; i < L.nu; ++i) {
      p->cgr_lin_idx.push_back(L.n_X + k * L.nu + i); p->cgr_lin_coef.push_back(2.f * L.smooth_weight);
    }
    if (L.use_slack) for (int k = 0; k <= L.N; ++k) { p->cgr_lin_idx.push_back(L.n_X + L.n_U + k); p->cgr_lin_coef.push_back(2.f * L.slack_penalty); }
    std::vector<int> ct(p->cg_rows);
    ct.insert(ct.end(), p->cj_nz.begin(), p->cj_nz.end());
    ct.insert(ct.end(), p->cgr_idx.begin(), p->cgr_idx.end());
    if (cudaMalloc(&p->d_compact, ct.size() * sizeof(int)) != cudaSuccess ||
        cudaMemcpy(p->d_compact, ct.data(), ct.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
      nlo_nlp_destroy(p); return nlo_fail("device allocation failed");
    }
  }
  if (cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) { delete p; return nlo_fail("cannot query device"); }
  const int nc = (int)copy_row.size();
  std::vector<int> tables(nzmap);
  tables.insert(tables.end(), copy_row.begin(), copy_row.end());
  tables.insert(tables.end(), copy_var.begin(), copy_var.end());
  for (int i = 0; i < nc; ++i) tables.push_back(nzmap[copy_emit[i]]);
  const size_t off_colind = tables.size();
  {
    std::vector<int> colind(p->L.n_w + 1, 0);
    for (int i = 0; i < p->L.nnz; ++i) colind[p->cols_ccs[i] + 1]++;
    for (int cidx = 0; cidx < p->L.n_w; ++cidx) colind[cidx + 1] += colind[cidx];
    tables.insert(tables.end(), colind.begin(), colind.end());
    tables.insert(tables.end(), p->rows_ccs.begin(), p->rows_ccs.end());
  }
  if (cudaMalloc(&p->d_tables, tables.size() * sizeof(int)) != cudaSuccess ||
      cudaMemcpy(p->d_tables, tables.data(), tables.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
    nlo_nlp_destroy(p); return nlo_fail("device allocation failed");
  }
  p->L.nzmap = p->d_tables; p->L.copy_row = p->d_tables + nzmap.size();
  p->L.copy_var = p->L.copy_row + nc; p->L.copy_nz = p->L.copy_var + nc; p->L.n_copy = nc;
  p->d_colind = p->d_tables + off_colind; p->d_row = p->d_colind + p->L.n_w + 1;
  {
    std::vector<int> hmap;
    if (nlo_nlp_build_hess_layout(p->L, &p->hrows_ccs, &p->hcols_ccs, &hmap)) { nlo_nlp_destroy(p); return 1; }
    if (cudaMalloc(&p->d_hmap, hmap.size() * sizeof(int)) != cudaSuccess ||
        cudaMemcpy(p->d_hmap, hmap.data(), hmap.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
      nlo_nlp_destroy(p); return nlo_fail("device allocation failed");
    }
  }
  for (auto& ln : p->lane)
    if (cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking) != cudaSuccess) { nlo_nlp_destroy(p); return nlo_fail("stream creation failed"); }
  *out = p;
  return 0;
}

static void free_scratch(NlpScratch& sc) {
  float** bufs[] = {&sc.px, &sc.py, &sc.s, &sc.jx, &sc.jy};
  for (float** b : bufs) { if (*b) cudaFree(*b); *b = nullptr; }
  sc.cap_P = 0;
}
static void free_lane_bufs(NlpLane& ln) {
  float** bufs[] = {&ln.d_in, &ln.d_w, &ln.d_g, &ln.d_jac, &ln.d_f, &ln.d_grad, &ln.d_og, &ln.d_ojac, &ln.d_ograd};
  for (float** b : bufs) { if (*b) cudaFree(*b); *b = nullptr; }
  free_scratch(ln.scratch);
  ln.cap_P = 0;
}

void nlo_nlp_destroy(nlo_nlp* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  free_scratch(p->scratch);
  for (auto& ln : p->lane) { free_lane_bufs(ln); if (ln.stream) cudaStreamDestroy(ln.stream); }
  if (p->d_tables) cudaFree(p->d_tables);
  if (p->d_hmap) cudaFree(p->d_hmap);
  if (p->d_compact) cudaFree(p->d_compact);
  if (p->zc) cudaFreeHost(p->zc);
  for (float*& b : p->d_hs) { if (b) cudaFree(b); b = nullptr; }
  delete p;
}

long long nlo_nlp_n_w(const nlo_nlp* p) { return p ? p->L.n_w : -1; }
long long nlo_nlp_n_g(const nlo_nlp* p) { return p ? p->L.n_g : -1; }
long long nlo_nlp_nnz_jac(const nlo_nlp* p) { return p ? p->L.nnz : -1; }
long long nlo_nlp_n_sdf_points(const nlo_nlp* p) { return p ? (long long)(p->L.N + 1) * p->L.nb : -1; }

int nlo_nlp_jac_sparsity(const nlo_nlp* p, int32_t* colind, int32_t* row) {
  if (!p || !colind || !row) return nlo_fail("null argument");
  const int n_w = p->L.n_w, nnz = p->L.nnz;
  for (int c = 0; c <= n_w; ++c) colind[c] = 0;
  for (int i = 0; i < nnz; ++i) { colind[p->cols_ccs[i] + 1]++; row[i] = p->rows_ccs[i]; }
  for (int c = 0; c < n_w; ++c) colind[c + 1] += colind[c];
  return 0;
}

static int ensure_scratch(const nlo_nlp* p, NlpScratch& sc, size_t P) {
  if (sc.cap_P >= P) return 0;
  free_scratch(sc);
  const size_t n = (size_t)(p->L.N + 1) * p->L.nb * P;
  float** bufs[] = {&sc.px, &sc.py, &sc.s, &sc.jx, &sc.jy};
  for (float** b : bufs) NLO_CUDA(cudaMalloc(b, n * sizeof(float)));
  sc.cap_P = P;
  return 0;
}

static int nlp_eval_on(nlo_nlp* p, NlpScratch& sc, const float* w, size_t P, size_t ld, float* g, float* jac, float* f, float* grad_f,
                       cudaStream_t st) {
  // K3 fused into K1: hard SDF rows of a footprint with a heading on the ReLU / ReLU tensor kernel (benchmark_6) - the footprint points are
  // formed inside the SDF kernel and its results go straight into g and dg/dw: two launches per evaluation, no point / value scratch
  const NlpDev& L = p->L;
  const bool fused = (g || jac) && L.sdf_mode == NLO_SDF_LEARNED && !L.use_slack && L.shape != NLO_SHAPE_DOT && L.nnz_sdf_row == 3 &&
                     L.rows_per_knot == L.nb && nlo_sdf_tc_rows_supported(p->model) && (size_t)(L.N + 1) * L.nb * P <= 0x7FFFFF00u;
  if (fused) {
    if (nlo_nlp_launch_assembly(p, sc, w, P, ld, g, jac, f, grad_f, st, 0, true)) return 1;
    return nlo_sdf_tc_rows_launch(p->model, w, P, ld, L.N + 1, L.nx, L.nb, L.bx, L.by, g ? g + (size_t)L.g_off_sdf * ld : nullptr, jac,
                                  L.nzmap + L.e_off_sdf, L.rows_per_knot * L.nnz_sdf_row, st);
  }
  if ((g || jac) && p->L.sdf_mode == NLO_SDF_LEARNED && ensure_scratch(p, sc, P)) return 1;   // (analytic obstacles need no scratch)
  if (nlo_nlp_launch_assembly(p, sc, w, P, ld, g, jac, f, grad_f, st, 0)) return 1;
  if (g || jac) {
    if (p->L.sdf_mode == NLO_SDF_LEARNED) {
      const size_t n = (size_t)(p->L.N + 1) * p->L.nb * P;
      if (nlo_sdf_eval(p->model, sc.px, sc.py, nullptr, n, sc.s, jac ? sc.jx : nullptr, jac ? sc.jy : nullptr, st)) return 1;
    }
    if (nlo_nlp_launch_assembly(p, sc, w, P, ld, g, jac, f, grad_f, st, 1)) return 1;
  }
  return 0;
}

static int ensure_hess_scratch(nlo_nlp* p, size_t P) {
  if (p->hs_cap_P >= P) return 0;
  const size_t n = (size_t)(p->L.N + 1) * p->L.nb * P;
  for (float*& b : p->d_hs) { if (b) cudaFree(b); b = nullptr; }
  p->hs_cap_P = 0;
  for (float*& b : p->d_hs) NLO_CUDA(cudaMalloc(&b, n * sizeof(float)));
  p->hs_cap_P = P;
  return 0;
}

int nlo_nlp_reserve(nlo_nlp* p, size_t P) {
  if (!p) return nlo_fail("null nlp");
  NLO_CUDA(cudaSetDevice(p->device));
  if (ensure_scratch(p, p->scratch, P)) return 1;
  return ensure_hess_scratch(p, P);
}

int nlo_nlp_eval(nlo_nlp* p, const float* w, size_t P, size_t ld, float* g, float* jac, float* f, float* grad_f, void* stream) {
  if (!p) return nlo_fail("null nlp");
  if (P == 0) return 0;
  if (!w) return nlo_fail("null w");
  if (ld < P) return nlo_fail("ld (%zu) < P (%zu)", ld, P);
  NLO_CUDA(cudaSetDevice(p->device));
  return nlp_eval_on(p, p->scratch, w, P, ld, g, jac, f, grad_f, (cudaStream_t)stream);
}

// ---- Hessian of the Lagrangian (nlp_hess.cu) ------------------------------------------------------------------
long long nlo_nlp_nnz_hess(const nlo_nlp* p) { return p ? (long long)p->hrows_ccs.size() : -1; }

int nlo_nlp_hess_sparsity(const nlo_nlp* p, int32_t* colind, int32_t* row) {
  if (!p || !colind || !row) return nlo_fail("null argument");
  const int n_w = p->L.n_w, nnz = (int)p->hrows_ccs.size();
  for (int c = 0; c <= n_w; ++c) colind[c] = 0;
  for (int i = 0; i < nnz; ++i) { colind[p->hcols_ccs[i] + 1]++; row[i] = p->hrows_ccs[i]; }
  for (int c = 0; c < n_w; ++c) colind[c + 1] += colind[c];
  return 0;
}

int nlo_nlp_hess(nlo_nlp* p, const float* w, const float* sigma, const float* lam, size_t P, size_t ld, float* hess, void* stream) {
  if (!p) return nlo_fail("null nlp");
  if (P == 0) return 0;
  if (!w || !lam || !hess) return nlo_fail("null argument");
  if (ld < P) return nlo_fail("ld (%zu) < P (%zu)", ld, P);
  NLO_CUDA(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (ensure_scratch(p, p->scratch, P)) return 1;
  const size_t n = (size_t)(p->L.N + 1) * p->L.nb * P;
  if (ensure_hess_scratch(p, P)) return 1;
  NlpScratch& sc = p->scratch;
  if (nlo_nlp_launch_points(p->L, w, P, ld, sc.px, sc.py, st)) return 1;
  if (p->L.sdf_mode == NLO_SDF_CIRCLES) {
    if (nlo_nlp_launch_circles_hess(p->L, sc.px, sc.py, n, sc.s, sc.jx, sc.jy, p->d_hs[0], p->d_hs[1], p->d_hs[2], p->sm_count, st)) return 1;
  } else {
    if (p->model->prec == NLO_PREC_TC_3XF16 && (nlo_sdf_tc_hess_supported(p->model) || nlo_sdf_tc_hess_gemm_supported(p->model))) {
      // one fused launch: value, Jacobian and Hessian of every footprint point
      if (nlo_sdf_tc_hess_supported(p->model)
              ? nlo_sdf_tc_hess_launch(p->model, sc.px, sc.py, nullptr, n, sc.s, sc.jx, sc.jy, p->d_hs[0], p->d_hs[1], p->d_hs[2], st)
              : nlo_sdf_tc_hess_gemm_launch(p->model, sc.px, sc.py, nullptr, n, sc.s, sc.jx, sc.jy, p->d_hs[0], p->d_hs[1], p->d_hs[2], st)) return 1;
      return nlo_nlp_launch_hess(p->L, p->d_hmap, w, sigma, lam, P, ld, sc.s, sc.jx, sc.jy, p->d_hs[0], p->d_hs[1], p->d_hs[2], hess, st);
    }
    if (nlo_sdf_eval(p->model, sc.px, sc.py, nullptr, n, sc.s, sc.jx, sc.jy, st)) return 1;
    if (sdf_is_piecewise_linear(p->model)) {
      for (float* b : p->d_hs) NLO_CUDA(cudaMemsetAsync(b, 0, n * sizeof(float), st));
    } else if (nlo_sdf_hess(p->model, sc.px, sc.py, nullptr, n, p->d_hs[0], p->d_hs[1], p->d_hs[2], st)) {
      return 1;
    }
  }
  return nlo_nlp_launch_hess(p->L, p->d_hmap, w, sigma, lam, P, ld, sc.s, sc.jx, sc.jy, p->d_hs[0], p->d_hs[1], p->d_hs[2], hess, st);
}

// Only the Euler defect rows of g and their Jacobian values (K2): what the HBM roofline of the dynamics kernel is measured on.
int nlo_nlp_eval_dynamics(nlo_nlp* p, const float* w, size_t P, size_t ld, float* g, float* jac, void* stream) {
  if (!p || !w) return nlo_fail("null argument");
  if (P == 0) return 0;
  if (ld < P) return nlo_fail("ld (%zu) < P (%zu)", ld, P);
  NLO_CUDA(cudaSetDevice(p->device));
  return nlo_nlp_launch_dynamics(p->L, w, P, ld, g, jac, (cudaStream_t)stream);
}

int nlo_nlp_violation(nlo_nlp* p, const float* g, const float* lbg, const float* ubg, size_t P, size_t ld, float* viol, void* stream) {
  if (!p || !g || !lbg || !ubg || !viol) return nlo_fail("null argument");
  NLO_CUDA(cudaSetDevice(p->device));
  return nlo_launch_violation(p->L.n_g, g, lbg, ubg, P, ld, viol, p->sm_count, (cudaStream_t)stream);
}

int nlo_nlp_jac_tvec(nlo_nlp* p, const float* jac, const float* y, const float* add, size_t P, size_t ld, float* out, void* stream) {
  if (!p || !jac || !y || !out) return nlo_fail("null argument");
  if (P == 0) return 0;
  NLO_CUDA(cudaSetDevice(p->device));
  return nlo_launch_jtv(p->L.n_w, p->d_colind, p->d_row, jac, y, P, ld, add, out, p->sm_count, (cudaStream_t)stream);
}

int nlo_transpose_to_soa(const float* aos, float* soa, size_t P, size_t rows, size_t ld, void* stream) {
  int dev = 0, sm = 148;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
  return nlo_launch_transpose(aos, soa, P, rows, rows, ld, sm, (cudaStream_t)stream);
}
int nlo_transpose_to_aos(const float* soa, float* aos, size_t P, size_t rows, size_t ld, void* stream) {
  int dev = 0, sm = 148;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
  return nlo_launch_transpose(soa, aos, rows, P, ld, rows, sm, (cudaStream_t)stream);
}

static int ensure_lane(nlo_nlp* p, NlpLane& ln, size_t P) {
  if (ln.cap_P >= P) return 0;
  cudaStream_t st = ln.stream;
  free_lane_bufs(ln);
  ln.stream = st;
  const NlpDev& L = p->L;
  NLO_CUDA(cudaMalloc(&ln.d_in, (size_t)L.n_w * P * sizeof(float)));
  NLO_CUDA(cudaMalloc(&ln.d_w, (size_t)L.n_w * P * sizeof(float)));
  NLO_CUDA(cudaMalloc(&ln.d_g, (size_t)L.n_g * P * sizeof(float)));
  NLO_CUDA(cudaMalloc(&ln.d_jac, (size_t)L.nnz * P * sizeof(float)));
  NLO_CUDA(cudaMalloc(&ln.d_f, P * sizeof(float)));
  NLO_CUDA(cudaMalloc(&ln.d_grad, (size_t)L.n_w * P * sizeof(float)));
  NLO_CUDA(cudaMalloc(&ln.d_og, (size_t)L.n_g * P * sizeof(float)));
  NLO_CUDA(cudaMalloc(&ln.d_ojac, (size_t)L.nnz * P * sizeof(float)));
  NLO_CUDA(cudaMalloc(&ln.d_ograd, (size_t)L.n_w * P * sizeof(float)));
  ln.cap_P = P;
  return 0;
}

// Host-buffer evaluation, problem-major rows.  The batch is cut into chunks that alternate between two lanes
// (streams): while one lane's results travel device->host, the other lane's inputs travel host->device and its
// kernels run, so PCIe is busy in both directions and the GPU work hides behind the copies.
// Small batches (the reference's own case is ONE problem per IPOPT callback): no memcpy and no device transpose.  The decision
// vectors are laid out variable-major in pinned, device-mapped memory by the CPU, the kernels read and write that memory directly
// (zero-copy), and the CPU scatters the results back: three launches and one stream synchronisation per call.
// compact: g / jac / grad records hold only the entries listed by nlo_nlp_compact_layout (cg_rows, cj_nz, cgr_idx).
static int nlp_eval_host_small(nlo_nlp* p, const float* w_host, size_t P, float* g_host, float* jac_host, float* f_host, float* grad_host,
                               bool compact) {
  const NlpDev& L = p->L;
  const size_t n_in = (size_t)L.n_w * P, n_out = ((size_t)L.n_g + L.nnz + L.n_w + 1) * P;
  if (p->zc_cap < n_in + n_out) {
    if (p->zc) cudaFreeHost(p->zc);
    p->zc = nullptr; p->zc_cap = 0;
    NLO_CUDA(cudaHostAlloc(&p->zc, (n_in + n_out) * sizeof(float), cudaHostAllocMapped | cudaHostAllocPortable));
    p->zc_cap = n_in + n_out;
  }
  float* w = p->zc; float* g = w + n_in; float* jac = g + (size_t)L.n_g * P; float* grad = jac + (size_t)L.nnz * P; float* f = grad + (size_t)L.n_w * P;
  for (size_t i = 0; i < P; ++i) for (int v = 0; v < L.n_w; ++v) w[(size_t)v * P + i] = w_host[i * L.n_w + v];
  cudaStream_t st = p->lane[0].stream;
  if (nlp_eval_on(p, p->lane[0].scratch, w, P, P, g_host ? g : nullptr, jac_host ? jac : nullptr, f_host ? f : nullptr, grad_host ? grad : nullptr, st)) return 1;
  NLO_CUDA(cudaStreamSynchronize(st));
  const int ng = compact ? (int)p->cg_rows.size() : L.n_g, nj = compact ? (int)p->cj_nz.size() : L.nnz, nr = compact ? (int)p->cgr_idx.size() : L.n_w;
  for (size_t i = 0; i < P; ++i) {
    if (g_host) for (int r = 0; r < ng; ++r) g_host[i * ng + r] = g[(size_t)(compact ? p->cg_rows[r] : r) * P + i];
    if (jac_host) for (int z = 0; z < nj; ++z) jac_host[i * nj + z] = jac[(size_t)(compact ? p->cj_nz[z] : z) * P + i];
    if (grad_host) for (int v = 0; v < nr; ++v) grad_host[i * nr + v] = grad[(size_t)(compact ? p->cgr_idx[v] : v) * P + i];
    if (f_host) f_host[i] = f[i];
  }
  return 0;
}

static int nlp_eval_host_impl(nlo_nlp* p, const float* w_host, size_t P, float* g_host, float* jac_host, float* f_host, float* grad_host,
                              bool compact) {
  if (!p) return nlo_fail("null nlp");
  if (P == 0) return 0;
  if (!w_host) return nlo_fail("null w");
  NLO_CUDA(cudaSetDevice(p->device));
  if (P <= 8) return nlp_eval_host_small(p, w_host, P, g_host, jac_host, f_host, grad_host, compact);
  const NlpDev& L = p->L;
  // chunks: enough of them that the first chunk's upload + kernels (the only part no copy hides) is a small share of the call,
  // large enough that a chunk's kernels fill the GPU.  Measured on B200 (benchmark_6 x 65,536, compact form, ms per call):
  // 8 chunks / 2 lanes 17.8, 16 / 2 18.1, 32 / 3 16.3, 64 / 4 15.0 (the raw device->host copy of the payload alone: 13.2)
  static const int want_chunks = [] { const char* e = getenv("NLO_B200_HOST_CHUNKS"); const int v = e ? atoi(e) : 64; return v < 1 ? 1 : v; }();
  static const int want_lanes = [] { const char* e = getenv("NLO_B200_HOST_LANES"); const int v = e ? atoi(e) : 4; return v < 1 ? 1 : (v > NLO_HOST_LANES ? NLO_HOST_LANES : v); }();
  size_t chunk = (P + want_chunks - 1) / want_chunks;
  if (chunk < 1024) chunk = 1024;
  if (chunk > P) chunk = P;
  const size_t n_chunks = (P + chunk - 1) / chunk;
  const int n_lanes = (int)std::min<size_t>(n_chunks, (size_t)want_lanes);
  const int ng = compact ? (int)p->cg_rows.size() : L.n_g, nj = compact ? (int)p->cj_nz.size() : L.nnz, nr = compact ? (int)p->cgr_idx.size() : L.n_w;
  const int* d_cg = p->d_compact; const int* d_cj = d_cg + p->cg_rows.size(); const int* d_cr = d_cj + p->cj_nz.size();
  for (int l = 0; l < n_lanes; ++l) if (ensure_lane(p, p->lane[l], chunk)) return 1;
  for (size_t c = 0; c < n_chunks; ++c) {
    NlpLane& ln = p->lane[c % n_lanes];
    cudaStream_t st = ln.stream;
    const size_t p0 = c * chunk, Pc = (p0 + chunk <= P) ? chunk : P - p0;
    NLO_CUDA(cudaMemcpyAsync(ln.d_in, w_host + p0 * L.n_w, (size_t)L.n_w * Pc * sizeof(float), cudaMemcpyHostToDevice, st));
    if (nlo_launch_transpose(ln.d_in, ln.d_w, Pc, L.n_w, L.n_w, Pc, p->sm_count, st)) return 1;
    if (nlp_eval_on(p, ln.scratch, ln.d_w, Pc, Pc, g_host ? ln.d_g : nullptr, jac_host ? ln.d_jac : nullptr, f_host ? ln.d_f : nullptr,
                    grad_host ? ln.d_grad : nullptr, st)) return 1;
    if (jac_host) {
      if (compact ? nlo_launch_pack_rows(ln.d_jac, Pc, Pc, d_cj, nj, ln.d_ojac, p->sm_count, st)
                  : nlo_launch_transpose(ln.d_jac, ln.d_ojac, L.nnz, Pc, Pc, L.nnz, p->sm_count, st)) return 1;
      NLO_CUDA(cudaMemcpyAsync(jac_host + p0 * nj, ln.d_ojac, (size_t)nj * Pc * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    if (g_host) {
      if (compact ? nlo_launch_pack_rows(ln.d_g, Pc, Pc, d_cg, ng, ln.d_og, p->sm_count, st)
                  : nlo_launch_transpose(ln.d_g, ln.d_og, L.n_g, Pc, Pc, L.n_g, p->sm_count, st)) return 1;
      NLO_CUDA(cudaMemcpyAsync(g_host + p0 * ng, ln.d_og, (size_t)ng * Pc * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    if (grad_host) {
      if (compact ? nlo_launch_pack_rows(ln.d_grad, Pc, Pc, d_cr, nr, ln.d_ograd, p->sm_count, st)
                  : nlo_launch_transpose(ln.d_grad, ln.d_ograd, L.n_w, Pc, Pc, L.n_w, p->sm_count, st)) return 1;
      NLO_CUDA(cudaMemcpyAsync(grad_host + p0 * nr, ln.d_ograd, (size_t)nr * Pc * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    if (f_host) NLO_CUDA(cudaMemcpyAsync(f_host + p0, ln.d_f, Pc * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  for (int l = 0; l < n_lanes; ++l) NLO_CUDA(cudaStreamSynchronize(p->lane[l].stream));
  return 0;
}

int nlo_nlp_eval_host(nlo_nlp* p, const float* w_host, size_t P, float* g_host, float* jac_host, float* f_host, float* grad_host) {
  return nlp_eval_host_impl(p, w_host, P, g_host, jac_host, f_host, grad_host, false);
}

int nlo_nlp_eval_host_compact(nlo_nlp* p, const float* w_host, size_t P, float* g_var_host, float* jac_var_host, float* f_host,
                              float* grad_var_host) {
  return nlp_eval_host_impl(p, w_host, P, g_var_host, jac_var_host, f_host, grad_var_host, true);
}

int nlo_nlp_compact_counts(const nlo_nlp* p, nlo_nlp_compact_counts_t* out) {
  if (!p || !out) return nlo_fail("null argument");
  out->n_g_var = (long long)p->cg_rows.size(); out->n_g_copy = (long long)p->cg_copy_row.size();
  out->n_jac_var = (long long)p->cj_nz.size(); out->n_jac_const = (long long)p->cj_const_nz.size();
  out->n_grad_var = (long long)p->cgr_idx.size(); out->n_grad_lin = (long long)p->cgr_lin_idx.size();
  return 0;
}

int nlo_nlp_compact_layout(const nlo_nlp* p, int32_t* g_var_rows, int32_t* g_copy_rows, int32_t* g_copy_vars, int32_t* jac_var_nz,
                           int32_t* jac_const_nz, float* jac_const_val, int32_t* grad_var_idx, int32_t* grad_lin_idx, float* grad_lin_coef) {
  if (!p) return nlo_fail("null nlp");
  auto put = [](int32_t* dst, const std::vector<int>& v) { if (dst) for (size_t i = 0; i < v.size(); ++i) dst[i] = v[i]; };
  auto putf = [](float* dst, const std::vector<float>& v) { if (dst) for (size_t i = 0; i < v.size(); ++i) dst[i] = v[i]; };
  put(g_var_rows, p->cg_rows); put(g_copy_rows, p->cg_copy_row); put(g_copy_vars, p->cg_copy_var);
  put(jac_var_nz, p->cj_nz); put(jac_const_nz, p->cj_const_nz); putf(jac_const_val, p->cj_const_val);
  put(grad_var_idx, p->cgr_idx); put(grad_lin_idx, p->cgr_lin_idx); putf(grad_lin_coef, p->cgr_lin_coef);
  return 0;
}

}  // extern "C"
