/*
 * nlo_b200.h - C ABI of libnlo_b200.so: the B200 (sm_100a) implementation of the
 * NLP-evaluation hot path of pyMSE/NLOTrajectories.
 *
 * Three groups of entry points:
 *
 *  (1) CasADi "external" ABI - the drop-in for the l4casadi-generated shim
 *      reference: _l4c_generated/nn_sdf.cpp  (symbol for symbol, same signature and sparsity):
 *        nn_sdf            nn_sdf.cpp:57-60     value          s = MLP(p)
 *        jac_nn_sdf        nn_sdf.cpp:64-70     Jacobian       ds/dp
 *        adj1_nn_sdf       nn_sdf.cpp:76-83     adjoint        sbar * ds/dp
 *        jac_adj1_nn_sdf   nn_sdf.cpp:88-104    d(adj1)/dp  =  sbar * Hessian(s)
 *      plus the *_n_in/_n_out/_sparsity_in/_sparsity_out companions (nn_sdf.cpp:36-55,64-65,76-77,88-89).
 *      And the batched forms (suffix _batch; P points per call) that the reference's call site
 *      core/sdf/l4casadi.py:247-255 already supports syntactically ((rows*cols) x 2 coords).
 *
 *  (2) Batched learned-SDF API on device-resident fp32 structure-of-arrays (and host-buffer forms).
 *
 *  (3) Batched NLP evaluation: g(w), the structural non-zeros of dg/dw in compressed-column
 *      order, f(w) and grad f(w) for P independent problems
 *      reference: src/nlotrajectories/core/runner.py:44-103 (assembly),
 *                 core/dynamics.py:33-148, core/geometry.py:59-144, core/utils.py:18-33,
 *                 core/sdf/casadi.py:27-45,377-390, core/sdf/l4casadi.py:241-257.
 *
 * All functions return 0 on success and non-zero on failure (never throw); the message of the
 * last failure on the calling thread is available from nlo_last_error().  There is no CPU
 * fallback: every compute entry point fails if no CUDA device is usable.
 */
#ifndef NLO_B200_H
#define NLO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NLO_API __attribute__((visibility("default")))
#else
#define NLO_API
#endif

/* ------------------------------------------------------------------------------------------ */
/* common                                                                                      */
/* ------------------------------------------------------------------------------------------ */
NLO_API int         nlo_version(void);
NLO_API const char* nlo_last_error(void);
NLO_API int         nlo_device_count(void);            /* <0 on error                          */
NLO_API int         nlo_device_sm_count(int device);   /* streaming multiprocessors, <0 on err */

/* activation ids (phi0 for the first layer, phi for hidden layers)                            */
enum {
  NLO_ACT_RELU = 0, NLO_ACT_TANH = 1, NLO_ACT_SIGMOID = 2, NLO_ACT_LEAKY_RELU = 3,
  NLO_ACT_SIN = 4,        /* sin(p * a)      SIREN      core/nn_architectures.py:25-26         */
  NLO_ACT_COS_SCALE = 5,  /* p * cos(a)      Fourier    core/nn_architectures.py:38            */
  NLO_ACT_IDENTITY = 6
};
enum { NLO_KIND_MLP = 0, NLO_KIND_FOURIER = 1, NLO_KIND_SIREN = 2 };

/* arithmetic of the H x H layers                                                              */
enum {
  NLO_PREC_FP32_SIMT = 0,   /* FP32 FMA pipe                                                   */
  NLO_PREC_TC_3XF16 = 1,    /* tcgen05 tensor tiles, error-compensated split-fp16 (3 products), FP32 accumulate */
  NLO_PREC_AUTO      = 2    /* tensor tiles when the shape supports them (H in {64,128})       */
};

/* Network:  a0 = W0 p + b0, h0 = phi0(a0);  a_l = W_l h_{l-1} + b_l, h_l = phi(a_l), l=1..M;
 *           s = w_out . h_M + b_out.
 * Flat fp32 weight blob order: W0[H][2], b0[H], { W_l[H][H] (row = output neuron), b_l[H] } l=1..M,
 * w_out[H], b_out[1].                                                                         */
typedef struct nlo_sdf_desc {
  uint32_t kind;            /* NLO_KIND_*  (informational)                                     */
  uint32_t hidden;          /* H                                                               */
  uint32_t n_hidden_mats;   /* M                                                               */
  uint32_t act0;            /* NLO_ACT_* of the first layer                                    */
  uint32_t act;             /* NLO_ACT_* of the hidden layers                                  */
  float    p0;              /* parameter of act0 (Fourier scale / SIREN omega0)                */
  float    p;               /* parameter of act                                                */
} nlo_sdf_desc;

typedef struct nlo_sdf_model nlo_sdf_model;

NLO_API size_t nlo_sdf_weight_count(const nlo_sdf_desc* desc);
NLO_API int  nlo_sdf_create(const nlo_sdf_desc* desc, const float* weights_host, size_t n_weights,
                            int device, nlo_sdf_model** out);
/* ".nlow" file: 64-byte header {"NLOW", u32 version=1, nlo_sdf_desc, u64 n_weights, pad} + fp32 blob */
NLO_API int  nlo_sdf_load(const char* path, int device, nlo_sdf_model** out);
NLO_API int  nlo_sdf_save(const char* path, const nlo_sdf_desc* desc, const float* weights_host, size_t n_weights);
NLO_API void nlo_sdf_destroy(nlo_sdf_model* m);
NLO_API int  nlo_sdf_set_precision(nlo_sdf_model* m, int prec);   /* NLO_PREC_*                */
NLO_API int  nlo_sdf_get_precision(const nlo_sdf_model* m);       /* resolved (never AUTO)      */
NLO_API int  nlo_sdf_describe(const nlo_sdf_model* m, nlo_sdf_desc* out);

/* Device-resident evaluation.  x, y: fp32[n] (== CasADi's column-major n x 2 coordinate matrix).
 * sbar: adjoint seeds fp32[n] or NULL (=1).  Outputs fp32[n], any may be NULL:
 *   s = MLP(p)            (nn_sdf)
 *   jx, jy = sbar * ds/dp (jac_nn_sdf when sbar==NULL, adj1_nn_sdf otherwise)
 * stream: a cudaStream_t (NULL = default stream).  Asynchronous.                              */
NLO_API int nlo_sdf_eval(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n,
                         float* s, float* jx, float* jy, void* stream);
/* sbar * Hessian: hxx, hxy, hyy fp32[n]  (jac_adj1_nn_sdf)                                     */
NLO_API int nlo_sdf_hess(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n,
                         float* hxx, float* hxy, float* hyy, void* stream);
/* Host-buffer forms: copies in, evaluates, copies out, synchronises.  Same layouts, host memory. */
NLO_API int nlo_sdf_eval_host(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n,
                              float* s, float* jx, float* jy);
NLO_API int nlo_sdf_hess_host(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n,
                              float* hxx, float* hxy, float* hyy);
/* kernels launched by this library on this thread since load (bench.py's gpu_launches)        */
NLO_API unsigned long long nlo_launch_count(void);

/* ------------------------------------------------------------------------------------------ */
/* CasADi external ABI (casadi_real = double, casadi_int = long long; nn_sdf.cpp:11-17)         */
/* The functions evaluate the process-global model: bound with nlo_casadi_bind(), or loaded on  */
/* first use from $NLO_B200_WEIGHTS (a .nlow file) on device $NLO_B200_DEVICE (default 0).      */
/* ------------------------------------------------------------------------------------------ */
NLO_API int  nlo_casadi_bind(nlo_sdf_model* m);         /* library does not take ownership      */
NLO_API int  nlo_casadi_set_batch(long long n_points);  /* P of the *_batch functions (default $NLO_B200_BATCH or 1) */

NLO_API int nn_sdf(const double** arg, double** res, long long* iw, double* w, int mem);
NLO_API long long nn_sdf_n_in(void);
NLO_API long long nn_sdf_n_out(void);
NLO_API const long long* nn_sdf_sparsity_in(long long i);
NLO_API const long long* nn_sdf_sparsity_out(long long i);
NLO_API void nn_sdf_incref(void);
NLO_API void nn_sdf_decref(void);

NLO_API int jac_nn_sdf(const double** arg, double** res, long long* iw, double* w, int mem);
NLO_API long long jac_nn_sdf_n_in(void);
NLO_API long long jac_nn_sdf_n_out(void);

NLO_API int adj1_nn_sdf(const double** arg, double** res, long long* iw, double* w, int mem);
NLO_API long long adj1_nn_sdf_n_in(void);
NLO_API long long adj1_nn_sdf_n_out(void);

NLO_API int jac_adj1_nn_sdf(const double** arg, double** res, long long* iw, double* w, int mem);
NLO_API long long jac_adj1_nn_sdf_n_in(void);
NLO_API long long jac_adj1_nn_sdf_n_out(void);

/* Batched externals: i0 is P x 2 dense (column-major: x[P] then y[P]), o0 is P x 1.
 * jac_nn_sdf_batch returns the P x 2P Jacobian on the sparsity of jac_nn_sdf_batch_sparsity_out(0)
 * (2 non-zeros per row: values ds/dx[P] then ds/dy[P]).  jac_adj1_nn_sdf_batch returns the
 * 2P x 2P block-sparse matrix (4 non-zeros per point, column-major: for column c<P {hxx,hxy}, for
 * column P+c {hxy,hyy}).                                                                       */
NLO_API int nn_sdf_batch(const double** arg, double** res, long long* iw, double* w, int mem);
NLO_API long long nn_sdf_batch_n_in(void);
NLO_API long long nn_sdf_batch_n_out(void);
NLO_API const long long* nn_sdf_batch_sparsity_in(long long i);
NLO_API const long long* nn_sdf_batch_sparsity_out(long long i);
NLO_API int jac_nn_sdf_batch(const double** arg, double** res, long long* iw, double* w, int mem);
NLO_API long long jac_nn_sdf_batch_n_in(void);
NLO_API long long jac_nn_sdf_batch_n_out(void);
NLO_API const long long* jac_nn_sdf_batch_sparsity_in(long long i);
NLO_API const long long* jac_nn_sdf_batch_sparsity_out(long long i);
NLO_API int adj1_nn_sdf_batch(const double** arg, double** res, long long* iw, double* w, int mem);
NLO_API long long adj1_nn_sdf_batch_n_in(void);
NLO_API long long adj1_nn_sdf_batch_n_out(void);
NLO_API const long long* adj1_nn_sdf_batch_sparsity_in(long long i);
NLO_API const long long* adj1_nn_sdf_batch_sparsity_out(long long i);
NLO_API int jac_adj1_nn_sdf_batch(const double** arg, double** res, long long* iw, double* w, int mem);
NLO_API long long jac_adj1_nn_sdf_batch_n_in(void);
NLO_API long long jac_adj1_nn_sdf_batch_n_out(void);
NLO_API const long long* jac_adj1_nn_sdf_batch_sparsity_in(long long i);
NLO_API const long long* jac_adj1_nn_sdf_batch_sparsity_out(long long i);

/* ------------------------------------------------------------------------------------------ */
/* batched NLP evaluation                                                                      */
/* ------------------------------------------------------------------------------------------ */
enum {  /* core/dynamics.py:151-158 */
  NLO_DYN_POINT_1ST = 0, NLO_DYN_POINT_2ND = 1, NLO_DYN_UNICYCLE = 2, NLO_DYN_UNICYCLE_2ND = 3,
  NLO_DYN_ACKERMANN = 4, NLO_DYN_ACKERMANN_2ND = 5
};
enum { NLO_SHAPE_DOT = 0, NLO_SHAPE_RECTANGLE = 1, NLO_SHAPE_TRIANGLE = 2 };   /* core/geometry.py:17-20 */
enum { NLO_SDF_LEARNED = 0, NLO_SDF_CIRCLES = 1 };      /* CIRCLES = analytic obstacles: soft-min union of circles / squares */
enum { NLO_OBST_CIRCLE = 0, NLO_OBST_SQUARE = 1 };
#define NLO_MAX_CIRCLES 8

typedef struct nlo_nlp_desc {
  uint32_t dynamics;        /* NLO_DYN_*                                                       */
  uint32_t shape;           /* NLO_SHAPE_*                                                     */
  uint32_t N;               /* control intervals (solver.N)                                    */
  uint32_t use_slack;       /* core/runner.py:67-71                                            */
  uint32_t use_smooth;      /* core/runner.py:92-96                                            */
  uint32_t enforce_heading; /* core/runner.py:51-56                                            */
  uint32_t sdf_mode;        /* NLO_SDF_LEARNED (solver.mode l4casadi) / NLO_SDF_CIRCLES (casadi) */
  uint32_t n_circles;
  float dt, slack_penalty, smooth_weight;
  float length, width, wheelbase;
  float circles[NLO_MAX_CIRCLES][4];   /* analytic obstacles: cx, cy, radius | size, margin       */
  uint32_t obstacle_kind[NLO_MAX_CIRCLES]; /* NLO_OBST_CIRCLE (core/sdf/casadi.py:27-41) / NLO_OBST_SQUARE (:48-115) */
} nlo_nlp_desc;

typedef struct nlo_nlp nlo_nlp;

/* model may be NULL when sdf_mode == NLO_SDF_CIRCLES; the nlp does not take ownership         */
NLO_API int  nlo_nlp_create(const nlo_nlp_desc* desc, nlo_sdf_model* model, int device, nlo_nlp** out);
NLO_API void nlo_nlp_destroy(nlo_nlp* p);
NLO_API long long nlo_nlp_n_w(const nlo_nlp* p);      /* decision variables         */
NLO_API long long nlo_nlp_n_g(const nlo_nlp* p);      /* constraint rows            */
NLO_API long long nlo_nlp_nnz_jac(const nlo_nlp* p);  /* structural nnz of dg/dw                */
NLO_API long long nlo_nlp_n_sdf_points(const nlo_nlp* p); /* learned-SDF points per evaluation  */
/* compressed-column pattern of dg/dw: colind[n_w+1], row[nnz]                                  */
NLO_API int  nlo_nlp_jac_sparsity(const nlo_nlp* p, int32_t* colind, int32_t* row);

/* Scratch of the device entry points (footprint points, SDF values / Jacobians / Hessians of (N+1)*nb*P points) is owned by the
 * handle: ONE nlo_nlp_eval / nlo_nlp_hess call per handle may be in flight at a time (calls on one stream are ordered and
 * therefore fine; use one handle per stream for concurrent evaluation).  The scratch grows inside the call when P exceeds
 * every earlier P (cudaFree + cudaMalloc: synchronises the device and cannot be captured in a CUDA graph);
 * nlo_nlp_reserve(p, P) sizes it up front so that later calls with <= P problems allocate nothing.                       */
NLO_API int  nlo_nlp_reserve(nlo_nlp* p, size_t P);

/* Device-resident evaluation of P problems.  Structure-of-arrays, variable-major:
 *   w      fp32 [n_w ][ld]   element (v, problem i) at w[v*ld + i]
 *   g      fp32 [n_g ][ld]
 *   jac    fp32 [nnz ][ld]   values in the compressed-column order of nlo_nlp_jac_sparsity
 *   f      fp32 [P]
 *   grad_f fp32 [n_w ][ld]
 * Any output may be NULL.  ld >= P.  Asynchronous on `stream`.                                 */
NLO_API int nlo_nlp_eval(nlo_nlp* p, const float* w, size_t P, size_t ld,
                         float* g, float* jac, float* f, float* grad_f, void* stream);
/* K2 alone: only the Euler defect rows of g (core/runner.py:59-64) and their dg/dw values are written - the HBM-bound kernel
 * bench.py times in isolation for its achieved-GB/s figure.  Same layouts as nlo_nlp_eval; g / jac may be NULL.              */
NLO_API int nlo_nlp_eval_dynamics(nlo_nlp* p, const float* w, size_t P, size_t ld, float* g, float* jac, void* stream);
/* out = add + (dg/dw)^T y for every problem: jac [nnz][ld] (CCS values from nlo_nlp_eval), y [n_g][ld], add [n_w][ld] or NULL,
 * out [n_w][ld].  The product a constrained solver needs for the gradient of its (augmented) Lagrangian.        */
NLO_API int nlo_nlp_jac_tvec(nlo_nlp* p, const float* jac, const float* y, const float* add, size_t P, size_t ld,
                             float* out, void* stream);
/* Hessian of the Lagrangian  sigma * hess f + sum_r lam_r * hess g_r  (what IPOPT's eval_h asks CasADi's nlp_hess_l
 * for; core/runner.py:113-125 leaves hessian_approximation at its default "exact").  Built from the second
 * derivatives of core/dynamics.py:59-148, core/geometry.py:78-117, core/utils.py:28-31, core/runner.py:80-98 and
 * jac_adj1_nn_sdf (_l4c_generated/nn_sdf.cpp:88-104).  Structural non-zeros of the UPPER triangle in
 * compressed-column order: colind[n_w+1], row[nnz_hess].
 *   sigma fp32 [P] or NULL (= 1)     lam fp32 [n_g][ld]     hess fp32 [nnz_hess][ld]                              */
NLO_API long long nlo_nlp_nnz_hess(const nlo_nlp* p);
NLO_API int  nlo_nlp_hess_sparsity(const nlo_nlp* p, int32_t* colind, int32_t* row);
NLO_API int  nlo_nlp_hess(nlo_nlp* p, const float* w, const float* sigma, const float* lam, size_t P, size_t ld,
                          float* hess, void* stream);
/* max constraint violation per problem given bounds lbg/ubg (fp32[n_g], device), for best-of selection */
NLO_API int nlo_nlp_violation(nlo_nlp* p, const float* g, const float* lbg, const float* ubg, size_t P, size_t ld,
                              float* viol, void* stream);
/* Host-buffer form: problem-major rows (w_host[P][n_w] etc., what a per-problem solver holds).
 * Copies in (pinned staging), transposes on device, evaluates, transposes back, copies out, syncs. */
NLO_API int nlo_nlp_eval_host(nlo_nlp* p, const float* w_host, size_t P,
                              float* g_host, float* jac_host, float* f_host, float* grad_f_host);
/* Compact host form.  Of what nlo_nlp_eval_host returns, much does not depend on w at all and the rest of it is a copy of w:
 *   g      rows that are copies of a variable (x_0 pin, terminal pin, slack >= 0, control box: core/runner.py:50-56,67-69,101-103)
 *   dg/dw  the +-1 / -dt entries of the Euler defects (core/runner.py:59-64), the 1s of the copy rows and of d(row)/d(slack)
 *   grad f zero except d/d(x_k, y_k) (path length) and the entries linear in their own variable (slack / control penalties,
 *          core/runner.py:85-96)
 * benchmark_6: 1,453 of 3,225 Jacobian values, 173 of 1,057 rows of g and 565 of 727 gradient entries.  A per-problem solver fills
 * those ONCE (nlo_nlp_compact_layout gives the index lists and values) and asks only for what varies:
 *   g_var_host    [P][n_g_var]    g[g_var_rows[c]]
 *   jac_var_host  [P][n_jac_var]  jac[jac_var_nz[c]]       (positions in the CCS order of nlo_nlp_jac_sparsity)
 *   grad_var_host [P][n_grad_var] grad_f[grad_var_idx[c]]
 *   f_host        [P]
 * Everything else follows from the layout:  g[g_copy_rows[i]] = w[g_copy_vars[i]];  jac[jac_const_nz[i]] = jac_const_val[i];
 * grad_f[grad_lin_idx[i]] = grad_lin_coef[i] * w[grad_lin_idx[i]] (fp32 product, bit-identical to the full form); all other
 * gradient entries are 0.  Same arithmetic and same kernels as nlo_nlp_eval_host; any output may be NULL.                  */
typedef struct nlo_nlp_compact_counts_t {
  long long n_g_var, n_g_copy, n_jac_var, n_jac_const, n_grad_var, n_grad_lin;
} nlo_nlp_compact_counts_t;
NLO_API int nlo_nlp_compact_counts(const nlo_nlp* p, nlo_nlp_compact_counts_t* out);
/* index lists sized by nlo_nlp_compact_counts; any pointer may be NULL                          */
NLO_API int nlo_nlp_compact_layout(const nlo_nlp* p, int32_t* g_var_rows, int32_t* g_copy_rows, int32_t* g_copy_vars,
                                   int32_t* jac_var_nz, int32_t* jac_const_nz, float* jac_const_val,
                                   int32_t* grad_var_idx, int32_t* grad_lin_idx, float* grad_lin_coef);
NLO_API int nlo_nlp_eval_host_compact(nlo_nlp* p, const float* w_host, size_t P,
                                      float* g_var_host, float* jac_var_host, float* f_host, float* grad_var_host);
/* ------------------------------------------------------------------------------------------ */
/* batched interior-point solver (the caller on both sides of the evaluation path)             */
/* The reference hands every problem to IPOPT (core/runner.py:112-133: tol 1e-4, exact Hessian, */
/* max_iter 1000); here P multi-start problems are solved at once on the device: primal-dual    */
/* interior point, slacks on the inequality rows, l1-merit line search, monotone barrier, and a  */
/* block-tridiagonal (stage-structured) Cholesky of the condensed KKT matrix per problem.        */
/* ------------------------------------------------------------------------------------------ */
typedef struct nlo_ip nlo_ip;
typedef struct nlo_ip_options {
  double tol;            /* scaled KKT error at which a problem is converged (core/runner.py:118: 1e-4)  */
  int    max_iter;       /* iteration limit                                                              */
  double mu0;            /* initial barrier parameter (0.1)                                               */
  int    ls_multipliers; /* re-estimate the equality multipliers by least squares after every step (1)    */
  int    compact;        /* finished problems leave the working set every 10 iterations (1)               */
  int    verbose;
} nlo_ip_options;
typedef struct nlo_ip_stats {
  int iterations, evaluations, hessians, trials, compactions;   /* batched calls of each kind                              */
  long long trial_problems;                                     /* problems evaluated by the line-search trials, in total   */
  double phase_ms[9];   /* device time per phase (CUDA events): evaluation, residuals + barrier update, Hessian, KKT assembly +
                           factorisation, step, line search, update, least-squares multipliers, compaction / output            */
  long long kkt_problems, kkt_retries;   /* Newton systems solved / of those, how many failed the inertia test at their first delta      */
  long long kkt_retry_hist[16];          /* of the retries: how many succeeded at attempt 1..15 of the search (slot 15: none)            */
} nlo_ip_stats;
/* lbg, ubg: fp64[n_g] bounds of g in Opti's canonical form (+-INFINITY for one-sided rows; rows with lbg == ubg are equalities).
 * max_problems: batch size the device buffers are sized for.  The solver keeps a pointer to `p` (not owned).              */
NLO_API int    nlo_ip_create(nlo_nlp* p, const double* lbg, const double* ubg, size_t max_problems, nlo_ip** out);
NLO_API void   nlo_ip_destroy(nlo_ip* s);
NLO_API size_t nlo_ip_capacity(const nlo_ip* s);
/* Solve P problems from the starts w0_host[P][n_w] (problem-major fp64).  Outputs (host, any may be NULL): w_host[P][n_w],
 * f, viol (max bound violation of g), kkt_err (scaled optimality error), iters, status (1 converged, 2 feasible with a
 * stationary objective but a KKT error above tol - typical on the kinks of a ReLU SDF -, 0 neither), lam_host[P][n_g].
 * opt may be NULL (defaults above).                                                                                      */
NLO_API int    nlo_ip_solve(nlo_ip* s, const double* w0_host, size_t P, const nlo_ip_options* opt, double* w_host, double* f_host,
                            double* viol_host, double* kkt_err_host, int* iters_host, int* status_host, double* lam_host,
                            nlo_ip_stats* stats);
/* One regularised Newton step of the condensed KKT system on device-resident inputs (all variable-major, leading dimension
 * ld <= capacity):  (H + J^T diag(omega) J + delta I) dw = rhs  per problem, with jac / hess the CCS values of
 * nlo_nlp_eval / nlo_nlp_hess, omega fp64 [n_g][ld], rhs / dw fp64 [n_w][ld], delta_in / delta_out fp64 [P]: delta starts at
 * delta_in and grows (x8 from 1e-4) until the block-tridiagonal Cholesky succeeds.                                         */
NLO_API int    nlo_ip_kkt_step(nlo_ip* s, const float* jac, const float* hess, const double* omega, const double* rhs,
                               const double* delta_in, size_t P, size_t ld, double* dw, double* delta_out, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* initial guesses: batched RRT against the exact obstacle SDF (core/trajectory_initialization.py:175-216) */
/* ------------------------------------------------------------------------------------------ */
enum { NLO_RRT_CIRCLE = 0, NLO_RRT_SQUARE = 1, NLO_RRT_POLYGON = 2 };
typedef struct nlo_rrt_obstacle {
  uint32_t kind;                      /* NLO_RRT_*  (trapezoids and elliptical half-rings are polygons: core/sdf/casadi.py:135-148,218-246) */
  uint32_t first_vertex, n_vertices;  /* polygon: range in the vertex array                                                              */
  uint32_t pad;
  double cx, cy, size, margin;        /* circle: centre, radius; square: centre, side length; every kind: margin                         */
} nlo_rrt_obstacle;
/* P planners, one per seed, each from start[2] towards goal[2] inside the box [lo, hi]: one warp per planner runs its whole search
 * (sample / nearest / steer by step_size / keep `inflation` of clearance at both ends and the midpoint of the new edge / stop within
 * step_size of the goal) in one kernel.  postprocess != 0: the same warp then inserts a midpoint before every corner sharper than
 * 60 degrees and shortcuts the path greedily through collision-free straight segments (the reference's insert_intermediate_points /
 * _shortcut_path, core/trajectory_initialization.py:222-224).  path_host[P][max_path][2] receives every path root first,
 * path_len_host[P] its number of nodes (-1: no path within max_iter iterations, -2: longer than max_path).  Synchronous.            */
NLO_API int nlo_rrt_paths(const nlo_rrt_obstacle* obstacles, int n_obstacles, const double* vertices, int n_vertices,
                          const double* start, const double* goal, const double* lo, const double* hi, const long long* seeds, size_t P,
                          double step_size, int max_iter, double inflation, double goal_sample_rate, int max_path, int postprocess,
                          int device, double* path_host, int* path_len_host);

/* layout helpers on device: [rows][ld] variable-major <-> [P][rows] problem-major              */
NLO_API int nlo_transpose_to_soa(const float* aos, float* soa, size_t P, size_t rows, size_t ld, void* stream);
NLO_API int nlo_transpose_to_aos(const float* soa, float* aos, size_t P, size_t rows, size_t ld, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NLO_B200_H */
