// b200bo.cu -- host driver + C ABI (include/b200bo.h) of the B200 GP-surrogate / acquisition engine.
// sm_100a only.  No CPU fallback: every entry point fails with B200BO_E_NODEVICE / B200BO_E_CUDA when the
// device is missing.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/b200bo.h"
#include "dgemm.cuh"
#include "fit_kernels.cuh"
#include "predict_kernels.cuh"

using namespace b2;

static thread_local std::string g_err;
static int set_err(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CU_TRY(expr)                                                                         \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      char _b[512];                                                                          \
      snprintf(_b, sizeof _b, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return set_err(B200BO_E_CUDA, _b);                                                     \
    }                                                                                        \
  } while (0)

#define CHECK_ARG(cond, msg) \
  do {                       \
    if (!(cond)) return set_err(B200BO_E_ARG, msg); \
  } while (0)

namespace {

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t reserve(size_t want) {
    if (want <= n) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) n = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

struct EvPool {
  std::vector<cudaEvent_t> ev;
  size_t used = 0;
  cudaEvent_t get() {
    if (used == ev.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      ev.push_back(e);
    }
    return ev[used++];
  }
  void reset() { used = 0; }
  void destroy() {
    for (auto e : ev) cudaEventDestroy(e);
    ev.clear();
  }
};

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

using GemmNT = GemmCore<64, 64, 32, 32, false, false, 3>;  // A (MxK), B (NxK)
using GemmNN = GemmCore<64, 64, 32, 32, false, true, 3>;   // A (MxK), B (KxN)
using GemmTN = GemmCore<64, 64, 32, 32, true, true, 3>;    // A (KxM), B (KxN)

}  // namespace

struct b200bo_ctx {
  int device = 0, num_sms = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  int prec = B200BO_PREC_FP64;
  bool keepR = false;
  // training data
  int N = 0, D = 0, ld = 0;
  DevBuf<double> Xt, y, F, theta;
  // factorisation
  DevBuf<double> A, W, S, Rkeep, Dinv, Yt, Ft, rho, gamma, part, scal;
  DevBuf<int> status;
  bool factored = false;
  int corr = 0, mode = 0, trend = 0, estimate_trend = 1, n_theta = 0;
  double sigma2 = NAN, noise_var = 0, beta = 0, G = NAN, llf = -INFINITY, par_last = NAN;
  // predict workspace
  int Mc = 0;
  DevBuf<double> Xc, Kst, yhat, sumsq, dotf, mse, params, part_val, best_val, vals;
  DevBuf<long long> part_idx, best_idx;
  EvPool evs;
  double timings[B200BO_N_TIMINGS] = {0};
  double fit_timings[B200BO_N_TIMINGS] = {0};
};

namespace {

template <typename Core, bool A_KM, bool B_KN>
cudaError_t launch_gemm(b200bo_ctx* h, const GemmArgs& g, int M, int N, int batch) {
  auto kern = dgemm_kernel<64, 64, 32, 32, A_KM, B_KN, 3>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Core::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid(N / 64, M / 64, batch);
  kern<<<grid, Core::NT, Core::SMEM_BYTES, h->stream>>>(g);
  return cudaGetLastError();
}

struct PhaseTimer {
  b200bo_ctx* h;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spans[4];
  void begin(int k) {
    cudaEvent_t e = h->evs.get();
    cudaEventRecord(e, h->stream);
    spans[k].push_back({e, nullptr});
  }
  void end(int k) {
    cudaEvent_t e = h->evs.get();
    cudaEventRecord(e, h->stream);
    spans[k].back().second = e;
  }
  double total(int k) {
    double t = 0;
    for (auto& s : spans[k]) {
      float ms = 0;
      if (s.second && cudaEventElapsedTime(&ms, s.first, s.second) == cudaSuccess) t += ms;
    }
    return t;
  }
};

int ensure_predict_ws(b200bo_ctx* h, int q, bool need_vals_stage) {
  if (h->Mc == 0) h->Mc = h->num_sms * PC_BM;
  size_t Mc = h->Mc;
  CU_TRY(h->Xc.reserve(Mc * h->D));
  CU_TRY(h->Kst.reserve(Mc * (size_t)h->ld));
  CU_TRY(h->yhat.reserve(Mc));
  CU_TRY(h->sumsq.reserve(Mc));
  CU_TRY(h->dotf.reserve(Mc));
  CU_TRY(h->mse.reserve(Mc));
  int qq = std::max(q, 1);
  CU_TRY(h->params.reserve(qq));
  CU_TRY(h->part_val.reserve((size_t)qq * h->num_sms));
  CU_TRY(h->part_idx.reserve((size_t)qq * h->num_sms));
  CU_TRY(h->best_val.reserve(qq));
  CU_TRY(h->best_idx.reserve(qq));
  if (need_vals_stage) CU_TRY(h->vals.reserve((size_t)qq * Mc));
  return 0;
}

}  // namespace

extern "C" {

const char* b200bo_last_error(void) { return g_err.c_str(); }
int b200bo_version(void) { return 100; }

int b200bo_create(int device, b200bo_handle* out) {
  CHECK_ARG(out != nullptr, "out is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return set_err(B200BO_E_NODEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
  CHECK_ARG(device >= 0 && device < n, "device index out of range");
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return set_err(B200BO_E_NODEVICE, "libb200bo is built for sm_100a (B200) only; found sm_" +
                                          std::to_string(prop.major) + std::to_string(prop.minor));
  CU_TRY(cudaSetDevice(device));
  b200bo_ctx* h = new b200bo_ctx();
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  CU_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CU_TRY(h->scal.reserve(16));
  CU_TRY(h->status.reserve(1));
  *out = h;
  return 0;
}

int b200bo_destroy(b200bo_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  h->Xt.release(); h->y.release(); h->F.release(); h->theta.release();
  h->A.release(); h->W.release(); h->S.release(); h->Rkeep.release(); h->Dinv.release();
  h->Yt.release(); h->Ft.release(); h->rho.release(); h->gamma.release(); h->part.release();
  h->scal.release(); h->status.release();
  h->Xc.release(); h->Kst.release(); h->yhat.release(); h->sumsq.release(); h->dotf.release();
  h->mse.release(); h->params.release(); h->part_val.release(); h->best_val.release(); h->vals.release();
  h->part_idx.release(); h->best_idx.release();
  h->evs.destroy();
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int b200bo_set_stream(b200bo_handle h, void* s) {
  CHECK_ARG(h, "handle is NULL");
  CU_TRY(cudaSetDevice(h->device));
  CU_TRY(cudaStreamSynchronize(h->stream));
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  h->stream = (cudaStream_t)s;
  h->own_stream = false;
  return 0;
}

int b200bo_set_precision(b200bo_handle h, int prec) {
  CHECK_ARG(h, "handle is NULL");
  CHECK_ARG(prec == B200BO_PREC_FP64, "only B200BO_PREC_FP64 is available in this build");
  h->prec = prec;
  return 0;
}

int b200bo_set_keep_R(b200bo_handle h, int keep) {
  CHECK_ARG(h, "handle is NULL");
  h->keepR = keep != 0;
  return 0;
}

int b200bo_set_train(b200bo_handle h, const double* X, const double* y, int N, int D) {
  CHECK_ARG(h && X && y, "NULL argument");
  CHECK_ARG(N >= 1 && D >= 1, "N and D must be positive");
  CHECK_ARG(D <= 1024, "D > 1024 is not supported");
  CU_TRY(cudaSetDevice(h->device));
  h->N = N;
  h->D = D;
  h->ld = round_up(N, 128);
  h->factored = false;
  const int ld = h->ld;
  std::vector<double> xt((size_t)D * ld, 0.0), yy(ld, 0.0), ff(ld, 0.0);
  for (int i = 0; i < N; ++i) {
    for (int d = 0; d < D; ++d) xt[(size_t)d * ld + i] = X[(size_t)i * D + d];
    yy[i] = y[i];
    ff[i] = 1.0;  // constant trend basis F = ones (trend.py:76-79); zero on padding rows
  }
  CU_TRY(h->Xt.reserve(xt.size()));
  CU_TRY(h->y.reserve(ld));
  CU_TRY(h->F.reserve(ld));
  CU_TRY(h->theta.reserve(D));
  CU_TRY(cudaMemcpyAsync(h->Xt.p, xt.data(), xt.size() * 8, cudaMemcpyHostToDevice, h->stream));
  CU_TRY(cudaMemcpyAsync(h->y.p, yy.data(), ld * 8, cudaMemcpyHostToDevice, h->stream));
  CU_TRY(cudaMemcpyAsync(h->F.p, ff.data(), ld * 8, cudaMemcpyHostToDevice, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  // predict workspaces depend on ld
  h->Kst.release();
  return 0;
}

int b200bo_factor(b200bo_handle h, int corr, const double* theta, int n_theta, int mode, double par_last,
                  double noise_var, int trend, const double* beta_or_null, double* out_llf,
                  double* out_sigma2, double* out_noise_var, int* out_status) {
  CHECK_ARG(h && theta, "NULL argument");
  CHECK_ARG(h->N > 0, "set_train first");
  CHECK_ARG(corr >= 0 && corr <= 5, "unknown correlation id");
  CHECK_ARG(mode >= 0 && mode <= 2, "unknown estimation mode");
  CHECK_ARG(trend == B200BO_TREND_CONSTANT, "only the constant trend is implemented on device");
  CHECK_ARG(n_theta == 1 || n_theta == h->D, "Length of theta must be 1 or D");
  CU_TRY(cudaSetDevice(h->device));
  const int N = h->N, D = h->D, ld = h->ld, nb = ld / NB;
  const size_t nn = (size_t)ld * ld;
  h->factored = false;
  CU_TRY(h->A.reserve(nn));
  CU_TRY(h->W.reserve(nn));
  CU_TRY(h->S.reserve(nn));
  CU_TRY(h->Dinv.reserve((size_t)nb * NB * NB));
  CU_TRY(h->Yt.reserve(ld));
  CU_TRY(h->Ft.reserve(ld));
  CU_TRY(h->rho.reserve(ld));
  CU_TRY(h->gamma.reserve(ld));
  const int gchunks = (ld + 255) / 256;
  CU_TRY(h->part.reserve((size_t)gchunks * ld));
  if (h->keepR) CU_TRY(h->Rkeep.reserve(nn));

  std::vector<double> th(D);
  for (int d = 0; d < D; ++d) th[d] = theta[n_theta == 1 ? 0 : d];
  cudaStream_t st = h->stream;
  h->evs.reset();
  PhaseTimer pt{h};
  int launches = 0;
  CU_TRY(cudaMemcpyAsync(h->theta.p, th.data(), D * 8, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemsetAsync(h->status.p, 0, sizeof(int), st));
  cudaEvent_t e0 = h->evs.get(), e1 = h->evs.get();
  CU_TRY(cudaEventRecord(e0, st));

  // ---- 1. kernel-matrix assembly -------------------------------------------------------------------
  pt.begin(0);
  {
    AssembleArgs a;
    a.Xt = h->Xt.p; a.R = h->A.p; a.theta = h->theta.p;
    a.N = N; a.D = D; a.ld = ld; a.corr = corr; a.mode = mode;
    a.sigma2 = mode == B200BO_MODE_NOISY ? par_last : 0.0;
    a.noise_var = mode == B200BO_MODE_NOISY ? noise_var : 0.0;
    a.alpha = mode == B200BO_MODE_NOISE_ESTIM ? par_last : 1.0;
    size_t smem = ((size_t)2 * D * NB + ((D + 1) & ~1) + NB * 66) * sizeof(double);
    CU_TRY(cudaFuncSetAttribute(kmat_assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kmat_assemble_kernel<<<nb * (nb + 1) / 2, 256, smem, st>>>(a);
    CU_TRY(cudaGetLastError());
    ++launches;
    if (h->keepR) CU_TRY(cudaMemcpyAsync(h->Rkeep.p, h->A.p, nn * 8, cudaMemcpyDeviceToDevice, st));
  }
  pt.end(0);

  // ---- 2. blocked right-looking Cholesky (lower), fp64 DMMA trailing updates -------------------------
  pt.begin(1);
  CU_TRY(cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHOL_DIAG_SMEM));
  for (int jb = 0; jb < nb; ++jb) {
    double* Ajj = h->A.p + (size_t)jb * NB * (ld + 1);
    double* Dj = h->Dinv.p + (size_t)jb * NB * NB;
    chol_diag_kernel<<<1, 256, CHOL_DIAG_SMEM, st>>>(Ajj, ld, Dj, h->status.p);
    CU_TRY(cudaGetLastError());
    ++launches;
    int mrem = ld - (jb + 1) * NB;
    if (mrem > 0) {
      double* P = h->A.p + (size_t)(jb + 1) * NB * ld + (size_t)jb * NB;
      GemmArgs g{};  // panel: P <- P * Ljj^-T   (C(m,n) = sum_k P(m,k) Dinv(n,k))
      g.A = P; g.B = Dj; g.C = P; g.lda = ld; g.ldb = NB; g.ldc = ld; g.K = NB; g.alpha = 1.0; g.beta = 0.0;
      CU_TRY((launch_gemm<GemmNT, false, false>(h, g, mrem, NB, 1)));
      GemmArgs s{};  // trailing update: A22 <- A22 - P P^T, lower tiles only
      s.A = P; s.B = P; s.C = h->A.p + (size_t)(jb + 1) * NB * (ld + 1);
      s.lda = ld; s.ldb = ld; s.ldc = ld; s.K = NB; s.alpha = -1.0; s.beta = 1.0; s.lower_only = 1;
      CU_TRY((launch_gemm<GemmNT, false, false>(h, s, mrem, mrem, 1)));
      launches += 2;
    }
  }
  zero_upper_kernel<<<h->num_sms * 4, 256, 0, st>>>(h->A.p, ld);
  CU_TRY(cudaGetLastError());
  ++launches;
  pt.end(1);

  // ---- 3. L^-1 by recursive doubling: [[A,0],[B,C]]^-1 = [[A^-1,0],[-C^-1 B A^-1, C^-1]] -------------
  pt.begin(2);
  CU_TRY(cudaMemsetAsync(h->W.p, 0, nn * 8, st));
  scatter_dinv_kernel<<<nb, 256, 0, st>>>(h->Dinv.p, h->W.p, ld);
  CU_TRY(cudaGetLastError());
  ++launches;
  for (int s = NB; s < ld; s *= 2) {
    auto merge = [&](int a0, int c, int batch) -> int {
      // T = B * W_A     (c x s) = (c x s)(s x s), W_A lower  -> k >= n0
      GemmArgs g1{};
      g1.A = h->A.p + (size_t)(a0 + s) * ld + a0; g1.lda = ld;
      g1.B = h->W.p + (size_t)a0 * (ld + 1); g1.ldb = ld;
      g1.C = h->S.p + (size_t)(a0 + s) * ld + a0; g1.ldc = ld;
      g1.sA = g1.sB = g1.sC = (long long)2 * s * (ld + 1);
      g1.K = s; g1.alpha = 1.0; g1.beta = 0.0; g1.kb_mode = 1;
      CU_TRY((launch_gemm<GemmNN, false, true>(h, g1, c, s, batch)));
      // W_B = -W_C * T  (c x s) = (c x c)(c x s), W_C lower  -> k < m0 + BM
      GemmArgs g2{};
      g2.A = h->W.p + (size_t)(a0 + s) * (ld + 1); g2.lda = ld;
      g2.B = h->S.p + (size_t)(a0 + s) * ld + a0; g2.ldb = ld;
      g2.C = h->W.p + (size_t)(a0 + s) * ld + a0; g2.ldc = ld;
      g2.sA = g2.sB = g2.sC = (long long)2 * s * (ld + 1);
      g2.K = c; g2.alpha = -1.0; g2.beta = 0.0; g2.ke_mode = 2;
      CU_TRY((launch_gemm<GemmNN, false, true>(h, g2, c, s, batch)));
      launches += 2;
      return 0;
    };
    int gf = ld / (2 * s);
    if (gf > 0) {
      int rc = merge(0, s, gf);
      if (rc) return rc;
    }
    int rem = ld - gf * 2 * s;
    if (rem > s) {
      int rc = merge(gf * 2 * s, rem - s, 1);
      if (rc) return rc;
    }
  }
  pt.end(2);

  // ---- 4. solves: Yt, Ft, rho, gamma and the likelihood scalars ---------------------------------------
  pt.begin(3);
  const int est = beta_or_null == nullptr;
  const double beta_fixed = est ? 0.0 : beta_or_null[0];
  tri_gemv2_kernel<<<(ld + 7) / 8, 256, 0, st>>>(h->W.p, ld, ld, h->y.p, h->F.p, h->Yt.p, h->Ft.p);
  CU_TRY(cudaGetLastError());
  fit_scalars_kernel<<<1, 1024, 0, st>>>(h->A.p, ld, ld, h->Ft.p, h->Yt.p, h->scal.p);
  CU_TRY(cudaGetLastError());
  rho_kernel<<<1, 1024, 0, st>>>(h->Yt.p, h->Ft.p, ld, est, beta_fixed, h->rho.p, h->scal.p);
  CU_TRY(cudaGetLastError());
  tri_gemvT_partial_kernel<<<dim3((ld + 255) / 256, gchunks), 256, 0, st>>>(h->W.p, ld, ld, h->rho.p, h->part.p, 256);
  CU_TRY(cudaGetLastError());
  colsum_partials_kernel<<<(ld + 255) / 256, 256, 0, st>>>(h->part.p, ld, gchunks, h->gamma.p);
  CU_TRY(cudaGetLastError());
  launches += 5;
  pt.end(3);
  CU_TRY(cudaEventRecord(e1, st));

  double sc[5];
  double ft0 = 0;
  int flag = 0;
  CU_TRY(cudaMemcpyAsync(sc, h->scal.p, sizeof sc, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaMemcpyAsync(&ft0, h->Ft.p, 8, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaMemcpyAsync(&flag, h->status.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  h->fit_timings[0] = ms;
  for (int k = 0; k < 4; ++k) h->fit_timings[1 + k] = pt.total(k);
  h->fit_timings[5] = launches;

  const double ff = sc[0], logdet = sc[2], rr = sc[3];
  double llf, s2, nv;
  const double two_pi = 6.283185307179586;
  if (mode == B200BO_MODE_NOISELESS) {  // gpr.py:932-945
    int k = est ? 1 : 0;                // rank(Q Q^T) for the p = 1 basis
    s2 = rr / (N - k);
    nv = 0.0;
    llf = -0.5 * (N * log(two_pi * s2) + 2.0 * logdet + N);
  } else if (mode == B200BO_MODE_NOISE_ESTIM) {  // gpr.py:949-959
    double s2t = rr / N;
    s2 = par_last * s2t;
    nv = (1.0 - par_last) * s2t;
    llf = -0.5 * (N * log(two_pi * s2t) + 2.0 * logdet + N);
  } else {  // gpr.py:963-977
    s2 = par_last;
    nv = noise_var;
    double s2t = s2 + nv;
    llf = -0.5 * (N * log(two_pi * s2t) + 2.0 * logdet + rr / s2t);
  }
  int status = B200BO_FIT_OK;
  if (flag || llf != llf) status = B200BO_FIT_NOT_SPD;
  else if (llf > 0) status = B200BO_FIT_REJECTED;  // gpr.py:981-982
  h->corr = corr; h->mode = mode; h->trend = trend; h->estimate_trend = est; h->n_theta = n_theta;
  h->par_last = par_last;
  h->sigma2 = s2; h->noise_var = nv;
  h->beta = est ? sc[4] : beta_fixed;
  // LAPACK dgeqrf sign convention for the 1x1 R factor: -sign(Ft[0]) * ||Ft||  (gpr.py:805)
  h->G = (ft0 >= 0 ? -1.0 : 1.0) * sqrt(ff);
  h->llf = status == B200BO_FIT_OK ? llf : -INFINITY;
  h->factored = status == B200BO_FIT_OK;
  if (out_llf) *out_llf = h->llf;
  if (out_sigma2) *out_sigma2 = s2;
  if (out_noise_var) *out_noise_var = nv;
  if (out_status) *out_status = status;
  return 0;
}

int b200bo_llf_grad(b200bo_handle h, double* out_grad, int n_par) {
  CHECK_ARG(h && out_grad, "NULL argument");
  if (!h->factored) return set_err(B200BO_E_STATE, "llf_grad before a successful factor()");
  const int N = h->N, D = h->D, ld = h->ld, nb = ld / NB, nt = h->n_theta;
  CHECK_ARG(n_par == nt + (h->mode == B200BO_MODE_NOISELESS ? 0 : 1), "n_par does not match the estimation mode");
  if (!corr_has_dtheta(h->corr))
    return set_err(B200BO_E_ARG, "the reference leaves this kernel's theta-gradient unimplemented (gpr.py:758-768)");
  CU_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  // Rinv = L^-T L^-1 (lower tiles): cho_solve((L, True), eye(N)) of gpr.py:997 as one TN GEMM on L^-1
  GemmArgs g{};
  g.A = h->W.p; g.B = h->W.p; g.C = h->S.p; g.lda = g.ldb = g.ldc = ld; g.K = ld; g.alpha = 1.0; g.beta = 0.0;
  g.lower_only = 1; g.kb_mode = 2;  // W(k,i) = 0 for k < i: start at the row-tile offset (i >= j)
  CU_TRY((launch_gemm<GemmTN, true, true>(h, g, ld, ld, 1)));
  const int ntiles = nb * (nb + 1) / 2, S = D + 4;
  CU_TRY(h->part.reserve((size_t)ntiles * S + S));
  double s2t = h->sigma2 + h->noise_var;
  GradArgs a;
  a.Xt = h->Xt.p; a.theta = h->theta.p; a.Rinv = h->S.p; a.gamma = h->gamma.p; a.partial = h->part.p;
  a.N = N; a.D = D; a.ld = ld; a.corr = h->corr;
  if (h->mode == B200BO_MODE_NOISELESS) { a.a = 1.0 / h->sigma2; a.b = 1.0; }              // gpr.py:1002-1010
  else if (h->mode == B200BO_MODE_NOISE_ESTIM) { a.a = h->par_last / s2t; a.b = h->par_last; }  // :1011-1019
  else { a.a = 1.0 / s2t; a.b = 1.0; }                                                       // :1026-1038 (quirk g2)
  size_t smem = ((size_t)2 * D * NB + ((D + 1) & ~1) + 2 * NB + 8 * S) * sizeof(double);
  CU_TRY(cudaFuncSetAttribute(llf_grad_traces_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  llf_grad_traces_kernel<<<ntiles, 256, smem, st>>>(a);
  CU_TRY(cudaGetLastError());
  double* dout = h->part.p + (size_t)ntiles * S;
  grad_reduce_kernel<<<S, 1024, 0, st>>>(h->part.p, ntiles, S, dout);
  CU_TRY(cudaGetLastError());
  std::vector<double> r(S);
  CU_TRY(cudaMemcpyAsync(r.data(), dout, S * 8, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  const double T1 = r[D], T2 = r[D + 1], trRinv = r[D + 2], gg = r[D + 3];
  double last = 0.0;
  if (h->mode == B200BO_MODE_NOISE_ESTIM) {
    // -0.5 (sum(Rinv * (R0 - I)) - gamma^T (R0 - I) gamma / s2t): off-diagonal pairs only   gpr.py:1021-1025
    last = -(T1 - T2 / s2t);
  } else if (h->mode == B200BO_MODE_NOISY) {
    // -0.5 (sum(Cinv * R0) - gamma_^T R0 gamma_), Cinv = Rinv / s2t, gamma_ = gamma / s2t    gpr.py:1027-1037
    last = -0.5 * ((2.0 * T1 + trRinv) / s2t - (2.0 * T2 + gg) / (s2t * s2t));
  }
  if (nt == D) {
    for (int d = 0; d < D; ++d) out_grad[d] = r[d];
    if (n_par > nt) out_grad[nt] = last;
  } else {
    // Isotropic theta with D > 1.  The reference indexes its (N,N,D[+1]) gradient tensor with the PARAMETER
    // index (gpr.py:1004-1005, :1034-1035), so parameter 0 sees only the feature-0 slice, and in "noisy" mode
    // parameter 1 (sigma2) sees the feature-1 slice instead of R0.  Reproduced as is.
    out_grad[0] = r[0];
    if (n_par > 1) out_grad[1] = (h->mode == B200BO_MODE_NOISY && D > 1) ? r[1] : last;
  }
  return 0;
}

int b200bo_get_state(b200bo_handle h, int what, double* out, size_t n_elems) {
  CHECK_ARG(h && out, "NULL argument");
  if (!(h->factored || (what == B200BO_STATE_R && h->keepR && h->Rkeep.p)))
    return set_err(B200BO_E_STATE, "no successful factor() yet");
  CU_TRY(cudaSetDevice(h->device));
  const int N = h->N, ld = h->ld;
  auto copy_mat = [&](const double* src) -> int {
    CHECK_ARG(n_elems == (size_t)N * N, "expected N*N elements");
    CU_TRY(cudaMemcpy2DAsync(out, (size_t)N * 8, src, (size_t)ld * 8, (size_t)N * 8, N, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    return 0;
  };
  auto copy_vec = [&](const double* src) -> int {
    CHECK_ARG(n_elems == (size_t)N, "expected N elements");
    CU_TRY(cudaMemcpyAsync(out, src, (size_t)N * 8, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    return 0;
  };
  switch (what) {
    case B200BO_STATE_L: return copy_mat(h->A.p);
    case B200BO_STATE_LINV: return copy_mat(h->W.p);
    case B200BO_STATE_R: return copy_mat(h->Rkeep.p);
    case B200BO_STATE_GAMMA: return copy_vec(h->gamma.p);
    case B200BO_STATE_YT: return copy_vec(h->Yt.p);
    case B200BO_STATE_FT: return copy_vec(h->Ft.p);
    case B200BO_STATE_RHO: return copy_vec(h->rho.p);
    case B200BO_STATE_BETA:
      CHECK_ARG(n_elems == 1, "expected 1 element");
      out[0] = h->beta;
      return 0;
    case B200BO_STATE_G:
      CHECK_ARG(n_elems == 1, "expected 1 element");
      out[0] = h->G;
      return 0;
  }
  return set_err(B200BO_E_ARG, "unknown state id");
}

// Shared driver of predict / acq.  acq_id < 0: predict only.
static int run_candidates(b200bo_handle h, const double* Xc, int64_t M, int loc, int eval_mse, double* yhat_out,
                          double* mse_out, int acq_id, int minimize, double plugin, const double* params, int q,
                          double* vals, double* best_val, int64_t* best_idx) {
  CHECK_ARG(h, "handle is NULL");
  if (!h->factored) return set_err(B200BO_E_STATE, "predict before a successful factor()");
  CHECK_ARG(M >= 0, "M < 0");
  CHECK_ARG(loc == B200BO_HOST || loc == B200BO_DEVICE, "bad loc");
  CU_TRY(cudaSetDevice(h->device));
  const bool do_acq = acq_id >= 0;
  const bool dev = loc == B200BO_DEVICE;
  int rc = ensure_predict_ws(h, q, do_acq && vals && !dev);
  if (rc) return rc;
  cudaStream_t st = h->stream;
  const int D = h->D, ld = h->ld, Mc = h->Mc;
  h->evs.reset();
  PhaseTimer pt{h};
  int launches = 0, contract_launches = 0;
  static bool attr_done = false;
  if (!attr_done) {
    CU_TRY(cudaFuncSetAttribute(contract_fp64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PredCore::SMEM_BYTES));
    attr_done = true;
  }
  const size_t ks_smem = ((size_t)KS_ROWS * D + D) * sizeof(double);
  if (ks_smem > 48 * 1024)
    CU_TRY(cudaFuncSetAttribute(kstar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ks_smem));
  if (do_acq) {
    std::vector<double> pr(q);
    for (int c = 0; c < q; ++c) pr[c] = params ? params[c] : 0.0;
    CU_TRY(cudaMemcpyAsync(h->params.p, pr.data(), q * 8, cudaMemcpyHostToDevice, st));
    std::vector<long long> bi(q, -1);
    CU_TRY(cudaMemsetAsync(h->best_val.p, 0, q * 8, st));
    CU_TRY(cudaMemcpyAsync(h->best_idx.p, bi.data(), q * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaStreamSynchronize(st));  // pr / bi are stack-owned
  }
  cudaEvent_t e0 = h->evs.get(), e1 = h->evs.get();
  CU_TRY(cudaEventRecord(e0, st));
  for (int64_t a = 0; a < M; a += Mc) {
    const int m = (int)std::min<int64_t>(Mc, M - a);
    const int mpad = round_up(m, PC_BM);
    const double* xc = Xc + (size_t)a * D;
    if (!dev) {
      CU_TRY(cudaMemcpyAsync(h->Xc.p, xc, (size_t)m * D * 8, cudaMemcpyHostToDevice, st));
      xc = h->Xc.p;
    }
    double* yh = (dev && yhat_out) ? yhat_out + a : h->yhat.p;
    pt.begin(0);
    {
      KstarArgs k;
      k.Xc = xc; k.Xt = h->Xt.p; k.theta = h->theta.p; k.gamma = h->gamma.p;
      k.Kst = eval_mse ? h->Kst.p : nullptr; k.yhat = yh;
      k.M = m; k.N = h->N; k.D = D; k.ld = ld; k.corr = h->corr; k.beta = h->beta;
      kstar_kernel<<<mpad / KS_ROWS, 256, ks_smem, st>>>(k);
      CU_TRY(cudaGetLastError());
      ++launches;
    }
    pt.end(0);
    if (eval_mse) {
      pt.begin(1);
      ContractArgs c;
      c.Kst = h->Kst.p; c.Linv = h->W.p; c.Ft = h->Ft.p; c.sumsq = h->sumsq.p; c.dotf = h->dotf.p; c.ld = ld;
      contract_fp64_kernel<<<mpad / PC_BM, PredCore::NT, PredCore::SMEM_BYTES, st>>>(c);
      CU_TRY(cudaGetLastError());
      ++launches;
      ++contract_launches;
      pt.end(1);
      pt.begin(2);
      AcqArgs g{};
      g.yhat = yh; g.sumsq = h->sumsq.p; g.dotf = h->dotf.p; g.mse_in = nullptr;
      g.mse_out = mse_out ? (dev ? mse_out + a : h->mse.p) : nullptr;
      g.M = m; g.estimate_trend = h->estimate_trend; g.sigma2 = h->sigma2; g.G = h->G;
      g.idx_base = a;
      if (do_acq) {
        g.vals = vals ? (dev ? vals : h->vals.p) : nullptr;
        g.vals_ld = dev ? M : Mc;
        g.vals_off = dev ? a : 0;
        g.params = h->params.p; g.part_val = h->part_val.p; g.part_idx = h->part_idx.p;
        g.acq = acq_id; g.minimize = minimize; g.q = q; g.plugin = plugin;
        int nblk = std::min(h->num_sms, (m + 255) / 256);
        acq_kernel<<<dim3(nblk, q), 256, 0, st>>>(g);
        CU_TRY(cudaGetLastError());
        argmax_merge_kernel<<<(q + 63) / 64, 64, 0, st>>>(h->part_val.p, h->part_idx.p, nblk, q, h->best_val.p, h->best_idx.p);
        CU_TRY(cudaGetLastError());
        launches += 2;
      } else {
        mse_kernel<<<std::min(h->num_sms, (m + 255) / 256), 256, 0, st>>>(g);
        CU_TRY(cudaGetLastError());
        ++launches;
      }
      pt.end(2);
    }
    if (!dev) {
      if (yhat_out) CU_TRY(cudaMemcpyAsync(yhat_out + a, h->yhat.p, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
      if (eval_mse && mse_out) CU_TRY(cudaMemcpyAsync(mse_out + a, h->mse.p, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
      if (do_acq && vals)
        CU_TRY(cudaMemcpy2DAsync(vals + a, (size_t)M * 8, h->vals.p, (size_t)Mc * 8, (size_t)m * 8, q, cudaMemcpyDeviceToHost, st));
    }
  }
  CU_TRY(cudaEventRecord(e1, st));
  if (do_acq) {
    std::vector<long long> bi(q);
    CU_TRY(cudaMemcpyAsync(best_val, h->best_val.p, q * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(bi.data(), h->best_idx.p, q * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    for (int c = 0; c < q; ++c) best_idx[c] = bi[c];
  }
  CU_TRY(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  h->timings[0] = ms;
  for (int k = 0; k < 3; ++k) h->timings[1 + k] = pt.total(k);
  h->timings[4] = contract_launches;
  h->timings[5] = launches;
  h->timings[6] = 0;
  return 0;
}

int b200bo_predict(b200bo_handle h, const double* Xc, int64_t M, int loc, int eval_mse, double* yhat, double* mse) {
  CHECK_ARG(Xc || M == 0, "Xc is NULL");
  CHECK_ARG(yhat, "yhat is NULL");
  CHECK_ARG(!eval_mse || mse, "mse is NULL with eval_mse");
  return run_candidates(h, Xc, M, loc, eval_mse, yhat, eval_mse ? mse : nullptr, -1, 1, 0.0, nullptr, 0, nullptr,
                        nullptr, nullptr);
}

int b200bo_acq(b200bo_handle h, const double* Xc, int64_t M, int loc, int acq_id, int minimize, double plugin,
               const double* params, int q, double* vals, double* best_val, int64_t* best_idx) {
  CHECK_ARG(Xc || M == 0, "Xc is NULL");
  CHECK_ARG(acq_id >= 0 && acq_id <= 3, "unknown acquisition id");
  CHECK_ARG(q >= 1 && q <= 4096, "q out of range");
  CHECK_ARG(best_val && best_idx, "best_val / best_idx are NULL");
  CHECK_ARG(params || acq_id == B200BO_ACQ_EI, "params is NULL");
  return run_candidates(h, Xc, M, loc, 1, nullptr, nullptr, acq_id, minimize, plugin, params, q, vals, best_val,
                        best_idx);
}

int b200bo_acq_from_moments(b200bo_handle h, const double* yhat, const double* mse, int64_t M, int loc, int acq_id,
                            int minimize, double plugin, const double* params, int q, double* vals,
                            double* best_val, int64_t* best_idx) {
  CHECK_ARG(h && yhat && mse, "NULL argument");
  if (!h->factored) return set_err(B200BO_E_STATE, "acquisition before a successful factor()");
  CHECK_ARG(acq_id >= 0 && acq_id <= 3, "unknown acquisition id");
  CHECK_ARG(q >= 1 && q <= 4096, "q out of range");
  CHECK_ARG(best_val && best_idx, "best_val / best_idx are NULL");
  CHECK_ARG(params || acq_id == B200BO_ACQ_EI, "params is NULL");
  CU_TRY(cudaSetDevice(h->device));
  const bool dev = loc == B200BO_DEVICE;
  int rc = ensure_predict_ws(h, q, vals && !dev);
  if (rc) return rc;
  cudaStream_t st = h->stream;
  const int Mc = h->Mc;
  std::vector<double> pr(q);
  for (int c = 0; c < q; ++c) pr[c] = params ? params[c] : 0.0;
  std::vector<long long> bi(q, -1);
  CU_TRY(cudaMemcpyAsync(h->params.p, pr.data(), q * 8, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemsetAsync(h->best_val.p, 0, q * 8, st));
  CU_TRY(cudaMemcpyAsync(h->best_idx.p, bi.data(), q * 8, cudaMemcpyHostToDevice, st));
  for (int64_t a = 0; a < M; a += Mc) {
    const int m = (int)std::min<int64_t>(Mc, M - a);
    const double *yh = yhat + a, *ms = mse + a;
    if (!dev) {
      CU_TRY(cudaMemcpyAsync(h->yhat.p, yh, (size_t)m * 8, cudaMemcpyHostToDevice, st));
      CU_TRY(cudaMemcpyAsync(h->mse.p, ms, (size_t)m * 8, cudaMemcpyHostToDevice, st));
      yh = h->yhat.p;
      ms = h->mse.p;
    }
    AcqArgs g{};
    g.yhat = yh; g.mse_in = ms; g.M = m; g.sigma2 = h->sigma2; g.idx_base = a;
    g.vals = vals ? (dev ? vals : h->vals.p) : nullptr;
    g.vals_ld = dev ? M : Mc; g.vals_off = dev ? a : 0;
    g.params = h->params.p; g.part_val = h->part_val.p; g.part_idx = h->part_idx.p;
    g.acq = acq_id; g.minimize = minimize; g.q = q; g.plugin = plugin;
    int nblk = std::min(h->num_sms, (m + 255) / 256);
    acq_kernel<<<dim3(nblk, q), 256, 0, st>>>(g);
    CU_TRY(cudaGetLastError());
    argmax_merge_kernel<<<(q + 63) / 64, 64, 0, st>>>(h->part_val.p, h->part_idx.p, nblk, q, h->best_val.p, h->best_idx.p);
    CU_TRY(cudaGetLastError());
    if (!dev && vals)
      CU_TRY(cudaMemcpy2DAsync(vals + a, (size_t)M * 8, h->vals.p, (size_t)Mc * 8, (size_t)m * 8, q, cudaMemcpyDeviceToHost, st));
  }
  CU_TRY(cudaMemcpyAsync(best_val, h->best_val.p, q * 8, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaMemcpyAsync(bi.data(), h->best_idx.p, q * 8, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  for (int c = 0; c < q; ++c) best_idx[c] = bi[c];
  return 0;
}

int b200bo_get_timings(b200bo_handle h, double* out, int n) {
  CHECK_ARG(h && out && n >= 1, "bad argument");
  for (int i = 0; i < n && i < B200BO_N_TIMINGS; ++i) out[i] = h->timings[i];
  return 0;
}
int b200bo_get_fit_timings(b200bo_handle h, double* out, int n) {
  CHECK_ARG(h && out && n >= 1, "bad argument");
  for (int i = 0; i < n && i < B200BO_N_TIMINGS; ++i) out[i] = h->fit_timings[i];
  return 0;
}

}  // extern "C"
