// trend_kernels.cuh -- linear / quadratic basis-expansion trends (p > 1 regression functions), float64.
//
// Reference: surrogate/gaussian_process/trend.py:94-142 (linear: [1, x]; quadratic: [1, x, {x_k x_j, j >= k}]),
// _compute_aux_var gpr.py:800-808 (Ft = L^-1 F, thin QR, rho), predict gpr.py:486-510:
//   yhat = f(x)^T beta + r gamma                                   :490
//   u = G^-T (Ft^T rt - f(x)),  MSE = sigma2 (1 - sum rt^2 + sum u^2), clipped at 0    :496-510
// Ft^T rt = (L^-T Ft)^T r = FV^T r, so the p dot products come from ONE GEMM of the k* block against FV (N x p) and
// no second triangular solve is needed.  The constant trend (p = 1) keeps its scalar fast path elsewhere.
#pragma once
#include <cuda_runtime.h>

namespace b2 {

constexpr int TR_PMAX = 64;  // basis functions (padded column count of F / Ft / FV)

__host__ __device__ inline int trend_p(int trend, int D) { return trend == 0 ? 1 : trend == 1 ? D + 1 : (D + 1) * (D + 2) / 2; }

// j-th basis function at x (strided access: x[d * stride])
__device__ __forceinline__ double trend_basis_at(int trend, const double* x, size_t stride, int D, int j) {
  if (j == 0) return 1.0;
  if (j <= D) return x[(size_t)(j - 1) * stride];
  // quadratic block: for k = 0..D-1: x_k * x_j, j = k..D-1      trend.py:133-134
  int t = j - D - 1;
  int k = 0;
  while (t >= D - k) {
    t -= D - k;
    ++k;
  }
  return x[(size_t)k * stride] * x[(size_t)(k + t) * stride];
}

// F (ld, TR_PMAX) row-major from the transposed training set; zero on padding rows / columns
__global__ void trend_basis_kernel(const double* __restrict__ Xt, int N, int D, int ld, int trend, int p, double* __restrict__ F) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ld * TR_PMAX) return;
  const int row = i / TR_PMAX, col = i % TR_PMAX;
  F[i] = (row < N && col < p) ? trend_basis_at(trend, Xt + row, (size_t)ld, D, col) : 0.0;
}

struct TrendMseArgs {
  const double* Xc;     // (M, D)
  const double* Vt;     // (Mpad, TR_PMAX)  FV^T r = Ft^T rt
  const double* sumsq;  // (M,)
  const double* Gm;     // (TR_PMAX, TR_PMAX) row-major upper-triangular R factor of the thin QR of Ft
  const double* beta;   // (TR_PMAX,)
  double* yhat;         // (M,) in: r gamma; out: + f(x)^T beta
  double* mse;          // (M,) or NULL
  int M, D, trend, p, estimate_trend;
  double sigma2;
};

__global__ void __launch_bounds__(128) trend_mse_kernel(TrendMseArgs a) {
  extern __shared__ double sm[];
  double* G = sm;                       // [p][p]
  double* be = G + a.p * a.p;           // [p]
  for (int e = threadIdx.x; e < a.p * a.p; e += blockDim.x) G[e] = a.Gm[(e / a.p) * TR_PMAX + e % a.p];
  for (int e = threadIdx.x; e < a.p; e += blockDim.x) be[e] = a.beta[e];
  __syncthreads();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= a.M) return;
  const double* x = a.Xc + (size_t)m * a.D;
  double u[TR_PMAX];
  double fb = 0.0, uu = 0.0;
  for (int j = 0; j < a.p; ++j) {
    const double f = trend_basis_at(a.trend, x, 1, a.D, j);
    fb += f * be[j];
    if (a.estimate_trend && a.mse) {
      // forward substitution with G^T (lower): u_j = (v_j - sum_{i<j} G_ij u_i) / G_jj        gpr.py:496-498
      double v = a.Vt[(size_t)m * TR_PMAX + j] - f;
      for (int i = 0; i < j; ++i) v -= G[i * a.p + j] * u[i];
      u[j] = v / G[j * a.p + j];
      uu += u[j] * u[j];
    }
  }
  a.yhat[m] += fb;                                             // gpr.py:490
  if (a.mse) {
    const double v = (1.0 - a.sumsq[m] + uu) * a.sigma2;      // gpr.py:502-505
    a.mse[m] = v < 0.0 ? 0.0 : v;                              // gpr.py:510
  }
}

}  // namespace b2
