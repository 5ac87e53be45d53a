// fit_kernels.cuh -- device side of the fixed-hyper-parameter fit:
//   kmat_assemble   R (N,N)                       gpr.py:772-782 (+ :935 / :952 / :966-967 per mode)
//   chol_diag       64x64 diagonal factor + inverse (the POTF2 / TRTI2 leaves of the blocked algorithms)
//   (panel / SYRK / TRTRI merges are launches of dgemm_kernel, see b200bo.cu)
//   gemv kernels + scalar reductions   Yt, Ft, rho, gamma, log|L|, rho^T rho   gpr.py:795-811, :784-788
// Paths are relative to /root/reference/bayes_optim/surrogate/gaussian_process/.
#pragma once
#include <cuda_runtime.h>

#include "gp_math.h"

namespace b2 {

constexpr int NB = 64;  // Cholesky / TRTRI block size
constexpr size_t CHOL_DIAG_SMEM = (2 * NB * (NB + 1) + 32 * 33 + 3 * NB) * sizeof(double);

// ---------------------------------------------------------------------------------------------------
// Kernel-matrix assembly.  One CTA = one 64x64 tile (ti >= tj) of the lower triangle, mirrored into the
// upper triangle through shared memory so both global writes are row-contiguous 16-byte stores.
// Xt is the transposed training set (D x ld), so the loads are coalesced along the point index.
// The pairwise-distance table of l1_cross_distances (gpr.py:48-61) is never materialised.
//   off-diagonal: noiseless r | noisy (sigma2*r)/(sigma2+tau2) | noise_estim alpha*r
//   diagonal    : 1           | (sigma2+tau2)/(sigma2+tau2)    | alpha + (1-alpha)
//   rows/cols >= N (padding up to the tile multiple): identity.
// ---------------------------------------------------------------------------------------------------
struct AssembleArgs {
  const double* Xt;  // (D, ld)
  double* R;         // (ld, ld)
  const double* theta;  // (D,)
  int N, D, ld, corr, mode;
  double sigma2, noise_var, alpha;
};

__global__ void __launch_bounds__(256) kmat_assemble_kernel(AssembleArgs p) {
  // triangular tile index -> (ti, tj), ti >= tj
  int t = blockIdx.x;
  int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  while (ti * (ti + 1) / 2 > t) --ti;
  int tj = t - ti * (ti + 1) / 2;
  const int i0 = ti * NB, j0 = tj * NB;
  extern __shared__ __align__(16) double sm[];
  double* xi = sm;                    // [D][64]
  double* xj = xi + p.D * NB;         // [D][64]
  double* th = xj + p.D * NB;         // [D]
  double* tile = th + ((p.D + 1) & ~1);  // [64][66]
  const int tid = threadIdx.x;
  for (int e = tid; e < p.D * NB; e += 256) {
    int d = e / NB, c = e % NB;
    xi[e] = p.Xt[(size_t)d * p.ld + i0 + c];
    xj[e] = p.Xt[(size_t)d * p.ld + j0 + c];
  }
  for (int d = tid; d < p.D; d += 256) th[d] = p.theta[d];
  const double pw = corr_has_extra_param(p.corr) ? p.theta[p.D] : 0.0;  // exponent (generalized_exponential) / nu (general Matern)
  __syncthreads();
  // thread -> 4 rows x 4 cols (cols interleaved by 16 so smem reads of xj are conflict-free)
  const int tr = (tid / 16) * 4, tc = tid % 16;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = corr_init(p.corr);
  for (int d = 0; d < p.D; ++d) {
    double xa[4], xb[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) xa[a] = xi[d * NB + tr + a];
#pragma unroll
    for (int b = 0; b < 4; ++b) xb[b] = xj[d * NB + tc + 16 * b];
    double thd = th[d];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = corr_accum_p(p.corr, acc[a][b], thd, xa[a] - xb[b], pw);
  }
  const double s2t = p.sigma2 + p.noise_var;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int gi = i0 + tr + a, gj = j0 + tc + 16 * b;
      double v;
      if (gi >= p.N || gj >= p.N) {
        v = gi == gj ? 1.0 : 0.0;
      } else if (gi == gj) {
        v = p.mode == 1 ? (p.sigma2 * 1.0 + p.noise_var) / s2t : (p.mode == 2 ? p.alpha * 1.0 + (1.0 - p.alpha) : 1.0);
      } else {
        double r = corr_finish_p(p.corr, acc[a][b], pw);
        v = p.mode == 1 ? (p.sigma2 * r) / s2t : (p.mode == 2 ? p.alpha * r : r);
      }
      tile[(tr + a) * 66 + tc + 16 * b] = v;
    }
  __syncthreads();
  // lower tile: rows i0.., cols j0..  -- each thread stores 16-byte pairs, a warp covers one 512 B row
  for (int e = tid; e < NB * NB / 2; e += 256) {
    int r = e / (NB / 2), c = (e % (NB / 2)) * 2;
    double2 v = make_double2(tile[r * 66 + c], tile[r * 66 + c + 1]);
    *reinterpret_cast<double2*>(p.R + (size_t)(i0 + r) * p.ld + j0 + c) = v;
  }
  if (ti != tj) {  // mirrored tile: rows j0.., cols i0..
    for (int e = tid; e < NB * NB / 2; e += 256) {
      int r = e / (NB / 2), c = (e % (NB / 2)) * 2;
      double2 v = make_double2(tile[c * 66 + r], tile[(c + 1) * 66 + r]);
      *reinterpret_cast<double2*>(p.R + (size_t)(j0 + r) * p.ld + i0 + c) = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Appending m <= 64 training points to a factored model (b200bo_append): the m new ROWS of R, mode scaling as
// in kmat_assemble_kernel.  Columns j < N0 (correlations with the old points) go to Tr (64, ld), zero-padded;
// the (m, m) block among the new points goes to Cb (64, 64) with an identity pad.  grid = ld / 256, 256 threads.
// ---------------------------------------------------------------------------------------------------
struct AppendRowsArgs {
  const double* Xt;     // (D, ld), already holding the new points in columns N0 .. N0 + m - 1
  const double* theta;  // (D [+1])
  double* Tr;           // (64, ld)
  double* Cb;           // (64, 64)
  int N0, m, D, ld, corr, mode;
  double sigma2, noise_var, alpha;
};

__global__ void __launch_bounds__(256) kmat_append_rows_kernel(AppendRowsArgs p) {
  extern __shared__ __align__(16) double sm_ar[];
  double* xn = sm_ar;              // [64][D] new points
  double* th = xn + NB * p.D;      // [D]
  const int tid = threadIdx.x;
  for (int e = tid; e < NB * p.D; e += 256) {
    const int i = e / p.D, d = e % p.D;
    xn[e] = i < p.m ? p.Xt[(size_t)d * p.ld + p.N0 + i] : 0.0;
  }
  for (int d = tid; d < p.D; d += 256) th[d] = p.theta[d];
  const double pw = corr_has_extra_param(p.corr) ? p.theta[p.D] : 0.0;
  __syncthreads();
  const int j = blockIdx.x * 256 + tid;  // column of R; the (64, 64) block Cb spans columns N0 .. N0 + 63, which may
  if (j >= max(p.ld, p.N0 + NB)) return;  // reach past the pitch (its identity padding must still be written)
  const double s2t = p.sigma2 + p.noise_var;
  const double diag = p.mode == 1 ? (p.sigma2 + p.noise_var) / s2t : (p.mode == 2 ? p.alpha + (1.0 - p.alpha) : 1.0);
  for (int i = 0; i < NB; ++i) {
    const int gi = p.N0 + i;
    double v = 0.0;
    if (i < p.m && j <= gi) {
      if (j == gi) {
        v = diag;
      } else {
        double acc = corr_init(p.corr);
        for (int d = 0; d < p.D; ++d) acc = corr_accum_p(p.corr, acc, th[d], xn[i * p.D + d] - p.Xt[(size_t)d * p.ld + j], pw);
        const double r = corr_finish_p(p.corr, acc, pw);
        v = p.mode == 1 ? (p.sigma2 * r) / s2t : (p.mode == 2 ? p.alpha * r : r);
      }
    }
    if (j < p.ld) p.Tr[(size_t)i * p.ld + j] = j < p.N0 ? v : 0.0;
    const int c = j - p.N0;
    if (c >= 0 && c < NB) p.Cb[i * NB + c] = (i < p.m && c < p.m) ? (c <= i ? v : 0.0) : (i == c ? 1.0 : 0.0);
  }
}

// symmetrise the (64, 64) block before its Cholesky update (the GEMM S S^T fills every entry; only the lower part of
// the R block was written)
__global__ void append_sym_kernel(double* __restrict__ Cb, int m) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= NB * NB) return;
  const int i = e / NB, c = e % NB;
  if (c > i && i < m && c < m) Cb[e] = Cb[c * NB + i];
}

// rows N0 .. N0 + m - 1 of L and L^-1 from the pieces: [S | L22 | 0] and [W21 | L22^-1 | 0]
__global__ void __launch_bounds__(256) append_scatter_kernel(double* __restrict__ A, double* __restrict__ W, int ld, int N0, int m,
                                                             const double* __restrict__ S, const double* __restrict__ W21,
                                                             const double* __restrict__ L22, const double* __restrict__ L22inv) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= ld) return;
  for (int i = 0; i < m; ++i) {
    double a = 0.0, w = 0.0;
    if (j < N0) {
      a = S[(size_t)i * ld + j];
      w = W21[(size_t)i * ld + j];
    } else if (j - N0 <= i) {
      a = L22[i * NB + (j - N0)];
      w = L22inv[i * NB + (j - N0)];
    }
    A[(size_t)(N0 + i) * ld + j] = a;
    W[(size_t)(N0 + i) * ld + j] = w;
  }
}

// new (ld1, ld1) buffer from an (ld0, ld0) one: the leading (n, n) block is copied, the rest is the identity
__global__ void grow_matrix_kernel(const double* __restrict__ src, int ld0, double* __restrict__ dst, int ld1, int n) {
  const size_t total = (size_t)ld1 * ld1;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / ld1), j = (int)(e % ld1);
    dst[e] = (i < n && j < n) ? src[(size_t)i * ld0 + j] : (i == j ? 1.0 : 0.0);
  }
}

// ---------------------------------------------------------------------------------------------------
// Diagonal block: factor A[jb,jb] = Ljj Ljj^T in shared memory, write Ljj back (lower; the strict upper
// triangle of the block is zeroed) and its inverse to Dinv (64x64 row-major, zero strict upper).
// status[0] |= 1 when a pivot is <= 0 or NaN (scipy.linalg.cholesky raises LinAlgError there, gpr.py:795).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) chol_diag_kernel(double* __restrict__ A, int ld, double* __restrict__ Dinv,
                                                        int* __restrict__ status) {
  // 64 x 64 diagonal block: factor + inverse, one CTA.  The sequential part is kept to ONE barrier and one
  // reciprocal per column: the trailing update uses the unscaled column and 1/d_j, the 1/sqrt(d_j) scaling of
  // all columns is applied once at the end, and the inverse is built by block recursion 16 -> 32 -> 64.
  extern __shared__ __align__(16) double sm_cd[];
  double(*s)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(sm_cd);
  double(*x)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(sm_cd + NB * (NB + 1));  // inverse, x[r][c]
  double(*t)[33] = reinterpret_cast<double(*)[33]>(sm_cd + 2 * NB * (NB + 1));     // 32 x 32 scratch
  double* rs = sm_cd + 2 * NB * (NB + 1) + 32 * 33;                                // 1 / sqrt(d_c) = 1 / L_cc
  const int tid = threadIdx.x;
  double* colbuf = rs + NB;  // [2][NB] column j of the running factorisation, double-buffered
  // thread (tr, tc) keeps the 16 elements (tr + 16 i, tc + 16 k) in registers for the whole factorisation
  const int tr = tid >> 4, tc = tid & 15;
  double a[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) a[i][k] = A[(size_t)(tr + 16 * i) * ld + tc + 16 * k];
  for (int e = tid; e < NB * NB; e += 256) x[e / NB][e % NB] = 0.0;
  // ---- right-looking factorisation on unscaled columns: a[r][c] -= a[r][j] a[c][j] / d_j  (r >= c > j) ----
#pragma unroll
  for (int kj = 0; kj < 4; ++kj) {  // unrolled so that a[][] is indexed statically (stays in registers)
#pragma unroll 1
    for (int jj = 0; jj < 16; ++jj) {
      const int j = 16 * kj + jj;
      double* cb = colbuf + (j & 1) * NB;
      if (tc == jj) {
#pragma unroll
        for (int i = 0; i < 4; ++i) cb[tr + 16 * i] = a[i][kj];
      }
      __syncthreads();
      const double dinv = __drcp_rn(cb[j]);
      double li[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) li[i] = cb[tr + 16 * i] * dinv;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < kj) continue;  // columns of earlier groups are final
        const int c = tc + 16 * k;
        if (c > j) {
          const double lc = cb[c];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (tr + 16 * i >= c) a[i][k] = fma(-li[i], lc, a[i][k]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) s[tr + 16 * i][tc + 16 * k] = a[i][k];
  __syncthreads();
  if (tid < NB) {
    const double d = s[tid][tid];
    // scipy's cholesky raises LinAlgError on a pivot <= 0 -> llf = -inf (gpr.py:946).  R has a unit diagonal in
    // every estimation mode, so a pivot below 16 eps has lost all its digits: exactly singular matrices (duplicate
    // points without nugget), whose computed pivot is +-O(eps) by rounding luck, are rejected deterministically.
    if (!(d > 16.0 * 2.220446049250313e-16)) atomicOr(status, 1);
    rs[tid] = 1.0 / sqrt(d);
  }
  __syncthreads();
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e / NB, c = e % NB;
    if (r > c) s[r][c] *= rs[c];
    else if (r == c) s[r][c] = s[r][c] * rs[c];  // d / sqrt(d)
    else s[r][c] = 0.0;
  }
  __syncthreads();
  // ---- inverse: four 16 x 16 diagonal blocks by forward substitution (one thread per column) ----
  if (tid < NB) {
    const int c = tid, b0 = c & ~15;
    x[c][c] = rs[c];
    for (int r = c + 1; r < b0 + 16; ++r) {
      double acc = 0.0;
      for (int k = c; k < r; ++k) acc += s[r][k] * x[k][c];
      x[r][c] = -acc * rs[r];
    }
  }
  __syncthreads();
  // ---- merges: [[A,0],[B,C]]^-1 = [[A^-1,0],[-C^-1 B A^-1, C^-1]] for half = 16, then 32 ----
  for (int half = 16; half < NB; half *= 2) {
    const int nblk = NB / (2 * half);  // independent merges at this level
    const int per = half * half;
    // T = B A^-1   (A^-1 lower: k >= col)
    for (int e = tid; e < nblk * per; e += 256) {
      const int q = e / per, i = (e % per) / half, jj = e % half;
      const int o = q * 2 * half;
      double acc = 0.0;
      for (int k = jj; k < half; ++k) acc += s[o + half + i][o + k] * x[o + k][o + jj];
      t[q * half + i][jj] = acc;  // (nblk * half) x half <= 32 x 32
    }
    __syncthreads();
    // X_B = -C^-1 T   (C^-1 lower: k <= row)
    for (int e = tid; e < nblk * per; e += 256) {
      const int q = e / per, i = (e % per) / half, jj = e % half;
      const int o = q * 2 * half;
      double acc = 0.0;
      for (int k = 0; k <= i; ++k) acc += x[o + half + i][o + half + k] * t[q * half + k][jj];
      x[o + half + i][o + jj] = -acc;
    }
    __syncthreads();
  }
  for (int e = tid; e < NB * NB; e += 256) {
    int r = e / NB, c = e % NB;
    A[(size_t)r * ld + c] = s[r][c];
    Dinv[e] = x[r][c];
  }
}

// ---------------------------------------------------------------------------------------------------
// Split form of the diagonal step for the look-ahead Cholesky: only the FACTORISATION of the 64 x 64 block and the
// panel solve against it sit on the critical path; the block inverse (level 0 of the recursive L^-1) runs on a helper
// stream whenever the factor is there.
// ---------------------------------------------------------------------------------------------------
constexpr size_t CHOL_FACTOR_SMEM = (NB * (NB + 1) + 3 * NB) * sizeof(double);

// factor A[jb,jb] = Ljj Ljj^T, write Ljj back (strict upper triangle zeroed).  Same arithmetic as chol_diag_kernel.
__global__ void __launch_bounds__(256) chol_factor_kernel(double* __restrict__ A, int ld, int* __restrict__ status) {
  extern __shared__ __align__(16) double sm_cf[];
  double(*s)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(sm_cf);
  double* rs = sm_cf + NB * (NB + 1);
  double* colbuf = rs + NB;  // [2][NB]
  const int tid = threadIdx.x;
  const int tr = tid >> 4, tc = tid & 15;
  double a[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) a[i][k] = A[(size_t)(tr + 16 * i) * ld + tc + 16 * k];
#pragma unroll
  for (int kj = 0; kj < 4; ++kj) {
#pragma unroll 1
    for (int jj = 0; jj < 16; ++jj) {
      const int j = 16 * kj + jj;
      double* cb = colbuf + (j & 1) * NB;
      if (tc == jj) {
#pragma unroll
        for (int i = 0; i < 4; ++i) cb[tr + 16 * i] = a[i][kj];
      }
      __syncthreads();
      const double dinv = __drcp_rn(cb[j]);
      double li[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) li[i] = cb[tr + 16 * i] * dinv;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < kj) continue;
        const int c = tc + 16 * k;
        if (c > j) {
          const double lc = cb[c];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (tr + 16 * i >= c) a[i][k] = fma(-li[i], lc, a[i][k]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) s[tr + 16 * i][tc + 16 * k] = a[i][k];
  __syncthreads();
  if (tid < NB) {
    const double d = s[tid][tid];
    if (!(d > 16.0 * 2.220446049250313e-16)) atomicOr(status, 1);  // see chol_diag_kernel
    rs[tid] = 1.0 / sqrt(d);
  }
  __syncthreads();
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e / NB, c = e % NB;
    A[(size_t)r * ld + c] = r >= c ? s[r][c] * rs[c] : 0.0;
  }
}

// inverse of the factored diagonal block (16 -> 32 -> 64 block recursion, as in chol_diag_kernel) -> Dinv
__global__ void __launch_bounds__(256) chol_inverse_kernel(const double* __restrict__ A, int ld, double* __restrict__ Dinv) {
  extern __shared__ __align__(16) double sm_ci[];
  double(*s)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(sm_ci);
  double(*x)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(sm_ci + NB * (NB + 1));
  double(*t)[33] = reinterpret_cast<double(*)[33]>(sm_ci + 2 * NB * (NB + 1));
  double* rs = sm_ci + 2 * NB * (NB + 1) + 32 * 33;
  const int tid = threadIdx.x;
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e / NB, c = e % NB;
    s[r][c] = A[(size_t)r * ld + c];
    x[r][c] = 0.0;
  }
  __syncthreads();
  if (tid < NB) rs[tid] = 1.0 / s[tid][tid];
  __syncthreads();
  if (tid < NB) {
    const int c = tid, b0 = c & ~15;
    x[c][c] = rs[c];
    for (int r = c + 1; r < b0 + 16; ++r) {
      double acc = 0.0;
      for (int k = c; k < r; ++k) acc += s[r][k] * x[k][c];
      x[r][c] = -acc * rs[r];
    }
  }
  __syncthreads();
  for (int half = 16; half < NB; half *= 2) {
    const int nblk = NB / (2 * half);
    const int per = half * half;
    for (int e = tid; e < nblk * per; e += 256) {
      const int q = e / per, i = (e % per) / half, jj = e % half;
      const int o = q * 2 * half;
      double acc = 0.0;
      for (int k = jj; k < half; ++k) acc += s[o + half + i][o + k] * x[o + k][o + jj];
      t[q * half + i][jj] = acc;
    }
    __syncthreads();
    for (int e = tid; e < nblk * per; e += 256) {
      const int q = e / per, i = (e % per) / half, jj = e % half;
      const int o = q * 2 * half;
      double acc = 0.0;
      for (int k = 0; k <= i; ++k) acc += x[o + half + i][o + half + k] * t[q * half + k][jj];
      x[o + half + i][o + jj] = -acc;
    }
    __syncthreads();
  }
  for (int e = tid; e < NB * NB; e += 256) Dinv[e] = x[e / NB][e % NB];
}

// panel solve P <- P Ljj^-T by substitution, one thread per row of P (rows are independent; right-looking form so the
// 63 - c updates after each pivot are independent FMAs).  grid = rows / 64.
constexpr size_t PANEL_TRSM_SMEM = (NB * (NB + 1) + NB) * sizeof(double);
__global__ void __launch_bounds__(64) panel_trsm_kernel(double* __restrict__ P, int ld, const double* __restrict__ Ljj) {
  extern __shared__ __align__(16) double sm_pt[];
  double(*L)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(sm_pt);
  double* rd = sm_pt + NB * (NB + 1);
  const int tid = threadIdx.x;
  for (int e = tid; e < NB * NB; e += 64) L[e % NB][e / NB] = Ljj[(size_t)(e / NB) * ld + e % NB];  // L^T: L[c][j] = Ljj[j][c]
  __syncthreads();
  if (tid < NB) rd[tid] = 1.0 / L[tid][tid];
  __syncthreads();
  double* row = P + (size_t)(blockIdx.x * 64 + tid) * ld;
  double x[NB];
#pragma unroll
  for (int c = 0; c < NB; c += 2) {
    const double2 v = *reinterpret_cast<const double2*>(row + c);
    x[c] = v.x;
    x[c + 1] = v.y;
  }
#pragma unroll
  for (int c = 0; c < NB; ++c) {
    x[c] *= rd[c];
#pragma unroll
    for (int j = c + 1; j < NB; ++j) x[j] = fma(-x[c], L[c][j], x[j]);  // L[c][j] = Ljj[j][c]: broadcast read
  }
#pragma unroll
  for (int c = 0; c < NB; c += 2) *reinterpret_cast<double2*>(row + c) = make_double2(x[c], x[c + 1]);
}

// ---------------------------------------------------------------------------------------------------
// One panel step of the look-ahead Cholesky as ONE kernel: the latency chain "update the next block column ->
// factor its diagonal block -> solve the panel against it" used to be three dependent launches per panel.
// CTA b of step j owns row block b of panel j (global row block j + 1 + b) and does, entirely on chip:
//   (a) the step-(j-1) update of what it needs of block column j:   Dg -= Q0 Q0^T,  Xb -= Qb Q0^T
//       (Q* = rows of the previous solved panel; the diagonal block Dg is updated redundantly by every CTA),
//   (f) the factorisation Dg = L L^T (redundantly: 64 sequential columns either way),
//   (s) its own rows  Pb = Xb L^-T  by substitution (one thread per row).
// Nothing a neighbour still has to read is overwritten: the factor goes to Lside[j] and the first solved row block to
// P0side[j] (chol_finish_kernel moves both into A once the step is over); row blocks b >= 1 go straight into A.
// ---------------------------------------------------------------------------------------------------
constexpr int PC_TS = NB * (NB + 1);  // one padded 64 x 64 tile
constexpr size_t PANEL_CHAIN_SMEM = (4 * PC_TS + 3 * NB) * sizeof(double);

__global__ void __launch_bounds__(256, 1)
panel_chain_kernel(double* __restrict__ A, int ld, int jb, int nrows, double* __restrict__ Lside,
                   double* __restrict__ P0side, int* __restrict__ status) {
  extern __shared__ __align__(16) double sm_pc[];
  double(*sD)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(sm_pc);               // Dg -> L
  double(*sQ0)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(sm_pc + PC_TS);      // previous panel, rows of block j
  double(*sX)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(sm_pc + 2 * PC_TS);   // own rows of panel j
  double(*sQ)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(sm_pc + 3 * PC_TS);   // previous panel, own rows
  double* rs = sm_pc + 4 * PC_TS;
  double* colbuf = rs + NB;  // [2][NB]
  const int tid = threadIdx.x, b = blockIdx.x;
  const bool has_rows = b < nrows;   // the last step has no panel: one CTA factors the last diagonal block
  const bool has_prev = jb > 0;
  const double* Dg = A + (size_t)jb * NB * (ld + 1);
  double* Xb = A + (size_t)(jb + 1 + b) * NB * ld + (size_t)jb * NB;
  const double* Q0 = P0side + (size_t)(jb - 1) * NB * NB;                                   // (64, 64) row-major
  const double* Qb = A + (size_t)(jb + 1 + b) * NB * ld + (size_t)(jb - 1) * NB;
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e / NB, c = e % NB;
    sD[r][c] = Dg[(size_t)r * ld + c];
    if (has_prev) sQ0[r][c] = Q0[e];
    if (has_rows) {
      sX[r][c] = Xb[(size_t)r * ld + c];
      if (has_prev) sQ[r][c] = Qb[(size_t)r * ld + c];
    }
  }
  __syncthreads();
  const int tr = tid >> 4, tc = tid & 15;
  // ---- (a) + register tile of the diagonal block: thread (tr, tc) holds (tr + 16 i, tc + 16 k)
  double a[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) a[i][k] = sD[tr + 16 * i][tc + 16 * k];
  if (has_prev) {
    double x[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) x[i][k] = has_rows ? sX[tr + 16 * i][tc + 16 * k] : 0.0;
#pragma unroll 4
    for (int kk = 0; kk < NB; ++kk) {
      double qr[4], qc[4], qx[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        qr[i] = sQ0[tr + 16 * i][kk];
        qc[i] = sQ0[tc + 16 * i][kk];
        qx[i] = has_rows ? sQ[tr + 16 * i][kk] : 0.0;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          a[i][k] = fma(-qr[i], qc[k], a[i][k]);
          x[i][k] = fma(-qx[i], qc[k], x[i][k]);
        }
    }
    __syncthreads();
    if (has_rows) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) sX[tr + 16 * i][tc + 16 * k] = x[i][k];
    }
  }
  // ---- (f) right-looking factorisation on unscaled columns (chol_factor_kernel's arithmetic)
#pragma unroll
  for (int kj = 0; kj < 4; ++kj) {
#pragma unroll 1
    for (int jj = 0; jj < 16; ++jj) {
      const int j = 16 * kj + jj;
      double* cb = colbuf + (j & 1) * NB;
      if (tc == jj) {
#pragma unroll
        for (int i = 0; i < 4; ++i) cb[tr + 16 * i] = a[i][kj];
      }
      __syncthreads();
      const double dinv = __drcp_rn(cb[j]);
      double li[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) li[i] = cb[tr + 16 * i] * dinv;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < kj) continue;
        const int c = tc + 16 * k;
        if (c > j) {
          const double lc = cb[c];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (tr + 16 * i >= c) a[i][k] = fma(-li[i], lc, a[i][k]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) sD[tr + 16 * i][tc + 16 * k] = a[i][k];
  __syncthreads();
  if (tid < NB) {
    const double d = sD[tid][tid];
    if (b == 0 && !(d > 16.0 * 2.220446049250313e-16)) atomicOr(status, 1);  // see chol_diag_kernel
    rs[tid] = 1.0 / sqrt(d);
  }
  __syncthreads();
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e / NB, c = e % NB;
    const double v = r >= c ? sD[r][c] * rs[c] : 0.0;
    sD[r][c] = v;
    if (b == 0) Lside[(size_t)jb * NB * NB + e] = v;
  }
  __syncthreads();
  if (!has_rows) return;
  // ---- (s) own rows: x L^T = p, right-looking.  Four threads per row (columns 4 m + q, interleaved so the work stays
  // balanced as the pivot advances); the pivot value travels by shuffle.  One thread per row would be 2016 straight-line
  // FMAs executed once -- instruction-fetch bound (measured 18 us for the 64 x 64 block).
  {
    const int r = tid >> 2, q = tid & 3, lane = tid & 31;
    double x[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) x[m] = sX[r][4 * m + q];
#pragma unroll
    for (int c = 0; c < NB; ++c) {
      const int mc = c >> 2, qc = c & 3;
      const double xc = __shfl_sync(0xffffffffu, x[mc] * rs[c], (lane & ~3) | qc);   // 1 / L_cc = rs[c]
      if (q == qc) x[mc] = xc;
      if (q > qc) x[mc] = fma(-xc, sD[4 * mc + q][c], x[mc]);
#pragma unroll
      for (int m = mc + 1; m < 16; ++m) x[m] = fma(-xc, sD[4 * m + q][c], x[m]);
    }
#pragma unroll
    for (int m = 0; m < 16; ++m) sX[r][4 * m + q] = x[m];
  }
  __syncthreads();
  double* out = b == 0 ? P0side + (size_t)jb * NB * NB : Xb;
  const int old = b == 0 ? NB : ld;
  for (int e = tid; e < NB * NB; e += 256) out[(size_t)(e / NB) * old + e % NB] = sX[e / NB][e % NB];
}

// after a panel step: the factor and the first solved row block move into A, and the block inverse (level 0 of the
// recursive L^-1) is formed -- all off the critical path
__global__ void __launch_bounds__(256) chol_finish_kernel(double* __restrict__ A, int ld, int jb, int nrows,
                                                          const double* __restrict__ Lside, const double* __restrict__ P0side,
                                                          double* __restrict__ Dinv) {
  extern __shared__ __align__(16) double sm_cf2[];
  double(*s)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(sm_cf2);
  double(*x)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(sm_cf2 + NB * (NB + 1));
  double(*t)[33] = reinterpret_cast<double(*)[33]>(sm_cf2 + 2 * NB * (NB + 1));
  double* rs = sm_cf2 + 2 * NB * (NB + 1) + 32 * 33;
  const int tid = threadIdx.x;
  double* Ajj = A + (size_t)jb * NB * (ld + 1);
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e / NB, c = e % NB;
    const double v = Lside[(size_t)jb * NB * NB + e];
    s[r][c] = v;
    x[r][c] = 0.0;
    Ajj[(size_t)r * ld + c] = v;
    if (nrows > 0) A[(size_t)((jb + 1) * NB + r) * ld + (size_t)jb * NB + c] = P0side[(size_t)jb * NB * NB + e];
  }
  __syncthreads();
  if (tid < NB) rs[tid] = 1.0 / s[tid][tid];
  __syncthreads();
  if (tid < NB) {
    const int c = tid, b0 = c & ~15;
    x[c][c] = rs[c];
    for (int r = c + 1; r < b0 + 16; ++r) {
      double acc = 0.0;
      for (int k = c; k < r; ++k) acc += s[r][k] * x[k][c];
      x[r][c] = -acc * rs[r];
    }
  }
  __syncthreads();
  for (int half = 16; half < NB; half *= 2) {
    const int nblk = NB / (2 * half);
    const int per = half * half;
    for (int e = tid; e < nblk * per; e += 256) {
      const int q = e / per, i = (e % per) / half, jj = e % half;
      const int o = q * 2 * half;
      double acc = 0.0;
      for (int k = jj; k < half; ++k) acc += s[o + half + i][o + k] * x[o + k][o + jj];
      t[q * half + i][jj] = acc;
    }
    __syncthreads();
    for (int e = tid; e < nblk * per; e += 256) {
      const int q = e / per, i = (e % per) / half, jj = e % half;
      const int o = q * 2 * half;
      double acc = 0.0;
      for (int k = 0; k <= i; ++k) acc += x[o + half + i][o + half + k] * t[q * half + k][jj];
      x[o + half + i][o + jj] = -acc;
    }
    __syncthreads();
  }
  for (int e = tid; e < NB * NB; e += 256) Dinv[(size_t)jb * NB * NB + e] = x[e / NB][e % NB];
}

// W[jb,jb] = Dinv[jb] for every diagonal block (level 0 of the recursive triangular inverse)
__global__ void scatter_dinv_kernel(const double* __restrict__ Dinv, double* __restrict__ W, int ld) {
  int jb = blockIdx.x;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    int r = e / NB, c = e % NB;
    W[(size_t)(jb * NB + r) * ld + jb * NB + c] = Dinv[(size_t)jb * NB * NB + e];
  }
}

// zero the strict upper triangle of a (ld x ld) matrix (the assembly wrote the symmetric R there)
__global__ void zero_upper_kernel(double* __restrict__ A, int ld) {
  size_t n = (size_t)ld * ld;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    size_t r = e / ld, c = e % ld;
    if (c > r) A[e] = 0.0;
  }
}

// ---------------------------------------------------------------------------------------------------
// y = W x (lower-triangular W, one warp per row, k <= row) for up to two right-hand sides at once:
// Yt = L^-1 y (gpr.py:799) and Ft = L^-1 F (gpr.py:804; constant trend: F = 1 on real rows, 0 on padding).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tri_gemv2_kernel(const double* __restrict__ W, int ld, int n,
                                                        const double* __restrict__ x0,
                                                        const double* __restrict__ x1, double* __restrict__ y0,
                                                        double* __restrict__ y1) {
  int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= n) return;
  const double* w = W + (size_t)row * ld;
  double a0 = 0.0, a1 = 0.0;
  for (int k = lane; k <= row; k += 32) {
    double wv = w[k];
    a0 += wv * x0[k];
    a1 += wv * x1[k];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
  }
  if (lane == 0) {
    y0[row] = a0;
    y1[row] = a1;
  }
}

// y = W^T x (gamma = L^-T rho, gpr.py:788): thread per column, rows j..n-1, coalesced across threads.
// Rows are split over blockIdx.y in chunks, partial sums land in part[(chunk, col)] and are summed in a
// fixed order by the caller's reduce kernel (deterministic).
__global__ void __launch_bounds__(256) tri_gemvT_partial_kernel(const double* __restrict__ W, int ld, int n,
                                                                const double* __restrict__ x,
                                                                double* __restrict__ part, int rows_per_chunk) {
  int col = blockIdx.x * blockDim.x + threadIdx.x;
  int r0 = blockIdx.y * rows_per_chunk;
  int r1 = min(n, r0 + rows_per_chunk);
  if (col >= n) return;
  double a = 0.0;
  for (int r = max(r0, col); r < r1; ++r) a += W[(size_t)r * ld + col] * x[r];
  part[(size_t)blockIdx.y * n + col] = a;
}

__global__ void colsum_partials_kernel(const double* __restrict__ part, int n, int chunks, double* __restrict__ y) {
  int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= n) return;
  double a = 0.0;
  for (int c = 0; c < chunks; ++c) a += part[(size_t)c * n + col];
  y[col] = a;
}

// Single-block deterministic reductions.  out[0] = sum Ft^2, out[1] = sum Ft*Yt, out[2] = sum log L_ii.
__device__ __forceinline__ double block_sum_1024(double v, double* sh) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double t = 0.0;
  if (w == 0) {
    t = lane < (int)(blockDim.x >> 5) ? sh[lane] : 0.0;
#pragma unroll
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;  // valid in warp 0
}

__global__ void __launch_bounds__(1024) fit_scalars_kernel(const double* __restrict__ L, int ld, int n,
                                                           const double* __restrict__ Ft,
                                                           const double* __restrict__ Yt, double* __restrict__ out) {
  __shared__ double sh[32];
  double ff = 0.0, fy = 0.0, ld_ = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double f = Ft[i];
    ff += f * f;
    fy += f * Yt[i];
    ld_ += log(L[(size_t)i * ld + i]);
  }
  double a = block_sum_1024(ff, sh);
  double b = block_sum_1024(fy, sh);
  double c = block_sum_1024(ld_, sh);
  if (threadIdx.x == 0) {
    out[0] = a;
    out[1] = b;
    out[2] = c;
  }
}

// rho = Yt - coef * Ft  with coef = (Ft.Yt)/(Ft.Ft) (ordinary kriging: Q Q^T Yt, gpr.py:805-806) or the
// fixed beta (simple kriging: L^-1 (F beta), gpr.py:808);  out[3] = rho^T rho, out[4] = coef used.
__global__ void __launch_bounds__(1024) rho_kernel(const double* __restrict__ Yt, const double* __restrict__ Ft,
                                                   int n, int estimate_trend, double beta_fixed,
                                                   double* __restrict__ rho, double* __restrict__ out) {
  __shared__ double sh[32];
  double coef = estimate_trend ? out[1] / out[0] : beta_fixed;
  double rr = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double r = Yt[i] - coef * Ft[i];
    rho[i] = r;
    rr += r * r;
  }
  double a = block_sum_1024(rr, sh);
  if (threadIdx.x == 0) {
    out[3] = a;
    out[4] = coef;
  }
}

// ---------------------------------------------------------------------------------------------------
// Likelihood-gradient traces (gpr.py:994-1038) without the (N,N,D) tensor of corr_grad_theta (gpr.py:745):
// for every pair i > j the correlation, its theta-derivative factor and the coefficient
//     c_ij = (a * gamma_i gamma_j - b * Rinv_ij) * dfactor_ij
// are formed once in registers, then  out[d] = sum_{i>j} c_ij * w_d(x_i - x_j)  for all d.
// Extra slots: [D] sum_{i>j} Rinv_ij R0_ij, [D+1] sum_{i>j} gamma_i gamma_j R0_ij, [D+2] trace(Rinv),
// [D+3] gamma^T gamma  (the sigma2 / alpha components).  One CTA per 64x64 lower tile; per-tile partials are
// summed in a fixed order by grad_reduce_kernel (deterministic).
// ---------------------------------------------------------------------------------------------------
struct GradArgs {
  const double* Xt;     // (D, ld)
  const double* theta;  // (D,)
  const double* Rinv;   // (ld, ld), lower triangle valid
  const double* gamma;  // (ld,)
  double* partial;      // (ntiles, D + 4)
  int N, D, ld, corr;
  double a, b;
};

__global__ void __launch_bounds__(256) llf_grad_traces_kernel(GradArgs p) {
  int t = blockIdx.x;
  int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  while (ti * (ti + 1) / 2 > t) --ti;
  int tj = t - ti * (ti + 1) / 2;
  const int i0 = ti * NB, j0 = tj * NB;
  extern __shared__ __align__(16) double sm[];
  double* xi = sm;                       // [D][64]
  double* xj = xi + p.D * NB;            // [D][64]
  double* th = xj + p.D * NB;            // [D]
  double* gi = th + ((p.D + 1) & ~1);    // [64]
  double* gj = gi + NB;                  // [64]
  double* wacc = gj + NB;                // [8][D+4]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = p.D + 4;
  for (int e = tid; e < p.D * NB; e += 256) {
    int d = e / NB, c = e % NB;
    xi[e] = p.Xt[(size_t)d * p.ld + i0 + c];
    xj[e] = p.Xt[(size_t)d * p.ld + j0 + c];
  }
  for (int d = tid; d < p.D; d += 256) th[d] = p.theta[d];
  if (tid < NB) gi[tid] = p.gamma[i0 + tid];
  else if (tid < 2 * NB) gj[tid - NB] = p.gamma[j0 + tid - NB];
  for (int e = tid; e < 8 * S; e += 256) wacc[e] = 0.0;
  __syncthreads();
  const int tr = (tid / 16) * 4, tc = tid % 16;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = corr_init(p.corr);
  for (int d = 0; d < p.D; ++d) {
    double xa[4], xb[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) xa[a] = xi[d * NB + tr + a];
#pragma unroll
    for (int b = 0; b < 4; ++b) xb[b] = xj[d * NB + tc + 16 * b];
    double thd = th[d];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = corr_accum(p.corr, acc[a][b], thd, xa[a] - xb[b]);
  }
  double c[4][4];
  double t1 = 0.0, t2 = 0.0, tr_ = 0.0, gg = 0.0;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int gi_ = i0 + tr + a, gj_ = j0 + tc + 16 * b;
      bool live = gi_ > gj_ && gi_ < p.N;  // strict lower triangle of the real matrix
      double rinv = p.Rinv[(size_t)gi_ * p.ld + gj_];
      if (live) {
        double r0 = corr_finish(p.corr, acc[a][b]);
        double g2 = gi[tr + a] * gj[tc + 16 * b];
        c[a][b] = (p.a * g2 - p.b * rinv) * corr_dtheta_factor(p.corr, acc[a][b]);
        t1 += rinv * r0;
        t2 += g2 * r0;
      } else {
        c[a][b] = 0.0;
        if (gi_ == gj_ && gi_ < p.N) {
          tr_ += rinv;
          gg += gi[tr + a] * gi[tr + a];
        }
      }
    }
  auto warp_add = [&](double v, int slot) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) wacc[warp * S + slot] += v;
  };
  for (int d = 0; d < p.D; ++d) {
    double xa[4], xb[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) xa[a] = xi[d * NB + tr + a];
#pragma unroll
    for (int b = 0; b < 4; ++b) xb[b] = xj[d * NB + tc + 16 * b];
    double pd = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) pd += c[a][b] * corr_dtheta_weight(p.corr, xa[a] - xb[b]);
    warp_add(pd, d);
  }
  warp_add(t1, p.D);
  warp_add(t2, p.D + 1);
  warp_add(tr_, p.D + 2);
  warp_add(gg, p.D + 3);
  __syncthreads();
  for (int sidx = tid; sidx < S; sidx += 256) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += wacc[w * S + sidx];
    p.partial[(size_t)t * S + sidx] = v;
  }
}

// out[s] = sum over tiles of partial[tile][s], fixed order; one block per slot
__global__ void __launch_bounds__(1024) grad_reduce_kernel(const double* __restrict__ partial, int ntiles, int S,
                                                           double* __restrict__ out) {
  __shared__ double sh[32];
  int s = blockIdx.x;
  double v = 0.0;
  for (int t = threadIdx.x; t < ntiles; t += blockDim.x) v += partial[(size_t)t * S + s];
  double a = block_sum_1024(v, sh);
  if (threadIdx.x == 0) out[s] = a;
}

}  // namespace b2
