// predict_kernels.cuh -- the M-candidate path, float64 parity flavour.
//   kstar_kernel          r = corr(theta, |Xc - X|) (M x N) and yhat = F(Xc) beta + r gamma   gpr.py:486-490
//   contract_fp64_kernel  rt = L^-1 r^T as a DMMA GEMM against L^-1, with the row reductions
//                         sum rt^2 and Ft^T rt fused into the epilogue (rt never leaves registers)
//                                                                                     gpr.py:494-498, :502
//   acq_kernel            MSE = sigma2 (1 - sum rt^2 + u^2) clipped at 0 (gpr.py:502-510), then q acquisition
//                         values per candidate (acquisition/acquisition_fun.py) + block arg-max
//   argmax_merge_kernel   fixed-order merge of block partials into the running per-criterion best
// Paths are relative to /root/reference/bayes_optim/ (gpr.py = surrogate/gaussian_process/gpr.py).
#pragma once
#include <cuda_runtime.h>

#include "dgemm.cuh"
#include "gp_math.h"

namespace b2 {

constexpr int KS_ROWS = 8;  // candidates per CTA in kstar_kernel

struct KstarArgs {
  const double* Xc;     // (M, D) row-major candidates of this chunk
  const double* Xt;     // (D, ld) transposed training set
  const double* theta;  // (D,)
  const double* gamma;  // (ld,), zero on padding
  double* Kst;          // (Mpad, ld) or NULL (eval_MSE = False: mean only)
  double* yhat;         // (M,)
  int M, N, D, ld, corr;
  double beta;  // constant trend: F(Xc) beta = beta
};

__global__ void __launch_bounds__(256) kstar_kernel(KstarArgs p) {
  extern __shared__ __align__(16) double sm[];
  double* xc = sm;                    // [KS_ROWS][D]
  double* th = xc + KS_ROWS * p.D;    // [D]
  __shared__ double red[KS_ROWS][8];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * KS_ROWS;
  for (int e = tid; e < KS_ROWS * p.D; e += 256) {
    int r = e / p.D, d = e % p.D;
    xc[e] = (m0 + r < p.M) ? p.Xc[(size_t)(m0 + r) * p.D + d] : 0.0;
  }
  for (int d = tid; d < p.D; d += 256) th[d] = p.theta[d];
  const double pw = corr_has_extra_param(p.corr) ? p.theta[p.D] : 0.0;  // exponent (generalized_exponential) / nu (general Matern)
  __syncthreads();
  double ysum[KS_ROWS];
#pragma unroll
  for (int r = 0; r < KS_ROWS; ++r) ysum[r] = 0.0;
  for (int n = tid; n < p.ld; n += 256) {
    double acc[KS_ROWS];
#pragma unroll
    for (int r = 0; r < KS_ROWS; ++r) acc[r] = corr_init(p.corr);
    if (n < p.N) {
      for (int d = 0; d < p.D; ++d) {
        double xd = p.Xt[(size_t)d * p.ld + n];
        double thd = th[d];
#pragma unroll
        for (int r = 0; r < KS_ROWS; ++r) acc[r] = corr_accum_p(p.corr, acc[r], thd, xc[r * p.D + d] - xd, pw);
      }
    }
    double g = p.gamma[n];
#pragma unroll
    for (int r = 0; r < KS_ROWS; ++r) {
      double k = n < p.N ? corr_finish_p(p.corr, acc[r], pw) : 0.0;  // padding columns contribute nothing
      if (p.Kst) p.Kst[(size_t)(m0 + r) * p.ld + n] = k;
      ysum[r] += k * g;
    }
  }
  // fixed-order block reduction of the 8 row sums
  const int lane = tid & 31, w = tid >> 5;
#pragma unroll
  for (int r = 0; r < KS_ROWS; ++r) {
    double v = ysum[r];
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[r][w] = v;
  }
  __syncthreads();
  if (tid < KS_ROWS && m0 + tid < p.M) {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += red[tid][k];
    p.yhat[m0 + tid] = p.beta + v;
  }
}

// ---------------------------------------------------------------------------------------------------
// rt = L^-1 r^T for a tile of 128 candidates, all N outputs, on the FP64 tensor pipe.
//   D(m, n) = sum_{k <= n} Kst(m, k) * Linv(n, k)      (Linv lower triangular: k-range stops at the tile end)
// The CTA walks the n-tiles itself, so the per-candidate reductions stay in registers:
//   sumsq(m) = sum_n D(m,n)^2          (rt**2).sum(axis=0)          gpr.py:502
//   dotf(m)  = sum_n Ft(n) D(m,n)      np.dot(Ft.T, rt)             gpr.py:498
// and are written once, in a fixed order (deterministic, no atomics).
// ---------------------------------------------------------------------------------------------------
constexpr int PC_BM = 128, PC_BN = 128, PC_WM = 32, PC_WN = 64, PC_STAGES = 3;
using PredCore = GemmCore<PC_BM, PC_BN, PC_WM, PC_WN, false, false, PC_STAGES>;

struct ContractArgs {
  const double* Kst;   // (Mpad, ld)
  const double* Linv;  // (ld, ld) lower
  const double* Ft;    // (ld,) zero on padding
  double* sumsq;       // (Mpad,)
  double* dotf;        // (Mpad,)
  int ld;
};

__global__ void __launch_bounds__(PredCore::NT, 1) contract_fp64_kernel(ContractArgs p) {
  extern __shared__ __align__(16) double smem_d[];
  __shared__ double red[2][PC_BM][PredCore::WARPS_N];
  const int m0 = blockIdx.x * PC_BM;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double rs[PredCore::TM], rf[PredCore::TM];
#pragma unroll
  for (int i = 0; i < PredCore::TM; ++i) rs[i] = rf[i] = 0.0;
  const int ntiles = p.ld / PC_BN;
  for (int nt = 0; nt < ntiles; ++nt) {
    const int n0 = nt * PC_BN;
    double acc[PredCore::TM][PredCore::TN][2];
#pragma unroll
    for (int i = 0; i < PredCore::TM; ++i)
#pragma unroll
      for (int j = 0; j < PredCore::TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    PredCore::run(acc, p.Kst, p.ld, m0, p.Linv, p.ld, n0, 0, n0 + PC_BN, smem_d);
#pragma unroll
    for (int j = 0; j < PredCore::TN; ++j) {
      int c = n0 + PredCore::col_of(j);
      double f0 = p.Ft[c], f1 = p.Ft[c + 1];
#pragma unroll
      for (int i = 0; i < PredCore::TM; ++i) {
        double v0 = acc[i][j][0], v1 = acc[i][j][1];
        rs[i] += v0 * v0;
        rs[i] += v1 * v1;
        rf[i] += f0 * v0;
        rf[i] += f1 * v1;
      }
    }
  }
  // rows are shared by the 4 lanes of a quad and by the WARPS_N warps along n
#pragma unroll
  for (int i = 0; i < PredCore::TM; ++i) {
    rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 1);
    rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 2);
    rf[i] += __shfl_xor_sync(0xffffffffu, rf[i], 1);
    rf[i] += __shfl_xor_sync(0xffffffffu, rf[i], 2);
    if ((lane & 3) == 0) {
      int r = PredCore::row_of(i);
      red[0][r][warp % PredCore::WARPS_N] = rs[i];
      red[1][r][warp % PredCore::WARPS_N] = rf[i];
    }
  }
  __syncthreads();
  for (int r = threadIdx.x; r < PC_BM; r += PredCore::NT) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int w = 0; w < PredCore::WARPS_N; ++w) {
      a += red[0][r][w];
      b += red[1][r][w];
    }
    p.sumsq[m0 + r] = a;
    p.dotf[m0 + r] = b;
  }
}

// ---------------------------------------------------------------------------------------------------
// Small-M flavour of the same contraction (M <= a few thousand: the reference's own calling pattern is M = 1,
// acquisition_fun.py:52-64, and the fast path re-scores a handful of band candidates).  With few candidates
// the work is one sweep over the lower triangle of L^-1, so it is parallelised over ROW blocks of L^-1
// instead of candidate tiles: CTA c owns the 32-row blocks c and nblk-1-c (balanced triangle), 32 candidates
// at a time, and writes per-CTA partial sums that rs_reduce_kernel adds in a fixed order (deterministic).
// ---------------------------------------------------------------------------------------------------
constexpr int RS_ROWS = 32, RS_KT = 64, RS_BC = 32, RS_LD = RS_KT + 1;

struct RescoreArgs {
  const double* Kst;   // (Mpad, ld) rows of r
  const double* Linv;  // (ld, ld) lower, zeros above the diagonal
  const double* Ft;    // (ld,)
  double* part;        // (chunks, gridDim.x, 2, RS_BC) partial sums: [0] sum rt^2, [1] Ft . rt
  int ld, M;
};

__global__ void __launch_bounds__(256) rs_contract_kernel(RescoreArgs p) {
  __shared__ double buf[2][RS_ROWS][RS_LD];
  double (*Ls)[RS_LD] = buf[0];
  double (*Ks)[RS_LD] = buf[1];
  double (*red)[RS_ROWS][RS_BC + 1] = (double (*)[RS_ROWS][RS_BC + 1]) & buf[0][0][0];  // reused after the k sweep
  const int tid = threadIdx.x;
  const int nl = tid >> 3, bg = tid & 7;  // row of the block, group of 4 candidates
  const int nblk = p.ld / RS_ROWS;
  const int cand0 = blockIdx.y * RS_BC;
  double tot_ss = 0.0, tot_df = 0.0;  // threads 0..31: totals of candidate tid
  for (int half = 0; half < 2; ++half) {
    const int rb = half == 0 ? (int)blockIdx.x : nblk - 1 - (int)blockIdx.x;
    if (rb >= nblk || rb < 0 || (half == 1 && rb <= (int)blockIdx.x)) continue;
    const int n0 = rb * RS_ROWS;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k0 = 0; k0 < n0 + RS_ROWS; k0 += RS_KT) {
      for (int e = tid; e < RS_ROWS * RS_KT; e += 256) {
        const int r = e / RS_KT, c = e % RS_KT;
        Ls[r][c] = p.Linv[(size_t)(n0 + r) * p.ld + k0 + c];
        Ks[r][c] = (cand0 + r < p.M) ? p.Kst[(size_t)(cand0 + r) * p.ld + k0 + c] : 0.0;
      }
      __syncthreads();
#pragma unroll 8
      for (int kk = 0; kk < RS_KT; ++kk) {
        const double l = Ls[nl][kk];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] += l * Ks[bg * 4 + j][kk];
      }
      __syncthreads();
    }
    const double f = p.Ft[n0 + nl];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red[0][nl][bg * 4 + j] = acc[j] * acc[j];
      red[1][nl][bg * 4 + j] = f * acc[j];
    }
    __syncthreads();
    if (tid < RS_BC) {
#pragma unroll 4
      for (int r = 0; r < RS_ROWS; ++r) {
        tot_ss += red[0][r][tid];
        tot_df += red[1][r][tid];
      }
    }
    __syncthreads();
  }
  if (tid < RS_BC) {
    double* o = p.part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2 * RS_BC;
    o[tid] = tot_ss;
    o[RS_BC + tid] = tot_df;
  }
}

__global__ void rs_reduce_kernel(const double* __restrict__ part, int nctas, int M, double* __restrict__ sumsq,
                                 double* __restrict__ dotf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const int chunk = i / RS_BC, b = i % RS_BC;
  double a = 0.0, d = 0.0;
  for (int c = 0; c < nctas; ++c) {
    const double* o = part + ((size_t)chunk * nctas + c) * 2 * RS_BC;
    a += o[b];
    d += o[RS_BC + b];
  }
  sumsq[i] = a;
  dotf[i] = d;
}

// ---------------------------------------------------------------------------------------------------
// MSE + acquisition + block arg-max.  grid = (candidate blocks, q criteria).
// ---------------------------------------------------------------------------------------------------
struct AcqArgs {
  const double* yhat;   // (M,)
  const double* sumsq;  // (M,) or NULL when mse_in is given
  const double* dotf;   // (M,)
  const double* mse_in; // (M,) precomputed MSE (acq_from_moments) or NULL
  double* mse_out;      // (M,) or NULL
  double* vals;         // (q, vals_ld) or NULL; this chunk starts at column vals_off
  long long vals_ld, vals_off;
  const double* params;  // (q,)
  double* part_val;      // (q, gridDim.x)
  long long* part_idx;   // (q, gridDim.x)
  long long idx_base;    // global index of candidate 0 of this chunk
  const long long* idx_map;  // NULL, or global index of candidate i (re-scored band of the fast path)
  const int* M_dev = nullptr;  // NULL, or the candidate count in device memory (then M is its cap)
  int M, acq, minimize, estimate_trend, q;
  double sigma2, plugin, G;  // G: the 1x1 triangular factor of the thin QR of Ft (|G| = ||Ft||)
};

__device__ __forceinline__ double mse_from_sums(const AcqArgs& p, int i) {
  if (p.mse_in) return p.mse_in[i];
  double u2 = 0.0;
  if (p.estimate_trend) {  // u = G^-T (Ft^T rt - f(x)), f = 1 for the constant trend  (gpr.py:496-498)
    double u = (p.dotf[i] - 1.0) / p.G;
    u2 = u * u;
  }
  double m = (1.0 - p.sumsq[i] + u2) * p.sigma2;  // gpr.py:502-505
  return m < 0.0 ? 0.0 : m;                       // gpr.py:510
}

__global__ void __launch_bounds__(256) acq_kernel(AcqArgs p) {
  __shared__ double sv[8];
  __shared__ long long si[8];
  const int c = blockIdx.y;
  const double par = p.acq == ACQ_MGFI ? fmin(p.params[c], 22.36) : p.params[c];  // acquisition_fun.py:262
  double bv = 0.0;
  long long bi = -1;
  const int M = p.M_dev ? min(*p.M_dev, p.M) : p.M;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
    double mse = mse_from_sums(p, i);
    if (c == 0 && p.mse_out) p.mse_out[i] = mse;
    double v = acq_value(p.acq, p.yhat[i], mse, p.sigma2, p.plugin, par, p.minimize);
    if (p.vals) p.vals[(size_t)c * p.vals_ld + p.vals_off + i] = v;
    long long gi = p.idx_map ? p.idx_map[i] : p.idx_base + i;
    if (bi < 0 || arg_better(v, gi, bv, bi)) {
      bv = v;
      bi = gi;
    }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi >= 0 && (bi < 0 || arg_better(ov, oi, bv, bi))) {
      bv = ov;
      bi = oi;
    }
  }
  if (lane == 0) {
    sv[w] = bv;
    si[w] = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k)
      if (si[k] >= 0 && (bi < 0 || arg_better(sv[k], si[k], bv, bi))) {
        bv = sv[k];
        bi = si[k];
      }
    p.part_val[(size_t)c * gridDim.x + blockIdx.x] = bv;
    p.part_idx[(size_t)c * gridDim.x + blockIdx.x] = bi;
  }
}

// MSE only (predict without acquisition)
__global__ void mse_kernel(AcqArgs p) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.M; i += gridDim.x * blockDim.x)
    p.mse_out[i] = mse_from_sums(p, i);
}

// best[c] <- merge(best[c], partials of this chunk), one thread per criterion, fixed order
__global__ void argmax_merge_kernel(const double* __restrict__ part_val, const long long* __restrict__ part_idx,
                                    int nblocks, int q, double* __restrict__ best_val,
                                    long long* __restrict__ best_idx) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= q) return;
  double bv = best_val[c];
  long long bi = best_idx[c];
  for (int b = 0; b < nblocks; ++b) {
    double v = part_val[(size_t)c * nblocks + b];
    long long i = part_idx[(size_t)c * nblocks + b];
    if (i >= 0 && (bi < 0 || arg_better(v, i, bv, bi))) {
      bv = v;
      bi = i;
    }
  }
  best_val[c] = bv;
  best_idx[c] = bi;
}

}  // namespace b2
