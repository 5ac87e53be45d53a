// band_kernels.cuh -- device-side pipeline of the arg-max band of the tensor-core path (sm_100a).
//
// After the fused tcgen05 pass every candidate carries approximate moments.  The band stage keeps the arg-max exact
// (fk::band_* in fast_kernels.cuh select the candidates that could still win once the pass's error is allowed for) and
// re-scores them in float64.  Round 1 drove that from the host: seven stream synchronisations per pass and a
// small-M contraction that streamed L^-1 through 64 CTAs (1.4 ms per step at C3 to re-score two candidates).  Here the
// whole stage is ONE stream-ordered sequence with device-side counts and a single device->host copy at its end:
//
//   band_reset -> thr0 -> scan -> refine -> filter -> gather -> kstar_small -> rowdot -> band_moments
//              -> acq (+ arg-max merge, global indices) -> band_check -> [one D2H of the control block]
//
// The kernels that work on the band read its size from device memory and are launched for the largest band the
// on-device route takes (BAND_DEV_MAX); the host looks at the control block afterwards and only then decides whether
// the (rare) wide-band / widen-and-repeat / escalation routes are needed.
//
// What the kernels restate: r and yhat of gpr.py:486-490 (kstar_small), rt = L^-1 r^T, sum rt^2, Ft^T rt of
// gpr.py:494-502 (rowdot), all float64 with fixed summation orders (bit-reproducible run to run).
#pragma once
#include <cuda_runtime.h>

#include "gp_math.h"

namespace b2 {
namespace bd {

constexpr int BAND_DEV_MAX = 1024;  // capacity of the on-device re-score buffers (the host picks the cut-over per N)
constexpr int KSS_COLS = 256;      // training columns per CTA of kstar_small_kernel
constexpr int KSS_ROWS = 8;        // candidates per CTA of kstar_small_kernel
constexpr int RD_WARPS = 16;       // warps per CTA of rowdot_kernel
constexpr int RD_CB = 8;           // candidates per register block of rowdot_kernel
constexpr int RD_UNROLL = 4;       // 32-wide k slices in flight per lane

// control block of one band pass: written on the device, copied to the host once per pass
struct BandCtl {
  int count0;      // scan survivors (may exceed the list capacity: then the pass is void)
  int count;       // band size
  int fused_err;   // error flag of the fused kernel (mbarrier timeout codes)
  int reserved;
  double err_y;    // max |yhat_fast - yhat_exact| inside the band
  double err_s;    // max |mse_fast - mse_exact| inside the band (unclipped)
  double ratio;    // max over the band of (observed error) / (half-width that was allowed for it)
  double pad;
};

// thr keys <- -inf, counts <- 0, best <- (0, -1), control block <- 0 (one launch instead of four copies + a sync)
__global__ void band_reset_kernel(long long* __restrict__ thr_key, int nkeys, long long ninf_key, int* __restrict__ counts,
                                  double* __restrict__ best_val, long long* __restrict__ best_idx, int q,
                                  BandCtl* __restrict__ ctl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nkeys) thr_key[i] = ninf_key;
  if (i < q) {
    best_val[i] = 0.0;
    best_idx[i] = -1;
  }
  if (i < 2) counts[i] = 0;
  if (i == 0) {
    ctl->count0 = ctl->count = ctl->fused_err = ctl->reserved = 0;
    ctl->err_y = ctl->err_s = ctl->ratio = ctl->pad = 0.0;
  }
}

// Xb[b, :] = Xc[list[b], :] for b < min(*count, cap)
__global__ void band_gather_dev_kernel(const double* __restrict__ Xc, const long long* __restrict__ list,
                                       const int* __restrict__ count, int cap, int D, double* __restrict__ Xb) {
  const int nb = min(*count, cap);
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nb * D) return;
  const int b = e / D, d = e % D;
  Xb[e] = Xc[(size_t)list[b] * D + d];
}

// ---------------------------------------------------------------------------------------------------------------
// r = corr(theta, |Xb - X|) for the band rows, parallel over training columns as well as rows (the tile-per-8-rows
// kstar_kernel of the float64 path leaves 147 SMs idle for a two-candidate band).  grid = (ld / KSS_COLS, cap / 8).
// Partial yhat sums per column slice go to ypart[(slice, row)]; band_moments_kernel adds them in a fixed order.
// ---------------------------------------------------------------------------------------------------------------
struct KstarSmallArgs {
  const double* Xb;     // (cap, D) band rows
  const double* Xt;     // (D, ld)
  const double* theta;  // (D [+1])
  const double* gamma;  // (ld,)
  double* Kst;          // (cap, ld)
  double* ypart;        // (ld / KSS_COLS, cap)
  const int* count;
  int cap, N, D, ld, corr;
};

__global__ void __launch_bounds__(KSS_COLS) kstar_small_kernel(KstarSmallArgs p) {
  extern __shared__ __align__(16) double kss_sm[];
  double* xc = kss_sm;               // [KSS_ROWS][D]
  double* th = xc + KSS_ROWS * p.D;  // [D]
  __shared__ double red[KSS_ROWS][KSS_COLS / 32];
  const int m = min(*p.count, p.cap);
  const int m0 = blockIdx.y * KSS_ROWS;
  if (m0 >= m) return;
  const int tid = threadIdx.x;
  for (int e = tid; e < KSS_ROWS * p.D; e += KSS_COLS) {
    const int r = e / p.D, d = e % p.D;
    xc[e] = (m0 + r < m) ? p.Xb[(size_t)(m0 + r) * p.D + d] : 0.0;
  }
  for (int d = tid; d < p.D; d += KSS_COLS) th[d] = p.theta[d];
  const double pw = corr_has_extra_param(p.corr) ? p.theta[p.D] : 0.0;
  __syncthreads();
  const int n = blockIdx.x * KSS_COLS + tid;
  double acc[KSS_ROWS];
#pragma unroll
  for (int r = 0; r < KSS_ROWS; ++r) acc[r] = corr_init(p.corr);
  if (n < p.N) {
    for (int d = 0; d < p.D; ++d) {
      const double xd = p.Xt[(size_t)d * p.ld + n];
      const double thd = th[d];
#pragma unroll
      for (int r = 0; r < KSS_ROWS; ++r) acc[r] = corr_accum_p(p.corr, acc[r], thd, xc[r * p.D + d] - xd, pw);
    }
  }
  // the last column slice may reach past the pitch (ld is a multiple of 128, the slice is 256 wide): those threads
  // only take part in the reductions
  const double g = n < p.ld ? p.gamma[n] : 0.0;
  const int lane = tid & 31, w = tid >> 5;
#pragma unroll
  for (int r = 0; r < KSS_ROWS; ++r) {
    const double k = n < p.N ? corr_finish_p(p.corr, acc[r], pw) : 0.0;
    if (m0 + r < m && n < p.ld) p.Kst[(size_t)(m0 + r) * p.ld + n] = k;
    double v = k * g;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[r][w] = v;
  }
  __syncthreads();
  if (tid < KSS_ROWS && m0 + tid < m) {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < KSS_COLS / 32; ++k) v += red[tid][k];
    p.ypart[(size_t)blockIdx.x * p.cap + m0 + tid] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// rt = L^-1 r^T for a handful of candidates: sweeps over the lower triangle of L^-1 at memory speed.
// A CTA owns 16 consecutive rows of L^-1 (one per warp: equal lengths inside the CTA; row blocks are dealt to the CTAs
// round-robin); its lanes stride over k, so a row is read in 256-byte coalesced pieces, RD_UNROLL of them in flight
// per lane and the next k-step's pieces requested before the current one is used.  The candidates' r rows are staged
// RD_CB candidates x RD_KT columns at a time in shared memory and shared by the 16 warps (per-warp reads of r from
// L2 made the first version of this kernel L2-bound: 0.4 ms for a 22-candidate band at N = 4096).  A wider band takes
// further passes over the row block, which by then sits in L2.  Summation orders are fixed: lanes -> shuffle tree,
// warps -> serial over the 16 rows, row blocks -> band_moments_kernel (bit-reproducible).
// ---------------------------------------------------------------------------------------------------------------
constexpr int RD_KT = 32 * RD_UNROLL;  // k columns per step

struct RowdotArgs {
  const double* Kst;   // (cap, ld)
  const double* Linv;  // (ld, ld) lower
  const double* Ft;    // (ld,)
  double* part;        // (ld / RD_WARPS, cap, 2): per row block [0] sum rt^2, [1] Ft . rt
  const int* count;
  int cap, ld;
};

__global__ void __launch_bounds__(32 * RD_WARPS) rowdot_kernel(RowdotArgs p) {
  __shared__ double Ks[RD_CB][RD_KT];
  __shared__ double red[2][RD_WARPS][RD_CB];
  const int m = min(*p.count, p.cap);
  if (m <= 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nblk = p.ld / RD_WARPS;
  for (int rb = blockIdx.x; rb < nblk; rb += gridDim.x) {
    const int n = rb * RD_WARPS + warp;          // this warp's row
    const int nmax = rb * RD_WARPS + RD_WARPS - 1;
    const double* wrow = p.Linv + (size_t)n * p.ld;
    const double f = p.Ft[n];
    for (int c0 = 0; c0 < m; c0 += RD_CB) {
      double acc[RD_CB];
#pragma unroll
      for (int j = 0; j < RD_CB; ++j) acc[j] = 0.0;
      double wv[RD_UNROLL], wn[RD_UNROLL];
#pragma unroll
      for (int u = 0; u < RD_UNROLL; ++u) {
        const int k = 32 * u + lane;
        wn[u] = k <= n ? wrow[k] : 0.0;
      }
      for (int k0 = 0; k0 <= nmax; k0 += RD_KT) {
        __syncthreads();  // the previous tile has been consumed
        for (int e = threadIdx.x; e < RD_CB * RD_KT; e += 32 * RD_WARPS) {
          const int c = e / RD_KT, kk = e % RD_KT;
          Ks[c][kk] = (c0 + c < m) ? p.Kst[(size_t)(c0 + c) * p.ld + k0 + kk] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < RD_UNROLL; ++u) wv[u] = wn[u];
        if (k0 + RD_KT <= nmax) {
#pragma unroll
          for (int u = 0; u < RD_UNROLL; ++u) {
            const int k = k0 + RD_KT + 32 * u + lane;
            wn[u] = k <= n ? wrow[k] : 0.0;
          }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < RD_CB; ++j)
#pragma unroll
          for (int u = 0; u < RD_UNROLL; ++u) acc[j] = fma(wv[u], Ks[j][32 * u + lane], acc[j]);
      }
#pragma unroll
      for (int j = 0; j < RD_CB; ++j) {
        double v = acc[j];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) {
          red[0][warp][j] = v * v;
          red[1][warp][j] = f * v;
        }
      }
      __syncthreads();
      if (threadIdx.x < RD_CB && c0 + threadIdx.x < m) {
        double a = 0.0, d = 0.0;
#pragma unroll
        for (int w = 0; w < RD_WARPS; ++w) {
          a += red[0][w][threadIdx.x];
          d += red[1][w][threadIdx.x];
        }
        double* o = p.part + ((size_t)rb * p.cap + c0 + threadIdx.x) * 2;
        o[0] = a;
        o[1] = d;
      }
    }
  }
}

// yhat = beta + sum of the column-slice partials; sum rt^2, Ft^T rt = sum of the row blocks' partials (fixed orders)
__global__ void band_moments_kernel(const double* __restrict__ ypart, int nslices, const double* __restrict__ part, int nwarps,
                                    const int* __restrict__ count, int cap, double beta, double* __restrict__ yhat,
                                    double* __restrict__ sumsq, double* __restrict__ dotf) {
  const int m = min(*count, cap);
  const int i = blockIdx.x;  // one candidate per block, 128 threads
  if (i >= m) return;
  __shared__ double s0[128], s1[128];
  double a = 0.0, d = 0.0;
  for (int w = threadIdx.x; w < nwarps; w += 128) {
    const double* o = part + ((size_t)w * cap + i) * 2;
    a += o[0];
    d += o[1];
  }
  s0[threadIdx.x] = a;
  s1[threadIdx.x] = d;
  __syncthreads();
  for (int o = 64; o; o >>= 1) {
    if (threadIdx.x < o) {
      s0[threadIdx.x] += s0[threadIdx.x + o];
      s1[threadIdx.x] += s1[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double y = 0.0;
    for (int s = 0; s < nslices; ++s) y += ypart[(size_t)s * cap + i];
    yhat[i] = beta + y;
    sumsq[i] = s0[0];
    dotf[i] = s1[0];
  }
}

// Observed errors of the fast pass on the band (exact moments of band entry b vs the fast moments of candidate
// list[b]) and their ratio to the half-widths the band was built with; also the bookkeeping the host reads once.
struct BandCheckArgs {
  const double *y_fast, *ss_fast, *df_fast;  // (M,)
  const double *y_ex, *ss_ex, *df_ex;        // (cap,)
  const long long* list;
  const int* counts;   // [0] scan survivors, [1] band size
  const int* fused_err;
  int cap, estimate_trend;
  double G, sigma2;
  double dy, ds, ds_abs, ds_rel, du;  // the half-widths of the pass (see fk::band_box)
  BandCtl* ctl;
};

__global__ void __launch_bounds__(256) band_check_kernel(BandCheckArgs p) {
  __shared__ double s0[8], s1[8], s2[8];
  const int n = min(p.counts[1], p.cap);
  double ey = 0.0, es = 0.0, ra = 0.0;
  for (int b = threadIdx.x; b < n; b += blockDim.x) {
    const long long i = p.list[b];
    const double dyv = fabs(p.y_fast[i] - p.y_ex[b]);
    double uf = 0.0, ue = 0.0;
    if (p.estimate_trend) {
      uf = (p.df_fast[i] - 1.0) / p.G;
      ue = (p.df_ex[b] - 1.0) / p.G;
    }
    const double dsv = fabs((uf * uf - p.ss_fast[i]) - (ue * ue - p.ss_ex[b])) * p.sigma2;
    const double ssf = fmax(p.ss_fast[i], 0.0);
    const double allowed = fmax(p.ds, p.ds_abs + p.ds_rel * sqrt(ssf + 1e-3)) + (2.0 * fabs(uf) * p.du + p.du * p.du) * p.sigma2;
    ey = fmax(ey, dyv);
    es = fmax(es, dsv);
    ra = fmax(ra, fmax(dyv / p.dy, dsv / allowed));
  }
  for (int o = 16; o; o >>= 1) {
    ey = fmax(ey, __shfl_xor_sync(0xffffffffu, ey, o));
    es = fmax(es, __shfl_xor_sync(0xffffffffu, es, o));
    ra = fmax(ra, __shfl_xor_sync(0xffffffffu, ra, o));
  }
  if ((threadIdx.x & 31) == 0) {
    s0[threadIdx.x >> 5] = ey;
    s1[threadIdx.x >> 5] = es;
    s2[threadIdx.x >> 5] = ra;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
      ey = fmax(ey, s0[k]);
      es = fmax(es, s1[k]);
      ra = fmax(ra, s2[k]);
    }
    p.ctl->count0 = p.counts[0];
    p.ctl->count = p.counts[1];
    p.ctl->fused_err = *p.fused_err;
    p.ctl->err_y = ey;
    p.ctl->err_s = es;
    p.ctl->ratio = ra;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// statistics of L^-1 the a-priori half-widths are built from (once per factorisation, with the fp16 split):
// per row n: sum_k W_nk^2 and sum_k |W_nk| (one warp per row, fixed order)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) linv_rowstats_kernel(const double* __restrict__ W, int ld, int n, double* __restrict__ rowsq,
                                                            double* __restrict__ rowl1) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const double* w = W + (size_t)row * ld;
  double a = 0.0, b = 0.0;
  for (int k = lane; k <= row; k += 32) {
    const double v = w[k];
    a = fma(v, v, a);
    b += fabs(v);
  }
  for (int o = 16; o; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane == 0) {
    rowsq[row] = a;
    rowl1[row] = b;
  }
}

// deterministic +-1 start vector of the power iteration (zero on padding rows)
__global__ void pm_init_kernel(double* __restrict__ v, int n, int ld) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ld) return;
  unsigned x = (unsigned)i * 2654435761u + 0x9E3779B9u;
  x ^= x >> 16;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  v[i] = i < n ? ((x & 1u) ? 1.0 : -1.0) : 0.0;
}

// out[slot] = sum v^2 (single block, fixed order)
__global__ void __launch_bounds__(1024) sumsq_kernel(const double* __restrict__ v, int n, double* __restrict__ out, int slot) {
  __shared__ double sh[32];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) a = fma(v[i], v[i], a);
  for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x < 32) {
    a = sh[threadIdx.x];
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (threadIdx.x == 0) out[slot] = a;
  }
}

}  // namespace bd
}  // namespace b2
