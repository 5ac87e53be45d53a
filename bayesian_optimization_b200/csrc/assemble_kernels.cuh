// assemble_kernels.cuh -- kernel-matrix assembly, TMA-staged (sm_100a).
//
//   R(i, j) = corr(theta, X_i - X_j), mode scaling fused       gpr.py:772-782 (+ :935 / :952 / :966-967 per mode)
//
// Second version of kmat_assemble_kernel (fit_kernels.cuh; same arithmetic, bit-identical R).  One CTA = one 64 x 64 tile
// (ti >= tj) of the lower triangle and its mirror image:
//   * the two (D, 64) slabs of the transposed training set arrive by TMA (cp.async.bulk.tensor.2d, one mbarrier) --
//     no thread spends instructions on staging;
//   * a thread owns 4 rows x 4 CONSECUTIVE columns, so a row segment of the tile is one 256-bit global store and a
//     column segment one 256-bit store of the mirrored tile: the tile never passes through shared memory (the first
//     version staged it in a padded buffer for the transpose: 34 KB of shared memory per CTA, two more barriers, and
//     16-byte stores);
//   * rows / columns >= N (padding up to the tile multiple) are the identity.
// The pairwise-distance table of l1_cross_distances (gpr.py:48-61) is never materialised.
// The kernel stays float64-ALU bound (distance: 3 D flops, then sqrt + exp per entry), not HBM bound: see DESIGN.md 4.1.
#pragma once
#include "fast_kernels.cuh"
#include "fit_kernels.cuh"

namespace b2 {

constexpr int KA_DMAX = 64;  // features per TMA box (shared memory: 2 x D x 64 doubles)

__device__ __forceinline__ void st_global_v4f64(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// CORR: the correlation id as a compile-time constant (the switches of gp_math.h fold away: the first version kept the
// run-time id inside its unrolled loops, and its largest stall reason was "no instruction" -- instruction-cache misses on a
// body that inlined every kernel including the Bessel-function Matern), or -1 = generic (p.corr at run time).
template <int CORR>
__global__ void __launch_bounds__(256, CORR >= 0 ? 3 : 2) kmat_assemble_tma_kernel(const __grid_constant__ CUtensorMap mapX, AssembleArgs p) {
  int t = blockIdx.x;
  int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  while (ti * (ti + 1) / 2 > t) --ti;
  const int tj = t - ti * (ti + 1) / 2;
  const int i0 = ti * NB, j0 = tj * NB;
  extern __shared__ __align__(128) double sm_ka[];
  double* xi = sm_ka;                 // [D][64]
  double* xj = xi + p.D * NB;         // [D][64]
  double* th = xj + p.D * NB;         // [D + 1]
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x;
  const int corr = CORR >= 0 ? CORR : p.corr;
  const uint32_t b = fk::smem_u32(&bar);
  if (tid == 0) {
    fk::mbar_init(b, 1);
    fk::fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
    fk::mbar_arrive_expect_tx(b, (uint32_t)(2 * p.D * NB * sizeof(double)));
    fk::tma_load_2d(fk::smem_u32(xi), &mapX, i0, 0, b);
    fk::tma_load_2d(fk::smem_u32(xj), &mapX, j0, 0, b);
  }
  for (int d = tid; d < p.D + (corr_has_extra_param(corr) ? 1 : 0); d += 256) th[d] = p.theta[d];
  __syncthreads();
  const double pw = corr_has_extra_param(corr) ? th[p.D] : 0.0;
  while (!fk::mbar_try_wait(b, 0)) {
  }
  // thread -> rows tr .. tr + 3, columns tc .. tc + 3
  const int tr = (tid >> 4) * 4, tc = (tid & 15) * 4;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = corr_init(corr);
  for (int d = 0; d < p.D; ++d) {
    const double2 a01 = *reinterpret_cast<const double2*>(xi + d * NB + tr), a23 = *reinterpret_cast<const double2*>(xi + d * NB + tr + 2);
    const double2 b01 = *reinterpret_cast<const double2*>(xj + d * NB + tc), b23 = *reinterpret_cast<const double2*>(xj + d * NB + tc + 2);
    const double xa[4] = {a01.x, a01.y, a23.x, a23.y}, xb[4] = {b01.x, b01.y, b23.x, b23.y};
    const double thd = th[d];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][c] = corr_accum_p(corr, acc[a][c], thd, xa[a] - xb[c], pw);
  }
  const double s2t = p.sigma2 + p.noise_var;
  double (&v)[4][4] = acc;   // finished in place
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int gi = i0 + tr + a, gj = j0 + tc + c;
      const double r = corr_finish_p(corr, acc[a][c], pw);
      double x = p.mode == 1 ? (p.sigma2 * r) / s2t : (p.mode == 2 ? p.alpha * r : r);
      if (gi == gj) x = p.mode == 1 ? (p.sigma2 * 1.0 + p.noise_var) / s2t : (p.mode == 2 ? p.alpha * 1.0 + (1.0 - p.alpha) : 1.0);
      if (gi >= p.N || gj >= p.N) x = gi == gj ? 1.0 : 0.0;
      v[a][c] = x;
    }
  // lower tile: a warp stores two 512-byte rows per instruction
#pragma unroll
  for (int a = 0; a < 4; ++a) st_global_v4f64(p.R + (size_t)(i0 + tr + a) * p.ld + j0 + tc, v[a][0], v[a][1], v[a][2], v[a][3]);
  if (ti != tj) {  // mirrored tile: row j0 + tc + c, columns i0 + tr .. tr + 3
#pragma unroll
    for (int c = 0; c < 4; ++c) st_global_v4f64(p.R + (size_t)(j0 + tc + c) * p.ld + i0 + tr, v[0][c], v[1][c], v[2][c], v[3][c]);
  }
}

}  // namespace b2
