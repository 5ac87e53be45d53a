// dgemm.cuh -- float64 tile GEMM on the FP64 tensor pipe (mma.sync.m8n8k4.f64, "DMMA").
//
// tcgen05.mma has no f64 kind, so every float64 contraction of the parity path (Cholesky trailing
// updates, triangular inverse, and the M-candidate contraction  w = L^-1 k*  that replaces the reference's
// solve_triangular(C, r.T), surrogate/gaussian_process/gpr.py:494) runs through this mainloop.
//
//   C(m,n) = sum_{k in [k_begin,k_end)} A(m,k) * B(n,k)
//
// Operand layouts (all row-major in global memory, 16-byte aligned, dimensions padded to tile multiples):
//   A_KM == false : A stored (M x K), k contiguous      A_KM == true : A stored (K x M), m contiguous
//   B_KN == false : B stored (N x K), k contiguous      B_KN == true : B stored (K x N), n contiguous
// Shared-memory tiles are filled with 16-byte cp.async (LDGSTS) in a multi-stage ring; the +4-double row
// padding makes every 8-byte fragment load of a half-warp hit 16 distinct bank pairs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2 {

constexpr int GEMM_BK = 16;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// One operand tile in shared memory: R rows of the m- (or n-) extent by GEMM_BK of k.
template <int R, bool RC>  // RC: global source is (K x R), r contiguous
struct OpTile {
  static constexpr int LD = RC ? (R + 4) : (GEMM_BK + 4);
  static constexpr int ELEMS = RC ? GEMM_BK * LD : R * LD;

  template <int NT>
  __device__ static __forceinline__ void load_async(double* s, const double* __restrict__ g, int ld, int r0,
                                                    int k0, int tid) {
    if (!RC) {
      constexpr int CH = R * (GEMM_BK / 2);
#pragma unroll
      for (int c = tid; c < CH; c += NT) {
        int r = c / (GEMM_BK / 2), kc = (c % (GEMM_BK / 2)) * 2;
        cp_async16(s + r * LD + kc, g + (size_t)(r0 + r) * ld + k0 + kc);
      }
    } else {
      constexpr int CPR = R / 2;
      constexpr int CH = GEMM_BK * CPR;
#pragma unroll
      for (int c = tid; c < CH; c += NT) {
        int k = c / CPR, rc = (c % CPR) * 2;
        cp_async16(s + k * LD + rc, g + (size_t)(k0 + k) * ld + r0 + rc);
      }
    }
  }
  // fragment element for DMMA: row (lane>>2) of an 8-row group starting at r, k = kk*4 + (lane&3)
  __device__ static __forceinline__ double frag(const double* s, int r, int kk, int lane) {
    if (!RC) return s[(r + (lane >> 2)) * LD + kk * 4 + (lane & 3)];
    return s[(kk * 4 + (lane & 3)) * LD + r + (lane >> 2)];
  }
};

template <int BM, int BN, int WM, int WN, bool A_KM, bool B_KN, int STAGES>
struct GemmCore {
  static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
  static constexpr int NT = WARPS_M * WARPS_N * 32;
  static constexpr int TM = WM / 8, TN = WN / 8;
  using TA = OpTile<BM, A_KM>;
  using TB = OpTile<BN, B_KN>;
  static constexpr int STAGE_ELEMS = TA::ELEMS + TB::ELEMS;
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_ELEMS * sizeof(double);

  // acc[i][j][0..1] += sum_k A(m0+.., k) B(n0+.., k) over k in [k_begin, k_end) (multiples of GEMM_BK).
  // All threads of the CTA must call this together; smem must hold SMEM_BYTES.
  __device__ static __forceinline__ void run(double (&acc)[TM][TN][2], const double* __restrict__ A, int lda,
                                             int m0, const double* __restrict__ B, int ldb, int n0,
                                             int k_begin, int k_end, double* smem) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm0 = (warp / WARPS_N) * WM, wn0 = (warp % WARPS_N) * WN;
    const int nk = (k_end - k_begin) / GEMM_BK;
    __syncthreads();  // previous users of smem are done
    // prologue: STAGES-1 tiles in flight
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
      if (s < nk) {
        double* st = smem + s * STAGE_ELEMS;
        TA::template load_async<NT>(st, A, lda, m0, k_begin + s * GEMM_BK, tid);
        TB::template load_async<NT>(st + TA::ELEMS, B, ldb, n0, k_begin + s * GEMM_BK, tid);
      }
      cp_async_commit();
    }
    for (int it = 0; it < nk; ++it) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();  // tile `it` landed for everyone; everyone finished computing tile it-1
      {
        int nx = it + STAGES - 1;
        if (nx < nk) {
          double* st = smem + (nx % STAGES) * STAGE_ELEMS;
          TA::template load_async<NT>(st, A, lda, m0, k_begin + nx * GEMM_BK, tid);
          TB::template load_async<NT>(st + TA::ELEMS, B, ldb, n0, k_begin + nx * GEMM_BK, tid);
        }
        cp_async_commit();
      }
      const double* sa = smem + (it % STAGES) * STAGE_ELEMS;
      const double* sb = sa + TA::ELEMS;
#pragma unroll
      for (int kk = 0; kk < GEMM_BK / 4; ++kk) {
        double a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = TA::frag(sa, wm0 + i * 8, kk, lane);
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = TB::frag(sb, wn0 + j * 8, kk, lane);
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    }
    cp_async_wait<0>();
  }

  // coordinates of acc[i][j][e] inside the CTA tile
  __device__ static __forceinline__ int row_of(int i) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    return (warp / WARPS_N) * WM + i * 8 + (lane >> 2);
  }
  __device__ static __forceinline__ int col_of(int j) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    return (warp % WARPS_N) * WN + j * 8 + 2 * (lane & 3);
  }
};

// ---------------------------------------------------------------------------------------------------
// Generic batched GEMM kernel:  C = alpha * A.B^T(+layout variants) + beta * C  with triangular k-ranges.
// ---------------------------------------------------------------------------------------------------
struct GemmArgs {
  const double* A;
  const double* B;
  double* C;
  int lda, ldb, ldc;
  long long sA, sB, sC;  // batch strides (elements), batch index = blockIdx.z
  int K;                 // full k extent
  double alpha, beta;
  int lower_only;  // 1: skip CTA tiles that lie entirely above the diagonal (n0 > m0 + BM - 1)
  int kb_mode;     // k_begin: 0 -> 0, 1 -> n0, 2 -> m0
  int ke_mode;     // k_end:   0 -> K, 1 -> n0 + BN, 2 -> m0 + BM      (clamped to K)
};

template <int BM, int BN, int WM, int WN, bool A_KM, bool B_KN, int STAGES>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32)
    dgemm_kernel(GemmArgs p) {
  using Core = GemmCore<BM, BN, WM, WN, A_KM, B_KN, STAGES>;
  extern __shared__ __align__(16) double smem_d[];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (p.lower_only && n0 > m0 + BM - 1) return;
  const double* A = p.A + (long long)blockIdx.z * p.sA;
  const double* B = p.B + (long long)blockIdx.z * p.sB;
  double* C = p.C + (long long)blockIdx.z * p.sC;
  int kb = p.kb_mode == 1 ? n0 : (p.kb_mode == 2 ? m0 : 0);
  int ke = p.ke_mode == 1 ? n0 + BN : (p.ke_mode == 2 ? m0 + BM : p.K);
  kb = (kb / GEMM_BK) * GEMM_BK;
  if (ke > p.K) ke = p.K;
  double acc[Core::TM][Core::TN][2];
#pragma unroll
  for (int i = 0; i < Core::TM; ++i)
#pragma unroll
    for (int j = 0; j < Core::TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  if (ke > kb) Core::run(acc, A, p.lda, m0, B, p.ldb, n0, kb, ke, smem_d);
#pragma unroll
  for (int i = 0; i < Core::TM; ++i) {
    int r = m0 + Core::row_of(i);
#pragma unroll
    for (int j = 0; j < Core::TN; ++j) {
      int c = n0 + Core::col_of(j);
      double2* dst = reinterpret_cast<double2*>(C + (size_t)r * p.ldc + c);
      double2 v;
      if (p.beta != 0.0) {
        double2 o = *dst;
        v.x = p.alpha * acc[i][j][0] + p.beta * o.x;
        v.y = p.alpha * acc[i][j][1] + p.beta * o.y;
      } else {
        v.x = p.alpha * acc[i][j][0];
        v.y = p.alpha * acc[i][j][1];
      }
      *dst = v;
    }
  }
}

}  // namespace b2
