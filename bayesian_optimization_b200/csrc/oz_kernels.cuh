// oz_kernels.cuh -- float64-accurate rank-64 trailing update of the Cholesky factorisation on the tcgen05 tensor cores.
//
//   C(r, c) -= sum_{k < 64} P(r, k) P(c, k)      for the lower tiles of the trailing matrix (scipy.linalg.cholesky,
//                                                 surrogate/gaussian_process/gpr.py:795, blocked right-looking form)
//
// tcgen05.mma has no f64 kind, so the float64 panel P is cut into S signed 8-bit digit planes (an error-free "Ozaki"
// splitting: every row of the panel is scaled by a power of two, rounded to a (8 S - 2)-bit integer and written in
// balanced base-256 digits, most significant first) and the product is the exact integer sum
//
//   P P^T (m, n) = 2^(e_m + e_n - 2 B) sum_{p, q} 2^(8 (2 S - 2 - p - q))  (digit_p(m, :) . digit_q(n, :))
//
// of int8 x int8 -> int32 tensor-core products (kind::i8: exact, no rounding anywhere).  Digit pairs with p + q >= S
// are dropped (below 2^-(8 S - 2) of the row-scale product: S = 7 gives float64-grade, S = 8 a sub-ulp update).  All
// pairs with the same p + q = d accumulate into ONE 64-column TMEM block (|acc_d| <= (d + 1) 64 2^14 < 2^24), so a tile
// costs S (S + 1) / 2 pairs x 2 k-steps of M = 128, N = 64, K = 32 MMAs.  The S blocks form two groups with their own
// TMEM half and barriers -- "hi" = diagonals 0..3 (10 pairs), "lo" = diagonals 4..S-1 -- so the epilogue re-assembles
// the hi group (one exact int64, < 2^48) while the lo group of the same tile is still being multiplied, and the lo
// group while the hi group of the NEXT tile runs.  C <- (C - hi s) - lo s with s = the power-of-two row x column scale:
// both terms are exact in float64, the update costs two roundings of C (the fp64 DMMA kernel it replaces rounds 64 x).
//
// The tile is TRANSPOSED with respect to C: the 128 TMEM lanes are 128 consecutive C columns (rows of the panel taken
// as the A operand), the 64 accumulator columns are 64 C rows (B operand), which makes every warp-wide access to C a
// 256-byte row segment: the epilogue reads and writes C straight from / to global memory, the loads of a tile issued
// before its MMAs are waited for (first version, profiles/r02/oz_v1_*: C tiles through a 3-stage TMA ring in shared
// memory -- load, update, bulk store -- 67 us for the first C3 panel against 54 us for DMMA: the ring's round trip,
// not the MMAs, set the pace).  The digit planes travel by TMA with the 128-byte swizzle the UMMA descriptors expect.
//
//   warp 0      TMA loads (A digits per work item, B digits per tile, 3-stage ring)
//   warp 1      MMA issuer (one elected lane)
//   warps 2-17  epilogue: C loads, tcgen05.ld, int64 re-assembly, C stores
//
// Work item = (128-column block cb, up to G row blocks of 64): the A digits of a column block are loaded once per item.
#pragma once
#include "fast_kernels.cuh"

namespace b2 {
namespace oz {

using namespace fk;

constexpr int TM = 128;                 // C columns per tile = UMMA M = TMEM lanes
constexpr int TN = 64;                  // C rows per tile = UMMA N = accumulator columns per digit diagonal
constexpr int KP = 64;                  // panel width (k extent)
constexpr int NPL = 4;                  // digit planes are stored in pairs: 2 x 64 bytes per row -> 128-byte swizzle rows
constexpr int A_PLANE = TM * 128;       // 16 KB
constexpr int B_PLANE = TN * 128;       // 8 KB
constexpr int A_BYTES = NPL * A_PLANE;  // 64 KB
constexpr int B_BYTES = NPL * B_PLANE;  // 32 KB
constexpr int B_STAGES = 3;
constexpr int GRP_COLS = 256;           // TMEM columns of one diagonal group (4 diagonals x 64)
constexpr int OFF_A = 0;
constexpr int OFF_B = OFF_A + A_BYTES;
constexpr int OFF_BAR = OFF_B + B_STAGES * B_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
constexpr int EPI_WARPS = 16;                 // 4 per TMEM lane quadrant, 16 of the tile's 64 rows each
constexpr int RPW = TN / (EPI_WARPS / 4);     // C rows per epilogue warp
constexpr int EB = 4;                         // accumulator columns per tcgen05.ld batch
constexpr int NT_OZ = 32 * (2 + EPI_WARPS);  // 576 threads

enum {
  BAR_A_FULL = 0,
  BAR_A_EMPTY = 1,
  BAR_B_FULL = 2,     // [3]
  BAR_B_EMPTY = 5,    // [3]
  BAR_X_FULL = 8,     // hi group multiplied
  BAR_X_EMPTY = 9,    // count EPI_WARPS: hi group read out
  BAR_Y_FULL = 10,
  BAR_Y_EMPTY = 11,
  SLOT_TMEM_OZ = 16
};

struct OzArgs {
  const double* scA;  // (rows_pad,) 2^(e - B + 8 (S - 1)) per panel row (taken as a C column)
  const double* scB;  // (rows_pad,) 2^(e - B)             per panel row (taken as a C row)
  double* C;          // trailing matrix, origin = its (0, 0)
  int ldc;
  int rows;           // panel rows = order of the trailing matrix (multiple of 64)
  int Rcap;           // rows per digit plane pair in the digit buffer
  int G;              // row blocks per work item
  int dbg;            // developer timing knob (wrong results): 1 = two MMAs per group only, 2 = epilogue skips the arithmetic
  int* err;
};

// instruction descriptor, kind::i8: A = B = signed 8 bit (K-major), D = int32
__host__ __device__ constexpr uint32_t umma_idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// ---------------------------------------------------------------------------------------------------------------
// digit planes of a panel: one warp per row (64 float64 = 2 per lane)
//   digits[(u * Rcap + row) * 128 + 64 h + k] = digit 2 u + h of P(row, k);  rows >= `rows` (padding) are zero
// ---------------------------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(256) oz_split_kernel(const double* __restrict__ P, int ldp, int rows, int rows_pad, int Rcap,
                                                       uint8_t* __restrict__ digits, double* __restrict__ scA,
                                                       double* __restrict__ scB) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows_pad) return;
  double v0 = 0.0, v1 = 0.0;
  if (row < rows) {
    const double2 v = *reinterpret_cast<const double2*>(P + (size_t)row * ldp + 2 * lane);
    v0 = v.x;
    v1 = v.y;
  }
  double mx = fmax(fabs(v0), fabs(v1));
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  constexpr int B = 8 * S - 2;
  const int e = (mx > 0.0 && mx < INFINITY) ? ilogb(mx) + 1 : 0;  // mx < 2^e
  long long x0 = llrint(scalbn(v0, B - e)), x1 = llrint(scalbn(v1, B - e));
  if (!(mx < INFINITY)) x0 = x1 = 0;  // non-finite panel: the factorisation has failed already (status flag); keep the MMAs defined
  signed char d0[S], d1[S];
#pragma unroll
  for (int p = S - 1; p >= 0; --p) {
    d0[p] = (signed char)(x0 & 0xff);
    x0 = (x0 - (long long)d0[p]) >> 8;
    d1[p] = (signed char)(x1 & 0xff);
    x1 = (x1 - (long long)d1[p]) >> 8;
  }
#pragma unroll
  for (int p = 0; p < 2 * NPL; ++p) {
    uchar2 o;
    o.x = p < S ? (unsigned char)d0[p < S ? p : 0] : 0;
    o.y = p < S ? (unsigned char)d1[p < S ? p : 0] : 0;
    *reinterpret_cast<uchar2*>(digits + ((size_t)(p >> 1) * Rcap + row) * 128 + (p & 1) * 64 + 2 * lane) = o;
  }
  if (lane == 0) {
    scA[row] = row < rows ? scalbn(1.0, e - B + 8 * (S - 1)) : 0.0;
    scB[row] = row < rows ? scalbn(1.0, e - B) : 0.0;
  }
}

// work item `it` -> column block cb and row blocks [rb0, rb1) (units of TN rows); false when past the end
__device__ __forceinline__ bool oz_item(int it, int ncb, int nrb, int G, int& cb, int& rb0, int& rb1) {
  for (cb = 0; cb < ncb; ++cb) {
    const int first = cb * (TM / TN);            // first row block that reaches the diagonal of this column block
    const int n = (nrb - first + G - 1) / G;     // items of this column block
    if (it < n) {
      rb0 = first + it * G;
      rb1 = min(rb0 + G, nrb);
      return true;
    }
    it -= n;
  }
  return false;
}

template <int S>
__global__ void __launch_bounds__(NT_OZ, 1)
oz_syrk_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const OzArgs p) {
  static_assert(S >= 5 && S <= 8, "4 hi diagonals + 1..4 lo diagonals");
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = (uint64_t*)(smem + OFF_BAR);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  uint32_t* tmem_slot = (uint32_t*)(bars + SLOT_TMEM_OZ);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncb = (p.rows + TM - 1) / TM, nrb = p.rows / TN;

  if (threadIdx.x == 0) {
    if (sbase & 1023u) {
      atomicExch(p.err, 98);
      __trap();
    }
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    mbar_init(BAR(BAR_A_FULL), 1);
    mbar_init(BAR(BAR_A_EMPTY), 1);
    for (int i = 0; i < B_STAGES; ++i) {
      mbar_init(BAR(BAR_B_FULL + i), 1);
      mbar_init(BAR(BAR_B_EMPTY + i), 1);
    }
    mbar_init(BAR(BAR_X_FULL), 1);
    mbar_init(BAR(BAR_X_EMPTY), EPI_WARPS);
    mbar_init(BAR(BAR_Y_FULL), 1);
    mbar_init(BAR(BAR_Y_EMPTY), EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA loads ================================
    if (lane == 0) {
      uint32_t t = 0, ni = 0;
      int cb, rb0, rb1;
      for (int it = blockIdx.x; oz_item(it, ncb, nrb, p.G, cb, rb0, rb1); it += gridDim.x, ++ni) {
        mbar_wait(BAR(BAR_A_EMPTY), (ni & 1) ^ 1, p.err, 21);
        mbar_arrive_expect_tx(BAR(BAR_A_FULL), A_BYTES);
        for (int u = 0; u < NPL; ++u) tma_load_2d(sbase + OFF_A + u * A_PLANE, &mapA, 0, u * p.Rcap + TM * cb, BAR(BAR_A_FULL));
        for (int rb = rb0; rb < rb1; ++rb, ++t) {
          const uint32_t sb = t % B_STAGES;
          mbar_wait(BAR(BAR_B_EMPTY + sb), ((t / B_STAGES) & 1) ^ 1, p.err, 23);
          mbar_arrive_expect_tx(BAR(BAR_B_FULL + sb), B_BYTES);
          for (int u = 0; u < NPL; ++u)
            tma_load_2d(sbase + OFF_B + sb * B_BYTES + u * B_PLANE, &mapB, 0, u * p.Rcap + TN * rb, BAR(BAR_B_FULL + sb));
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_i8(TM, TN);
      const uint64_t da0 = umma_desc_sw128(sbase + OFF_A);
      uint32_t t = 0, ni = 0;
      int cb, rb0, rb1;
      for (int it = blockIdx.x; oz_item(it, ncb, nrb, p.G, cb, rb0, rb1); it += gridDim.x, ++ni) {
        mbar_wait(BAR(BAR_A_FULL), ni & 1, p.err, 24);
        for (int rb = rb0; rb < rb1; ++rb, ++t) {
          const uint32_t sb = t % B_STAGES;
          mbar_wait(BAR(BAR_B_FULL + sb), (t / B_STAGES) & 1, p.err, 25);
          const uint64_t db0 = umma_desc_sw128(sbase + OFF_B + sb * B_BYTES);
#pragma unroll
          for (int grp = 0; grp < 2; ++grp) {
            mbar_wait(BAR(grp == 0 ? BAR_X_EMPTY : BAR_Y_EMPTY), (t & 1) ^ 1, p.err, 26);
            tc_fence_after();
#pragma unroll
            for (int d = 4 * grp; d < (grp == 0 ? 4 : S); ++d) {
              if ((p.dbg & 1) && d > 4 * grp) break;
              const uint32_t td = tmem_base + (uint32_t)(grp * GRP_COLS + (d - 4 * grp) * TN);
#pragma unroll
              for (int pp = 0; pp <= d; ++pp) {
                const int q = d - pp;
                // digit plane x: pair plane x >> 1, byte offset 64 (x & 1) inside the 128-byte row; k-step: + 32 bytes
                const uint64_t da = da0 + (uint64_t)((pp >> 1) * (A_PLANE >> 4) + (pp & 1) * 4);
                const uint64_t db = db0 + (uint64_t)((q >> 1) * (B_PLANE >> 4) + (q & 1) * 4);
                umma_i8(td, da, db, idesc, pp != 0);
                umma_i8(td, da + 2, db + 2, idesc, 1);
              }
            }
            umma_commit(BAR(grp == 0 ? BAR_X_FULL : BAR_Y_FULL));
          }
          umma_commit(BAR(BAR_B_EMPTY + sb));
        }
        umma_commit(BAR(BAR_A_EMPTY));
      }
    }
  } else {
    // ================================ epilogue ================================
    const int ew = warp - 2;
    const int quad = warp & 3;        // TMEM lane quadrant this warp may read
    const int part = ew >> 2;         // which RPW of the tile's 64 rows
    const int lc = quad * 32 + lane;  // column inside the tile = TMEM lane
    constexpr double HI_SCALE = (double)(1ull << (8 * (S - 4)));
    const uint32_t tl = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(RPW * part);
    uint32_t t = 0;
    int cb, rb0, rb1;
    for (int it = blockIdx.x; oz_item(it, ncb, nrb, p.G, cb, rb0, rb1); it += gridDim.x) {
      const int c = TM * cb + lc;
      const bool cvalid = c < p.rows;
      const double sA = cvalid ? p.scA[c] : 0.0;
      for (int rb = rb0; rb < rb1; ++rb, ++t) {
        const int row0 = TN * rb + RPW * part;
        double* cp = p.C + (size_t)row0 * p.ldc + (cvalid ? c : 0);
        double cv[RPW], hi[RPW];
        if (cvalid) {
#pragma unroll
          for (int i = 0; i < RPW; ++i) cv[i] = cp[(size_t)i * p.ldc];
        }
        // ---- hi group: diagonals 0..3 -> one exact integer below 2^48
        mbar_wait(BAR(BAR_X_FULL), t & 1, p.err, 27);
        tc_fence_after();
#pragma unroll
        for (int bt = 0; bt < RPW / EB; ++bt) {
          uint32_t a[4][EB];
#pragma unroll
          for (int d = 0; d < 4; ++d) tmem_ld4(tl + (uint32_t)(d * TN + EB * bt), a[d]);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < EB; ++j) {
            long long th = 0;
#pragma unroll
            for (int d = 0; d < 4; ++d) th = th * 256 + (long long)(int)a[d][j];
            hi[EB * bt + j] = (p.dbg & 2) ? 0.0 : (double)th * HI_SCALE;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(BAR_X_EMPTY));
        // ---- lo group: diagonals 4..S-1
        mbar_wait(BAR(BAR_Y_FULL), t & 1, p.err, 28);
        tc_fence_after();
#pragma unroll
        for (int bt = 0; bt < RPW / EB; ++bt) {
          uint32_t a[S - 4][EB];
#pragma unroll
          for (int d = 0; d < S - 4; ++d) tmem_ld4(tl + (uint32_t)(GRP_COLS + d * TN + EB * bt), a[d]);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < EB; ++j) {
            long long tq = 0;
#pragma unroll
            for (int d = 0; d < S - 4; ++d) tq = tq * 256 + (long long)(int)a[d][j];
            const double s = sA * __ldg(p.scB + row0 + EB * bt + j);
            const int i = EB * bt + j;
            if (cvalid) cp[(size_t)i * p.ldc] = (cv[i] - hi[i] * s) - (double)tq * s;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(BAR_Y_EMPTY));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

}  // namespace oz
}  // namespace b2
