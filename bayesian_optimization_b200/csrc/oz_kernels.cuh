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
// pairs with the same p + q = d accumulate into ONE 32-column TMEM block (|acc_d| <= (d + 1) 64 2^14 < 2^24), so a tile
// costs S (S + 1) / 2 pairs x 2 k-steps of M = 128, N = 32, K = 32 MMAs and S TMEM blocks; the epilogue re-assembles
// the S int32 blocks into two int64 halves (diagonals 0..3 and 4..S-1, each below 2^48), converts each ONCE to float64
// (exact) and applies the power-of-two row / column scales -- exact multiplications: C <- (C - hi s) - lo s costs two
// roundings of C (the fp64 DMMA kernel it replaces rounds 64 times).
//
// The tile is TRANSPOSED with respect to C: the 128 TMEM lanes are 128 consecutive C columns (rows of the panel taken
// as the A operand), the 32 accumulator columns are 32 C rows (B operand), which makes every warp-wide access to the
// C tile a 256-byte row segment.  C tiles travel by TMA both ways (cp.async.bulk.tensor load into a 4-stage ring with
// an L2 prefetch a few tiles ahead, update in shared memory, bulk-group store back), the digit planes by TMA with the
// 128-byte swizzle the UMMA descriptors expect.  (Variant measured and dropped, profiles/r02/oz_v2_*: 64-row tiles with
// the diagonals in two TMEM groups and C read / written straight from the epilogue's registers -- the global
// read-modify-write alone took 58 us of the 67 us the TMA ring needs for everything.)
//
//   warp 0      TMA loads of the digit planes (A per work item, B per tile)
//   warp 1      MMA issuer (one elected lane), TMEM double-buffered: tile t+1 is multiplied while tile t is drained
//   warps 2-9   epilogue: tcgen05.ld, int64 re-assembly, C update in shared memory
//   warp 10     TMA stores of finished C tiles
//   warp 11     TMA loads (+ L2 prefetches) of the C tiles
//
// Work item = (128-column block cb, up to G row blocks of 32): the A digits of a column block are loaded once per item.
#pragma once
#include "fast_kernels.cuh"

namespace b2 {
namespace oz {

using namespace fk;

constexpr int TM = 128;                 // C columns per tile = UMMA M = TMEM lanes
constexpr int TN = 32;                  // C rows per tile = UMMA N = accumulator columns per digit diagonal
constexpr int KP = 64;                  // panel width (k extent)
constexpr int NPL = 4;                  // digit planes are stored in pairs: 2 x 64 bytes per row -> 128-byte swizzle rows
constexpr int A_PLANE = TM * 128;       // 16 KB
constexpr int B_PLANE = TN * 128;       // 4 KB
constexpr int A_BYTES = NPL * A_PLANE;  // 64 KB
constexpr int B_BYTES = NPL * B_PLANE;  // 16 KB
constexpr int C_BYTES = TN * TM * 8;    // 32 KB
constexpr int B_STAGES = 2, C_STAGES = 4;
constexpr int C_AHEAD = 4;              // L2 prefetch distance of the C tiles (tiles beyond the one being loaded)
constexpr int ACC_COLS = 256;           // TMEM columns of one accumulator buffer (8 diagonals x 32)
constexpr int OFF_A = 0;
constexpr int OFF_B = OFF_A + A_BYTES;
constexpr int OFF_C = OFF_B + B_STAGES * B_BYTES;
constexpr int OFF_BAR = OFF_C + C_STAGES * C_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
constexpr int EPI_WARPS = 8;
constexpr int NT_OZ = 32 * (2 + EPI_WARPS + 2);  // 384 threads

enum {
  BAR_A_FULL = 0,
  BAR_A_EMPTY = 1,
  BAR_B_FULL = 2,     // [2]
  BAR_B_EMPTY = 4,    // [2]
  BAR_ACC_FULL = 6,   // [2]
  BAR_ACC_EMPTY = 8,  // [2] count EPI_WARPS
  BAR_C_FULL = 10,    // [4]
  BAR_C_DONE = 14,    // [4] count EPI_WARPS: the tile has been updated in shared memory
  BAR_C_EMPTY = 18,   // [4] the store has read the stage
  SLOT_TMEM_OZ = 24
};

struct OzArgs {
  const double* scA;  // (rows_pad,) 2^(e - B + 8 (S - 1)) per panel row (taken as a C column)
  const double* scB;  // (rows_pad,) 2^(e - B)             per panel row (taken as a C row)
  int rows;           // panel rows = order of the trailing matrix (multiple of 64)
  int Rcap;           // rows per digit plane pair in the digit buffer
  int G;              // row blocks per work item
  int dbg;            // developer timing knob (wrong results): 1 = two MMAs per tile only, 2 = epilogue skips the arithmetic
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
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(map), "r"(c0), "r"(c1), "r"(src)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

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
oz_syrk_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapC, const OzArgs p) {
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
    tma_prefetch_desc(&mapC);
    mbar_init(BAR(BAR_A_FULL), 1);
    mbar_init(BAR(BAR_A_EMPTY), 1);
    for (int i = 0; i < B_STAGES; ++i) {
      mbar_init(BAR(BAR_B_FULL + i), 1);
      mbar_init(BAR(BAR_B_EMPTY + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(BAR(BAR_ACC_FULL + i), 1);
      mbar_init(BAR(BAR_ACC_EMPTY + i), EPI_WARPS);
    }
    for (int i = 0; i < C_STAGES; ++i) {
      mbar_init(BAR(BAR_C_FULL + i), 1);
      mbar_init(BAR(BAR_C_DONE + i), EPI_WARPS);
      mbar_init(BAR(BAR_C_EMPTY + i), 1);
    }
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
    // ================================ TMA loads: digit planes ================================
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
  } else if (warp == 11) {
    // ================================ TMA loads: C tiles, L2 prefetch C_AHEAD tiles further ================
    if (lane == 0) {
      uint32_t t = 0;
      int cb, rb0, rb1;
      for (int it = blockIdx.x; oz_item(it, ncb, nrb, p.G, cb, rb0, rb1); it += gridDim.x) {
        for (int rb = rb0; rb < rb1; ++rb, ++t) {
          const uint32_t sc = t % C_STAGES;
          if (rb + C_AHEAD < rb1) tma_prefetch_l2_2d(&mapC, TM * cb, TN * (rb + C_AHEAD));
          mbar_wait(BAR(BAR_C_EMPTY + sc), ((t / C_STAGES) & 1) ^ 1, p.err, 22);
          mbar_arrive_expect_tx(BAR(BAR_C_FULL + sc), C_BYTES);
          tma_load_2d(sbase + OFF_C + sc * C_BYTES, &mapC, TM * cb, TN * rb, BAR(BAR_C_FULL + sc));
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
          const uint32_t sb = t % B_STAGES, ab = t & 1;
          mbar_wait(BAR(BAR_B_FULL + sb), (t / B_STAGES) & 1, p.err, 25);
          mbar_wait(BAR(BAR_ACC_EMPTY + ab), ((t >> 1) & 1) ^ 1, p.err, 26);
          tc_fence_after();
          const uint64_t db0 = umma_desc_sw128(sbase + OFF_B + sb * B_BYTES);
          const uint32_t td = tmem_base + ab * ACC_COLS;
#pragma unroll
          for (int d = 0; d < S; ++d) {
            if ((p.dbg & 1) && d > 0) break;
#pragma unroll
            for (int pp = 0; pp <= d; ++pp) {
              const int q = d - pp;
              // digit plane x: pair plane x >> 1, byte offset 64 (x & 1) inside the 128-byte row; k-step: + 32 bytes
              const uint64_t da = da0 + (uint64_t)((pp >> 1) * (A_PLANE >> 4) + (pp & 1) * 4);
              const uint64_t db = db0 + (uint64_t)((q >> 1) * (B_PLANE >> 4) + (q & 1) * 4);
              umma_i8(td + (uint32_t)(d * TN), da, db, idesc, pp != 0);
              umma_i8(td + (uint32_t)(d * TN), da + 2, db + 2, idesc, 1);
            }
          }
          umma_commit(BAR(BAR_B_EMPTY + sb));
          umma_commit(BAR(BAR_ACC_FULL + ab));
        }
        umma_commit(BAR(BAR_A_EMPTY));
      }
    }
  } else if (warp < 2 + EPI_WARPS) {
    // ================================ epilogue ================================
    const int ew = warp - 2;
    const int quad = warp & 3;        // TMEM lane quadrant this warp may read
    const int half = ew >> 2;         // which 16 of the tile's 32 rows
    const int lc = quad * 32 + lane;  // column inside the tile = TMEM lane
    constexpr double HI_SCALE = (double)(1ull << (8 * (S - 4)));
    uint32_t t = 0;
    int cb, rb0, rb1;
    for (int it = blockIdx.x; oz_item(it, ncb, nrb, p.G, cb, rb0, rb1); it += gridDim.x) {
      const int c = TM * cb + lc;
      const double sA = c < p.rows ? p.scA[c] : 0.0;
      for (int rb = rb0; rb < rb1; ++rb, ++t) {
        const uint32_t sc = t % C_STAGES, ab = t & 1;
        double* Cs = (double*)(smem + OFF_C + sc * C_BYTES);
        mbar_wait(BAR(BAR_ACC_FULL + ab), (t >> 1) & 1, p.err, 28);
        tc_fence_after();
        double hs[16], ls[16];
#pragma unroll
        for (int bt = 0; bt < 2; ++bt) {
          const int j0 = 16 * half + 8 * bt;
          uint32_t a[S][8];
#pragma unroll
          for (int d = 0; d < S; ++d)
            tmem_ld8(tmem_base + ((uint32_t)(quad * 32) << 16) + ab * ACC_COLS + (uint32_t)(d * TN + j0), a[d]);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            long long thi = 0, tlo = 0;
#pragma unroll
            for (int d = 0; d < 4; ++d) thi = thi * 256 + (long long)(int)a[d][j];
#pragma unroll
            for (int d = 4; d < S; ++d) tlo = tlo * 256 + (long long)(int)a[d][j];
            const double s = (p.dbg & 2) ? 0.0 : sA * __ldg(p.scB + TN * rb + j0 + j);
            hs[8 * bt + j] = ((double)thi * HI_SCALE) * s;   // exact: an integer below 2^48 times powers of two
            ls[8 * bt + j] = (double)tlo * s;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(BAR_ACC_EMPTY + ab));   // the accumulators are free before C has even arrived
        mbar_wait(BAR(BAR_C_FULL + sc), (t / C_STAGES) & 1, p.err, 27);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          double* cp = Cs + (16 * half + i) * TM + lc;
          *cp = (*cp - hs[i]) - ls[i];
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(BAR_C_DONE + sc));
      }
    }
  } else if (warp == 10) {
    // ================================ TMA stores ================================
    if (lane == 0) {
      uint32_t t = 0;
      int cb, rb0, rb1;
      for (int it = blockIdx.x; oz_item(it, ncb, nrb, p.G, cb, rb0, rb1); it += gridDim.x) {
        for (int rb = rb0; rb < rb1; ++rb, ++t) {
          const uint32_t sc = t % C_STAGES;
          mbar_wait(BAR(BAR_C_DONE + sc), (t / C_STAGES) & 1, p.err, 29);
          fence_proxy_async();
          tma_store_2d(&mapC, TM * cb, TN * rb, sbase + OFF_C + sc * C_BYTES);
          bulk_commit();
          bulk_wait_read0();
          mbar_arrive(BAR(BAR_C_EMPTY + sc));
        }
      }
      bulk_wait0();
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
