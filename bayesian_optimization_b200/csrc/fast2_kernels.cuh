// fast2_kernels.cuh -- second-generation fused tensor-core kernel of the M-candidate path (sm_100a).
//
// Same contract as predict_fused_tc_kernel (fast_kernels.cuh) but the squared distance is taken from a Gram
// product ON THE TENSOR CORES as well, so the CUDA cores only do ~15 instructions per cross-correlation value:
//
//   |x - x'|^2_theta = a_m + b_j - 2 <xs_m, Xs_j>,     xs = sqrt(c theta) (x - mu)   (centred => small terms)
//
//   warp 0      TMA: L^-1 (hi, lo) 128 x 64 fp16 boxes -> 3-slot ring                        (B of the main MMA)
//   warp 3      TMA: training block Xs (hi, lo) 64 x 64 fp16 + (b_j, gamma_j, f_j) -> 2-stage ring  (B of the Gram MMA)
//   warp 1      MMA issuer: Gram MMA of chunk i+1 (M=128, N=64) into a double-buffered 64-column TMEM block,
//               then the main MMAs of chunk i (M=128, N=128, three split products) into the 384-column
//               accumulator super-tile; tcgen05.commit hands stages back / results on
//   warps 4-7   epilogue: tcgen05.ld the accumulators, sum rt^2 per candidate                  gpr.py:502
//   warps 8-23  producers: tcgen05.ld the Gram block, r = corr(...) in fp32 (gpr.py:486-488), fp16 (hi, lo) split,
//               store as the K-major SWIZZLE_128B A operand; dot products r.gamma (gpr.py:490), r.f (gpr.py:498)
// Kernels: RBF and Matern-1/2, -3/2, -5/2 (anything that is a function of the theta-weighted L2 distance).
#pragma once
#include "fast_kernels.cuh"

namespace b2 {
namespace fk2 {

using namespace fk;

constexpr int NB = 128;                     // columns per main MMA / per B slot
constexpr int WC = 384;                     // accumulator super-tile width (TMEM columns 0..383)
constexpr int G_COL0 = 384;                 // Gram blocks at TMEM columns 384..447 and 448..511
constexpr int B_SLOTS = 3;
constexpr int B_PLANE = NB * KC * 2;        // 16 KB
constexpr int B_SLOT_BYTES = 2 * B_PLANE;   // 32 KB (hi, lo)
constexpr int X_STAGES = 2;
constexpr int X_PLANE = KC * 128;           // 64 training rows x 128 B = 8 KB
constexpr int AUX_BYTES = 3 * KC * 4;       // b_j, gamma_j, f_j
constexpr int X_STAGE_BYTES = 2 * X_PLANE;
constexpr int AX_PLANE = BM * 128;          // candidate tile, 128 rows x 128 B = 16 KB
constexpr int NPW = 16;                     // producer warps
constexpr int PW0 = 8;
constexpr int NT2 = 32 * (PW0 + NPW);       // 768 threads
constexpr int X_SCALE_LOG2 = 4;
constexpr int TRACE_CHUNKS = 256;             // coordinates x 16 before the fp16 split

constexpr int OFF_A = 0;
constexpr int OFF_B = OFF_A + A_STAGES * A_STAGE_BYTES;          // 64 KB
constexpr int OFF_X = OFF_B + B_SLOTS * B_SLOT_BYTES;            // +96 KB
constexpr int OFF_AX = OFF_X + X_STAGES * X_STAGE_BYTES;         // +34 KB
constexpr int OFF_AUX = OFF_AX + 2 * AX_PLANE;                   // +32 KB
constexpr int AUX_STAGES = 3;                                    // own ring: released by the producers, not by the Gram MMA
constexpr int OFF_BAR = OFF_AUX + AUX_STAGES * AUX_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 320;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");

struct Fused2Args {
  const double* Xc;      // (M, D)
  const double* cscale;  // (D,) sqrt(c theta_d)
  const double* cmean;   // (D,) centre mu_d
  const float* aux;      // (ld / 64, 3, 64): b_j (1e30 on padding), gamma_j 2^-14, f_j 2^-14
  double* yhat;
  double* sumsq;
  double* dotf;
  float2* exch;          // (gridDim.x, 3, 128) scratch: partial dot products of the column quarters 1..3
  float* dbg_w;
  long long* trace;      // NULL, or (TRACE_CHUNKS, 8) clock64 stamps of CTA 0 (developer timeline, B200BO_TRACE=1)
  int* err;
  long long M;
  int N, D, ld, corr, dk_steps;  // dk_steps = ceil(D / 16): k-steps of the Gram MMA
  double beta;
  float out_scale;
};

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// barrier indices
enum {
  BAR_FULL_A = 0,    // [4] producers -> MMA
  BAR_EMPTY_A = 4,   // [4] MMA commit -> producers
  BAR_FULL_B = 8,    // [6] TMA -> MMA
  BAR_EMPTY_B = 14,  // [6] MMA commit -> TMA
  BAR_FULL_X = 20,   // [2] TMA -> MMA, producers
  BAR_EMPTY_X = 22,  // [2] MMA commit -> TMA
  BAR_FULL_G = 24,   // [2] MMA commit -> producers
  BAR_FULL_AX = 26,  // producers -> MMA: candidate operand of the Gram MMA written
  BAR_ACC_FULL = 27, // MMA commit -> epilogue
  BAR_ACC_EMPTY = 28, // epilogue -> MMA (count 4)
  BAR_FULL_AUX = 29,  // [3] TMA -> producers: (b_j, gamma_j, f_j) of a chunk
  BAR_EMPTY_AUX = 32  // [3] producers -> TMA
};

// r * 2^14 from the (clamped) scaled squared distance; CORR is a template parameter so the loop body is branch-free
template <int CORR>
__device__ __forceinline__ float corr_from_acc(float acc) {
  if (CORR == RBF) return ex2_approx((float)A_SCALE_LOG2 - acc);
  const float t = sqrt_approx(acc);
  const float e = ex2_approx(fmaf(t, -1.4426950408889634f, (float)A_SCALE_LOG2));
  if (CORR == MATERN12) return e;
  if (CORR == MATERN32) return fmaf(t, e, e);
  return fmaf(acc, 1.0f / 3.0f, 1.0f + t) * e;  // MATERN52
}

// NPROD = 3: hi*hi + hi*lo + lo*hi (~2^-22);  NPROD = 1: hi*hi only (fp16 operands, ~2^-11) -- the cheap first pass
// whose arg-max band is then re-scored exactly
template <int CORR, int NPROD>
__global__ void __launch_bounds__(NT2, 1)
predict_fused_tc2_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                         const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
                         const Fused2Args p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = (uint64_t*)(smem + OFF_BAR);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  uint32_t* tmem_slot = (uint32_t*)(bars + 36);
  // A ring: the one-product pass stores only the hi plane, so the same 64 KB hold four stages instead of two
  constexpr int A_ST = NPROD == 1 ? 4 : 2;
  constexpr int A_STRIDE = NPROD == 1 ? A_HALF_BYTES : A_STAGE_BYTES;
  // B ring likewise: six 16 KB slots (each lasts 256 tensor cycles -- the ring must cover the L2 latency) or three 32 KB
  constexpr int B_ST = NPROD == 1 ? 6 : 3;
  constexpr int B_STRIDE = NPROD == 1 ? B_PLANE : B_SLOT_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = p.ld;
  const int n_super = (ld + WC - 1) / WC;
  const long long n_tiles = (p.M + BM - 1) / BM;

  if (threadIdx.x == 0) {
    if (sbase & 1023u) {  // the swizzle atoms need a 1024-byte aligned base
      atomicExch(p.err, 99);
      __trap();
    }
    tma_prefetch_desc(&map_hi);
    tma_prefetch_desc(&map_lo);
    tma_prefetch_desc(&map_xh);
    tma_prefetch_desc(&map_xl);
    for (int i = 0; i < A_ST; ++i) {
      mbar_init(BAR(BAR_FULL_A + i), 1);   // one elected producer thread arrives after the producers' named barrier
      mbar_init(BAR(BAR_EMPTY_A + i), 1);
    }
    for (int i = 0; i < B_ST; ++i) {
      mbar_init(BAR(BAR_FULL_B + i), 1);
      mbar_init(BAR(BAR_EMPTY_B + i), 1);
    }
    for (int i = 0; i < X_STAGES; ++i) {
      mbar_init(BAR(BAR_FULL_X + i), 1);
      mbar_init(BAR(BAR_EMPTY_X + i), 1);
      mbar_init(BAR(BAR_FULL_G + i), 1);
    }
    for (int i = 0; i < AUX_STAGES; ++i) {
      mbar_init(BAR(BAR_FULL_AUX + i), 1);
      mbar_init(BAR(BAR_EMPTY_AUX + i), 1);
    }
    mbar_init(BAR(BAR_FULL_AX), 1);
    mbar_init(BAR(BAR_ACC_FULL), 1);
    mbar_init(BAR(BAR_ACC_EMPTY), 4);
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA: L^-1 (hi, lo) slots ================================
    if (lane == 0) {
      uint32_t it = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        for (int s = 0; s < n_super; ++s) {
          const int kext = min(ld, WC * (s + 1));
          for (int k0 = 0; k0 < kext; k0 += KC)
            for (int j = 0; j < WC / NB; ++j) {
              const int n0 = WC * s + NB * j;
              if (n0 >= ld || k0 >= n0 + NB) continue;  // beyond the matrix / above the diagonal
              const uint32_t b = it % B_ST, ph = (it / B_ST) & 1;
              mbar_wait(BAR(BAR_EMPTY_B + b), ph ^ 1, p.err, 1);
              mbar_arrive_expect_tx(BAR(BAR_FULL_B + b), NPROD == 3 ? B_SLOT_BYTES : B_PLANE);
              tma_load_2d(sbase + OFF_B + b * B_STRIDE, &map_hi, k0, n0, BAR(BAR_FULL_B + b));
              if (NPROD == 3) tma_load_2d(sbase + OFF_B + b * B_STRIDE + B_PLANE, &map_lo, k0, n0, BAR(BAR_FULL_B + b));
              ++it;
            }
        }
    }
  } else if (warp == 3) {
    // ================================ TMA: training block of the Gram MMA ================================
    if (lane == 0) {
      uint32_t it = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        for (int s = 0; s < n_super; ++s) {
          const int kext = min(ld, WC * (s + 1));
          for (int k0 = 0; k0 < kext; k0 += KC) {
            const uint32_t x = it % X_STAGES, ph = (it / X_STAGES) & 1;
            const uint32_t ax = it % AUX_STAGES, pax = (it / AUX_STAGES) & 1;
            mbar_wait(BAR(BAR_EMPTY_AUX + ax), pax ^ 1, p.err, 12);
            mbar_arrive_expect_tx(BAR(BAR_FULL_AUX + ax), AUX_BYTES);
            bulk_load_1d(sbase + OFF_AUX + ax * AUX_BYTES, p.aux + (size_t)(k0 / KC) * 3 * KC, AUX_BYTES, BAR(BAR_FULL_AUX + ax));
            mbar_wait(BAR(BAR_EMPTY_X + x), ph ^ 1, p.err, 7);
            mbar_arrive_expect_tx(BAR(BAR_FULL_X + x), 2 * X_PLANE);
            const uint32_t dst = sbase + OFF_X + x * X_STAGE_BYTES;
            tma_load_2d(dst, &map_xh, 0, k0, BAR(BAR_FULL_X + x));
            tma_load_2d(dst + X_PLANE, &map_xl, 0, k0, BAR(BAR_FULL_X + x));
            ++it;
          }
        }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      const uint32_t idesc_main = umma_idesc_f16(BM, NB);
      const uint32_t idesc_gram = umma_idesc_f16(BM, KC);  // N = 64 training points
      const uint64_t dax_hi = umma_desc_sw128(sbase + OFF_AX);
      const uint64_t dax_lo = umma_desc_sw128(sbase + OFF_AX + AX_PLANE);
      uint32_t ic = 0;   // global chunk counter (A stage, X stage, Gram buffer all advance with it)
      uint32_t ib = 0, ist = 0, itile = 0;
      auto issue_gram = [&](uint32_t i) {  // Gram MMA of global chunk i into TMEM block i & 1
        const uint32_t x = i % X_STAGES, ph = (i / X_STAGES) & 1;
        mbar_wait(BAR(BAR_FULL_X + x), ph, p.err, 8);
        tc_fence_after();
        const uint64_t dx_hi = umma_desc_sw128(sbase + OFF_X + x * X_STAGE_BYTES);
        const uint64_t dx_lo = umma_desc_sw128(sbase + OFF_X + x * X_STAGE_BYTES + X_PLANE);
        const uint32_t tg = tmem_base + (uint32_t)(G_COL0 + KC * (i & 1));
        for (int ks = 0; ks < p.dk_steps; ++ks) {
          const uint64_t o = (uint64_t)(ks * 2);
          umma_f16(tg, dax_hi + o, dx_hi + o, idesc_gram, ks != 0);
          umma_f16(tg, dax_hi + o, dx_lo + o, idesc_gram, 1);
          umma_f16(tg, dax_lo + o, dx_hi + o, idesc_gram, 1);
        }
        umma_commit(BAR(BAR_FULL_G + (i & 1)));
        umma_commit(BAR(BAR_EMPTY_X + x));
      };
      int chunks_per_tile = 0;
      for (int s = 0; s < n_super; ++s) chunks_per_tile += min(ld, WC * (s + 1)) / KC;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++itile) {
        // candidate operand of this tile is in place (and every Gram MMA of the previous tile has retired)
        mbar_wait(BAR(BAR_FULL_AX), itile & 1, p.err, 9);
        tc_fence_after();
        issue_gram(ic);
        if (chunks_per_tile > 1) issue_gram(ic + 1);
        int lc = 0;  // chunk index within the tile
        for (int s = 0; s < n_super; ++s) {
          const int kext = min(ld, WC * (s + 1));
          mbar_wait(BAR(BAR_ACC_EMPTY), (ist & 1) ^ 1, p.err, 2);
          tc_fence_after();
          for (int k0 = 0; k0 < kext; k0 += KC, ++lc) {
            const uint32_t a = ic % A_ST, pha = (ic / A_ST) & 1;
            const bool tr = p.trace && blockIdx.x == 0 && ic < TRACE_CHUNKS;
            if (tr) p.trace[ic * 8 + 0] = clock64();
            mbar_wait(BAR(BAR_FULL_A + a), pha, p.err, 3);
            tc_fence_after();
            if (tr) p.trace[ic * 8 + 1] = clock64();
            // run two chunks ahead: the producers have consumed Gram block ic (they signalled FULL_A), so its TMEM
            // columns take the Gram product of chunk ic+2 -- queued BEFORE main(ic), it is ready when main(ic) starts
            if (lc + 2 < chunks_per_tile) issue_gram(ic + 2);
            const uint64_t da_hi = umma_desc_sw128(sbase + OFF_A + a * A_STRIDE);
            const uint64_t da_lo = umma_desc_sw128(sbase + OFF_A + a * A_STRIDE + A_HALF_BYTES);
            for (int j = 0; j < WC / NB; ++j) {
              const int n0 = WC * s + NB * j;
              if (n0 >= ld || k0 >= n0 + NB) continue;
              const uint32_t b = ib % B_ST, phb = (ib / B_ST) & 1;
              mbar_wait(BAR(BAR_FULL_B + b), phb, p.err, 4);
              tc_fence_after();
              const uint64_t db_hi = umma_desc_sw128(sbase + OFF_B + b * B_STRIDE);
              const uint64_t db_lo = umma_desc_sw128(sbase + OFF_B + b * B_STRIDE + B_PLANE);
              const uint32_t td = tmem_base + (uint32_t)(NB * j);
#pragma unroll
              for (int ks = 0; ks < KC / 16; ++ks) {
                const uint64_t o = (uint64_t)(ks * 2);
                umma_f16(td, da_hi + o, db_hi + o, idesc_main, (k0 | ks) != 0);
                if (NPROD == 3) {
                  umma_f16(td, da_hi + o, db_lo + o, idesc_main, 1);
                  umma_f16(td, da_lo + o, db_hi + o, idesc_main, 1);
                }
              }
              umma_commit(BAR(BAR_EMPTY_B + b));
              ++ib;
            }
            umma_commit(BAR(BAR_EMPTY_A + a));
            if (tr) p.trace[ic * 8 + 2] = clock64();
            ++ic;
          }
          umma_commit(BAR(BAR_ACC_FULL));
          ++ist;
        }
      }
    }
  } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + 4) {
    // ================================ epilogue ================================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    uint32_t ist = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      double ss = 0.0;
      for (int s = 0; s < n_super; ++s) {
        const int ncols = min(WC, ld - WC * s);
        mbar_wait(BAR(BAR_ACC_FULL), ist & 1, p.err, 5);
        tc_fence_after();
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, r);
          tmem_ld_wait();
          float part = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float v = __uint_as_float(r[j]) * p.out_scale;
            part = fmaf(v, v, part);
          }
          ss += (double)part;
          if (p.dbg_w) {
            float* o = p.dbg_w + (size_t)(tile * BM + row) * ld + WC * s + c0;
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(r[j]) * p.out_scale;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(BAR_ACC_EMPTY));
        ++ist;
      }
      p.sumsq[tile * BM + row] = ss;
    }
  } else if (warp >= PW0) {
    // ================================ producers ================================
    const int pw = warp - PW0;
    const int quad = pw & 3;            // TMEM lane quadrant this warp may read (warp % 4)
    const int kq = pw >> 2;             // which 16-wide quarter of the 64-column chunk
    const int m = quad * 32 + lane;     // row of the tile = TMEM lane
    const uint32_t row_off = (uint32_t)((m >> 3) * 1024 + (m & 7) * 128);
    float2* exch = p.exch + (size_t)blockIdx.x * 3 * BM;  // [3][128] partial dot products of quarters 1..3
    const float CG = -2.0f / (float)(1 << (2 * X_SCALE_LOG2));
    uint32_t ic = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      // ---- candidate operand of the Gram MMA: this thread writes 32 B (16 features) of its row, hi and lo ----
      float am;
      {
        const long long gm = tile * BM + m;
        double a2 = 0.0;
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) hi[i] = lo[i] = 0u;
        for (int d = 0; d < p.D; ++d) {  // every thread of the row needs a_m; only its own quarter is stored
          const double v = gm < p.M ? (p.Xc[gm * p.D + d] - p.cmean[d]) * p.cscale[d] : 0.0;
          a2 += v * v;
          if ((d >> 4) == kq) {
            const float vs = (float)(v * (double)(1 << X_SCALE_LOG2));
            const __half h = __float2half_rn(vs);
            const __half l = __float2half_rn(vs - __half2float(h));
            const int e = d & 15;
            hi[e >> 1] |= (uint32_t)__half_as_ushort(h) << (16 * (e & 1));
            lo[e >> 1] |= (uint32_t)__half_as_ushort(l) << (16 * (e & 1));
          }
        }
        am = (float)a2;
        // the previous tile's Gram MMAs have retired: every producer warp consumed the last Gram block
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint32_t off = row_off + (uint32_t)((((kq * 2 + c) ^ (m & 7)) & 7) * 16);
          *(uint4*)(smem + OFF_AX + off) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
          *(uint4*)(smem + OFF_AX + AX_PLANE + off) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, %0;" ::"n"(32 * NPW) : "memory");
        if (warp == PW0 && lane == 0) mbar_arrive(BAR(BAR_FULL_AX));
      }
      float ysum = 0.f, fsum = 0.f;
      double ysum_d = 0.0, fsum_d = 0.0;
      for (int s = 0; s < n_super; ++s) {
        const int kext = min(ld, WC * (s + 1));
        const bool last = s == n_super - 1;
        for (int k0 = 0; k0 < kext; k0 += KC) {
          const uint32_t g = ic & 1;
          const uint32_t ax = ic % AUX_STAGES;
          const bool tr = p.trace && blockIdx.x == 0 && ic < TRACE_CHUNKS && warp == PW0 && lane == 0;
          if (tr) p.trace[ic * 8 + 3] = clock64();
          mbar_wait(BAR(BAR_FULL_AUX + ax), (ic / AUX_STAGES) & 1, p.err, 11);  // aux block landed (TMA -> this thread)
          mbar_wait(BAR(BAR_FULL_G + g), (ic / 2) & 1, p.err, 10);
          tc_fence_after();
          if (tr) p.trace[ic * 8 + 4] = clock64();
          uint32_t gr[16];
          tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(G_COL0 + KC * g + 16 * kq), gr);
          const float* aux = (const float*)(smem + OFF_AUX + ax * AUX_BYTES) + 16 * kq;
          float bj[16], gj[16], fj[16];
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            *(float4*)&bj[i] = *(const float4*)(aux + i);
            if (last) {
              *(float4*)&gj[i] = *(const float4*)(aux + KC + i);
              *(float4*)&fj[i] = *(const float4*)(aux + 2 * KC + i);
            }
          }
          tmem_ld_wait();
          const uint32_t a = ic % A_ST, pha = (ic / A_ST) & 1;
          mbar_wait(BAR(BAR_EMPTY_A + a), pha ^ 1, p.err, 6);
          if (tr) p.trace[ic * 8 + 5] = clock64();
          float kv[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float acc = fmaxf(fmaf(CG, __uint_as_float(gr[i]), am + bj[i]), 0.f);
            kv[i] = corr_from_acc<CORR>(acc);
          }
          if (last) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              ysum = fmaf(kv[i], gj[i], ysum);
              fsum = fmaf(kv[i], fj[i], fsum);
            }
            ysum_d += (double)ysum;
            fsum_d += (double)fsum;
            ysum = fsum = 0.f;
          }
          uint8_t* a_hi = smem + OFF_A + a * A_STRIDE;
          uint8_t* a_lo = a_hi + A_HALF_BYTES;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float v0 = kv[8 * c + 2 * i], v1 = kv[8 * c + 2 * i + 1];
              const __half2 h = __floats2half2_rn(v0, v1);
              hi[i] = *(const uint32_t*)&h;
              if (NPROD == 3) {
                const float2 hf = __half22float2(h);
                const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
                lo[i] = *(const uint32_t*)&l;
              }
            }
            const uint32_t off = row_off + (uint32_t)((((kq * 2 + c) ^ (m & 7)) & 7) * 16);
            *(uint4*)(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (NPROD == 3) *(uint4*)(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          tc_fence_before();    // our tcgen05.ld of the Gram block is ordered before the MMA that reuses it
          fence_proxy_async();  // generic-proxy stores -> visible to the tensor core
          if (tr) p.trace[ic * 8 + 6] = clock64();
          // ONE arrival per barrier per chunk: every mbarrier event wakes all sleeping warps of the CTA
          asm volatile("bar.sync 1, %0;" ::"n"(32 * NPW) : "memory");
          if (tr) p.trace[ic * 8 + 7] = clock64();
          if (warp == PW0 && lane == 0) {
            mbar_arrive(BAR(BAR_FULL_A + a));
            mbar_arrive(BAR(BAR_EMPTY_AUX + ax));  // the aux block has been read by every producer
          }
          ++ic;
        }
      }
      // combine the four quarters of a row: quarters 1..3 hand their partial sums to quarter 0
      if (kq > 0) exch[(kq - 1) * BM + m] = make_float2((float)ysum_d, (float)fsum_d);
      asm volatile("bar.sync 1, %0;" ::"n"(32 * NPW) : "memory");
      if (kq == 0) {
        double y = ysum_d, f = fsum_d;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float2 e = exch[k * BM + m];
          y += (double)e.x;
          f += (double)e.y;
        }
        p.yhat[tile * BM + m] = p.beta + y;
        p.dotf[tile * BM + m] = f;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * NPW) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// training operand of the Gram MMA: Xh, Xl (ld, 64) fp16 = 16 sqrt(c theta_d) (X_jd - mu_d) split; aux blocks
__global__ void xs2_prep_kernel(const double* __restrict__ Xt, const double* __restrict__ cscale,
                                const double* __restrict__ cmean, const double* __restrict__ gamma,
                                const double* __restrict__ fvec, int N, int D, int ld, __half* __restrict__ Xh,
                                __half* __restrict__ Xl, float* __restrict__ aux) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ld) return;
  double b = 0.0;
  for (int d = 0; d < 64; ++d) {
    double v = 0.0;
    if (d < D && j < N) {
      v = (Xt[(size_t)d * ld + j] - cmean[d]) * cscale[d];
      b += v * v;
    }
    const float vs = (float)(v * (double)(1 << X_SCALE_LOG2));
    const __half h = __float2half_rn(vs);
    Xh[(size_t)j * 64 + d] = h;
    Xl[(size_t)j * 64 + d] = __float2half_rn(vs - __half2float(h));
  }
  const double inv = 1.0 / (double)(1 << A_SCALE_LOG2);
  float* blk = aux + (size_t)(j / KC) * 3 * KC + (j % KC);
  blk[0] = j < N ? (float)b : 1e30f;  // padding points: distance "infinite" => r = 0
  blk[KC] = (float)(gamma[j] * inv);
  blk[2 * KC] = (float)(fvec[j] * inv);
}

}  // namespace fk2
}  // namespace b2
