// fast_kernels.cuh -- the M-candidate path on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   predict_fused_tc_kernel   one persistent, warp-specialised kernel per launch:
//     * 8 producer warps build the cross-correlation tile r = corr(theta, |Xc - X|) (gpr.py:486-488) in fp32
//       on the CUDA cores, split every value into an fp16 (hi, lo) pair and store it straight into shared
//       memory in the canonical K-major SWIZZLE_128B operand layout -- r never touches HBM;
//       while the full K range is in flight they also take the two dot products  yhat - beta = r . gamma
//       (gpr.py:490) and  Ft^T rt = r . (L^-T Ft)  (gpr.py:498);
//     * 1 TMA thread streams the pre-split fp16 (hi, lo) copy of L^-1 (row-major = K-major "B" operand);
//     * 1 MMA thread issues tcgen05.mma.kind::f16 with fp32 accumulators in TMEM: three products per
//       k-step (hi*hi + hi*lo + lo*hi) give ~2^-22 relative accuracy at a third of the fp16 tensor rate;
//       rt = L^-1 r^T (gpr.py:494) is accumulated for a 512-column super-tile, skipping the blocks above
//       the diagonal of the triangular L^-1;
//     * 4 epilogue warps read the accumulators with tcgen05.ld and reduce sum(rt^2) per candidate
//       (gpr.py:502) in registers -- rt never leaves the SM either.
//   Outputs per candidate: yhat, sum rt^2, Ft^T rt (float64), consumed by the band / acquisition kernels.
//
// The result is approximate (~1e-6); the arg-max returned by b200bo_acq in B200BO_PREC_FAST is made exact by
// re-scoring, on the fp64 path, every candidate whose error interval reaches the best lower bound
// (band_* kernels below).
// Paths are relative to /root/reference/bayes_optim/ (gpr.py = surrogate/gaussian_process/gpr.py).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gp_math.h"

namespace b2 {
namespace fk {

constexpr int BM = 128;   // candidates per CTA tile = UMMA M = TMEM lanes
constexpr int KC = 64;    // k-chunk: 64 fp16 = one 128-byte swizzle atom row
constexpr int BN = 256;   // columns per MMA instruction / per B stage
constexpr int WCOLS = 512;  // super-tile width = all TMEM columns
constexpr int A_STAGES = 2, B_STAGES = 2;
constexpr int A_HALF_BYTES = BM * KC * 2;          // one fp16 plane (hi or lo) of an A stage
constexpr int A_STAGE_BYTES = 2 * A_HALF_BYTES;    // 32 KB
constexpr int B_HALF_BYTES = BN * KC * 2;          // 32 KB
constexpr int B_STAGE_BYTES = 2 * B_HALF_BYTES;    // 64 KB
constexpr int NUM_PRODUCER_WARPS = 8;
constexpr int PRODUCER_WARP0 = 8;                  // warps 8..15
constexpr int EPI_WARP0 = 4;                       // warps 4..7 (warp % 4 = TMEM lane quadrant)
constexpr int NT = 32 * (PRODUCER_WARP0 + NUM_PRODUCER_WARPS);  // 512 threads
constexpr int A_SCALE_LOG2 = 14;                   // r in [0,1] -> [0, 2^14] before the fp16 split
constexpr uint32_t WAIT_MAX_SPINS = 1u << 24;
constexpr long long WAIT_TIMEOUT_CYCLES = 8000000000LL;  // ~4-5 s: a pipeline bug traps instead of hanging the GPU

__host__ __device__ constexpr int smem_off_A(int s) { return s * A_STAGE_BYTES; }
__host__ __device__ constexpr int smem_off_B(int s) { return A_STAGES * A_STAGE_BYTES + s * B_STAGE_BYTES; }
__host__ __device__ constexpr int smem_off_X() { return A_STAGES * A_STAGE_BYTES + B_STAGES * B_STAGE_BYTES; }
// X chunk staging: [2][DP + 2][KC] floats (rows: scaled features, gamma, f = L^-T Ft), then the barriers
__host__ __device__ constexpr int smem_x_bytes(int DP) { return 2 * (DP + 2) * KC * 4; }
__host__ __device__ constexpr int smem_total(int DP) { return smem_off_X() + smem_x_bytes(DP) + 256 + 1024; }

struct FusedArgs {
  const double* Xc;     // (M, D) candidates, row-major float64
  const float* Xs;      // (DP + 2, ld) fp32: scaled training features (transposed), gamma * 2^-14, f * 2^-14
  const double* cscale; // (D,) per-feature coordinate scale (sqrt(c theta_d) or c theta_d)
  double* yhat;         // (Mpad,)  beta + r . gamma
  double* sumsq;        // (Mpad,)  sum rt^2
  double* dotf;         // (Mpad,)  Ft^T rt
  float* dbg_w;         // NULL or (Mpad, ld): rt as the tensor cores produced it (tests)
  int* err;             // device flag: non-zero when a pipeline wait timed out
  long long M;
  int N, D, ld, corr;
  double beta;
  float out_scale;      // 2^-(A_SCALE_LOG2 + b_scale_log2)
};

// ------------------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (woken by the arrive)
// or the hint expires -- no instruction issue while waiting, which matters under the 1 kW power cap
__device__ __forceinline__ uint32_t mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok;
}
// bounded wait: a pipeline bug must not hang the GPU -- flag it and trap (~2 s: 2^17 sleeps of <= 16 us)
__device__ uint32_t g_wait_hint_ns = 16000u;  // tunable (B200BO_WAIT_HINT_NS); 0 = plain try_wait polling
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  long long t0 = 0;
  const uint32_t hint = g_wait_hint_ns;
  while (!(hint ? mbar_try_wait_hint(bar, parity, hint) : mbar_try_wait(bar, parity))) {
    if ((++spins & 63u) == 0) {  // look at the clock only now and then
      const long long t = clock64();
      if (t0 == 0) t0 = t;
      if (t - t0 > WAIT_TIMEOUT_CYCLES || spins > WAIT_MAX_SPINS) {
        atomicExch(err, code);
        __threadfence_system();
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (sm_100 version bit)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;  // stride byte offset
  d |= 1ull << 46;                    // descriptor version (Blackwell)
  d |= 2ull << 61;                    // SWIZZLE_128B
  return d;
}
// instruction descriptor, kind::f16: A = B = fp16 (K-major), D = fp32, M = 128, N = 256
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void producer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(32 * NUM_PRODUCER_WARPS) : "memory"); }

// correlation value * 2^14 from the accumulated (pre-scaled) distance -- fp32 restatement of kernel.py:
//   RBF / abs-exp : coordinates carry sqrt(theta log2 e) (resp. theta log2 e): r = 2^-acc       :329, :286
//   Matern nu     : coordinates carry sqrt(2 nu theta): t = sqrt(acc) = sqrt(2 nu) h             :184-200
__device__ __forceinline__ float corr_finish_scaled(int corr, float acc) {
  const float LOG2E = 1.4426950408889634f;
  if (corr == RBF || corr == ABSEXP) return ex2_approx((float)A_SCALE_LOG2 - acc);
  const float t = acc > 0.f ? acc * rsqrtf(acc) : 0.f;
  const float e = ex2_approx((float)A_SCALE_LOG2 - t * LOG2E);
  if (corr == MATERN12) return e;
  if (corr == MATERN32) return fmaf(t, e, e);
  return (1.0f + t + acc * (1.0f / 3.0f)) * e;  // MATERN52
}

// ------------------------------------------------------------------------------------------------------
// the kernel.  DP = feature count padded to {8, 16, 32, 64} (padding coordinates are 0 on both sides).
// ------------------------------------------------------------------------------------------------------
template <int DP, bool ABS>
__global__ void __launch_bounds__(NT, 1)
predict_fused_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                        const FusedArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic shared memory is only guaranteed 16-byte aligned: round up to the 1024 B the swizzle atoms need
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  float* xstage = (float*)(smem + smem_off_X());
  uint64_t* bars = (uint64_t*)(smem + smem_off_X() + smem_x_bytes(DP));
  // barrier map: [0,2) full_A  [2,4) empty_A  [4,6) full_B  [6,8) empty_B  [8] tmem_full  [9] tmem_empty
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  uint32_t* tmem_slot = (uint32_t*)(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = p.ld;
  const int n_super = (ld + WCOLS - 1) / WCOLS;
  const long long n_tiles = (p.M + BM - 1) / BM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_hi);
    tma_prefetch_desc(&map_lo);
    for (int i = 0; i < A_STAGES; ++i) {
      mbar_init(BAR(0 + i), NUM_PRODUCER_WARPS);
      mbar_init(BAR(2 + i), 1);
    }
    for (int i = 0; i < B_STAGES; ++i) {
      mbar_init(BAR(4 + i), 1);
      mbar_init(BAR(6 + i), 1);
    }
    mbar_init(BAR(8), 1);
    mbar_init(BAR(9), 4);
    fence_barrier_init();
  }
  if (warp == 2) {  // TMEM: all 512 columns (one CTA per SM: shared memory allows no second CTA)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(WCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer: L^-1 (hi, lo) blocks ================================
    if (lane == 0) {
      uint32_t it = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int s = 0; s < n_super; ++s) {
          const int kext = min(ld, WCOLS * (s + 1));
          for (int k0 = 0; k0 < kext; k0 += KC) {
            for (int hf = 0; hf < 2; ++hf) {
              const int n0 = WCOLS * s + BN * hf;
              if (n0 >= ld || k0 >= n0 + BN) continue;  // beyond the matrix / above the diagonal
              const uint32_t b = it % B_STAGES, ph = (it / B_STAGES) & 1;
              mbar_wait(BAR(6 + b), ph ^ 1, p.err, 1);
              mbar_arrive_expect_tx(BAR(4 + b), B_STAGE_BYTES);
              tma_load_2d(sbase + smem_off_B(b), &map_hi, k0, n0, BAR(4 + b));
              tma_load_2d(sbase + smem_off_B(b) + B_HALF_BYTES, &map_lo, k0, n0, BAR(4 + b));
              ++it;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(BM, BN);
      uint32_t ita = 0, itb = 0, ist = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int s = 0; s < n_super; ++s) {
          const int kext = min(ld, WCOLS * (s + 1));
          mbar_wait(BAR(9), (ist & 1) ^ 1, p.err, 2);  // epilogue has drained the accumulators
          tc_fence_after();
          for (int k0 = 0; k0 < kext; k0 += KC) {
            const uint32_t a = ita % A_STAGES, pha = (ita / A_STAGES) & 1;
            mbar_wait(BAR(0 + a), pha, p.err, 3);
            tc_fence_after();
            const uint64_t da_hi = umma_desc_sw128(sbase + smem_off_A(a));
            const uint64_t da_lo = umma_desc_sw128(sbase + smem_off_A(a) + A_HALF_BYTES);
            for (int hf = 0; hf < 2; ++hf) {
              const int n0 = WCOLS * s + BN * hf;
              if (n0 >= ld || k0 >= n0 + BN) continue;
              const uint32_t b = itb % B_STAGES, phb = (itb / B_STAGES) & 1;
              mbar_wait(BAR(4 + b), phb, p.err, 4);
              tc_fence_after();
              const uint64_t db_hi = umma_desc_sw128(sbase + smem_off_B(b));
              const uint64_t db_lo = umma_desc_sw128(sbase + smem_off_B(b) + B_HALF_BYTES);
              const uint32_t td = tmem_base + (uint32_t)(BN * hf);
#pragma unroll
              for (int ks = 0; ks < KC / 16; ++ks) {
                const uint64_t o = (uint64_t)(ks * 2);  // 16 fp16 = 32 B = 2 descriptor units
                umma_f16(td, da_hi + o, db_hi + o, idesc, (k0 | ks) != 0);
                umma_f16(td, da_hi + o, db_lo + o, idesc, 1);
                umma_f16(td, da_lo + o, db_hi + o, idesc, 1);
              }
              umma_commit(BAR(6 + b));  // B stage free once these MMAs retire
              ++itb;
            }
            umma_commit(BAR(2 + a));  // A stage free
            ++ita;
          }
          umma_commit(BAR(8));  // accumulators of this super-tile complete
          ++ist;
        }
      }
    }
  } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + 4) {
    // ================================ epilogue: sum rt^2 per candidate ================================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    uint32_t ist = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      double ss = 0.0;
      for (int s = 0; s < n_super; ++s) {
        const int ncols = min(WCOLS, ld - WCOLS * s);
        mbar_wait(BAR(8), ist & 1, p.err, 5);
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
            float* o = p.dbg_w + (size_t)(tile * BM + row) * ld + WCOLS * s + c0;
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(r[j]) * p.out_scale;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(9));
        ++ist;
      }
      p.sumsq[tile * BM + row] = ss;
    }
  } else if (warp >= PRODUCER_WARP0) {
    // ================================ producers: r tile -> fp16 (hi, lo) in shared memory ================
    const int pw = warp - PRODUCER_WARP0;
    const int ptid = threadIdx.x - 32 * PRODUCER_WARP0;  // 0..255
    const int m = 16 * pw + (lane & 15);                  // row of the tile
    const int kh = lane >> 4;                             // which 32-wide half of the k-chunk
    const uint32_t row_off = (uint32_t)((m >> 3) * 1024 + (m & 7) * 128);
    constexpr int XROWS = DP + 2;
    auto issue_x = [&](int buf, int chunk) {  // cp.async the (DP+2) x 64 fp32 block of training columns
      for (int e = ptid; e < XROWS * (KC / 4); e += 32 * NUM_PRODUCER_WARPS) {
        const int rr = e / (KC / 4), q4 = e % (KC / 4);
        cp_async16(smem_u32(xstage + (size_t)buf * XROWS * KC + rr * KC + q4 * 4), p.Xs + (size_t)rr * ld + chunk * KC + q4 * 4);
      }
    };
    issue_x(0, 0);
    cp_async_commit_wait_all();
    producer_bar();
    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      float x[DP];
      {
        const long long gm = tile * BM + m;
#pragma unroll
        for (int d = 0; d < DP; ++d) x[d] = (d < p.D && gm < p.M) ? (float)(p.Xc[gm * p.D + d] * p.cscale[d]) : 0.f;
      }
      double ysum = 0.0, fsum = 0.0;
      for (int s = 0; s < n_super; ++s) {
        const int kext = min(ld, WCOLS * (s + 1));
        const bool last = s == n_super - 1;
        for (int k0 = 0; k0 < kext; k0 += KC) {
          const int buf = it & 1;
          {  // prefetch the training block of the next chunk in program order
            int nk = k0 + KC;
            if (nk >= kext) nk = 0;  // next super-tile / next tile restarts at column 0
            issue_x(buf ^ 1, nk / KC);
          }
          const uint32_t a = it % A_STAGES, pha = (it / A_STAGES) & 1;
          mbar_wait(BAR(2 + a), pha ^ 1, p.err, 6);
          const float* xb = xstage + (size_t)buf * XROWS * KC;
          uint8_t* a_hi = smem + smem_off_A(a);
          uint8_t* a_lo = a_hi + A_HALF_BYTES;
#pragma unroll 1
          for (int g = 0; g < 4; ++g) {
            const int j8 = kh * 32 + g * 8;
            float acc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
            for (int d = 0; d < DP; ++d) {
              const float4 u = *(const float4*)(xb + d * KC + j8);
              const float4 v = *(const float4*)(xb + d * KC + j8 + 4);
              const float xd = x[d];
              float t;
              if (ABS) {
                t = xd - u.x; acc[0] += fabsf(t);
                t = xd - u.y; acc[1] += fabsf(t);
                t = xd - u.z; acc[2] += fabsf(t);
                t = xd - u.w; acc[3] += fabsf(t);
                t = xd - v.x; acc[4] += fabsf(t);
                t = xd - v.y; acc[5] += fabsf(t);
                t = xd - v.z; acc[6] += fabsf(t);
                t = xd - v.w; acc[7] += fabsf(t);
              } else {
                t = xd - u.x; acc[0] = fmaf(t, t, acc[0]);
                t = xd - u.y; acc[1] = fmaf(t, t, acc[1]);
                t = xd - u.z; acc[2] = fmaf(t, t, acc[2]);
                t = xd - u.w; acc[3] = fmaf(t, t, acc[3]);
                t = xd - v.x; acc[4] = fmaf(t, t, acc[4]);
                t = xd - v.y; acc[5] = fmaf(t, t, acc[5]);
                t = xd - v.z; acc[6] = fmaf(t, t, acc[6]);
                t = xd - v.w; acc[7] = fmaf(t, t, acc[7]);
              }
            }
            float kv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) kv[i] = (k0 + j8 + i < p.N) ? corr_finish_scaled(p.corr, acc[i]) : 0.f;
            if (last) {  // the last super-tile sweeps the whole K range once: take the two dot products here
              const float* gm_ = xb + DP * KC + j8;
              const float* fv_ = xb + (DP + 1) * KC + j8;
              float py = 0.f, pf = 0.f;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                py = fmaf(kv[i], gm_[i], py);
                pf = fmaf(kv[i], fv_[i], pf);
              }
              ysum += (double)py;
              fsum += (double)pf;
            }
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const __half2 h = __floats2half2_rn(kv[2 * i], kv[2 * i + 1]);
              const float2 hf = __half22float2(h);
              const __half2 l = __floats2half2_rn(kv[2 * i] - hf.x, kv[2 * i + 1] - hf.y);
              hi[i] = *(const uint32_t*)&h;
              lo[i] = *(const uint32_t*)&l;
            }
            const uint32_t off = row_off + (uint32_t)((((kh * 4 + g) ^ (m & 7)) & 7) * 16);
            *(uint4*)(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *(uint4*)(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          fence_proxy_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(0 + a));
          cp_async_commit_wait_all();
          producer_bar();
          ++it;
        }
      }
      // combine the two k-halves of a row (lanes l and l ^ 16) and write
      ysum += __shfl_xor_sync(0xffffffffu, ysum, 16);
      fsum += __shfl_xor_sync(0xffffffffu, fsum, 16);
      if (kh == 0) {
        p.yhat[tile * BM + m] = p.beta + ysum;
        p.dotf[tile * BM + m] = fsum;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(WCOLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------
// one-off state preparation after factor()
// ------------------------------------------------------------------------------------------------------
// max |W| over the lower triangle (block partials; host takes the max and picks the power-of-two scale)
__global__ void absmax_kernel(const double* __restrict__ W, size_t n, double* __restrict__ partial) {
  double m = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    m = fmax(m, fabs(W[i]));
  __shared__ double sm[32];
  for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.0;
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) partial[blockIdx.x] = m;
  }
}

// L^-1 (float64) -> fp16 hi / lo planes of  2^sb * L^-1  (hi + lo carries ~22 significant bits)
__global__ void linv_split_kernel(const double* __restrict__ W, size_t n, double scale, __half* __restrict__ hi,
                                  __half* __restrict__ lo) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = W[i] * scale;
    const __half h = __double2half(v);
    hi[i] = h;
    lo[i] = __double2half(v - (double)__half2float(h));
  }
}

// Xs rows: [0, D) scaled features, [D, DP) zero, DP: gamma 2^-14, DP+1: f 2^-14   (all fp32, ld columns)
__global__ void xs_prep_kernel(const double* __restrict__ Xt, const double* __restrict__ cscale,
                               const double* __restrict__ gamma, const double* __restrict__ fvec, int D, int DP,
                               int ld, float* __restrict__ Xs) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ld) return;
  const double inv = 1.0 / (double)(1 << A_SCALE_LOG2);
  for (int d = 0; d < DP; ++d) Xs[(size_t)d * ld + j] = d < D ? (float)(Xt[(size_t)d * ld + j] * cscale[d]) : 0.f;
  Xs[(size_t)DP * ld + j] = (float)(gamma[j] * inv);
  Xs[(size_t)(DP + 1) * ld + j] = (float)(fvec[j] * inv);
}

// ------------------------------------------------------------------------------------------------------
// arg-max band: which candidates could still be the maximiser once the fast pass's error is allowed for.
//
// Every candidate carries a box [y - dy, y + dy] x [mse - ds, mse + ds] around its fast moments.  A criterion's
// lower / upper bound over the box follows from monotonicity (EI, PI, MGFI decrease with the signed mean; EI and UCB
// increase with s; PI and MGFI are evaluated at the s-end their numerator's sign selects, or at both ends).
//   1. band_thr0_kernel   : thr0_c = max lower bound over a strided SUBSAMPLE -- any lower bound of the true
//                           maximum is a valid threshold
//   2. band_scan_kernel   : all candidates; upper bound first in fp32 (directed-rounded box, explicit margins), in
//                           fp64 only when the fp32 screen cannot reject; survivors -> list0 (~ stride x q entries)
//   3. band_refine_kernel : exact bounds for list0; thr1_c = max lower bound (list0 contains the global best)
//   4. band_filter_kernel : list0 entries whose upper bound reaches thr1_c for some criterion -> the band
// ------------------------------------------------------------------------------------------------------
struct BandArgs {
  const double* yhat;    // (M,) fast
  const double* sumsq;   // (M,)
  const double* dotf;    // (M,)
  const double* params;  // (q,)
  long long M;
  int acq, minimize, estimate_trend, q;
  double sigma2, plugin, G;
  // half-widths of the error intervals.  yhat: dy.  mse of candidate i:
  //   max(ds, ds_abs + ds_rel * sqrt(sum rt_i^2)) + sigma2 (2 |u_i| du + du^2)
  // ds = the calibrated (observed) width, ds_abs / ds_rel = the a-priori rounding model of the tensor-core pass (its
  // standard deviation grows with ||rt||), du = half-width of u = (Ft^T rt - 1) / G        (all in mse units but du)
  double dy, ds;
  double ds_abs = 0.0, ds_rel = 0.0, du = 0.0;
};

struct Box {
  double ylo, yhi, s0, s1;  // signed mean (negated when maximising) and standard deviation, both ends
};

__device__ __forceinline__ Box band_box(const BandArgs& p, long long i) {
  double u2 = 0.0;
  if (p.estimate_trend) {
    const double u = (p.dotf[i] - 1.0) / p.G;
    u2 = u * u;
  }
  const double ss = p.sumsq[i];
  const double mse = (1.0 - ss + u2) * p.sigma2;  // unclipped; the interval ends are clipped (gpr.py:510)
  const double y = p.minimize ? p.yhat[i] : -p.yhat[i];
  const double ds = fmax(p.ds, p.ds_abs + p.ds_rel * sqrt(fmax(ss, 0.0) + 1e-3)) + (2.0 * sqrt(u2) * p.du + p.du * p.du) * p.sigma2;
  Box b;
  b.ylo = y - p.dy;
  b.yhi = y + p.dy;
  b.s0 = sqrt(fmax(mse - ds, 0.0));
  b.s1 = sqrt(fmax(mse + ds, 0.0));
  return b;
}

__device__ __forceinline__ double pi_coef(double y, double eps) { return y > 0 ? 1.0 - eps : 1.0 + eps; }

__device__ __forceinline__ double band_hi(const BandArgs& p, const Box& b, double par) {
  double v;
  switch (p.acq) {
    case ACQ_EI:
      v = acq_ei(b.ylo, b.s1, p.sigma2, p.plugin);
      break;
    case ACQ_UCB:
      v = acq_ucb(b.yhi, b.s1, par);
      break;
    case ACQ_PI:
      if (par < 1.0) {
        v = acq_pi(b.ylo, (p.plugin - pi_coef(b.ylo, par) * b.ylo) >= 0 ? b.s0 : b.s1, p.plugin, par);
      } else {
        v = fmax(fmax(acq_pi(b.ylo, b.s0, p.plugin, par), acq_pi(b.ylo, b.s1, p.plugin, par)),
                 fmax(acq_pi(b.yhi, b.s0, p.plugin, par), acq_pi(b.yhi, b.s1, p.plugin, par)));
      }
      break;
    default: {  // MGFI
      if (par * (p.plugin - b.ylo - 1.0) + par * par * b.s1 * b.s1 / 2.0 > 700.0) return INFINITY;  // exp overflow -> 0 quirk
      v = acq_mgfi(b.ylo, b.s1, p.plugin, par);
      if (p.plugin - b.ylo > 0) v = fmax(v, acq_mgfi(b.ylo, b.s0, p.plugin, par));
    }
  }
  if (v != v) return INFINITY;
  return v + 1e-12 * fabs(v);
}

__device__ __forceinline__ double band_lo(const BandArgs& p, const Box& b, double par) {
  double v;
  switch (p.acq) {
    case ACQ_EI:
      v = acq_ei(b.yhi, b.s0, p.sigma2, p.plugin);
      break;
    case ACQ_UCB:
      v = acq_ucb(b.ylo, b.s0, par);
      break;
    case ACQ_PI:
      if (par < 1.0) {
        v = acq_pi(b.yhi, (p.plugin - pi_coef(b.yhi, par) * b.yhi) >= 0 ? b.s1 : b.s0, p.plugin, par);
      } else {
        v = fmin(fmin(acq_pi(b.ylo, b.s0, p.plugin, par), acq_pi(b.ylo, b.s1, p.plugin, par)),
                 fmin(acq_pi(b.yhi, b.s0, p.plugin, par), acq_pi(b.yhi, b.s1, p.plugin, par)));
      }
      break;
    default: {
      v = acq_mgfi(b.yhi, b.s0, p.plugin, par);
      if (p.plugin - b.yhi > 0) v = fmin(v, acq_mgfi(b.yhi, b.s1, p.plugin, par));
    }
  }
  if (v != v) return -INFINITY;
  return v - 1e-12 * fabs(v);
}

// fp32 over-estimate of band_hi (or +inf when fp32 cannot decide).  The box is rounded outwards, every product
// formula gets a relative margin far above fp32 rounding, EI an absolute margin for its cancellation.
__device__ __forceinline__ float norm_cdf_f(float z) { return 0.5f * erfcf(-z * 0.70710678f); }
__device__ __forceinline__ float band_hi_f32(int acq, float ylo, float s0, float s1, float plugin, float par,
                                            float inv_sqrt_sigma2) {
  if (acq == ACQ_EI) {
    if (s1 * inv_sqrt_sigma2 < 0.99e-6f) return 0.f;  // acq_ei's early-out, with slack
    const float d = plugin - ylo, z = d / s1;
    const float a = d * norm_cdf_f(z), b = s1 * 0.39894228f * __expf(-0.5f * z * z);
    return a + b + 2e-6f * (fabsf(a) + b);
  }
  if (acq == ACQ_PI) {
    const float coef = ylo > 0 ? 1.0f - par : 1.0f + par;
    const float num = plugin - coef * ylo;
    const float sd = num >= 0 ? s0 : s1;
    if (!(sd > 0.f)) return INFINITY;
    return norm_cdf_f(num / sd + 1e-5f * fabsf(num / sd) + 1e-6f) * 1.0001f;
  }
  // MGFI (par already capped at 22.36); the caller handles the plugin - ylo > 0 branch in fp64
  const float e = par * (plugin - ylo - 1.0f) + 0.5f * par * par * s1 * s1;
  if (e > 80.f) return INFINITY;
  const float bp = (plugin - ylo) / s1 + par * s1;
  return norm_cdf_f(bp + 1e-5f * fabsf(bp) + 1e-6f) * __expf(e + 1e-5f * fabsf(e) + 1e-6f) * 1.0001f;
}

// monotone map double -> signed 64-bit key, so atomicMax on the key is a max on the value
__host__ __device__ __forceinline__ long long ord_key(double v) {
#if defined(__CUDA_ARCH__)
  long long k = __double_as_longlong(v);
#else
  long long k;
  memcpy(&k, &v, 8);
#endif
  return k >= 0 ? k : k ^ 0x7FFFFFFFFFFFFFFFLL;
}
__host__ __device__ __forceinline__ double ord_val(long long k) {
  k = k >= 0 ? k : k ^ 0x7FFFFFFFFFFFFFFFLL;
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(k);
#else
  double v;
  memcpy(&v, &k, 8);
  return v;
#endif
}

constexpr int BAND_MAX_Q = 512;  // criteria per call on the band path (parameters live in shared memory)

__device__ __forceinline__ double band_par(const BandArgs& p, int c) {
  return p.acq == ACQ_MGFI ? fmin(p.params[c], 22.36) : p.params[c];  // acquisition_fun.py:262
}

// 1. threshold from a strided subsample: thr_key[c] = max_c lower bound
__global__ void __launch_bounds__(256) band_thr0_kernel(BandArgs p, int stride, long long* __restrict__ thr_key) {
  __shared__ double par_s[BAND_MAX_Q];
  for (int c = threadIdx.x; c < p.q; c += blockDim.x) par_s[c] = band_par(p, c);
  __syncthreads();
  const long long ns = (p.M + stride - 1) / stride;
  for (long long k0 = (long long)blockIdx.x * blockDim.x; k0 < ns; k0 += (long long)gridDim.x * blockDim.x) {
    const long long k = k0 + threadIdx.x;
    const bool on = k < ns;
    Box b;
    if (on) b = band_box(p, k * stride);
    for (int c = 0; c < p.q; ++c) {
      double lo = on ? band_lo(p, b, par_s[c]) : -INFINITY;
      for (int o = 16; o; o >>= 1) lo = fmax(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      if ((threadIdx.x & 31) == 0 && lo > -INFINITY) atomicMax(thr_key + c, ord_key(lo));
    }
  }
}

// 2. scan: candidates whose upper bound reaches thr0 for some criterion -> list (positions >= cap are dropped and
//    reported through the count)
__global__ void __launch_bounds__(256) band_scan_kernel(BandArgs p, const long long* __restrict__ thr_key,
                                                        long long* __restrict__ list, int cap, int* __restrict__ count) {
  __shared__ double par_s[BAND_MAX_Q], thr_s[BAND_MAX_Q];
  for (int c = threadIdx.x; c < p.q; c += blockDim.x) {
    par_s[c] = band_par(p, c);
    thr_s[c] = ord_val(thr_key[c]);
  }
  __syncthreads();
  const float plf = (float)p.plugin, iss = (float)(1.0 / sqrt(p.sigma2));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.M; i += (long long)gridDim.x * blockDim.x) {
    const Box b = band_box(p, i);
    const float ylo = __double2float_rd(b.ylo), s0 = __double2float_rd(b.s0), s1 = __double2float_ru(b.s1);
    bool in = false;
    for (int c = 0; c < p.q && !in; ++c) {
      const double thr = thr_s[c];
      if (p.acq == ACQ_UCB) {
        in = band_hi(p, b, par_s[c]) >= thr;
        continue;
      }
      // fp32 screen only where fp32 resolves the threshold (far from underflow) and the criterion is monotone in s
      const bool screen = thr > 1e-20 && thr < 1e30 && !(p.acq == ACQ_MGFI && p.plugin - b.ylo > 0) &&
                          !(p.acq == ACQ_PI && par_s[c] >= 1.0);
      if (screen && (double)band_hi_f32(p.acq, ylo, s0, s1, plf, (float)par_s[c], iss) < thr) continue;
      in = band_hi(p, b, par_s[c]) >= thr;
    }
    if (in) {
      const int pos = atomicAdd(count, 1);
      if (pos < cap) list[pos] = i;
    }
  }
}

// 3. exact bounds of the listed candidates: hiB (n, q) and thr1_key[c] = max lower bound.  One thread per (entry,
//    criterion) pair -- the list holds tens of entries, a loop over the criteria would leave the GPU idle.
__global__ void __launch_bounds__(256) band_refine_kernel(BandArgs p, const long long* __restrict__ list,
                                                          const int* __restrict__ count, int cap,
                                                          double* __restrict__ hiB, long long* __restrict__ thr_key) {
  const int n = min(*count, cap);
  const long long total = (long long)n * p.q;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int bi = (int)(e / p.q), c = (int)(e % p.q);
    const Box b = band_box(p, list[bi]);
    const double par = band_par(p, c);
    hiB[e] = band_hi(p, b, par);
    const double lo = band_lo(p, b, par);
    if (lo > -INFINITY) atomicMax(thr_key + c, ord_key(lo));
  }
}

// 4. the band: listed candidates that reach the refined threshold of some criterion
__global__ void __launch_bounds__(256) band_filter_kernel(const long long* __restrict__ list, const int* __restrict__ count,
                                                          int cap, const double* __restrict__ hiB,
                                                          const long long* __restrict__ thr_key, int q,
                                                          long long* __restrict__ out, int* __restrict__ out_count) {
  const int n = min(*count, cap);
  for (int bi = blockIdx.x * blockDim.x + threadIdx.x; bi < n; bi += gridDim.x * blockDim.x) {
    bool in = false;
    for (int c = 0; c < q && !in; ++c) in = hiB[(size_t)bi * q + c] >= ord_val(thr_key[c]);
    if (in) out[atomicAdd(out_count, 1)] = list[bi];
  }
}

// list[i] = i * stride (the strided calibration / check samples)
__global__ void iota_stride_kernel(long long* __restrict__ list, int n, long long stride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) list[i] = (long long)i * stride;
}

__global__ void list_offset_kernel(long long* __restrict__ list, int n, long long off) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) list[i] += off;
}

// Xb[b, :] = Xc[list[b] - idx_base, :]
__global__ void band_gather_kernel(const double* __restrict__ Xc, const long long* __restrict__ list, int nb,
                                   long long idx_base, int D, double* __restrict__ Xb) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nb * D) return;
  const int b = e / D, d = e % D;
  Xb[e] = Xc[(size_t)(list[b] - idx_base) * D + d];
}

// max |fast - exact| of yhat and of the (unclipped) mse over n candidates -> out[0], out[1]
__global__ void band_err_kernel(const double* __restrict__ y_fast, const double* __restrict__ ss_fast,
                                const double* __restrict__ df_fast, const double* __restrict__ y_ex,
                                const double* __restrict__ ss_ex, const double* __restrict__ df_ex,
                                const long long* __restrict__ list, long long idx_base, int n, int estimate_trend,
                                double G, double sigma2, double* __restrict__ out) {
  __shared__ double s0[8], s1[8];
  double ey = 0.0, es = 0.0;
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < n; b += gridDim.x * blockDim.x) {
    const long long i = list ? list[b] - idx_base : b;
    ey = fmax(ey, fabs(y_fast[i] - y_ex[b]));
    double uf = 0.0, ue = 0.0;
    if (estimate_trend) {
      uf = (df_fast[i] - 1.0) / G;
      ue = (df_ex[b] - 1.0) / G;
    }
    es = fmax(es, fabs((uf * uf - ss_fast[i]) - (ue * ue - ss_ex[b])) * sigma2);
  }
  for (int o = 16; o; o >>= 1) {
    ey = fmax(ey, __shfl_xor_sync(0xffffffffu, ey, o));
    es = fmax(es, __shfl_xor_sync(0xffffffffu, es, o));
  }
  if ((threadIdx.x & 31) == 0) {
    s0[threadIdx.x >> 5] = ey;
    s1[threadIdx.x >> 5] = es;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
      ey = fmax(ey, s0[k]);
      es = fmax(es, s1[k]);
    }
    // single block launch
    out[0] = ey;
    out[1] = es;
  }
}

}  // namespace fk
}  // namespace b2
