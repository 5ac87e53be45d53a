// fast5_kernels.cuh -- CTA-pair fused kernel, producers DECOUPLED from the MMA ring (sm_100a).  Fifth generation.
//
// Generation 4 (fast4_kernels.cuh) computes every cross-correlation chunk r[:, k0:k0+64] of a candidate tile once and
// replays it by TMA from an L2-resident scratch, but its producers still write first uses straight into the A ring:
// they work in bursts of six chunks per accumulator super-tile, cannot start before the ring frees a stage, and the
// MMA issuer idles through every burst (profiles/r01/trace_replay_gen4_mb160.txt: 2800 clk per computed chunk against
// 985 clk per replayed one), and the whole pair waits ~3600 clk per super-tile while the epilogue drains 384 TMEM
// columns at the 64 B/clk TMEM read rate.  Here:
//
//   * the producers write r ONLY to the scratch (row-major fp16, (CTA, chunk, plane) blocks of 128 x 64) and publish
//     a per-group counter; EVERY A operand reaches shared memory by TMA (SWIZZLE_128B box of the scratch), first uses
//     included, so A and B of a chunk complete ONE barrier and the MMA issuer pays one wait per chunk;
//   * the producers run ahead of the MMA by up to a whole tile: the slot of chunk c is free for tile t+1 as soon as
//     the last super-tile of tile t has consumed its replay of c (the TMA thread publishes the consumed-chunk count
//     it learns from the EMPTY barriers), so the next tile's r is built underneath the current tile's MMAs;
//   * the 384-column accumulator is three 128-column blocks with their own FULL / EMPTY barriers: block b is final
//     once the chunk that holds its diagonal has been issued (L^-1 is triangular), so blocks 0 and 1 are drained while
//     the last four / two chunks of the super-tile still run, the diagonal chunks skip the finished blocks, and the
//     next super-tile only waits for the 128 columns of block 2.
//
// Scratch ordering: writer = generic-proxy st.global + fence.proxy.async + CTA barrier + st.release of the counter;
// reader = ld.acquire of the counter + fence.proxy.async + cp.async.bulk.tensor.  Everything else (cta_group::2 M=256
// MMAs, B halves by TMA, Gram product on the tensor cores with its own issuer, two alternating producer groups) is
// generation 3/4.  Needs ld >= 512 (at least two super-tiles and nchunks >= ring depth); the host falls back otherwise.
#pragma once
#include "fast4_kernels.cuh"

namespace b2 {
namespace fk5 {

using namespace fk3;

constexpr int TRACE5_CHUNKS = 800;
constexpr int NBLK = 3;  // 128-column accumulator blocks of a super-tile

enum {
  BAR_FULL = 0,        // [4] TMA of both CTAs (A from the scratch + B halves) -> leader MMA thread (count 2 + tx)
  BAR_EMPTY_ST = 4,    // [4] commit multicast -> TMA thread of both CTAs
  BAR_FULL_X = 8,      // [4] TMA of both CTAs -> leader Gram thread (count 2 + tx)
  BAR_EMPTY_X = 12,    // [4] commit multicast -> training-block TMA thread
  BAR_FULL_G = 16,     // [2] commit multicast -> producers
  BAR_EMPTY_G = 18,    // [2] producer groups of both CTAs -> leader Gram thread (count 2)
  BAR_FULL_AX = 20,    // producers of both CTAs -> leader Gram thread (count 2)
  BAR_ACC_FULL = 21,   // [3] commit multicast -> epilogue, one per accumulator block
  BAR_ACC_EMPTY = 24,  // [3] epilogue warps of both CTAs -> leader MMA thread (count 8)
  BAR_FULL_AUX = 27,   // [3] local TMA -> producers
  BAR_EMPTY_AUX = 30,  // [3] producers -> local TMA
  SLOT_TMEM = 36,
  SLOT_PROD = 37,      // two u32: chunks finished by producer group 0 / 1
  SLOT_CONS = 38       // u32: chunk uses whose MMAs have retired
};

__device__ __forceinline__ uint32_t ld_acquire_cta(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_cta(uint32_t saddr, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
// wait until the monotonic counter at saddr reaches `need` (wrap-safe compare); traps instead of hanging
__device__ __forceinline__ void spin_until_ge(uint32_t saddr, uint32_t need, int* err, int code) {
  if ((int32_t)(ld_acquire_cta(saddr) - need) >= 0) return;
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_cta(saddr) - need) < 0) {
    __nanosleep(40);
    if (clock64() - t0 > WAIT_TIMEOUT_CYCLES) {
      atomicExch(err, code);
      __threadfence_system();
      __trap();
    }
  }
}

// One lane of a converged warp (CUTLASS' elect_one_sync).  The issuing warps run their loops with ALL lanes -- every
// address / descriptor is then a warp-uniform value the compiler keeps in uniform registers -- and only the
// tcgen05 / TMA / arrive instructions sit under this predicate.  A loop that runs inside `if (lane == 0)` instead
// makes the compiler wrap every UTCHMMA / UTMALDG in an ELECT + 7 x R2UR.BROADCAST waterfall (~18 instructions per
// MMA, ~250 dependent single-thread instructions per chunk: that, not the tensor pipe, paced generations 3 and 4).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred px;\n\t"
      "elect.sync _|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}

// issue-under-elect forms (pred = 1 on the elected lane).  They are BRANCHES on purpose: for `if (elected) asm(...)` the
// compiler moves the (warp-uniform) operands into uniform registers directly, while an instruction predicated inside the
// asm on the divergent elect value gets the ELECT + R2UR.BROADCAST waterfall again.
__device__ __forceinline__ void umma_f16_pair_p(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate, uint32_t pred) {
  if (pred) umma_f16_pair(tmem_d, da, db, idesc, accumulate);
}
__device__ __forceinline__ void umma_commit_pair_p(uint32_t bar, uint32_t pred) {
  if (pred) umma_commit_pair(bar);
}
__device__ __forceinline__ void tma_load_2d_pair_p(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster, uint32_t pred) {
  if (pred) tma_load_2d_pair(dst, map, c0, c1, bar_cluster);
}
// TMA load with an L2 eviction-priority hint (createpolicy): the fp16 L^-1 is shared by every SM and re-read 66 times per
// tile -> evict_last; a scratch chunk is dead after its replay in the last super-tile of the tile -> evict_first there.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_2d_pair_hint(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster,
                                                      uint64_t pol, uint32_t pred) {
  if (pred)
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar_cluster), "l"(pol)
        : "memory");
}
// arrive.expect_tx on the leader's copy of a barrier (address already mapped into the leader's window for the peer)
__device__ __forceinline__ void mbar_expect_tx_cluster_p(uint32_t bar_cluster, uint32_t bytes, uint32_t pred) {
  if (pred) asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(bar_cluster), "r"(bytes) : "memory");
}
// first probe inline, the slow path (sleeping try_wait + timeout trap) out of line
__device__ __forceinline__ void mbar_wait_fast(uint32_t bar, uint32_t parity, int* err, int code) {
  if (!mbar_try_wait(bar, parity)) mbar_wait(bar, parity, err, code);
}

// corr_from_acc (fast2_kernels.cuh) on two values: the polynomial part in packed fp32, the MUFU calls per value
template <int CORR>
__device__ __forceinline__ float2 corr2_from_acc(float2 acc) {
  const float2 c14 = make_float2((float)A_SCALE_LOG2, (float)A_SCALE_LOG2);
  if (CORR == RBF) {
    const float2 a = __ffma2_rn(acc, make_float2(-1.f, -1.f), c14);
    return make_float2(ex2_approx(a.x), ex2_approx(a.y));
  }
  const float2 t = make_float2(sqrt_approx(acc.x), sqrt_approx(acc.y));
  const float2 a = __ffma2_rn(t, make_float2(-1.4426950408889634f, -1.4426950408889634f), c14);
  const float2 e = make_float2(ex2_approx(a.x), ex2_approx(a.y));
  if (CORR == MATERN12) return e;
  if (CORR == MATERN32) return __ffma2_rn(t, e, e);
  const float2 q = __ffma2_rn(acc, make_float2(1.0f / 3.0f, 1.0f / 3.0f), __fadd2_rn(make_float2(1.f, 1.f), t));
  return __fmul2_rn(q, e);  // MATERN52
}

template <int CORR, int NPROD, bool TRACE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT2, 1)
predict_fused_decoupled_kernel(const __grid_constant__ fk4::ReplayMaps rmaps, const Fused2Args p, const fk4::ReplayArgs ra) {
  const PairMaps& maps = rmaps.pm;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = (uint64_t*)(smem + OFF_BAR);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  uint32_t* tmem_slot = (uint32_t*)(bars + SLOT_TMEM);
  const uint32_t prod_cnt = BAR(SLOT_PROD), cons_cnt = BAR(SLOT_CONS);
  constexpr int ST = NPROD == 1 ? 4 : 2;                                  // stages (A + B of a chunk)
  constexpr int PLANES = NPROD == 1 ? 1 : 2;
  constexpr int A_STRIDE = NPROD == 1 ? A_HALF_BYTES : A_STAGE_BYTES;    // 16 / 32 KB
  constexpr int B_STRIDE = NPROD == 1 ? BA_PLANE + BB_PLANE : 2 * (BA_PLANE + BB_PLANE);  // 24 / 48 KB
  // B stage layout: [slot A hi][slot A lo (three products)][slot B hi][slot B lo]; slot A holds this CTA's half of the
  // 256-row block (128 rows) or of block 1 alone (64 rows), slot B this CTA's half of block 2 (64 rows)
  constexpr int BOFF_A_LO = BA_PLANE;
  constexpr int BOFF_B_HI = NPROD == 1 ? BA_PLANE : 2 * BA_PLANE;
  constexpr int BOFF_B_LO = BOFF_B_HI + BB_PLANE;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // warp index as a uniform value
  const uint32_t rank = __shfl_sync(0xffffffffu, cluster_ctarank(), 0);
  const bool leader = rank == 0;
  const int ld = p.ld;
  const int n_super = (ld + WC - 1) / WC;
  const int nch = ld / KC;                              // chunks (= scratch slots) per tile; even, >= 8
  int upt = 0;                                          // chunk uses per tile
  for (int s = 0; s < n_super; ++s) upt += min(ld, WC * (s + 1)) / KC;
  const long long n_tiles = (p.M + BM - 1) / BM;
  const long long n_ptiles = (n_tiles + 1) / 2;        // the pair works on tiles 2 pt and 2 pt + 1
  const long long pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  auto LBAR = [&](int i) { return map_to_cta(BAR(i), 0); };  // the leader's copy of barrier i
  auto scr_row = [&](int kc, int plane) { return (((int)blockIdx.x * nch + kc) * PLANES + plane) * BM; };

  if (threadIdx.x == 0) {
    if ((sbase & 1023u) || ra.n_store < nch || nch < 8) {
      atomicExch(p.err, 99);
      __trap();
    }
    tma_prefetch_desc(&rmaps.scr);
    tma_prefetch_desc(&maps.hi128);
    tma_prefetch_desc(&maps.hi64);
    tma_prefetch_desc(&maps.xh32);
    for (int i = 0; i < ST; ++i) {
      mbar_init(BAR(BAR_FULL + i), 2);
      mbar_init(BAR(BAR_EMPTY_ST + i), 1);
    }
    for (int i = 0; i < X_STAGES; ++i) {
      mbar_init(BAR(BAR_FULL_X + i), 2);
      mbar_init(BAR(BAR_EMPTY_X + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(BAR(BAR_FULL_G + i), 1);
      mbar_init(BAR(BAR_EMPTY_G + i), 2);
    }
    for (int i = 0; i < NBLK; ++i) {
      mbar_init(BAR(BAR_ACC_FULL + i), 1);
      mbar_init(BAR(BAR_ACC_EMPTY + i), 8);
    }
    for (int i = 0; i < AUX_STAGES; ++i) {
      mbar_init(BAR(BAR_FULL_AUX + i), 1);
      mbar_init(BAR(BAR_EMPTY_AUX + i), 1);
    }
    mbar_init(BAR(BAR_FULL_AX), 2);
    ((volatile uint32_t*)(bars + SLOT_PROD))[0] = 0u;
    ((volatile uint32_t*)(bars + SLOT_PROD))[1] = 0u;
    ((volatile uint32_t*)(bars + SLOT_CONS))[0] = 0u;
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();  // both CTAs' barriers are initialised and TMEM is allocated before anything crosses over
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ================================ TMA: A (scratch) + this CTA's halves of the L^-1 blocks of every chunk use ======
    // All lanes walk the loop, the elected lane issues (predicated, branch-free).  Replayed chunks below the diagonal
    // blocks (the bulk of the uses) take a short straight-line path.
    {
      const uint32_t el = elect_one_sync() ? 1u : 0u;
      const uint32_t a_bytes = (uint32_t)(PLANES * A_HALF_BYTES);
      const uint32_t full0 = LBAR(BAR_FULL);
      const uint64_t pol_keep = l2_policy_evict_last(), pol_norm = l2_policy_evict_normal(), pol_dead = l2_policy_evict_first();
      uint32_t it = 0, tl = 0, st = 0, ph = 0;
      for (long long pt = pair; pt < n_ptiles; pt += n_pairs, ++tl)
        for (int s = 0; s < n_super; ++s) {
          const int kext = min(ld, WC * (s + 1));
          const int n0 = WC * s;
          const bool act_2 = n0 + NBA < ld;
          const int n_bulk = n0 / KC;                               // replayed chunks under full blocks
          const bool desc = ((n_super - 1 - s) & 1) != 0;           // sawtooth: see the main-MMA warp
          for (int ci = 0; ci < kext / KC; ++ci, ++it) {
            const int kc = ci < n_bulk ? (desc ? n_bulk - 1 - ci : ci) : ci;
            const int k0 = kc * KC;
            mbar_wait_fast(BAR(BAR_EMPTY_ST + st), ph ^ 1, p.err, 1);
            // the MMAs of use it - ST have retired: tell the producers (they reuse scratch slots of the previous tile)
            if (el && it >= (uint32_t)ST) st_release_cta(cons_cnt, it - ST + 1);
            const uint32_t fb = full0 + 8u * st;
            const uint32_t da = sbase + OFF_A + st * A_STRIDE;
            const uint32_t dst = sbase + OFF_B + st * B_STRIDE;
            if (k0 < n0 && act_2) {
              // replay under full blocks: A + 128 rows of blocks 0+1 + 64 rows of block 2
              const uint64_t pol_a = s == n_super - 1 ? pol_dead : pol_norm;   // last replay of this chunk in this tile
              mbar_expect_tx_cluster_p(fb, a_bytes + (uint32_t)((BA_PLANE + BB_PLANE) * PLANES), el);
              tma_load_2d_pair_hint(da, &rmaps.scr, 0, scr_row(kc, 0), fb, pol_a, el);
              if (NPROD == 3) tma_load_2d_pair_hint(da + A_HALF_BYTES, &rmaps.scr, 0, scr_row(kc, 1), fb, pol_a, el);
              tma_load_2d_pair_hint(dst, &maps.hi128, k0, n0 + (int)rank * (NBA / 2), fb, pol_keep, el);
              if (NPROD == 3) tma_load_2d_pair_hint(dst + BOFF_A_LO, &maps.lo128, k0, n0 + (int)rank * (NBA / 2), fb, pol_keep, el);
              tma_load_2d_pair_hint(dst + BOFF_B_HI, &maps.hi64, k0, n0 + NBA + (int)rank * (NBB / 2), fb, pol_keep, el);
              if (NPROD == 3) tma_load_2d_pair_hint(dst + BOFF_B_LO, &maps.lo64, k0, n0 + NBA + (int)rank * (NBB / 2), fb, pol_keep, el);
            } else {
              if (k0 >= n0) {  // first use of this chunk in this tile: the producers must have stored it
                spin_until_ge(prod_cnt + 4u * (uint32_t)(kc & 1), tl * (uint32_t)(nch / 2) + (uint32_t)(kc / 2) + 1u, p.err, 14);
                asm volatile("fence.proxy.async.global;" ::: "memory");
              }
              // which B boxes: k0 < n0+128: blocks 0+1 as one 256-row block (128 rows per CTA) | n0+128 <= k0 < n0+256:
              // block 1 alone (64 rows per CTA) | block 2 (64 rows per CTA) whenever it exists (k0 < kext covers the rest)
              const bool act_01 = k0 < n0 + 128;
              const bool act_1 = !act_01 && k0 < n0 + 256;
              const uint32_t bbytes = (uint32_t)((act_01 ? BA_PLANE : 0) + (act_1 ? BB_PLANE : 0) + (act_2 ? BB_PLANE : 0)) * (uint32_t)PLANES;
              mbar_expect_tx_cluster_p(fb, bbytes + a_bytes, el);
              tma_load_2d_pair_p(da, &rmaps.scr, 0, scr_row(kc, 0), fb, el);
              if (NPROD == 3) tma_load_2d_pair_p(da + A_HALF_BYTES, &rmaps.scr, 0, scr_row(kc, 1), fb, el);
              if (act_01) {
                const int row0 = n0 + (int)rank * (NBA / 2);
                tma_load_2d_pair_p(dst, &maps.hi128, k0, row0, fb, el);
                if (NPROD == 3) tma_load_2d_pair_p(dst + BOFF_A_LO, &maps.lo128, k0, row0, fb, el);
              } else if (act_1) {
                const int row0 = n0 + 128 + (int)rank * 64;
                tma_load_2d_pair_p(dst, &maps.hi64, k0, row0, fb, el);
                if (NPROD == 3) tma_load_2d_pair_p(dst + BOFF_A_LO, &maps.lo64, k0, row0, fb, el);
              }
              if (act_2) {
                const int row0 = n0 + NBA + (int)rank * (NBB / 2);
                tma_load_2d_pair_p(dst + BOFF_B_HI, &maps.hi64, k0, row0, fb, el);
                if (NPROD == 3) tma_load_2d_pair_p(dst + BOFF_B_LO, &maps.lo64, k0, row0, fb, el);
              }
            }
            st = (st + 1) & (ST - 1);
            ph ^= (st == 0);
          }
        }
    }
  } else if (warp == 3) {
    // ================================ TMA: training block halves + aux, once per (tile, chunk) ================
    {
      const bool el = elect_one_sync();
      uint32_t it = 0;
      for (long long pt = pair; pt < n_ptiles; pt += n_pairs)
        for (int kc = 0; kc < nch; ++kc, ++it) {
          const int k0 = kc * KC;
          const uint32_t x = it % X_STAGES, ph = (it / X_STAGES) & 1;
          const uint32_t ax = it % AUX_STAGES, pax = (it / AUX_STAGES) & 1;
          mbar_wait(BAR(BAR_EMPTY_AUX + ax), pax ^ 1, p.err, 12);
          if (el) {
            mbar_arrive_expect_tx(BAR(BAR_FULL_AUX + ax), AUX_BYTES);
            fk2::bulk_load_1d(sbase + OFF_AUX + ax * AUX_BYTES, p.aux + (size_t)kc * 3 * KC, AUX_BYTES, BAR(BAR_FULL_AUX + ax));
          }
          mbar_wait(BAR(BAR_EMPTY_X + x), ph ^ 1, p.err, 7);
          const uint32_t fx = LBAR(BAR_FULL_X + x);
          const uint32_t dst = sbase + OFF_X + x * X_STAGE_BYTES;
          if (el) {
            mbar_arrive_expect_tx_leader(BAR(BAR_FULL_X + x), leader, 2 * XH_PLANE);
            tma_load_2d_pair(dst, &maps.xh32, 0, k0 + 32 * (int)rank, fx);
            tma_load_2d_pair(dst + XH_PLANE, &maps.xl32, 0, k0 + 32 * (int)rank, fx);
          }
          __syncwarp();
        }
    }
  } else if (warp == 2) {
    // ================================ Gram-MMA issuer (leader CTA only) ================================
    if (leader) {
      const bool el = elect_one_sync();
      const uint32_t idesc_gram = umma_idesc_f16(2 * BM, KC);
      const uint64_t dax_hi = umma_desc_sw128(sbase + OFF_AX);
      const uint64_t dax_lo = umma_desc_sw128(sbase + OFF_AX + AX_PLANE);
      uint32_t i = 0, itile = 0;
      for (long long pt = pair; pt < n_ptiles; pt += n_pairs, ++itile) {
        // candidate operands of this tile are in place (every Gram MMA of the previous tile has retired: the producers
        // consumed its last block before they wrote the new operand)
        mbar_wait(BAR(BAR_FULL_AX), itile & 1, p.err, 9);
        for (int lc = 0; lc < nch; ++lc, ++i) {
          const uint32_t x = i % X_STAGES, ph = (i / X_STAGES) & 1;
          mbar_wait(BAR(BAR_FULL_X + x), ph, p.err, 8);
          // Gram block i & 1 is free once the producer groups of both CTAs have read out chunk i - 2
          if (i >= 2) mbar_wait(BAR(BAR_EMPTY_G + (i & 1)), ((i - 2) / 2) & 1, p.err, 13);
          tc_fence_after();
          const uint64_t dx_hi = umma_desc_sw128(sbase + OFF_X + x * X_STAGE_BYTES);
          const uint64_t dx_lo = umma_desc_sw128(sbase + OFF_X + x * X_STAGE_BYTES + XH_PLANE);
          const uint32_t tg = tmem_base + (uint32_t)(G_COL0 + KC * (i & 1));
          if (el) {
            for (int ks = 0; ks < p.dk_steps; ++ks) {
              const uint64_t o = (uint64_t)(ks * 2);
              umma_f16_pair(tg, dax_hi + o, dx_hi + o, idesc_gram, ks != 0);
              umma_f16_pair(tg, dax_hi + o, dx_lo + o, idesc_gram, 1);
              umma_f16_pair(tg, dax_lo + o, dx_hi + o, idesc_gram, 1);
            }
            umma_commit_pair(BAR(BAR_FULL_G + (i & 1)));
            umma_commit_pair(BAR(BAR_EMPTY_X + x));
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ================================ main-MMA issuer (leader CTA only) ================================
    // All 32 lanes walk the loop (uniform descriptors), one elected lane issues (predicated).  The issuer is bound by
    // the LATENCY of its own instruction stream (ncu source counters: its stall samples spread evenly over the loop
    // body, and it slows down further whenever the producers compete for issue slots), so the loop body is kept as short
    // as it goes: running stage / parity counters, descriptors by one multiply-add, a straight-line path for the bulk
    // of the chunks.  The wait for the NEXT chunk's stage sits between the two MMA groups of the current chunk and the
    // commit right behind the last MMA, so the tensor pipe's queue does not run dry inside a barrier instruction.
    if (leader) {
      const uint32_t el = elect_one_sync() ? 1u : 0u;
      const uint32_t idesc_256 = umma_idesc_f16(2 * BM, NBA);
      const uint32_t idesc_128 = umma_idesc_f16(2 * BM, NBB);
      const uint64_t da0 = umma_desc_sw128(sbase + OFF_A), db0 = umma_desc_sw128(sbase + OFF_B);
      constexpr uint64_t A_STEP = A_STRIDE >> 4, B_STEP = B_STRIDE >> 4;     // descriptor start-address units (16 B)
      uint32_t n_my = 0;
      for (long long pt = pair; pt < n_ptiles; pt += n_pairs) ++n_my;
      uint32_t left = n_my * (uint32_t)upt;      // chunk uses still to issue
      uint32_t ic = 0, ist = 0, st = 0, ph = 0;
      if (left) mbar_wait(BAR(BAR_FULL + 0), 0, p.err, 3);
      for (long long pt = pair; pt < n_ptiles; pt += n_pairs) {
        for (int s = 0; s < n_super; ++s, ++ist) {
          const int kext = min(ld, WC * (s + 1));
          const int n0 = WC * s;
          const bool has2 = n0 + NBA < ld;
          const uint32_t pe = (ist & 1) ^ 1;
          // Sawtooth over the replayed chunks: every other super-tile walks them downwards, so that a super-tile starts
          // with the chunks the previous one touched last and the scratch (148 MB at N = 4096, cyclic reads thrash an
          // LRU-like 126 MB L2) is re-read in most-recently-used order.  The diagonal chunks stay last and ascending (the
          // block-final commits depend on it) and the LAST super-tile ascends, so that the slots of the low chunks are
          // released first and the producers keep their head start on the next tile.
          const int n_bulk = n0 / KC;
          const bool desc = ((n_super - 1 - s) & 1) != 0;
          for (int ci = 0; ci < kext / KC; ++ci, ++ic) {
            const int k0 = (ci < n_bulk ? (desc ? n_bulk - 1 - ci : ci) : ci) * KC;
            const bool first = ci == 0;   // first chunk of the super-tile: overwrite the accumulators
            const uint32_t nst = (st + 1) & (ST - 1), nph = ph ^ (nst == 0 ? 1u : 0u);
            const bool tr = TRACE && p.trace && blockIdx.x == 0 && ic < (uint32_t)TRACE5_CHUNKS && lane == 0;
            if (tr) p.trace[ic * 8 + 0] = clock64();
            const uint64_t da_hi = da0 + A_STEP * st, da_lo = da_hi + (A_HALF_BYTES >> 4);
            const uint64_t db_hi = db0 + B_STEP * st, db_lo = db_hi + (BOFF_A_LO >> 4);
            const uint64_t db2_hi = db_hi + (BOFF_B_HI >> 4), db2_lo = db_hi + (BOFF_B_LO >> 4);
            --left;
            if (!first && k0 < n0 && has2) {
              // ---- bulk: a replayed chunk under three full blocks
              tc_fence_after();
              if (el) {
#pragma unroll
                for (int ks = 0; ks < KC / 16; ++ks) {
                  const uint64_t o = (uint64_t)(ks * 2);
                  umma_f16_pair(tmem_base, da_hi + o, db_hi + o, idesc_256, 1);
                  if (NPROD == 3) {
                    umma_f16_pair(tmem_base, da_hi + o, db_lo + o, idesc_256, 1);
                    umma_f16_pair(tmem_base, da_lo + o, db_hi + o, idesc_256, 1);
                  }
                }
              }
              if (left) mbar_wait_fast(BAR(BAR_FULL + nst), nph, p.err, 3);
              if (el) {
#pragma unroll
                for (int ks = 0; ks < KC / 16; ++ks) {
                  const uint64_t o = (uint64_t)(ks * 2);
                  umma_f16_pair(tmem_base + (uint32_t)NBA, da_hi + o, db2_hi + o, idesc_128, 1);
                  if (NPROD == 3) {
                    umma_f16_pair(tmem_base + (uint32_t)NBA, da_hi + o, db2_lo + o, idesc_128, 1);
                    umma_f16_pair(tmem_base + (uint32_t)NBA, da_lo + o, db2_hi + o, idesc_128, 1);
                  }
                }
                umma_commit_pair(BAR(BAR_EMPTY_ST + st));
              }
            } else {
              // ---- first chunk of a super-tile, diagonal chunks, partial last super-tile
              if (first) {  // blocks 0 and 1 of the previous super-tile have been drained (long ago, normally)
                mbar_wait(BAR(BAR_ACC_EMPTY + 0), pe, p.err, 2);
                mbar_wait(BAR(BAR_ACC_EMPTY + 1), pe, p.err, 2);
              }
              tc_fence_after();
              if (tr) p.trace[ic * 8 + 1] = clock64();
              if (k0 < n0 + NBA) {
                // blocks 0+1 (N = 256) below the diagonal of block 0, block 1 alone (N = 128) next to it
                const bool both = k0 < n0 + 128;
                const uint32_t td = tmem_base + (both ? 0u : 128u);
                const uint32_t idesc = both ? idesc_256 : idesc_128;
#pragma unroll
                for (int ks = 0; ks < KC / 16; ++ks) {
                  const uint64_t o = (uint64_t)(ks * 2);
                  umma_f16_pair_p(td, da_hi + o, db_hi + o, idesc, !first || ks != 0, el);
                  if (NPROD == 3) {
                    umma_f16_pair_p(td, da_hi + o, db_lo + o, idesc, 1, el);
                    umma_f16_pair_p(td, da_lo + o, db_hi + o, idesc, 1, el);
                  }
                }
                if (k0 + KC == min(n0 + 128, kext)) umma_commit_pair_p(BAR(BAR_ACC_FULL + 0), el);     // block 0 is final
                if (k0 + KC == min(n0 + 256, kext)) umma_commit_pair_p(BAR(BAR_ACC_FULL + 1), el);     // block 1 is final
              }
              // the next chunk's operands: wait for them now, underneath the MMAs just queued
              if (left) mbar_wait_fast(BAR(BAR_FULL + nst), nph, p.err, 3);
              if (has2) {
                if (first) mbar_wait(BAR(BAR_ACC_EMPTY + 2), pe, p.err, 2);
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < KC / 16; ++ks) {
                  const uint64_t o = (uint64_t)(ks * 2);
                  umma_f16_pair_p(tmem_base + (uint32_t)NBA, da_hi + o, db2_hi + o, idesc_128, !first || ks != 0, el);
                  if (NPROD == 3) {
                    umma_f16_pair_p(tmem_base + (uint32_t)NBA, da_hi + o, db2_lo + o, idesc_128, 1, el);
                    umma_f16_pair_p(tmem_base + (uint32_t)NBA, da_lo + o, db2_hi + o, idesc_128, 1, el);
                  }
                }
              }
              umma_commit_pair_p(BAR(BAR_EMPTY_ST + st), el);  // one commit frees the A and the B stage in both CTAs
            }
            if (tr) p.trace[ic * 8 + 2] = clock64();
            st = nst;
            ph = nph;
          }
          if (!has2) {  // keep the phases of block 2's barriers in step with the super-tile count
            mbar_wait(BAR(BAR_ACC_EMPTY + 2), pe, p.err, 2);
          }
          umma_commit_pair_p(BAR(BAR_ACC_FULL + 2), el);
        }
      }
    }
  } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + 4) {
    // ================================ epilogue (own 128 candidates) ================================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    uint32_t ist = 0;
    for (long long pt = pair; pt < n_ptiles; pt += n_pairs) {
      const long long tile = 2 * pt + rank;
      double ss = 0.0;
      for (int s = 0; s < n_super; ++s, ++ist) {
        const int ncols = min(WC, ld - WC * s);
        for (int b = 0; b < NBLK; ++b) {
          mbar_wait(BAR(BAR_ACC_FULL + b), ist & 1, p.err, 5);
          tc_fence_after();
          if (128 * b < ncols) {
#pragma unroll 1
            for (int c0 = 128 * b; c0 < 128 * b + 128; c0 += 32) {
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
              if (p.dbg_w && tile < n_tiles) {
                float* o = p.dbg_w + (size_t)(tile * BM + row) * ld + WC * s + c0;
#pragma unroll
                for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(r[j]) * p.out_scale;
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(BAR(BAR_ACC_EMPTY + b), leader);
        }
      }
      if (tile < n_tiles) p.sumsq[tile * BM + row] = ss;
    }
  } else if (warp >= PW0) {
    // ================================ producers (own 128 candidates) ================================
    // Two groups of 8 warps take alternate chunks (group = chunk parity = Gram block).  Within a group: 2 warps per TMEM
    // lane quadrant, each thread 32 of the chunk's 64 columns in two passes of 16.  r goes to the scratch only.
    const int pw = warp - PW0;
    const int grp = pw >> 3;
    const int quad = pw & 3;
    const int ch = (pw >> 2) & 1;       // 32-column half of the chunk this thread builds
    const int kq = pw >> 2;             // 16-feature quarter of the candidate operand this thread writes (0..3)
    const int m = quad * 32 + lane;
    const uint32_t row_off = (uint32_t)((m >> 3) * 1024 + (m & 7) * 128);
    float2* exch = p.exch + (size_t)blockIdx.x * 3 * BM;
    const float CG = -2.0f / (float)(1 << (2 * X_SCALE_LOG2));
    const bool elected = (pw & 7) == 0 && lane == 0;
    uint8_t* const scr = (uint8_t*)ra.scratch;
    uint32_t j = 0, tl = 0, done = 0;   // computed chunks so far (all groups) / local tile index / chunks finished by this group
    for (long long pt = pair; pt < n_ptiles; pt += n_pairs, ++tl) {
      const long long tile = 2 * pt + rank;
      float am;
      {
        const long long gm = tile * BM + m;
        double a2 = 0.0;
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) hi[i] = lo[i] = 0u;
        for (int d = 0; d < p.D; ++d) {
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
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint32_t off = row_off + (uint32_t)((((kq * 2 + c) ^ (m & 7)) & 7) * 16);
          *(uint4*)(smem + OFF_AX + off) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
          *(uint4*)(smem + OFF_AX + AX_PLANE + off) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
        }
        fence_proxy_async();
        asm volatile("bar.sync 3, %0;" ::"n"(32 * NPW) : "memory");
        if (pw == 0 && lane == 0) mbar_arrive_leader(BAR(BAR_FULL_AX), leader);
      }
      double ysum_d = 0.0, fsum_d = 0.0;
      for (int kc = 0; kc < nch; ++kc, ++j) {
        if ((kc & 1) != grp) continue;  // the other group's chunk (nch is even: chunk parity = j parity)
        const uint32_t ax = j % AUX_STAGES;
        const bool tr = TRACE && p.trace && blockIdx.x == 0 && j < (uint32_t)TRACE5_CHUNKS && pw == 0 && lane == 0;
        if (tr) p.trace[j * 8 + 3] = clock64();
        mbar_wait(BAR(BAR_FULL_AUX + ax), (j / AUX_STAGES) & 1, p.err, 11);
        mbar_wait(BAR(BAR_FULL_G + grp), (j / 2) & 1, p.err, 10);
        tc_fence_after();
        if (tr) p.trace[j * 8 + 4] = clock64();
        float2 ys2 = make_float2(0.f, 0.f), fs2 = make_float2(0.f, 0.f);
        // read the whole Gram row segment first and hand the TMEM block back: the Gram MMA of chunk j + 2 then
        // runs while this group is still computing
        uint32_t gr0[16], gr1[16];
        fk2::tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(G_COL0 + KC * grp + 32 * ch), gr0);
        fk2::tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(G_COL0 + KC * grp + 32 * ch + 16), gr1);
        tmem_ld_wait();
        tc_fence_before();
        // the scratch slot of this chunk is free once the previous tile's last replay of it has been consumed
        if (elected && tl > 0) spin_until_ge(cons_cnt, tl * (uint32_t)upt - (uint32_t)nch + (uint32_t)kc + 1u, p.err, 15);
        if (grp == 0) asm volatile("bar.sync 1, %0;" ::"n"(32 * NPW / 2) : "memory");
        else asm volatile("bar.sync 2, %0;" ::"n"(32 * NPW / 2) : "memory");
        if (elected) mbar_arrive_leader(BAR(BAR_EMPTY_G + grp), leader);
        if (tr) p.trace[j * 8 + 5] = clock64();
        uint8_t* g_hi = scr + ((size_t)scr_row(kc, 0) + (size_t)m) * 128;
        uint8_t* g_lo = scr + ((size_t)scr_row(kc, PLANES - 1) + (size_t)m) * 128;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int col0 = 32 * ch + 16 * h;
          const uint32_t(&gr)[16] = h == 0 ? gr0 : gr1;
          const float* aux = (const float*)(smem + OFF_AUX + ax * AUX_BYTES) + col0;
          float bj[16];
#pragma unroll
          for (int i = 0; i < 16; i += 4) *(float4*)&bj[i] = *(const float4*)(aux + i);
          // two values per instruction (FFMA2 / FADD2 / FMUL2), as in generation 6
          float kv[16];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float2 g2 = make_float2(__uint_as_float(gr[i]), __uint_as_float(gr[i + 1]));
            float2 a2 = __ffma2_rn(make_float2(CG, CG), g2, __fadd2_rn(make_float2(am, am), make_float2(bj[i], bj[i + 1])));
            a2.x = fmaxf(a2.x, 0.f);
            a2.y = fmaxf(a2.y, 0.f);
            const float2 k2 = corr2_from_acc<CORR>(a2);
            kv[i] = k2.x;
            kv[i + 1] = k2.y;
          }
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 gj = *(const float4*)(aux + KC + i);
            const float4 fj = *(const float4*)(aux + 2 * KC + i);
            const float2 ka = make_float2(kv[i], kv[i + 1]), kb = make_float2(kv[i + 2], kv[i + 3]);
            ys2 = __ffma2_rn(ka, make_float2(gj.x, gj.y), __ffma2_rn(kb, make_float2(gj.z, gj.w), ys2));
            fs2 = __ffma2_rn(ka, make_float2(fj.x, fj.y), __ffma2_rn(kb, make_float2(fj.z, fj.w), fs2));
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float v0 = kv[8 * c + 2 * i], v1 = kv[8 * c + 2 * i + 1];
              const __half2 hh = __floats2half2_rn(v0, v1);
              hi[i] = *(const uint32_t*)&hh;
              if (NPROD == 3) {
                const float2 hf = __half22float2(hh);
                const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
                lo[i] = *(const uint32_t*)&l;
              }
            }
            const uint32_t goff = (uint32_t)(((col0 >> 3) + c) * 16);  // plain row-major: the TMA load applies the swizzle
            *(uint4*)(g_hi + goff) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (NPROD == 3) *(uint4*)(g_lo + goff) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        ysum_d += (double)(ys2.x + ys2.y);
        fsum_d += (double)(fs2.x + fs2.y);
        asm volatile("fence.proxy.async.global;" ::: "memory");  // scratch stores -> visible to the TMA loads
        if (tr) p.trace[j * 8 + 6] = clock64();
        if (grp == 0) asm volatile("bar.sync 1, %0;" ::"n"(32 * NPW / 2) : "memory");
        else asm volatile("bar.sync 2, %0;" ::"n"(32 * NPW / 2) : "memory");
        if (tr) p.trace[j * 8 + 7] = clock64();
        ++done;
        if (elected) {
          st_release_cta(prod_cnt + 4u * (uint32_t)grp, done);
          mbar_arrive(BAR(BAR_EMPTY_AUX + ax));
        }
      }
      // combine the four (group, column-half) partial dot products of a row
      const int part = grp * 2 + ch;
      if (part > 0) exch[(part - 1) * BM + m] = make_float2((float)ysum_d, (float)fsum_d);
      asm volatile("bar.sync 3, %0;" ::"n"(32 * NPW) : "memory");
      if (part == 0 && tile < n_tiles) {
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
      asm volatile("bar.sync 3, %0;" ::"n"(32 * NPW) : "memory");
    }
  }
  tc_fence_before();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the peer may still touch this CTA
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

}  // namespace fk5
}  // namespace b2
