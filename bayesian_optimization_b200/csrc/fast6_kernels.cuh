// fast6_kernels.cuh -- generation 6 of the fused tensor-core kernel: two CTA pairs SHARE a candidate tile (sm_100a).
//
// Generation 5 (fast5_kernels.cuh) keeps the fp16 cross-correlation r of the 128 candidates each SM is working on in a
// global scratch (1 MB per SM at N = 4096) and replays it by TMA for every accumulator super-tile.  148 SMs x 1 MB does
// not fit the 126 MB L2 next to the fp16 L^-1: ncu shows 35.6 KB of DRAM traffic per candidate (L2 hit rate 69 %), 234 x
// the algorithmic bytes.  Here the candidates in flight are halved instead of the scratch being shrunk:
//
//   * CTA pairs p and p ^ 1 ("sides" 0 and 1 of a group) work on the SAME 256 candidates.  The accumulator super-tiles
//     of the tile are dealt to the two sides so that both carry the same number of chunk uses (host-side greedy split,
//     `side_mask`), every side sums rt^2 over its own super-tiles and both add into the output (atomicAdd of two
//     addends on a zeroed array: commutative, hence bit-reproducible);
//   * the r chunks are PRODUCED once per group: side s builds the chunk pairs with ((kc >> 1) & 1) == s for its
//     128-candidate half (the Gram MMAs, the MUFU work and the yhat / Ft^T rt dot products per SM are halved too) and
//     stores them into the group's scratch, 74 MB in total at N = 4096 -- L2-resident;
//   * both sides replay every chunk from that scratch by TMA.  Chunks of the partner side are awaited through
//     monotonic counters in global memory (writer: st.global, fence.proxy.async, CTA barrier, st.release.gpu of the
//     counter; reader: ld.acquire.gpu, fence.proxy.async, cp.async.bulk.tensor), chunks of the own side through the
//     shared-memory counters of generation 5.  A slot is rewritten for the next tile once BOTH sides have retired
//     their last replay of it (own consumed-use counter in shared memory, the partner's in global memory);
//   * none of the working warps touches the global counters: a 25th "mailbox" warp publishes this side's three
//     shared-memory counters (st.release.gpu) and mirrors the partner's into shared memory (ld.acquire.gpu ->
//     st.release.cta).  The first version had the producers' elected thread fence + release to global memory and poll
//     the partner's counter once per chunk, inside the 256-thread barrier of its group: the tensor pipe fell to 41 %
//     (profiles/r02/gen6_first_*).  The hand-over latency (~3 us) is covered by the producers' head start.
//
// The partner pairs poll each other, so all CTAs of the grid must be co-resident: the host launches at most one CTA
// per SM and falls back to generation 5 when the occupancy query does not confirm it.  Everything else (cta_group::2
// M = 256 MMAs, B halves by TMA, Gram product on the tensor cores with its own issuer, two alternating producer groups,
// per-block accumulator drain, warp-uniform issue under elect.sync) is generation 5.
// Needs ld % 256 == 0 (the chunk pairs alternate between the sides), ld >= 1024 and an even number of CTA pairs.
#pragma once
#include "fast5_kernels.cuh"

namespace b2 {
namespace fk6 {

using namespace fk3;
using namespace fk5;
// the barrier map is generation 5's (fk3 declares an older one under the same names)
using fk5::BAR_FULL; using fk5::BAR_EMPTY_ST; using fk5::BAR_FULL_X; using fk5::BAR_EMPTY_X; using fk5::BAR_FULL_G;
using fk5::BAR_EMPTY_G; using fk5::BAR_FULL_AX; using fk5::BAR_ACC_FULL; using fk5::BAR_ACC_EMPTY; using fk5::BAR_FULL_AUX;
using fk5::BAR_EMPTY_AUX; using fk5::SLOT_TMEM; using fk5::SLOT_PROD; using fk5::SLOT_CONS;

constexpr int NT6 = NT2 + 32;               // + the mailbox warp
constexpr int MIRROR_OFF = 320;             // byte offset of the mirror words behind the barrier block
constexpr int SMEM6_BYTES = fk3::SMEM_BYTES + 64;
static_assert(SMEM6_BYTES <= 232448, "shared memory budget");

struct ShareArgs {
  uint32_t* flags;     // (groups, 2 ranks, 2 sides, 3 counters) x 8 words (one 32-byte sector per counter), zeroed per launch
  uint32_t side_mask;  // bit s = side that owns accumulator super-tile s
  int dead_hint;       // 1: the last replay of a chunk is loaded with L2::evict_first (generation 5's habit)
  int trace_cta;       // developer timeline: which CTA writes it (0 = side 0 of group 0, 2 = side 1)
  int kdb;             // the first kdb chunks of a tile have TWO scratch slots (tile parity): their production for the next
                       // tile does not wait for this tile's replays (0 = every chunk has one slot)
  uint32_t* smid_out;  // NULL, or (gridDim.x,): the SM every CTA ran on (developer: placement of the partner pairs)
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_global_256(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// wait until the monotonic global counter reaches `need`; `seen` caches the last value read (the counters only grow, so
// a cached value that already suffices saves the ~700-cycle round trip to L2)
__device__ __forceinline__ void spin_until_ge_gpu(const uint32_t* p, uint32_t need, uint32_t& seen, int* err, int code) {
  if ((int32_t)(seen - need) >= 0) return;
  seen = ld_acquire_gpu(p);
  if ((int32_t)(seen - need) >= 0) return;
  const long long t0 = clock64();
  while ((int32_t)((seen = ld_acquire_gpu(p)) - need) < 0) {
    __nanosleep(100);
    if (clock64() - t0 > WAIT_TIMEOUT_CYCLES) {
      atomicExch(err, code);
      __threadfence_system();
      __trap();
    }
  }
}

// what one side does per tile: its super-tiles (bits of `mine`), chunk uses per tile, uses of its last super-tile
struct SidePlan {
  uint32_t mine;  // bit s set = super-tile s is this side's
  int n;          // number of super-tiles
  int upt;        // chunk uses per tile
  int ul;         // chunk uses of the LAST super-tile (walked in ascending chunk order)
};
__device__ __forceinline__ SidePlan side_plan(uint32_t side_mask, uint32_t side, int n_super, int ld) {
  SidePlan sp;
  sp.mine = 0;
  sp.n = sp.upt = sp.ul = 0;
  for (int s = 0; s < n_super; ++s)
    if (((side_mask >> s) & 1u) == side) {
      sp.mine |= 1u << s;
      ++sp.n;
      sp.ul = min(ld, WC * (s + 1)) / KC;
      sp.upt += sp.ul;
    }
  return sp;
}
// consumed-use count of a side after which chunk kc of tile (tl - 1) is dead for that side (tl >= 1 tiles issued)
__device__ __forceinline__ uint32_t dead_after(const SidePlan& sp, uint32_t tl, int kc) {
  return tl * (uint32_t)sp.upt - (uint32_t)sp.ul + (kc < sp.ul ? (uint32_t)kc + 1u : 0u);
}

template <int CORR, int NPROD, bool TRACE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT6, 1)
predict_fused_shared_kernel(const __grid_constant__ fk4::ReplayMaps rmaps, const Fused2Args p, const fk4::ReplayArgs ra,
                            const ShareArgs sh) {
  const PairMaps& maps = rmaps.pm;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = (uint64_t*)(smem + OFF_BAR);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  uint32_t* tmem_slot = (uint32_t*)(bars + SLOT_TMEM);
  const uint32_t prod_cnt = BAR(SLOT_PROD), cons_cnt = BAR(SLOT_CONS);
  constexpr int ST = NPROD == 1 ? 4 : 2;
  const uint32_t mir = bar0 + (uint32_t)MIRROR_OFF;   // partner's counters, mirrored by the mailbox warp: prod 0, prod 1, consumed
  constexpr int PLANES = NPROD == 1 ? 1 : 2;
  constexpr int A_STRIDE = NPROD == 1 ? A_HALF_BYTES : A_STAGE_BYTES;
  constexpr int B_STRIDE = NPROD == 1 ? BA_PLANE + BB_PLANE : 2 * (BA_PLANE + BB_PLANE);
  constexpr int BOFF_A_LO = BA_PLANE;
  constexpr int BOFF_B_HI = NPROD == 1 ? BA_PLANE : 2 * BA_PLANE;
  constexpr int BOFF_B_LO = BOFF_B_HI + BB_PLANE;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const uint32_t rank = __shfl_sync(0xffffffffu, cluster_ctarank(), 0);
  const bool leader = rank == 0;
  const int ld = p.ld;
  const int n_super = (ld + WC - 1) / WC;
  const int nch = ld / KC;                      // chunks per tile; multiple of 4, >= 16
  const int nmy = nch / 2;                      // chunks this side produces per tile
  const uint32_t side = (blockIdx.x >> 1) & 1u;
  const long long group = blockIdx.x >> 2, n_groups = gridDim.x >> 2;
  const SidePlan me = side_plan(sh.side_mask, side, n_super, ld), ot = side_plan(sh.side_mask, side ^ 1u, n_super, ld);
  const long long n_tiles = (p.M + BM - 1) / BM;
  const long long n_ptiles = (n_tiles + 1) / 2;  // the group works on tiles 2 pt and 2 pt + 1
  auto LBAR = [&](int i) { return map_to_cta(BAR(i), 0); };
  const int scr_cta = (int)group * 2 + (int)rank;   // the scratch region of this 128-candidate half (shared by both sides)
  // scratch slot of chunk kc in local tile tl: the first kdb chunks alternate between two slots
  const int kdb = sh.kdb, nslot = nch + kdb;
  auto scr_row = [&](int kc, int plane, uint32_t tl) {
    const int slot = kc < kdb ? 2 * kc + (int)(tl & 1u) : kc + kdb;
    return ((scr_cta * nslot + slot) * PLANES + plane) * BM;
  };
  // global counters of the two sides for this half: [0] chunks by producer group 0, [1] by group 1, [2] consumed uses
  uint32_t* const fl_me = sh.flags + (size_t)((scr_cta * 2 + (int)side) * 3) * 8;
  uint32_t* const fl_ot = sh.flags + (size_t)((scr_cta * 2 + (int)(side ^ 1u)) * 3) * 8;
  // chunk index of this side's lc-th chunk: pairs (0,1), (4,5), ... for side 0, (2,3), (6,7), ... for side 1
  auto my_chunk = [&](int lc) { return ((lc >> 1) << 2) + ((int)side << 1) + (lc & 1); };

  if (threadIdx.x == 0 && sh.smid_out) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    sh.smid_out[blockIdx.x] = smid;
  }
  if (threadIdx.x == 0) {
    if ((sbase & 1023u) || ra.n_store < nslot || nch < 16 || (nch & 3) || me.n == 0 || ot.n == 0 || kdb < 0 || kdb > nch) {
      atomicExch(p.err, 97);
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
    for (int i = 0; i < 3; ++i) ((volatile uint32_t*)(smem + OFF_BAR + MIRROR_OFF))[i] = 0u;
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ================================ TMA: A (scratch) + this CTA's halves of the L^-1 blocks of every chunk use ======
    {
      const uint32_t el = elect_one_sync() ? 1u : 0u;
      const uint32_t a_bytes = (uint32_t)(PLANES * A_HALF_BYTES);
      const uint32_t full0 = LBAR(BAR_FULL);
      const uint64_t pol_keep = l2_policy_evict_last(), pol_norm = l2_policy_evict_normal(), pol_dead = l2_policy_evict_first();
      // chunk kc of local tile tl has been stored (by whichever side owns it)
      auto wait_chunk = [&](int kc, uint32_t tl) {
        const uint32_t need = tl * (uint32_t)(nch / 4) + (uint32_t)(kc >> 2) + 1u;
        if ((uint32_t)((kc >> 1) & 1) == side) spin_until_ge(prod_cnt + 4u * (uint32_t)(kc & 1), need, p.err, 14);
        else spin_until_ge(mir + 4u * (uint32_t)(kc & 1), need, p.err, 16);
      };
      uint32_t it = 0, tl = 0, st = 0, ph = 0;
      for (long long pt = group; pt < n_ptiles; pt += n_groups, ++tl) {
        int verified = 0;  // chunks [0, verified) of this tile are known to be in the scratch
        int idx = 0;
        for (int s = 0; s < n_super; ++s) {
          if (!((me.mine >> s) & 1u)) continue;
          const int kext = min(ld, WC * (s + 1));
          const int n0 = WC * s;
          const bool act_2 = n0 + NBA < ld;
          const int n_bulk = n0 / KC;
          const bool desc = ((me.n - 1 - idx) & 1) != 0;   // sawtooth; this side's LAST super-tile ascends
          ++idx;
          if (verified < n_bulk) {
            // replayed chunks this side has not touched yet (the partner's super-tiles in between used them first): the
            // four (side, producer group) counters are monotonic, so the highest chunk of each class covers the range
            for (int kc = max(verified, n_bulk - 4); kc < n_bulk; ++kc) wait_chunk(kc, tl);
            asm volatile("fence.proxy.async.global;" ::: "memory");
            verified = n_bulk;
          }
          // ---- replayed chunks under the diagonal blocks: ST uses per iteration with a compile-time stage (the TMA warp
          // shares its scheduler with four producer warps: the fewer instructions per use, the further it stays ahead)
          const int row_b1 = n0 + (int)rank * (NBA / 2), row_b2 = n0 + NBA + (int)rank * (NBB / 2);
          const uint64_t pol_a = (s == n_super - 1 && sh.dead_hint) ? pol_dead : pol_norm;
          const uint32_t bulk_bytes = a_bytes + (uint32_t)((BA_PLANE + (act_2 ? BB_PLANE : 0)) * PLANES);
          auto bulk = [&](const uint32_t stg, const uint32_t parity, const int kc) {
            const int k0 = kc * KC;
            mbar_wait_fast(BAR(BAR_EMPTY_ST + stg), parity, p.err, 1);
            if (el && it >= (uint32_t)ST) st_release_cta(cons_cnt, it - ST + 1);
            const uint32_t fb = full0 + 8u * stg;
            const uint32_t da = sbase + OFF_A + stg * A_STRIDE;
            const uint32_t dst = sbase + OFF_B + stg * B_STRIDE;
            mbar_expect_tx_cluster_p(fb, bulk_bytes, el);
            const int srow = scr_row(kc, 0, tl);
            tma_load_2d_pair_hint(da, &rmaps.scr, 0, srow, fb, pol_a, el);
            if (NPROD == 3) tma_load_2d_pair_hint(da + A_HALF_BYTES, &rmaps.scr, 0, srow + BM, fb, pol_a, el);
            tma_load_2d_pair_hint(dst, &maps.hi128, k0, row_b1, fb, pol_keep, el);
            if (NPROD == 3) tma_load_2d_pair_hint(dst + BOFF_A_LO, &maps.lo128, k0, row_b1, fb, pol_keep, el);
            if (act_2) {
              tma_load_2d_pair_hint(dst + BOFF_B_HI, &maps.hi64, k0, row_b2, fb, pol_keep, el);
              if (NPROD == 3) tma_load_2d_pair_hint(dst + BOFF_B_LO, &maps.lo64, k0, row_b2, fb, pol_keep, el);
            }
            ++it;
          };
          int ci = 0;
          const int kstep = desc ? -1 : 1;
          int kc = desc ? n_bulk - 1 : 0;
          while (ci < n_bulk && st != 0) {
            bulk(st, ph ^ 1u, kc);
            st = (st + 1) & (ST - 1);
            ph ^= (st == 0);
            ++ci;
            kc += kstep;
          }
          while (ci + ST <= n_bulk) {   // st == 0 here
#pragma unroll
            for (int stg = 0; stg < ST; ++stg) bulk((uint32_t)stg, ph ^ 1u, kc + stg * kstep);
            ph ^= 1u;
            ci += ST;
            kc += ST * kstep;
          }
          while (ci < n_bulk) {
            bulk(st, ph ^ 1u, kc);
            st = (st + 1) & (ST - 1);
            ph ^= (st == 0);
            ++ci;
            kc += kstep;
          }
          // ---- diagonal chunks (ascending): first touch of a chunk by this side waits for its producer
          for (; ci < kext / KC; ++ci, ++it) {
            const int kc = ci;
            const int k0 = kc * KC;
            mbar_wait_fast(BAR(BAR_EMPTY_ST + st), ph ^ 1, p.err, 1);
            if (el && it >= (uint32_t)ST) st_release_cta(cons_cnt, it - ST + 1);
            const uint32_t fb = full0 + 8u * st;
            const uint32_t da = sbase + OFF_A + st * A_STRIDE;
            const uint32_t dst = sbase + OFF_B + st * B_STRIDE;
            if (kc >= verified) {
              wait_chunk(kc, tl);
              asm volatile("fence.proxy.async.global;" ::: "memory");
              verified = kc + 1;
            }
            const bool act_01 = k0 < n0 + 128;
            const bool act_1 = !act_01 && k0 < n0 + 256;
            const uint32_t bbytes = (uint32_t)((act_01 ? BA_PLANE : 0) + (act_1 ? BB_PLANE : 0) + (act_2 ? BB_PLANE : 0)) * (uint32_t)PLANES;
            mbar_expect_tx_cluster_p(fb, bbytes + a_bytes, el);
            tma_load_2d_pair_p(da, &rmaps.scr, 0, scr_row(kc, 0, tl), fb, el);
            if (NPROD == 3) tma_load_2d_pair_p(da + A_HALF_BYTES, &rmaps.scr, 0, scr_row(kc, 1, tl), fb, el);
            if (act_01) {
              tma_load_2d_pair_p(dst, &maps.hi128, k0, row_b1, fb, el);
              if (NPROD == 3) tma_load_2d_pair_p(dst + BOFF_A_LO, &maps.lo128, k0, row_b1, fb, el);
            } else if (act_1) {
              const int row0 = n0 + 128 + (int)rank * 64;
              tma_load_2d_pair_p(dst, &maps.hi64, k0, row0, fb, el);
              if (NPROD == 3) tma_load_2d_pair_p(dst + BOFF_A_LO, &maps.lo64, k0, row0, fb, el);
            }
            if (act_2) {
              tma_load_2d_pair_p(dst + BOFF_B_HI, &maps.hi64, k0, row_b2, fb, el);
              if (NPROD == 3) tma_load_2d_pair_p(dst + BOFF_B_LO, &maps.lo64, k0, row_b2, fb, el);
            }
            st = (st + 1) & (ST - 1);
            ph ^= (st == 0);
          }
        }
      }
      // the last ST uses: publish them once their MMAs have retired (the mailbox warp leaves when the counters have
      // reached their final values)
      for (int k = 0; k < ST && it > 0; ++k) {
        mbar_wait_fast(BAR(BAR_EMPTY_ST + st), ph ^ 1, p.err, 1);
        st = (st + 1) & (ST - 1);
        ph ^= (st == 0);
      }
      if (el && it > 0) st_release_cta(cons_cnt, it);
    }
  } else if (warp == 3) {
    // ================================ TMA: training block halves + aux, once per (tile, own chunk) ================
    {
      const bool el = elect_one_sync();
      uint32_t it = 0;
      for (long long pt = group; pt < n_ptiles; pt += n_groups)
        for (int lc = 0; lc < nmy; ++lc, ++it) {
          const int kc = my_chunk(lc);
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
      for (long long pt = group; pt < n_ptiles; pt += n_groups, ++itile) {
        mbar_wait(BAR(BAR_FULL_AX), itile & 1, p.err, 9);
        for (int lc = 0; lc < nmy; ++lc, ++i) {
          const uint32_t x = i % X_STAGES, ph = (i / X_STAGES) & 1;
          mbar_wait(BAR(BAR_FULL_X + x), ph, p.err, 8);
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
    if (leader) {
      const uint32_t el = elect_one_sync() ? 1u : 0u;
      const uint32_t idesc_256 = umma_idesc_f16(2 * BM, NBA);
      const uint32_t idesc_128 = umma_idesc_f16(2 * BM, NBB);
      const uint64_t da0 = umma_desc_sw128(sbase + OFF_A), db0 = umma_desc_sw128(sbase + OFF_B);
      constexpr uint64_t A_STEP = A_STRIDE >> 4, B_STEP = B_STRIDE >> 4;
      uint32_t n_my = 0;
      for (long long pt = group; pt < n_ptiles; pt += n_groups) ++n_my;
      uint32_t left = n_my * (uint32_t)me.upt;
      uint32_t ist = 0, st = 0, ph = 0, ic = 0;
      if (left) mbar_wait(BAR(BAR_FULL + 0), 0, p.err, 3);
      for (long long pt = group; pt < n_ptiles; pt += n_groups) {
        for (int s = 0; s < n_super; ++s) {
          if (!((me.mine >> s) & 1u)) continue;
          const int kext = min(ld, WC * (s + 1));
          const int n0 = WC * s;
          const bool has2 = n0 + NBA < ld;
          const uint32_t pe = (ist & 1) ^ 1;
          const int n_bulk = n0 / KC;
          // The issuer is paced by the LENGTH of its own instruction stream whenever the producers share its scheduler
          // (developer timeline, profiles/r02/trace_gen6_*: 1190 clk per chunk use in a side's last super-tile against
          // 815 clk while the producers idle; ~100 SASS instructions per chunk use, 39 of them loop-header index
          // arithmetic and 20 R2UR moves of descriptors that lived in vector registers).  So: the order of the replayed
          // chunks is the TMA warp's business only (every chunk under the diagonal blocks takes the same MMAs), and the
          // bulk of a super-tile runs ST chunk uses per iteration with the stage a COMPILE-TIME constant -- descriptors
          // are then uniform-register adds of an immediate.
          const int nuse = kext / KC;
          // ---- general chunk use: first of the super-tile (overwrites the accumulators), diagonal chunks
          auto general = [&](int ci) {
            const int k0 = ci < n_bulk ? 0 : ci * KC;     // any chunk under the diagonal blocks behaves like k0 = 0
            const bool first = ci == 0;
            const uint32_t nst = (st + 1) & (ST - 1), nph = ph ^ (nst == 0 ? 1u : 0u);
            const bool tr = TRACE && p.trace && (int)blockIdx.x == sh.trace_cta && ic < (uint32_t)TRACE5_CHUNKS && lane == 0;
            if (tr) p.trace[ic * 8 + 0] = clock64();
            const uint64_t da_hi = da0 + A_STEP * st, da_lo = da_hi + (A_HALF_BYTES >> 4);
            const uint64_t db_hi = db0 + B_STEP * st, db_lo = db_hi + (BOFF_A_LO >> 4);
            const uint64_t db2_hi = db_hi + (BOFF_B_HI >> 4), db2_lo = db_hi + (BOFF_B_LO >> 4);
            --left;
            if (first) {
              mbar_wait(BAR(BAR_ACC_EMPTY + 0), pe, p.err, 2);
              mbar_wait(BAR(BAR_ACC_EMPTY + 1), pe, p.err, 2);
            }
            tc_fence_after();
            if (tr) p.trace[ic * 8 + 1] = clock64();
            if (k0 < n0 + NBA) {
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
              if (k0 + KC == min(n0 + 128, kext)) umma_commit_pair_p(BAR(BAR_ACC_FULL + 0), el);
              if (k0 + KC == min(n0 + 256, kext)) umma_commit_pair_p(BAR(BAR_ACC_FULL + 1), el);
            }
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
            umma_commit_pair_p(BAR(BAR_EMPTY_ST + st), el);
            if (tr) p.trace[ic * 8 + 2] = clock64();
            ++ic;
            st = nst;
            ph = nph;
          };
          // ---- one replayed chunk under the diagonal blocks; stage `stg` (compile-time in the unrolled loop below)
          auto bulk = [&](const uint32_t stg, const uint32_t wait_parity) {
            const bool tr = TRACE && p.trace && (int)blockIdx.x == sh.trace_cta && ic < (uint32_t)TRACE5_CHUNKS && lane == 0;
            if (tr) p.trace[ic * 8 + 0] = clock64();
            const uint64_t da_hi = da0 + A_STEP * stg, da_lo = da_hi + (A_HALF_BYTES >> 4);
            const uint64_t db_hi = db0 + B_STEP * stg, db_lo = db_hi + (BOFF_A_LO >> 4);
            const uint64_t db2_hi = db_hi + (BOFF_B_HI >> 4), db2_lo = db_hi + (BOFF_B_LO >> 4);
            --left;
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
            if (left) mbar_wait_fast(BAR(BAR_FULL + ((stg + 1) & (ST - 1))), wait_parity, p.err, 3);
            if (el) {
              if (has2) {
#pragma unroll
                for (int ks = 0; ks < KC / 16; ++ks) {
                  const uint64_t o = (uint64_t)(ks * 2);
                  umma_f16_pair(tmem_base + (uint32_t)NBA, da_hi + o, db2_hi + o, idesc_128, 1);
                  if (NPROD == 3) {
                    umma_f16_pair(tmem_base + (uint32_t)NBA, da_hi + o, db2_lo + o, idesc_128, 1);
                    umma_f16_pair(tmem_base + (uint32_t)NBA, da_lo + o, db2_hi + o, idesc_128, 1);
                  }
                }
              }
              umma_commit_pair(BAR(BAR_EMPTY_ST + stg));
            }
            if (tr) p.trace[ic * 8 + 2] = clock64();
            ++ic;
          };
          auto bulk_rt = [&]() {   // runtime stage: the few uses that bring the ring back to stage 0 / the remainder
            const uint32_t nst = (st + 1) & (ST - 1), nph = ph ^ (nst == 0 ? 1u : 0u);
            bulk(st, nph);
            st = nst;
            ph = nph;
          };
          int ci = 0;
          general(ci++);
          while (ci < n_bulk && st != 0) {
            bulk_rt();
            ++ci;
          }
          while (ci + ST <= n_bulk) {   // st == 0 here
#pragma unroll
            for (int stg = 0; stg < ST; ++stg) bulk((uint32_t)stg, stg == ST - 1 ? ph ^ 1u : ph);
            ph ^= 1u;
            ci += ST;
          }
          while (ci < n_bulk) {
            bulk_rt();
            ++ci;
          }
          for (; ci < nuse; ++ci) general(ci);
          if (!has2) {
            mbar_wait(BAR(BAR_ACC_EMPTY + 2), pe, p.err, 2);
          }
          umma_commit_pair_p(BAR(BAR_ACC_FULL + 2), el);
          ++ist;
        }
      }
    }
  } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + 4) {
    // ================================ epilogue (own 128 candidates, this side's super-tiles) ==========================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    uint32_t ist = 0;
    for (long long pt = group; pt < n_ptiles; pt += n_groups) {
      const long long tile = 2 * pt + rank;
      double ss = 0.0;
      for (int s = 0; s < n_super; ++s) {
        if (!((me.mine >> s) & 1u)) continue;
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
        ++ist;
      }
      if (tile < n_tiles) atomicAdd(p.sumsq + tile * BM + row, ss);   // two addends (one per side) on a zeroed array
    }
  } else if (warp >= PW0 && warp < PW0 + NPW) {
    // ================================ producers (own 128 candidates, this side's chunks) ================================
    const int pw = warp - PW0;
    const int grp = pw >> 3;
    const int quad = pw & 3;
    const int ch = (pw >> 2) & 1;
    const int kq = pw >> 2;
    const int m = quad * 32 + lane;
    const uint32_t row_off = (uint32_t)((m >> 3) * 1024 + (m & 7) * 128);
    float2* exch = p.exch + (size_t)blockIdx.x * 3 * BM;
    const float CG = -2.0f / (float)(1 << (2 * X_SCALE_LOG2));
    const bool elected = (pw & 7) == 0 && lane == 0;
    uint8_t* const scr = (uint8_t*)ra.scratch;
    uint32_t j = 0, tl = 0, done = 0;
    for (long long pt = group; pt < n_ptiles; pt += n_groups, ++tl) {
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
      for (int lc = 0; lc < nmy; ++lc, ++j) {
        if ((lc & 1) != grp) continue;  // the other group's chunk (nmy is even: chunk parity = j parity)
        const int kc = my_chunk(lc);
        const uint32_t ax = j % AUX_STAGES;
        const bool tr = TRACE && p.trace && (int)blockIdx.x == sh.trace_cta && j < (uint32_t)TRACE5_CHUNKS && pw == 0 && lane == 0;
        if (tr) p.trace[j * 8 + 3] = clock64();
        mbar_wait(BAR(BAR_FULL_AUX + ax), (j / AUX_STAGES) & 1, p.err, 11);
        mbar_wait(BAR(BAR_FULL_G + grp), (j / 2) & 1, p.err, 10);
        tc_fence_after();
        if (tr) p.trace[j * 8 + 4] = clock64();
        float2 ys2 = make_float2(0.f, 0.f), fs2 = make_float2(0.f, 0.f);
        uint32_t gr0[16], gr1[16];
        fk2::tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(G_COL0 + KC * grp + 32 * ch), gr0);
        fk2::tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(G_COL0 + KC * grp + 32 * ch + 16), gr1);
        tmem_ld_wait();
        tc_fence_before();
        // the scratch slot of this chunk is free once BOTH sides have retired the previous tile's last replay of it
        // (a double-buffered chunk reuses the slot of tile tl - 2)
        const uint32_t tw = kc < kdb ? tl - 1u : tl;
        if (elected && tl > (kc < kdb ? 1u : 0u)) {
          spin_until_ge(cons_cnt, dead_after(me, tw, kc), p.err, 15);
          spin_until_ge(mir + 8u, dead_after(ot, tw, kc), p.err, 17);
        }
        if (grp == 0) asm volatile("bar.sync 1, %0;" ::"n"(32 * NPW / 2) : "memory");
        else asm volatile("bar.sync 2, %0;" ::"n"(32 * NPW / 2) : "memory");
        if (elected) mbar_arrive_leader(BAR(BAR_EMPTY_G + grp), leader);
        if (tr) p.trace[j * 8 + 5] = clock64();
        uint8_t* g_hi = scr + ((size_t)scr_row(kc, 0, tl) + (size_t)m) * 128;
        uint8_t* g_lo = scr + ((size_t)scr_row(kc, PLANES - 1, tl) + (size_t)m) * 128;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int col0 = 32 * ch + 16 * h;
          const uint32_t(&gr)[16] = h == 0 ? gr0 : gr1;
          const float* aux = (const float*)(smem + OFF_AUX + ax * AUX_BYTES) + col0;
          float bj[16];
#pragma unroll
          for (int i = 0; i < 16; i += 4) *(float4*)&bj[i] = *(const float4*)(aux + i);
          // two values per instruction (FFMA2 / FADD2 / FMUL2): the producers share their schedulers with the issuing
          // warps, so their instruction count is what the MMA issuer feels (~15 -> ~8 instructions per value)
          float kv[16];
#ifndef B200BO_FK6_SCALAR
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
#else  /* scalar form (A/B builds) */
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float acc = fmaxf(fmaf(CG, __uint_as_float(gr[i]), am + bj[i]), 0.f);
            kv[i] = corr_from_acc<CORR>(acc);
          }
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 gj = *(const float4*)(aux + KC + i);
            const float4 fj = *(const float4*)(aux + 2 * KC + i);
            ys2.x = fmaf(kv[i], gj.x, fmaf(kv[i + 1], gj.y, fmaf(kv[i + 2], gj.z, fmaf(kv[i + 3], gj.w, ys2.x))));
            fs2.x = fmaf(kv[i], fj.x, fmaf(kv[i + 1], fj.y, fmaf(kv[i + 2], fj.z, fmaf(kv[i + 3], fj.w, fs2.x))));
          }
#endif
          // 16 fp16 values = one 32-byte sector per plane, written with ONE 256-bit store (two 16-byte stores are two
          // partial-sector writes for L2: ncu showed a DRAM fill read for every scratch line)
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float v0 = kv[2 * i], v1 = kv[2 * i + 1];
            const __half2 hh = __floats2half2_rn(v0, v1);
            hi[i] = *(const uint32_t*)&hh;
            if (NPROD == 3) {
              const float2 hf = __half22float2(hh);
              const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
              lo[i] = *(const uint32_t*)&l;
            }
          }
          const uint32_t goff = (uint32_t)((col0 >> 3) * 16);  // plain row-major: the TMA load applies the swizzle
          st_global_256(g_hi + goff, hi);
          if (NPROD == 3) st_global_256(g_lo + goff, lo);
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
      // combine the four (group, column-half) partial dot products of a row; the two sides add into the output
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
        atomicAdd(p.yhat + tile * BM + m, side == 0 ? p.beta + y : y);
        atomicAdd(p.dotf + tile * BM + m, f);
      }
      asm volatile("bar.sync 3, %0;" ::"n"(32 * NPW) : "memory");
    }
  }
  else if (warp == PW0 + NPW) {
    // ================================ mailbox: this side's counters -> global, the partner's -> shared memory ==========
    // acquire.cta on a local counter orders the producers' scratch stores (st.global + fence.proxy.async + their CTA
    // barrier + st.release.cta) before the release.gpu below; on the other side acquire.gpu + release.cta hand the same
    // guarantee to the TMA warp, which adds its own fence.proxy.async before the cp.async.bulk.tensor.
    if (lane == 0) {
      uint32_t n_my = 0;
      for (long long pt = group; pt < n_ptiles; pt += n_groups) ++n_my;
      const uint32_t fin_prod = n_my * (uint32_t)(nch / 4), fin_cons = n_my * (uint32_t)me.upt;
      uint32_t pub0 = 0, pub1 = 0, pubc = 0;
      const long long t0 = clock64();
      while (n_my) {
        const uint32_t p0 = ld_acquire_cta(prod_cnt), p1 = ld_acquire_cta(prod_cnt + 4u), pc = ld_acquire_cta(cons_cnt);
        if (p0 != pub0) st_release_gpu(fl_me + 0, pub0 = p0);
        if (p1 != pub1) st_release_gpu(fl_me + 8, pub1 = p1);
        if (pc != pubc) st_release_gpu(fl_me + 16, pubc = pc);
        const uint32_t q0 = ld_acquire_gpu(fl_ot + 0), q1 = ld_acquire_gpu(fl_ot + 8), qc = ld_acquire_gpu(fl_ot + 16);
        st_release_cta(mir + 0u, q0);
        st_release_cta(mir + 4u, q1);
        st_release_cta(mir + 8u, qc);
        if (pub0 == fin_prod && pub1 == fin_prod && pubc == fin_cons) break;
        __nanosleep(64);
        if (clock64() - t0 > 16 * WAIT_TIMEOUT_CYCLES) {
          atomicExch(p.err, 18);
          __threadfence_system();
          __trap();
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

}  // namespace fk6
}  // namespace b2
