// fast4_kernels.cuh -- CTA-pair fused kernel with r REPLAY (sm_100a).  Fourth generation of the M-candidate path.
//
// Generation 3 (fast3_kernels.cuh) recomputes the cross-correlation chunk r[:, k0:k0+64] of a candidate tile for every
// 384-column accumulator super-tile that needs it (TMEM holds 384 fp32 accumulators per candidate, L^-1 is triangular:
// a chunk is needed by (N - k0) / 384 super-tiles, x5.8 on average at N = 4096) and the MUFU / issue slots of that
// recomputation, not the tensor pipe, bound the kernel (profiles/r01/pair_p1_ncu_summary.txt: XU 59 %, tensor 44 %).
// Here a chunk is computed ONCE per tile, on its first use; the producers store the fp16 A operand to shared memory
// (for the MMA of this super-tile) AND to a per-CTA scratch in global memory (L2-resident: 16 KB per chunk and plane),
// and every later super-tile gets the chunk back by TMA (SWIZZLE_128B box of the row-major scratch == the layout the
// producers write by hand) straight into the A stage.  Only the first `n_store` chunks of a tile (the most reused
// ones) are stored -- the host sizes the scratch to stay inside L2 next to the fp16 L^-1 -- the rest is recomputed
// as before.  Everything else (cta_group::2 M=256 MMAs, B halves by TMA, Gram product on the tensor cores, separate
// Gram issuer, two alternating producer groups, epilogue) is generation 3.
//
// Ordering of the scratch: a replay of chunk c is issued by the TMA thread only after it has seen the EMPTY barrier of
// a stage that is at least 6 chunks younger than the chunk's first use (a super-tile opens 6 new chunks, the ring has
// <= 4 stages), i.e. after the MMA that consumed the first use retired -- the producers' global stores, their
// fence.proxy.async and their CTA barrier all precede the FULL arrival that MMA waited for.  The next tile overwrites
// a scratch chunk only after the MMA fed by its last replay retired (same argument).
// n_store must be even: chunk parity then equals computed-chunk parity, so group g keeps stages {g, g+2} and Gram block g.
#pragma once
#include "fast3_kernels.cuh"

namespace b2 {
namespace fk4 {

using namespace fk3;

struct ReplayMaps {
  fk3::PairMaps pm;
  CUtensorMap scr;  // r scratch as a 2-D fp16 tensor (rows x 64), box 64 x 128 rows, SWIZZLE_128B
};

struct ReplayArgs {
  __half* scratch;  // (gridDim.x, n_store, planes, 128 rows, 64) fp16, row-major; planes = 1 (one product) or 2 (hi, lo)
  int n_store;      // chunks of a tile kept in the scratch (even; 0 = recompute everything, i.e. generation 3)
};


template <int CORR, int NPROD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT2, 1)
predict_fused_replay_kernel(const __grid_constant__ ReplayMaps rmaps, const Fused2Args p, const ReplayArgs ra) {
  const PairMaps& maps = rmaps.pm;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = (uint64_t*)(smem + OFF_BAR);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  uint32_t* tmem_slot = (uint32_t*)(bars + 36);
  constexpr int ST = NPROD == 1 ? 4 : 2;                                  // stages (A ring and B ring alike)
  constexpr int A_STRIDE = NPROD == 1 ? A_HALF_BYTES : A_STAGE_BYTES;    // 16 / 32 KB
  constexpr int B_STRIDE = NPROD * 0 + (NPROD == 1 ? BA_PLANE + BB_PLANE : 2 * (BA_PLANE + BB_PLANE));  // 24 / 48 KB
  // B stage layout: [256-block hi][256-block lo (three products)][128-block hi][128-block lo]
  constexpr int BOFF_A_LO = BA_PLANE;
  constexpr int BOFF_B_HI = NPROD == 1 ? BA_PLANE : 2 * BA_PLANE;
  constexpr int BOFF_B_LO = BOFF_B_HI + BB_PLANE;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int ld = p.ld;
  const int n_super = (ld + WC - 1) / WC;
  const long long n_tiles = (p.M + BM - 1) / BM;
  const long long n_ptiles = (n_tiles + 1) / 2;        // the pair works on tiles 2 pt and 2 pt + 1
  const long long pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  auto LBAR = [&](int i) { return map_to_cta(BAR(i), 0); };  // the leader's copy of barrier i
  constexpr int PLANES = NPROD == 1 ? 1 : 2;
  const int n_store = ra.n_store;
  // chunk k0 of super-tile s comes back from the scratch (it was computed and stored by an earlier super-tile of this tile)
  auto replayed = [&](int s, int k0) { return k0 < WC * s && (k0 / KC) < n_store; };
  // scratch row of (this CTA, chunk, plane): 128 rows of 64 fp16 each
  auto scr_row = [&](int kc, int plane) { return (((int)blockIdx.x * n_store + kc) * PLANES + plane) * BM; };

  if (threadIdx.x == 0) {
    if ((sbase & 1023u) || (n_store & 1)) {
      atomicExch(p.err, 99);
      __trap();
    }
    tma_prefetch_desc(&rmaps.scr);
    tma_prefetch_desc(&maps.hi128);
    tma_prefetch_desc(&maps.hi64);
    tma_prefetch_desc(&maps.xh32);
    for (int i = 0; i < ST; ++i) {
      mbar_init(BAR(BAR_FULL_A + i), 2);
      mbar_init(BAR(BAR_FULL_B + i), 2);
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
    for (int i = 0; i < AUX_STAGES; ++i) {
      mbar_init(BAR(BAR_FULL_AUX + i), 1);
      mbar_init(BAR(BAR_EMPTY_AUX + i), 1);
    }
    mbar_init(BAR(BAR_FULL_AX), 2);
    mbar_init(BAR(BAR_ACC_FULL), 1);
    mbar_init(BAR(BAR_ACC_EMPTY), 8);
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();  // both CTAs' barriers are initialised and TMEM is allocated before anything crosses over
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA: this CTA's halves of the L^-1 blocks of every chunk ================
    if (lane == 0) {
      uint32_t it = 0;
      for (long long pt = pair; pt < n_ptiles; pt += n_pairs)
        for (int s = 0; s < n_super; ++s) {
          const int kext = min(ld, WC * (s + 1));
          const int n0 = WC * s;
          for (int k0 = 0; k0 < kext; k0 += KC, ++it) {
            const uint32_t st = it % ST, ph = (it / ST) & 1;
            const bool act_a = k0 < n0 + NBA;                        // 256-column block below / on the diagonal
            const bool act_b = n0 + NBA < ld;                        // 128-column block exists (k0 < kext covers the rest)
            mbar_wait(BAR(BAR_EMPTY_ST + st), ph ^ 1, p.err, 1);
            const uint32_t fb = LBAR(BAR_FULL_B + st);
            mbar_arrive_expect_tx_leader(BAR(BAR_FULL_B + st), leader, (uint32_t)((act_a ? BA_PLANE : 0) + (act_b ? BB_PLANE : 0)) * (NPROD == 3 ? 2u : 1u));
            const uint32_t dst = sbase + OFF_B + st * B_STRIDE;
            if (act_a) {
              const int row0 = n0 + (int)rank * (NBA / 2);
              tma_load_2d_pair(dst, &maps.hi128, k0, row0, fb);
              if (NPROD == 3) tma_load_2d_pair(dst + BOFF_A_LO, &maps.lo128, k0, row0, fb);
            }
            if (act_b) {
              const int row0 = n0 + NBA + (int)rank * (NBB / 2);
              tma_load_2d_pair(dst + BOFF_B_HI, &maps.hi64, k0, row0, fb);
              if (NPROD == 3) tma_load_2d_pair(dst + BOFF_B_LO, &maps.lo64, k0, row0, fb);
            }
            if (replayed(s, k0)) {
              // A operand of this chunk back from the scratch; this thread stands in for the producers' arrival
              const uint32_t fa = LBAR(BAR_FULL_A + st);
              mbar_arrive_expect_tx_leader(BAR(BAR_FULL_A + st), leader, (uint32_t)(PLANES * A_HALF_BYTES));
              const uint32_t da = sbase + OFF_A + st * A_STRIDE;
              tma_load_2d_pair(da, &rmaps.scr, 0, scr_row(k0 / KC, 0), fa);
              if (NPROD == 3) tma_load_2d_pair(da + A_HALF_BYTES, &rmaps.scr, 0, scr_row(k0 / KC, 1), fa);
            }
          }
        }
    }
  } else if (warp == 3) {
    // ================================ TMA: training block halves + aux ================================
    if (lane == 0) {
      uint32_t it = 0;
      for (long long pt = pair; pt < n_ptiles; pt += n_pairs)
        for (int s = 0; s < n_super; ++s) {
          const int kext = min(ld, WC * (s + 1));
          for (int k0 = 0; k0 < kext; k0 += KC) {
            if (replayed(s, k0)) continue;  // no Gram product, no aux block: `it` counts computed chunks
            const uint32_t x = it % X_STAGES, ph = (it / X_STAGES) & 1;
            const uint32_t ax = it % AUX_STAGES, pax = (it / AUX_STAGES) & 1;
            mbar_wait(BAR(BAR_EMPTY_AUX + ax), pax ^ 1, p.err, 12);
            mbar_arrive_expect_tx(BAR(BAR_FULL_AUX + ax), AUX_BYTES);
            fk2::bulk_load_1d(sbase + OFF_AUX + ax * AUX_BYTES, p.aux + (size_t)(k0 / KC) * 3 * KC, AUX_BYTES, BAR(BAR_FULL_AUX + ax));
            mbar_wait(BAR(BAR_EMPTY_X + x), ph ^ 1, p.err, 7);
            const uint32_t fx = LBAR(BAR_FULL_X + x);
            mbar_arrive_expect_tx_leader(BAR(BAR_FULL_X + x), leader, 2 * XH_PLANE);
            const uint32_t dst = sbase + OFF_X + x * X_STAGE_BYTES;
            tma_load_2d_pair(dst, &maps.xh32, 0, k0 + 32 * (int)rank, fx);
            tma_load_2d_pair(dst + XH_PLANE, &maps.xl32, 0, k0 + 32 * (int)rank, fx);
            ++it;
          }
        }
    }
  } else if (warp == 2) {
    // ================================ Gram-MMA issuer (leader CTA only) ================================
    if (lane == 0 && leader) {
      const uint32_t idesc_gram = umma_idesc_f16(2 * BM, KC);
      const uint64_t dax_hi = umma_desc_sw128(sbase + OFF_AX);
      const uint64_t dax_lo = umma_desc_sw128(sbase + OFF_AX + AX_PLANE);
      int chunks_per_tile = 0;
      for (int s = 0; s < n_super; ++s)
        for (int k0 = 0; k0 < min(ld, WC * (s + 1)); k0 += KC) chunks_per_tile += replayed(s, k0) ? 0 : 1;
      uint32_t i = 0, itile = 0;
      for (long long pt = pair; pt < n_ptiles; pt += n_pairs, ++itile) {
        // candidate operands of this tile are in place (every Gram MMA of the previous tile has retired: the producers
        // consumed its last block before they wrote the new operand)
        mbar_wait(BAR(BAR_FULL_AX), itile & 1, p.err, 9);
        for (int lc = 0; lc < chunks_per_tile; ++lc, ++i) {
          const uint32_t x = i % X_STAGES, ph = (i / X_STAGES) & 1;
          mbar_wait(BAR(BAR_FULL_X + x), ph, p.err, 8);
          // Gram block i & 1 is free once the producer groups of both CTAs have read out chunk i - 2
          if (i >= 2) mbar_wait(BAR(BAR_EMPTY_G + (i & 1)), ((i - 2) / 2) & 1, p.err, 13);
          tc_fence_after();
          const uint64_t dx_hi = umma_desc_sw128(sbase + OFF_X + x * X_STAGE_BYTES);
          const uint64_t dx_lo = umma_desc_sw128(sbase + OFF_X + x * X_STAGE_BYTES + XH_PLANE);
          const uint32_t tg = tmem_base + (uint32_t)(G_COL0 + KC * (i & 1));
          for (int ks = 0; ks < p.dk_steps; ++ks) {
            const uint64_t o = (uint64_t)(ks * 2);
            umma_f16_pair(tg, dax_hi + o, dx_hi + o, idesc_gram, ks != 0);
            umma_f16_pair(tg, dax_hi + o, dx_lo + o, idesc_gram, 1);
            umma_f16_pair(tg, dax_lo + o, dx_hi + o, idesc_gram, 1);
          }
          umma_commit_pair(BAR(BAR_FULL_G + (i & 1)));
          umma_commit_pair(BAR(BAR_EMPTY_X + x));
        }
      }
    }
  } else if (warp == 1) {
    // ================================ main-MMA issuer (leader CTA only) ================================
    if (lane == 0 && leader) {
      const uint32_t idesc_a = umma_idesc_f16(2 * BM, NBA);
      const uint32_t idesc_b = umma_idesc_f16(2 * BM, NBB);
      uint32_t ic = 0, ist = 0;
      for (long long pt = pair; pt < n_ptiles; pt += n_pairs) {
        for (int s = 0; s < n_super; ++s) {
          const int kext = min(ld, WC * (s + 1));
          const int n0 = WC * s;
          mbar_wait(BAR(BAR_ACC_EMPTY), (ist & 1) ^ 1, p.err, 2);
          for (int k0 = 0; k0 < kext; k0 += KC) {
            const uint32_t st = ic % ST, ph = (ic / ST) & 1;
            const bool tr = p.trace && blockIdx.x == 0 && ic < TRACE_CHUNKS;
            if (tr) p.trace[ic * 8 + 0] = clock64();
            mbar_wait(BAR(BAR_FULL_A + st), ph, p.err, 3);
            mbar_wait(BAR(BAR_FULL_B + st), ph, p.err, 4);
            tc_fence_after();
            if (tr) p.trace[ic * 8 + 1] = clock64();
            const uint64_t da_hi = umma_desc_sw128(sbase + OFF_A + st * A_STRIDE);
            const uint64_t da_lo = umma_desc_sw128(sbase + OFF_A + st * A_STRIDE + A_HALF_BYTES);
            const uint32_t bb = sbase + OFF_B + st * B_STRIDE;
            if (k0 < n0 + NBA) {
              const uint64_t db_hi = umma_desc_sw128(bb), db_lo = umma_desc_sw128(bb + BOFF_A_LO);
#pragma unroll
              for (int ks = 0; ks < KC / 16; ++ks) {
                const uint64_t o = (uint64_t)(ks * 2);
                umma_f16_pair(tmem_base, da_hi + o, db_hi + o, idesc_a, (k0 | ks) != 0);
                if (NPROD == 3) {
                  umma_f16_pair(tmem_base, da_hi + o, db_lo + o, idesc_a, 1);
                  umma_f16_pair(tmem_base, da_lo + o, db_hi + o, idesc_a, 1);
                }
              }
            }
            if (n0 + NBA < ld) {
              const uint64_t db_hi = umma_desc_sw128(bb + BOFF_B_HI), db_lo = umma_desc_sw128(bb + BOFF_B_LO);
              const uint32_t td = tmem_base + (uint32_t)NBA;
#pragma unroll
              for (int ks = 0; ks < KC / 16; ++ks) {
                const uint64_t o = (uint64_t)(ks * 2);
                umma_f16_pair(td, da_hi + o, db_hi + o, idesc_b, (k0 | ks) != 0);
                if (NPROD == 3) {
                  umma_f16_pair(td, da_hi + o, db_lo + o, idesc_b, 1);
                  umma_f16_pair(td, da_lo + o, db_hi + o, idesc_b, 1);
                }
              }
            }
            umma_commit_pair(BAR(BAR_EMPTY_ST + st));  // one commit frees the A and the B stage in both CTAs
            if (tr) p.trace[ic * 8 + 2] = clock64();
            ++ic;
          }
          umma_commit_pair(BAR(BAR_ACC_FULL));
          ++ist;
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
          if (p.dbg_w && tile < n_tiles) {
            float* o = p.dbg_w + (size_t)(tile * BM + row) * ld + WC * s + c0;
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(r[j]) * p.out_scale;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(BAR(BAR_ACC_EMPTY), leader);
        ++ist;
      }
      if (tile < n_tiles) p.sumsq[tile * BM + row] = ss;
    }
  } else if (warp >= PW0) {
    // ================================ producers (own 128 candidates) ================================
    // Two groups of 8 warps take alternate chunks (group = chunk parity = Gram block), so one group's barrier waits,
    // TMEM loads and stores overlap the other group's MUFU-bound compute.  Within a group: 2 warps per TMEM lane
    // quadrant, each thread 32 of the chunk's 64 columns in two passes of 16.
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
    uint32_t ic = 0, jc = 0;  // chunks / computed chunks so far (n_store is even: ic and jc have the same parity)
    uint8_t* const scr = (uint8_t*)ra.scratch;
    for (long long pt = pair; pt < n_ptiles; pt += n_pairs) {
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
      for (int s = 0; s < n_super; ++s) {
        const int kext = min(ld, WC * (s + 1));
        for (int k0 = 0; k0 < kext; k0 += KC, ++ic) {
          const bool rep = replayed(s, k0);
          const uint32_t j = jc;
          if (!rep) ++jc;
          if ((int)(ic & 1) != grp) continue;  // the other group's chunk
          const uint32_t a = ic % ST, pha = (ic / ST) & 1;
          if (rep) {
            // the TMA thread fills this stage; watch its EMPTY phase all the same, so that this group sees every
            // phase of its own stages in order (a parity wait that skipped two phases would pass too early)
            mbar_wait(BAR(BAR_EMPTY_ST + a), pha ^ 1, p.err, 6);
            continue;
          }
          const bool first = k0 >= WC * s;                                   // first use in this tile: take the dot products
          const int kc = k0 / KC;
          const bool keep = kc < n_store && k0 < WC * (n_super - 1);        // a later super-tile replays it
          const uint32_t ax = j % AUX_STAGES;
          const bool tr = p.trace && blockIdx.x == 0 && ic < TRACE_CHUNKS && pw == 0 && lane == 0;
          if (tr) p.trace[ic * 8 + 3] = clock64();
          mbar_wait(BAR(BAR_FULL_AUX + ax), (j / AUX_STAGES) & 1, p.err, 11);
          mbar_wait(BAR(BAR_FULL_G + grp), (j / 2) & 1, p.err, 10);
          tc_fence_after();
          if (tr) p.trace[ic * 8 + 4] = clock64();
          float ysum = 0.f, fsum = 0.f;
          // read the whole Gram row segment first and hand the TMEM block back: the Gram MMA of chunk ic + 2 then
          // runs while this group is still computing
          uint32_t gr0[16], gr1[16];
          fk2::tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(G_COL0 + KC * grp + 32 * ch), gr0);
          fk2::tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(G_COL0 + KC * grp + 32 * ch + 16), gr1);
          tmem_ld_wait();
          tc_fence_before();
          if (grp == 0) asm volatile("bar.sync 1, %0;" ::"n"(32 * NPW / 2) : "memory");
          else asm volatile("bar.sync 2, %0;" ::"n"(32 * NPW / 2) : "memory");
          if (elected) mbar_arrive_leader(BAR(BAR_EMPTY_G + grp), leader);
          mbar_wait(BAR(BAR_EMPTY_ST + a), pha ^ 1, p.err, 6);
          if (tr) p.trace[ic * 8 + 5] = clock64();
          uint8_t* a_hi = smem + OFF_A + a * A_STRIDE;
          uint8_t* a_lo = a_hi + A_HALF_BYTES;
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
            float kv[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float acc = fmaxf(fmaf(CG, __uint_as_float(gr[i]), am + bj[i]), 0.f);
              kv[i] = corr_from_acc<CORR>(acc);
            }
            if (first) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const float4 gj = *(const float4*)(aux + KC + i);
                const float4 fj = *(const float4*)(aux + 2 * KC + i);
                ysum = fmaf(kv[i], gj.x, fmaf(kv[i + 1], gj.y, fmaf(kv[i + 2], gj.z, fmaf(kv[i + 3], gj.w, ysum))));
                fsum = fmaf(kv[i], fj.x, fmaf(kv[i + 1], fj.y, fmaf(kv[i + 2], fj.z, fmaf(kv[i + 3], fj.w, fsum))));
              }
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
              const uint32_t off = row_off + (uint32_t)((((col0 >> 3) + c) ^ (m & 7)) * 16);
              *(uint4*)(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              if (NPROD == 3) *(uint4*)(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              if (keep) {  // plain row-major copy for the replays (the TMA load re-applies the swizzle)
                const uint32_t goff = (uint32_t)(((col0 >> 3) + c) * 16);
                *(uint4*)(g_hi + goff) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                if (NPROD == 3) *(uint4*)(g_lo + goff) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              }
            }
          }
          if (first) {
            ysum_d += (double)ysum;
            fsum_d += (double)fsum;
          }
          tc_fence_before();
          if (keep) asm volatile("fence.proxy.async.global;" ::: "memory");  // scratch stores -> visible to the TMA replays
          fence_proxy_async();
          if (tr) p.trace[ic * 8 + 6] = clock64();
          if (grp == 0) asm volatile("bar.sync 1, %0;" ::"n"(32 * NPW / 2) : "memory");
          else asm volatile("bar.sync 2, %0;" ::"n"(32 * NPW / 2) : "memory");
          if (tr) p.trace[ic * 8 + 7] = clock64();
          if (elected) {
            mbar_arrive_leader(BAR(BAR_FULL_A + a), leader);  // both CTAs' stages must be written before the pair MMA
            mbar_arrive(BAR(BAR_EMPTY_AUX + ax));
          }
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

}  // namespace fk4
}  // namespace b2
