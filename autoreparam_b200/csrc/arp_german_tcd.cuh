// tcgen05 engine, DUAL-TILE streaming variant (F <= 32).  Same algorithm and arithmetic as
// arp_german_tcs.cuh, but the CTA's 128 chains are two independent 64-chain tiles, each with its own
// MMA issuer, producer, chunk ring and mbarrier pipeline, running out of phase on the same SM:
//
//  * every MMA is an M = 64 instruction; its 64 rows occupy TMEM lanes 0-15 of each 32-lane quarter
//    (tile 0) or lanes 16-31 (tile 1), so both tiles share the same TMEM columns;
//  * a worker warp serves 16 chains of one tile with two workers per chain: lanes 0-15 take one
//    32-observation column block of the chunk, lanes 16-31 the next one (tcgen05.ld/st shape .16x32bx2,
//    whose immediate is the column offset of the upper half-warp);
//  * each SM sub-partition (= TMEM lane quarter) therefore hosts two warps of tile 0 and two of tile 1.
//    While one tile sits in its serial phases (site forward / reverse, barrier and MMA round trips,
//    accept), the other tile's epilogue keeps the issue slots, the XU pipe and the tensor pipe busy --
//    the single-tile kernel leaves ~75 % of a leapfrog step to such latency-bound phases.
//
// Roles (640 threads): warps 0-15 workers, 16/17 MMA issuers (tile 0/1, one lane each),
// 18/19 producers (tile 0/1, one lane each).
#pragma once
#include "arp_german_tcs.cuh"

namespace arp {

#ifndef TCD_PROFILE
#define TCD_PROFILE 0   // 1: one thread per tile accumulates clock() per phase and printf()s it (timing study only)
#endif
#define TCD_THREADS (TC_WORKERS + 128)
#define TCD_TILE 64
#define TCD_TW (TC_NQ * TCD_TILE)     // worker threads per tile

struct Tcd {
  static constexpr int NF = 32;
  static constexpr uint32_t SF = 128, SG = 512;
  static constexpr uint32_t XCHUNK = (TC_CHUNK / 8) * SG;          // 8192
  static constexpr uint32_t STAGE = 2 * XCHUNK;                    // head | tail of the t-scaled rows
  static constexpr uint32_t AIMG = (TCD_TILE / 8) * SG;            // 4096 per part per tile
  static constexpr int FPW = 8, NLOC = 17;
  static constexpr uint32_t RING = 0;                               // [2 tiles][NSTAGE][STAGE]
  static constexpr uint32_t A1 = RING + 2 * TCS_NSTAGE * STAGE;     // [2 tiles][AIMG]
  static constexpr uint32_t A2 = A1 + 2 * AIMG;
  static constexpr uint32_t XCH = A2 + 2 * AIMG;                    // float[2][4][TC_NQ][128]
  static constexpr uint32_t XS = XCH + 2 * 4 * TC_NQ * TC_CHAINS * 4;
  static constexpr uint32_t PAR = XS + NLOC * TC_WORKERS * 4;
  static constexpr uint32_t BAR = PAR + 4 * (2 * NF + 4) * 4;       // [2 tiles][16] mbarriers
  static constexpr uint32_t TMEM_PTR = BAR + 2 * 16 * 8;
  static constexpr uint32_t BYTES = TMEM_PTR + 16;
  static constexpr uint32_t COL_H = 0, COL_G = 256, COL_R2 = 320;
  static constexpr uint32_t IDESC_G1 = (1u << 4) | ((uint32_t)(TC_CHUNK >> 3) << 17) | ((64u >> 4) << 24);
  static constexpr uint32_t IDESC_G2 = (1u << 4) | (1u << 16) | ((uint32_t)(NF >> 3) << 17) | ((64u >> 4) << 24);
  static_assert(BYTES <= 232448, "shared memory budget");
};

// half-warp TMEM accesses: lanes 0-15 of the warp get columns [c, c + n), lanes 16-31 columns [c + n, c + 2n)
// of TMEM lanes base .. base + 15
#define TCD_LD32(taddr, v)                                                                                 \
  asm volatile("tcgen05.ld.sync.aligned.16x32bx2.x32.b32 "                                                 \
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"                                    \
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32], 32;"               \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) \
               : "r"(taddr) : "memory")
#define TCD_LD8(taddr, v)                                                                                  \
  asm volatile("tcgen05.ld.sync.aligned.16x32bx2.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], 8;"                \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) \
               : "r"(taddr) : "memory")
#define TCD_ST16(taddr, IMM, v)                                                                            \
  asm volatile("tcgen05.st.sync.aligned.16x32bx2.x16.b32 [%0], " #IMM ", "                                 \
               "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"                                  \
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), \
                 "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory")

// named barriers: 1, 2 = tile_bar of tile 0 / 1; tile t: 3 + 3t = "A operand written", 4 + 3t + b = "residual buffer b
// stored" (workers bar.arrive, the tile's issuer warp bar.sync's: see arp_german_tcs.cuh)
#define TCD_NB_COUNT (TCD_TW + 32)
__device__ __forceinline__ void tcd_nb_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(TCD_NB_COUNT) : "memory"); }
__device__ __forceinline__ void tcd_nb_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(TCD_NB_COUNT) : "memory"); }
__device__ __forceinline__ void tile_bar(int tile) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + tile), "n"(TCD_TW) : "memory");
}

template <bool GAMMA>
__global__ void __launch_bounds__(TCD_THREADS, 1)
k_german_tcd_hmc(TcsParams tp, HmcWs ws, HmcArgs p) {
  using K = Tcd;
  constexpr int NF = K::NF, FPW = K::FPW, NLOC = K::NLOC;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = smem_u32(smem);
  // role -> tile: workers (warp >> 2) & 1, issuers 16/17, producers 18/19
  const int tile = warp < 16 ? ((warp >> 2) & 1) : (warp & 1);
  const uint32_t bar_a = sbase + K::BAR + tile * 128, bar_h0 = bar_a + 8, bar_r0 = bar_a + 24, bar_g = bar_a + 40;
  const uint32_t bar_xf = bar_a + 48, bar_xe = bar_xf + 8 * TCS_NSTAGE;
  const uint32_t ring = sbase + K::RING + tile * TCS_NSTAGE * K::STAGE;
  float* xch = reinterpret_cast<float*>(smem + K::XCH);
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + K::TMEM_PTR);
  {
    float* par = reinterpret_cast<float*>(smem + K::PAR);
    for (int i = tid; i < p.D; i += TCD_THREADS) {
      par[i] = p.a[i];
      par[(2 * NF + 4) + i] = p.b[i];
      par[2 * (2 * NF + 4) + i] = p.eps0[i];
    }
  }
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)),
                 "r"((uint32_t)TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int tl = 0; tl < 2; ++tl) {
      const uint32_t ba = sbase + K::BAR + tl * 128;
      mbar_init(ba, TCD_TW);
      mbar_init(ba + 8, 1); mbar_init(ba + 16, 1);
      mbar_init(ba + 24, TCD_TW); mbar_init(ba + 32, TCD_TW);
      mbar_init(ba + 40, 1);
      for (int s = 0; s < 2 * TCS_NSTAGE; ++s) mbar_init(ba + 48 + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_s + ((uint32_t)(16 * tile) << 16);   // this tile's half of every lane quarter
  const int n_lf = p.T * p.L;
  const int NCH = tp.nchunk;

  if (tp.skew < 0 && tile == 1) {
    // timing study (ARP_TCD_SKEW=-1): tile 1 idle, its chains are not sampled
  } else if (warp >= 18) {
    // =========================== producers: chunk images L2 -> this tile's ring ===========================
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int s = 0; s < n_lf; ++s)
        for (int c = 0; c < NCH; ++c, ++cnt) {
          const uint32_t st = cnt % TCS_NSTAGE, use = cnt / TCS_NSTAGE;
          if (use > 0) mbar_wait(bar_xe + 8 * st, (use - 1) & 1);
          mbar_expect_tx(bar_xf + 8 * st, K::STAGE);
          bulk_g2s(ring + st * K::STAGE, tp.img + (size_t)c * K::STAGE, K::STAGE, bar_xf + 8 * st);
        }
    }
    __syncwarp();
  } else if (warp >= 16) {
    // =========================== MMA issuers (warp-uniform loop, lane 0 issues) ===========================
    {
      const uint32_t issue = lane == 0 ? 1u : 0u;   // all lanes run the loop; lane 0 issues
      // warp-uniform copies of everything that depends on the tile (derived from the warp index)
      const uint32_t tl = __shfl_sync(0xffffffffu, (uint32_t)tile, 0);
      const uint32_t bar_a = sbase + K::BAR + tl * 128, bar_h0 = bar_a + 8, bar_r0 = bar_a + 24, bar_g = bar_a + 40;
      const uint32_t bar_xf = bar_a + 48, bar_xe = bar_xf + 8 * TCS_NSTAGE;
      const uint32_t ring = sbase + K::RING + tl * TCS_NSTAGE * K::STAGE;
      const uint32_t sA[2] = {sbase + K::A1 + tl * K::AIMG, sbase + K::A2 + tl * K::AIMG};
      const int pa_sel[3] = {0, 0, 1}, pb_sel[3] = {0, 1, 0};
      const uint32_t tmu = __shfl_sync(0xffffffffu, tmem, 0);   // warp-uniform copy for the uniform datapath
      uint32_t cnt = 0;
      auto stage_of = [&](uint32_t k) { return ring + (k % TCS_NSTAGE) * K::STAGE; };
      const int nb0 = 3 + 3 * (int)tl;
      auto issue_g1 = [&](int c, uint32_t k, uint32_t commit_bar) {
        mbar_wait(bar_xf + 8 * (k % TCS_NSTAGE), (k / TCS_NSTAGE) & 1);
        tc_fence_after();
        const uint32_t d = tmu + K::COL_H + (uint32_t)(c & 1) * TC_CHUNK;
        const uint32_t xs = __shfl_sync(0xffffffffu, stage_of(k), 0);
        if (elect_one()) {
          const uint64_t a_base[2] = {tc_desc(sA[0], K::SF, K::SG), tc_desc(sA[1], K::SF, K::SG)};
          const uint64_t b_base = tc_desc(xs, K::SF, K::SG);
#pragma unroll
          for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int ks = 0; ks < NF / 16; ++ks)
              mma_ss(d, a_base[pa_sel[q]] + (uint64_t)((ks * 2 * K::SF) >> 4),
                     b_base + (uint64_t)((pb_sel[q] * K::XCHUNK + ks * 2 * K::SF) >> 4), K::IDESC_G1, (q | ks) ? 1u : 0u);
          if (commit_bar) tc_commit(commit_bar);
        }
        __syncwarp();
      };
      auto issue_g2 = [&](int c, uint32_t k, uint32_t stage_free_bar) {
        const uint32_t b = (uint32_t)(c & 1);
        const uint32_t xs = __shfl_sync(0xffffffffu, stage_of(k), 0);
        if (elect_one()) {
          const uint64_t b_base = tc_desc(xs, K::SG, K::SF);
          const uint32_t a_h = tmu + K::COL_H + b * TC_CHUNK, a_r2 = tmu + K::COL_R2 + b * 64;
#pragma unroll
          for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int w = 0; w < TC_NQ; ++w)
#pragma unroll
              for (int kk = 0; kk < 2; ++kk) {
                const uint32_t a_t = pa_sel[q] == 0 ? a_h + 32 * w + 8 * kk : a_r2 + 16 * w + 8 * kk;
                const uint32_t og = 4 * w + 2 * kk;
                mma_ts(tmu + K::COL_G, a_t, b_base + (uint64_t)((pb_sel[q] * K::XCHUNK + og * K::SG) >> 4), K::IDESC_G2,
                       (c | q | w | kk) ? 1u : 0u);
              }
          tc_commit(stage_free_bar);
        }
        __syncwarp();
      };
      for (int s = 0; s < n_lf; ++s) {
        const uint32_t k0 = cnt;
        tcd_nb_sync(nb0);
        tc_fence_after();
        issue_g1(0, k0, bar_h0);
        if (NCH > 1) issue_g1(1, k0 + 1, bar_h0 + 8);
        for (int c = 0; c < NCH; ++c) {
          const int b = c & 1;
          tcd_nb_sync(nb0 + 1 + b);
          tc_fence_after();
          issue_g2(c, k0 + c, bar_xe + 8 * ((k0 + c) % TCS_NSTAGE));
          if (c + 2 < NCH) issue_g1(c + 2, k0 + c + 2, bar_h0 + 8 * b);
          if (c == NCH - 1) tc_commit_if(issue, bar_g);
        }
        cnt += NCH;
      }
    }
    __syncwarp();
  } else {
    // ====================== chain workers: 16 chains x 2 workers per warp ======================
    const int qk = warp & 3;                 // TMEM lane quarter == SM sub-partition
    const int pp = warp >> 3;                // which pair of column blocks
    const int w = 2 * pp + (lane >> 4);      // worker quarter: features [8w, 8w+8), chunk columns [32w, 32w+32)
    const int r = tile * TCD_TILE + 16 * qk + (lane & 15);   // chain slot in the CTA; row 16 qk + (lane & 15) of the tile
    const int r64 = r - tile * TCD_TILE;
    const int chain = blockIdx.x * TC_CHAINS + r;
    const bool valid = chain < p.C;
    const int D = p.D, F = tp.F;
    // features are dealt to the four workers of a chain in contiguous, balanced ranges (25 -> 7, 6, 6, 6); worker w
    // owns K-slots [FPW w, FPW w + nf) of the A operand / X images and the matching columns of G
    const int nf = F / TC_NQ + (w < F % TC_NQ ? 1 : 0);
    const int fstart = w * (F / TC_NQ) + min(w, F % TC_NQ);
    const uint32_t tq = tmem + ((uint32_t)(32 * qk) << 16);   // lane base of this warp's 16 rows
    const size_t co = (size_t)chain * ws.sc;
    Vec Z{ws.z + co, ws.sd}, G{ws.g + co, ws.sd}, XC{ws.xc + co, ws.sd};
    float* xs = reinterpret_cast<float*>(smem + K::XS) + tid;
    const float* pa_s = reinterpret_cast<const float*>(smem + K::PAR);
    const float* pb_s = pa_s + (2 * NF + 4);
    const float* pe_s = pb_s + (2 * NF + 4);
    float lp_cur = ws.lp[chain], Hc = ws.H[chain], lavg = ws.lavg[chain], mult = ws.mult[chain];
    int nacc = ws.nacc[chain];
    const unsigned int gchain = p.chain_offset + (unsigned int)chain;
    uint32_t ph[2] = {0, 0}, pg = 0;
    const float a0 = pa_s[0], b0 = pb_s[0];
    // coordinate 0 (overall_log_scale) is replicated in all four workers of a chain.  Each keeps its own copy of
    // the current z / gradient in registers (all four take identical accept decisions), so no worker ever reads
    // what another worker of the chain writes to the global workspace.
    float z0_cur = Z(0), g0_cur = G(0), g0_prop = 0.f;
    uint8_t* a_row1 = smem + K::A1 + tile * K::AIMG + (r64 >> 3) * K::SG + (r64 & 7) * 16 + w * K::SF;
    uint8_t* a_row2 = smem + K::A2 + tile * K::AIMG + (r64 >> 3) * K::SG + (r64 & 7) * 16 + w * K::SF;
    const float LOG2E = 1.4426950408889634f;
    auto dof = [&](int i) { return i == 0 ? 0 : (i <= FPW ? fstart + i : F + fstart + i - FPW); };
    auto owned = [&](int i) { return i == 0 || (i <= FPW ? (i - 1) < nf : (i - 1 - FPW) < nf); };
    uint32_t par = 0;
    auto xch_at = [&](int slot, int q) -> float& { return xch[((par * 4 + slot) * TC_NQ + q) * TC_CHAINS + r]; };
    auto xch_sum = [&](int slot) { return (xch_at(slot, 0) + xch_at(slot, 1)) + (xch_at(slot, 2) + xch_at(slot, 3)); };
    float vreg[NLOC] = {};
#if TCD_PROFILE
    uint32_t pt[14] = {}, plast = (uint32_t)clock();
#define TCD_TICK(i) { const uint32_t now_ = (uint32_t)clock(); pt[i] += now_ - plast; plast = now_; }
#else
#define TCD_TICK(i)
#endif
    if (tile == 1 && tp.skew > 0) {   // start half a leapfrog step late: the two tiles then alternate phases
      const long long t0 = clock64();
      while (clock64() - t0 < tp.skew) {}
    }

    for (int t = 0; t < p.T; ++t) {
      const int tg = p.t_begin + t;
      TCD_TICK(9)
      if (p.ext_momenta) {
        const float* mom = p.ext_momenta + ((size_t)tg * p.C + (valid ? chain : 0)) * D;
#pragma unroll
        for (int i = 0; i < NLOC; ++i)
          if (owned(i)) xs[i * TC_WORKERS] = mom[dof(i)];
      } else {
        // Philox block j holds coordinates 4j .. 4j+3; my ranges are d = 0, [1+FPW w, ..+nf), [1+F+FPW w, ..+nf).
        // All blocks are generated unconditionally in unrolled loops (independent chains the scheduler can
        // interleave); only the stores are predicated.
        {
          float n4[4];
          philox_normal4_fast(p.seed, gchain, (unsigned int)tg, 0u, n4);
          xs[0] = n4[0];
        }
#pragma unroll
        for (int seg = 1; seg < 3; ++seg) {
          const int d_lo = seg == 1 ? 1 + fstart : 1 + F + fstart;
          const int d_hi = d_lo + nf;
          const int i_lo = seg == 1 ? 1 : 1 + FPW;
#pragma unroll
          for (int jj = 0; jj < FPW / 4 + 1; ++jj) {
            const int j = (d_lo >> 2) + jj;
            float n4[4];
            philox_normal4_fast(p.seed, gchain, (unsigned int)tg, (unsigned int)j, n4);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int d = 4 * j + q;
              if (d >= d_lo && d < d_hi) xs[(i_lo + d - d_lo) * TC_WORKERS] = n4[q];
            }
          }
        }
      }
      TCD_TICK(10)  // Philox momenta
      float ke0 = 0.f, ke1 = 0.f, ke0_tot = 0.f, ke1_tot = 0.f;
      {
        // all global loads first: with the loads inside the update loop every iteration waited a full L2
        // round trip (load -> FMA -> store -> next load cannot be hoisted above the store)
        float gq[NLOC], zq[NLOC];
#pragma unroll
        for (int i = 0; i < NLOC; ++i) {
          gq[i] = 0.f; zq[i] = 0.f;
          if (i == 0) { gq[0] = g0_cur; zq[0] = z0_cur; }
          else if (owned(i)) { const int d = dof(i); gq[i] = G(d); zq[i] = Z(d); }
        }
#pragma unroll
        for (int i = 0; i < NLOC; ++i) {
          if (owned(i)) {
            const int d = dof(i);
            float vi = xs[i * TC_WORKERS];
            if (i > 0 || w == 0) ke0 = fmaf(vi, vi, ke0);
            const float e = pe_s[d] * mult;
            vi = vi + 0.5f * e * gq[i];
            vreg[i] = vi;
            xs[i * TC_WORKERS] = zq[i] + e * vi;
          }
        }
      }
      float lpx = 0.f;
      // a coefficient that leaves the fp16 range of the A operand (only on wildly diverging trajectories) would make
      // GEMM1 return inf / NaN and the clamp below would swallow it: such a trajectory is rejected outright
      bool ovf = false;
      TCD_TICK(0)   // Philox + first kick
      for (int l = 0; l < p.L; ++l) {
        const bool last = (l == p.L - 1);
        float lp_top = 0.f;
        const Site s0 = site_fwd_fast(xs[0], 0.f, ARP_LOG_10, a0, b0, lp_top);
        {
          float be[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            be[k] = 0.f;
            if (k < nf) {
              const int f = fstart + k;
              float dummy = 0.f;
              float ls;
              if (GAMMA) ls = s0.x + xs[(1 + k) * TC_WORKERS];
              else ls = site_fwd_unit(xs[(1 + k) * TC_WORKERS], s0.x, pa_s[1 + f], dummy).x;
              const Site sb = site_fwd_fast(xs[(1 + FPW + k) * TC_WORKERS], 0.f, ls, pa_s[1 + F + f], pb_s[1 + F + f], dummy);
              be[k] = sb.x * LOG2E;   // GEMM1 then yields log2(e) t eta
              ovf |= !(fabsf(be[k]) < 60000.f);   // outside the fp16 range of the A operand (or NaN)
            }
          }
          uint4 hi, lo;
          split_pack(be[0], be[1], hi.x, lo.x);
          split_pack(be[2], be[3], hi.y, lo.y);
          split_pack(be[4], be[5], hi.z, lo.z);
          split_pack(be[6], be[7], hi.w, lo.w);
          *reinterpret_cast<uint4*>(a_row1) = hi;
          *reinterpret_cast<uint4*>(a_row2) = lo;
        }
        fence_async_smem();
        tc_fence_before();
        tcd_nb_arrive(3 + 3 * tile);
        TCD_TICK(1)   // site forward + A operand
        float lik = 0.f;
        for (int c = 0; c < NCH; ++c) {
          const int b = c & 1;
          mbar_wait(bar_h0 + 8 * b, ph[b]); ph[b] ^= 1;
          TCD_TICK(2)   // wait for H
          tc_fence_after();
          uint32_t hv[32];
          TCD_LD32(tq + K::COL_H + b * TC_CHUNK + 64 * pp, hv);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          TCD_TICK(3)   // tcgen05.ld
          uint32_t r1[16], r2[16];
          if (last) tcs_epilogue32<true, true>(hv, r1, r2, lik);
          else tcs_epilogue32<true, false>(hv, r1, r2, lik);
          TCD_ST16(tq + K::COL_H + b * TC_CHUNK + 64 * pp, 32, r1);   // packed head over the first half of each worker's H block
          TCD_ST16(tq + K::COL_R2 + b * 64 + 32 * pp, 16, r2);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
          tcd_nb_arrive(4 + 3 * tile + b);
          TCD_TICK(4)   // sigmoid + split + st
        }
        mbar_wait(bar_g, pg); pg ^= 1;
        TCD_TICK(5)   // wait for G
        tc_fence_after();
        uint32_t gv[FPW];
        TCD_LD8(tq + K::COL_G + 16 * pp, gv);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float acc0 = 0.f, lps = 0.f;
#pragma unroll
        for (int k = 0; k < FPW; ++k) {
          if (k < nf) {
            const int f = fstart + k;
            const float af = pa_s[1 + f], ab_ = pa_s[1 + F + f], bb_ = pb_s[1 + F + f];
            const float xs_s = xs[(1 + k) * TC_WORKERS], xs_b = xs[(1 + FPW + k) * TC_WORKERS];
            Site ss;
            if (GAMMA) {
              ss.x = xs_s;
              lps += 0.5f * xs_s - 0.5f * exp_fast(xs_s) + (float)(-0.34657359027997264 - 0.57236494292470008);
            } else {
              ss = site_fwd_unit(xs_s, s0.x, af, lps);
            }
            const Site sb = site_fwd_fast(xs_b, 0.f, GAMMA ? s0.x + xs_s : ss.x, ab_, bb_, lps);
            float gb, mb, lb, ab;
            site_rev(sb, __uint_as_float(gv[k]), 0.f, ab_, bb_, gb, mb, lb, ab);
            float gs, mb2, lb2, ab2;
            if (GAMMA) { gs = 0.5f - 0.5f * exp_fast(xs_s) + lb; mb2 = lb; }
            else site_rev(ss, lb, s0.x, af, 1.f, gs, mb2, lb2, ab2);
            acc0 += mb2;
            const float es = pe_s[1 + f] * mult, eb = pe_s[1 + F + f] * mult;
            float vs = vreg[1 + k] + 0.5f * es * gs;
            float vb = vreg[1 + FPW + k] + 0.5f * eb * gb;
            if (last) {
              ke1 = fmaf(vs, vs, ke1);
              ke1 = fmaf(vb, vb, ke1);
              ws.gx[co + (size_t)(1 + f) * ws.sd] = gs; ws.gx[co + (size_t)(1 + F + f) * ws.sd] = gb;
              ws.xcx[co + (size_t)(1 + f) * ws.sd] = ss.x; ws.xcx[co + (size_t)(1 + F + f) * ws.sd] = sb.x;
            } else {
              vs = vs + 0.5f * es * gs;
              vb = vb + 0.5f * eb * gb;
              xs[(1 + k) * TC_WORKERS] = xs_s + es * vs;
              xs[(1 + FPW + k) * TC_WORKERS] = xs_b + eb * vb;
            }
            vreg[1 + k] = vs;
            vreg[1 + FPW + k] = vb;
          }
        }
        if (last) {
          lik = 0.69314718055994531f * (lik + (w == 0 ? (float)(NCH * TC_CHUNK - tp.N) : 0.f));
        }
        xch_at(0, w) = acc0;
        xch_at(1, w) = ovf ? -INFINITY : lik + lps;
        xch_at(2, w) = ke0;
        xch_at(3, w) = ke1;
        TCD_TICK(6)   // G load + site reverse + kicks
        tile_bar(tile);
        TCD_TICK(7)   // named barrier
        const float acc0_t = xch_sum(0);
        lpx = xch_sum(1) + lp_top;
        if (l == 0) ke0_tot = xch_sum(2);
        if (last) ke1_tot = xch_sum(3);
        {
          float g0, mb, lb, ab;
          site_rev(s0, acc0_t, 0.f, a0, b0, g0, mb, lb, ab);
          const float e = pe_s[0] * mult;
          float v0 = vreg[0] + 0.5f * e * g0;
          if (last) {
            ke1_tot = fmaf(v0, v0, ke1_tot);
            g0_prop = g0;
            if (w == 0) ws.xcx[co] = s0.x;
          } else {
            v0 = v0 + 0.5f * e * g0;
            xs[0] = xs[0] + e * v0;
          }
          vreg[0] = v0;
        }
        par ^= 1;
        TCD_TICK(8)   // top-site reverse
      }
      float log_alpha = lpx - lp_cur + 0.5f * ke0_tot - 0.5f * ke1_tot;
      if (!(log_alpha == log_alpha) || log_alpha == -INFINITY) log_alpha = -INFINITY;
      float log_u;
      if (p.ext_log_u) log_u = p.ext_log_u[(size_t)tg * p.C + (valid ? chain : 0)];
      else log_u = philox_log_uniform(p.seed, gchain, (unsigned int)tg);
      const bool acc = log_u < log_alpha;
      TCD_TICK(11)  // log alpha + uniform
      if (acc) {
        float gq[NLOC], xq[NLOC];   // loads first, then stores (see the first kick)
#pragma unroll
        for (int i = 0; i < NLOC; ++i) {
          gq[i] = 0.f; xq[i] = 0.f;
          if (i == 0) {
            gq[0] = g0_prop;
            if (w == 0) xq[0] = ws.xcx[co];
          } else if (owned(i)) {
            const int d = dof(i);
            gq[i] = ws.gx[co + (size_t)d * ws.sd];
            xq[i] = ws.xcx[co + (size_t)d * ws.sd];
          }
        }
        z0_cur = xs[0];
        g0_cur = g0_prop;
#pragma unroll
        for (int i = 0; i < NLOC; ++i)
          if (owned(i) && (i > 0 || w == 0)) {
            const int d = dof(i);
            Z(d) = xs[i * TC_WORKERS];
            G(d) = gq[i];
            XC(d) = xq[i];
          }
        lp_cur = lpx;
        ++nacc;
      }
      TCD_TICK(12)  // accept copy
      const int t1 = tg + 1;
      if (t1 <= p.num_adapt) {
        const float ft = (float)t1;
        Hc += p.target_accept - expf(log_alpha < 0.f ? log_alpha : 0.f);
        const float log_step = ARP_LOG_10 - Hc * sqrtf(ft) / ((ft + 10.f) * 0.05f);
        const float eta = powf(ft, -0.75f);
        lavg = eta * log_step + (1.f - eta) * lavg;
        mult = (t1 < p.num_adapt) ? expf(log_step) : expf(lavg);
      }
      TCD_TICK(13)  // step-size adaptation
      const int since = tg - p.num_burnin;
      if (since >= 0 && (since % p.stride) == 0 && valid) {
        const int s = since / p.stride;
        if (s < p.S) {
          const size_t o = ((size_t)s * p.C + chain) * D;
          float xq[NLOC], zq[NLOC];
#pragma unroll
          for (int i = 0; i < NLOC; ++i) {
            xq[i] = 0.f; zq[i] = 0.f;
            if (owned(i) && (i > 0 || w == 0)) {
              const int d = dof(i);
              if (p.samples) xq[i] = XC(d);
              if (p.samples_orig) zq[i] = Z(d);
            }
          }
#pragma unroll
          for (int i = 0; i < NLOC; ++i)
            if (owned(i) && (i > 0 || w == 0)) {
              const int d = dof(i);
              if (p.samples) p.samples[o + d] = xq[i];
              if (p.samples_orig) p.samples_orig[o + d] = zq[i];
            }
          if (p.is_accepted && w == 0) p.is_accepted[(size_t)s * p.C + chain] = acc ? 1 : 0;
        }
      }
    }
#if TCD_PROFILE
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 4 || warp == 8))
      printf("warp %d: philox %u kick0 %u fwd %u waitH %u ld %u epi %u waitG %u rev %u bar %u top %u | alpha+u %u accept %u adapt %u store %u\n",
             warp, pt[10], pt[0], pt[1], pt[2], pt[3], pt[4], pt[5], pt[6], pt[7], pt[8], pt[11], pt[12], pt[13], pt[9]);
#endif
    if (w == 0) {
      ws.lp[chain] = lp_cur; ws.H[chain] = Hc; ws.lavg[chain] = lavg; ws.mult[chain] = mult; ws.nacc[chain] = nacc;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_ptr_s), "r"((uint32_t)TC_TMEM_COLS) : "memory");
  }
}

template <bool GAMMA>
static inline cudaError_t tcd_launch(dim3 grid, cudaStream_t st, const TcsParams& tp, const HmcWs& ws, const HmcArgs& p) {
  cudaError_t e = cudaFuncSetAttribute(k_german_tcd_hmc<GAMMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tcd::BYTES);
  if (e != cudaSuccess) return e;
  k_german_tcd_hmc<GAMMA><<<grid, TCD_THREADS, Tcd::BYTES, st>>>(tp, ws, p);
  return cudaGetLastError();
}

}  // namespace arp
