// tcgen05 engine for German credit (the only dense contraction of the hot path; primitives in arp_german_tc.cuh).
// One CTA owns 128 chains (= the 128 TMEM lanes).  X (fp16 head + tail of the label-scaled rows) is cut into 128-observation chunk
// images that a producer warp ring-buffers from L2 into shared memory with bulk asynchronous copies
// (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP); each chunk image feeds GEMM1 (K-major B) and,
// one epilogue later, GEMM2 (MN-major B) before its stage is released by tcgen05.commit.  This lifts
// the limits of the resident kernel: up to 64 features with full tails (the reference's real German
// credit data is 1000 x 62) and any number of observations.  L2 -> SM traffic is one chunk image per
// 128 observations per gradient evaluation of 128 chains (~1 % of L2 bandwidth at the measured rate).
//
// Roles (576 threads): warps 0-15 workers (4 per chain, as in the resident kernel), warp 16 MMA issuer
// (one lane), warp 17 producer (one lane).  For NF = 64 a worker owns 16 features, so the momentum moves
// from registers to the [d][chain] global workspace (coalesced, L2 resident).
#pragma once
#include "arp_german_tc.cuh"

namespace arp {

#define TCS_NSTAGE 3
// Worker -> MMA-issuer hand-offs ("A operand written", "residual of chunk c stored") use hardware NAMED barriers:
// the 512 workers bar.arrive, the issuer warp bar.sync's and is descheduled until the last arrival.  With an
// mbarrier the issuer's try_wait returned immediately and its BRA / YIELD / TRYWAIT spin loop was 13.6 % of all
// instructions the kernel executed (ncu source page), all of them on the one SM sub-partition that hosts the issuer
// warp next to four worker warps -- and the slowest sub-partition paces every chunk.
#define TCS_NB_A 2          // named barrier ids (1 = epi_bar among the workers); 3, 4 = residual buffers 0, 1
#define TCS_NB_R 3
#define TCS_NB_COUNT (TC_WORKERS + 32)
__device__ __forceinline__ void nb_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(TCS_NB_COUNT) : "memory"); }
__device__ __forceinline__ void nb_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(TCS_NB_COUNT) : "memory"); }
#ifndef TCS_PROFILE
#define TCS_PROFILE 0   // 1: one lane per worker warp accumulates clock() per phase and printf()s it (timing study only)
#endif
// TCS_GSPLIT 1: GEMM2 keeps SEPARATE TMEM accumulators for the head x head pass (q1 X1: the big products) and for the two
// tail passes (q1 X2 + q2 X1: 2^-11 of it), and the workers add them in fp32 at the end of the step.  The tensor core's
// fp32 accumulator is what limits the accuracy of G = X^T q: every MMA truncates at the magnitude of the running sum, and
// with one accumulator for all three passes that is 192 dependent accumulations per step on partial sums ~10 x the
// result (gradient at 1e-5 of the fp64 oracle, max-norm, typical-set states; the SIMT engine: 2e-7 .. 7e-7).  Split, the
// head accumulator sees 64 accumulations (32 with NF = 32, where TMEM has room for one head accumulator per chunk
// parity) and the tail sums are formed at their own scale.  Costs two (one) extra tcgen05.ld per step, no extra MMA.
// (Measured and removed in round 2: a fresh accumulator per chunk / per group of 2 or 4 chunks with the running sum
// kept by the workers in TMEM: 4 x more accurate, but 12 - 15 % slower -- the extra tcgen05.ld / st in the chunk loop.)
// Default (1): split for NF = 64 only -- the real 1000 x 62 data, where the elementwise 1e-5 gradient test needs it and
// the second accumulator costs 1 % (one extra load per step); for NF = 32 (two head accumulators + tails: three loads
// per step, more spills at the 96-register cap) it was measured 5 % slower (158.8 vs 150.8 ms per bench step) for a
// gradient that already passes the test.  2: split for both.  0: off.
#ifndef TCS_GSPLIT
#define TCS_GSPLIT 1
#endif
#define TCS_THREADS (TC_WORKERS + 64)
#define TCS_MMA_WARP (TC_WORKERS / 32)
#define TCS_PROD_WARP (TC_WORKERS / 32 + 1)

template <int NF>
struct Tcs {
  static constexpr uint32_t NFC = NF / 8;                 // feature chunks of 8
  static constexpr uint32_t SF = 128;                     // bytes between feature chunks
  static constexpr uint32_t SG = NFC * 128;               // bytes between 8-row groups
  static constexpr uint32_t XCHUNK = (TC_CHUNK / 8) * SG; // one part of one 128-observation chunk
  static constexpr uint32_t STAGE = 2 * XCHUNK;                  // head | tail of the t-scaled rows
  static constexpr uint32_t AIMG = (TC_CHAINS / 8) * SG;
  static constexpr int FPW = NF / TC_NQ;                  // features per worker
  static constexpr int NLOC = 1 + 2 * FPW;
  static constexpr bool RCP_SHARE = (NF == 32);
  // shared memory
  static constexpr uint32_t RING = 0;
  static constexpr uint32_t A1 = RING + TCS_NSTAGE * STAGE;
  static constexpr uint32_t A2 = A1 + AIMG;
  static constexpr uint32_t XCH = A2 + AIMG;                        // float[2][4][TC_NQ][128] (double-buffered by step parity)
  static constexpr uint32_t XS = XCH + 2 * 4 * TC_NQ * TC_CHAINS * 4;   // float[NLOC][512]
  static constexpr uint32_t PAR = XS + NLOC * TC_WORKERS * 4;       // float[3][2 NF + 4]
  static constexpr uint32_t BAR = PAR + 4 * (2 * NF + 4) * 4;       // 6 + 2 * NSTAGE mbarriers (PAR: a, b, eps0)
  static constexpr uint32_t TMEM_PTR = BAR + 16 * 8;
  // NF = 32: momenta of the NEXT transition, float[NLOC][512] (drawn while the workers wait for the first GEMM1 of the
  // last leapfrog step; NF = 64 has no shared memory left and draws them at the start of the transition)
  static constexpr bool MOM_AHEAD = (NF == 32);
  static constexpr uint32_t MOM = TMEM_PTR + 16;
  static constexpr uint32_t BYTES = MOM + (MOM_AHEAD ? NLOC * TC_WORKERS * 4 : 0);
  // TMEM columns: H 0 .. 255 (two chunk buffers; the fp16 head of q overwrites H in place), G accumulators, fp16 tail
  // of q 320 .. 447.  NF = 32: G head accumulators at 256 / 288 (even / odd chunks), tail accumulator at 448;
  // NF = 64: one head accumulator at 256 .. 319, tail accumulator at 448 .. 511.
  static constexpr uint32_t COL_H = 0, COL_G = 256, COL_R2 = 320, COL_GB = 448;
  static constexpr bool GSPLIT = (TCS_GSPLIT == 2) || (TCS_GSPLIT == 1 && NF == 64);
  static constexpr int NHEAD = (NF == 32) ? 2 : 1;
  static constexpr uint32_t COL_G1 = COL_G + 32;   // NHEAD == 2 only
  static constexpr uint32_t IDESC_G1 = (1u << 4) | ((uint32_t)(TC_CHUNK >> 3) << 17) | ((128u >> 4) << 24);
  static constexpr uint32_t IDESC_G2 = (1u << 4) | (1u << 16) | ((uint32_t)(NF >> 3) << 17) | ((128u >> 4) << 24);
  static_assert(BYTES <= 232448, "shared memory budget");
};

struct TcsParams {
  const uint8_t* img;   // nchunk stage images
  int N, F, nchunk;
};

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Likelihood epilogue of one worker: 32 elements of one chunk.  The chunk images hold the rows of X scaled by
// t_n = 2 y_n - 1 (exact in fp16) and the A operand is +log2(e) beta, so GEMM1 yields h_n = log2(e) t_n eta_n and
//   q_n = sigmoid(-t_n eta_n) = 1 / (1 + 2^h_n),   y_n - sigmoid(eta_n) = t_n q_n,
// i.e. GEMM2 over the SAME scaled rows needs q itself: X^T (y - sigmoid(eta)) = (t X)^T q.  No y in the epilogue.
// log-likelihood (last leapfrog step only): ln sigmoid(t eta) = ln(1 - q) = ln2 (h + log2 q).
// SHARE: four sigmoids share ONE MUFU.RCP (Montgomery's trick: 1/d_k = (1 / prod d) * prod_{j != k} d_j), i.e.
// 1.25 MUFU + 2.25 FMUL per element instead of 2 MUFU.  h is clamped at 30 (q < 1e-9 there) so the product of four
// denominators stays below 2^124; with m = min(h, 30) the log-likelihood term is m + log2 q(m) (= 0 to fp32 accuracy
// beyond the clamp, as it should).
// The eight groups of four elements run through a three-stage software pipeline (A: clamp + EX2, B: denominators
// + shared RCP, C: sigmoids + fp16 head / tail split), group g + 3 in A while g + 2 is in B and g in C: every MUFU
// result is consumed ~40 instructions after its issue, and the MUFU ops (the binding pipe: 8 clk per warp
// instruction per SM sub-partition, profiles/micro/pipes.cu) are spread evenly over the chunk instead of arriving
// as one burst per warp.  (Written group by group, ptxas put each MUFU.RCP directly in front of its consumers:
// 0.5 IPC inside the epilogue with four worker warps per scheduler, measured with clock().)
// (A clamp-free fast path -- one chained FSETP per group on the product of the denominators instead of the 32
// half-rate FMNMX, with a clamped cold path -- was measured 4 % SLOWER end to end and is not kept.)
// fp32 pair -> packed fp16 head / tail by TRUNCATION: head = the top 10 mantissa bits (one full-rate LOP3 instead of
// the half-rate F2FP -> HADD2.F32 round trip of split_pack), tail = x - head (exact), both then packed with F2FP.
// |head| < 2^-14 (fp16 subnormal) rounds in the pack: an absolute error below 2^-25 on a sigmoid value.
__device__ __forceinline__ void split_pack_trunc(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const float h0 = __uint_as_float(__float_as_uint(x0) & 0xffffe000u), h1 = __uint_as_float(__float_as_uint(x1) & 0xffffe000u);
  const __half2 h = __floats2half2_rn(h0, h1);
  const __half2 l = __floats2half2_rn(x0 - h0, x1 - h1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
#ifdef TCS_TRUNC_SPLIT
#define TCS_SPLIT split_pack_trunc
#else
#define TCS_SPLIT split_pack
#endif

// (Measured and removed in round 2: folding the clamp into GEMM1 through a bias K-slot -- h' = h - 31, one FADD.SAT
// instead of the half-rate FMNMX + FADD per element -- changed nothing (153.1 vs 153.2 ms), and carrying both epilogue
// variants in the kernel cost 4 %.  The epilogue is not dispatch-bound; see DESIGN.md.)
template <bool SHARE, bool LAST>
__device__ __forceinline__ void tcs_epilogue32(const uint32_t* hv, uint32_t* r1, uint32_t* r2, float& lik) {
  float m[8][4], e[8][4], p01[8], p23[8], inv[8];
  auto stage_a = [&](int g) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      m[g][q] = fminf(__uint_as_float(hv[4 * g + q]), 30.f);
      e[g][q] = ex2_approx(m[g][q]);
    }
  };
  auto stage_b = [&](int g) {
#pragma unroll
    for (int q = 0; q < 4; ++q) e[g][q] += 1.0f;
    if constexpr (SHARE) {
      p01[g] = e[g][0] * e[g][1];
      p23[g] = e[g][2] * e[g][3];
      inv[g] = rcp_approx(p01[g] * p23[g]);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) e[g][q] = rcp_approx(e[g][q]);
    }
  };
  auto stage_c = [&](int g) {
    float qv[4];
    if constexpr (SHARE) {
      const float i01 = inv[g] * p23[g], i23 = inv[g] * p01[g];
      qv[0] = i01 * e[g][1]; qv[1] = i01 * e[g][0]; qv[2] = i23 * e[g][3]; qv[3] = i23 * e[g][2];
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) qv[q] = e[g][q];
    }
    if (LAST) {
      // ln sigmoid(t eta) / ln2 = h + log2 q.  The two terms cancel for well-predicted observations, so the
      // difference is formed per group and only the (small) differences are accumulated.  With the shared
      // reciprocal sum_group log2 q = log2 prod q = log2(inv): ONE MUFU.LG2 per four observations.
      if constexpr (SHARE) {
        lik += ((m[g][0] + m[g][1]) + (m[g][2] + m[g][3])) + lg2_approx(inv[g]);
      } else {
        const float t0 = m[g][0] + lg2_approx(qv[0]), t1 = m[g][1] + lg2_approx(qv[1]);
        const float t2 = m[g][2] + lg2_approx(qv[2]), t3 = m[g][3] + lg2_approx(qv[3]);
        lik += (t0 + t1) + (t2 + t3);
      }
    }
    TCS_SPLIT(qv[0], qv[1], r1[2 * g], r2[2 * g]);
    TCS_SPLIT(qv[2], qv[3], r1[2 * g + 1], r2[2 * g + 1]);
  };
  stage_a(0); stage_a(1); stage_b(0); stage_a(2); stage_b(1);
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    if (g + 3 < 8) stage_a(g + 3);
    if (g + 2 < 8) stage_b(g + 2);
    stage_c(g);
  }
}

// The same epilogue on PACKED fp32 pairs (sm_100 FADD2 / FMUL2 / FFMA2: one instruction, two elements).  Pair i =
// elements (2i, 2i + 1) -- adjacent TMEM columns, so the registers tcgen05.ld delivers are already aligned pairs and the
// fp16 head / tail of a pair is exactly one packed half2 of the A operand.  Each lane of a pair belongs to its own group
// of four (pairs 4G .. 4G + 3), so the shared-reciprocal trick runs on both lanes at once: per 32 elements 16 FADD2 +
// 36 FMUL2 + 16 FFMA2 replace 64 FADD + 72 FMUL (204 instead of 272 instructions).  TCS_PACKED selects it.
// (bit mask for A/B measurements: 1 = packed adds, 2 = packed multiplies, 4 = packed head subtraction; 7 = all)
#ifndef TCS_PACKED
#define TCS_PACKED 7
#endif
__device__ __forceinline__ uint64_t f2_pack(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void f2_unpack(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t r;
  if constexpr (TCS_PACKED & 2) { asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); }
  else { float a0, a1, b0, b1; f2_unpack(a, a0, a1); f2_unpack(b, b0, b1); r = f2_pack(a0 * b0, a1 * b1); }
  return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  if constexpr (TCS_PACKED & 1) { asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); }
  else { float a0, a1, b0, b1; f2_unpack(a, a0, a1); f2_unpack(b, b0, b1); r = f2_pack(a0 + b0, a1 + b1); }
  return r;
}
// a * b + c, used only as c - a (b = -1): exact either way
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  if constexpr (TCS_PACKED & 4) { asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); }
  else { float a0, a1, b0, b1, c0, c1; f2_unpack(a, a0, a1); f2_unpack(b, b0, b1); f2_unpack(c, c0, c1); r = f2_pack(fmaf(a0, b0, c0), fmaf(a1, b1, c1)); }
  return r;
}

template <bool LAST>
__device__ __forceinline__ void tcs_epilogue32_packed(const uint32_t* hv, uint32_t* r1, uint32_t* r2, float& lik) {
  uint64_t d[4][4], p01[4], p23[4], inv[4], ms[4];
  const uint64_t one2 = f2_pack(1.0f, 1.0f), mone2 = f2_pack(-1.0f, -1.0f);
  auto stage_a = [&](int G) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = 4 * G + k;
      const float m0 = fminf(__uint_as_float(hv[2 * i]), 30.f), m1 = fminf(__uint_as_float(hv[2 * i + 1]), 30.f);
      d[G][k] = f2_pack(ex2_approx(m0), ex2_approx(m1));
      if (LAST) ms[G] = k == 0 ? f2_pack(m0, m1) : f2_add(ms[G], f2_pack(m0, m1));
    }
  };
  auto stage_b = [&](int G) {
#pragma unroll
    for (int k = 0; k < 4; ++k) d[G][k] = f2_add(d[G][k], one2);
    p01[G] = f2_mul(d[G][0], d[G][1]);
    p23[G] = f2_mul(d[G][2], d[G][3]);
    float x, y;
    f2_unpack(f2_mul(p01[G], p23[G]), x, y);
    inv[G] = f2_pack(rcp_approx(x), rcp_approx(y));
  };
  auto stage_c = [&](int G) {
    const uint64_t i01 = f2_mul(inv[G], p23[G]), i23 = f2_mul(inv[G], p01[G]);
    const uint64_t q[4] = {f2_mul(i01, d[G][1]), f2_mul(i01, d[G][0]), f2_mul(i23, d[G][3]), f2_mul(i23, d[G][2])};
    if (LAST) {   // sum over the 8 elements of (m + log2 q) = sum m + log2(inv) of both lanes (see tcs_epilogue32)
      float mx, my, ix, iy;
      f2_unpack(ms[G], mx, my);
      f2_unpack(inv[G], ix, iy);
      lik += (mx + my) + (lg2_approx(ix) + lg2_approx(iy));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float q0, q1, t0, t1;
      f2_unpack(q[k], q0, q1);
      const __half2 h = __floats2half2_rn(q0, q1);
      const float2 hf = __half22float2(h);
      f2_unpack(f2_fma(f2_pack(hf.x, hf.y), mone2, q[k]), t0, t1);     // q - head: exact
      const __half2 l = __floats2half2_rn(t0, t1);
      r1[4 * G + k] = *reinterpret_cast<const uint32_t*>(&h);
      r2[4 * G + k] = *reinterpret_cast<const uint32_t*>(&l);
    }
  };
  stage_a(0); stage_a(1); stage_b(0);
#pragma unroll
  for (int G = 0; G < 4; ++G) {
    if (G + 2 < 4) stage_a(G + 2);
    if (G + 1 < 4) stage_b(G + 1);
    stage_c(G);
  }
}

// Standard-normal momenta of transition `tgen` for the coordinates of one worker, written to dst[i * TC_WORKERS].
// Philox block j holds coordinates 4j .. 4j+3; the worker's ranges are d = 0, [1 + fstart, .. + nf) and
// [1 + F + fstart, .. + nf).  All blocks are generated unconditionally in unrolled loops (independent chains the
// scheduler can interleave); only the stores are predicated.  (Skipping the blocks that lie entirely outside the
// ranges with a warp-uniform test was measured 1.5 % slower: the branches serialise the blocks.)
// NOINLINE: a real call keeps the ~40 registers of the Philox rounds out of the register allocation of the leapfrog
// loop it is called from (inlined there it caused spills in the hot loop: -6 %).
template <int FPW, bool NOINLINE>
__device__ __forceinline__ void tcs_draw_momenta_body(uint64_t seed, uint32_t gchain, uint32_t tgen, int fstart, int nf, int F,
                                                      float* dst, int parts) {
  // parts: bit 0 = coordinate 0 and the log-scale range, bit 1 = the coefficient range (the draw is split over the wait
  // windows of the last two leapfrog steps)
  if (parts & 1) {
    float n4[4];
    philox_normal4_fast(seed, gchain, tgen, 0u, n4);
    dst[0] = n4[0];
  }
#pragma unroll
  for (int seg = 1; seg < 3; ++seg) {
    if (!(parts & seg)) continue;
    const int d_lo = seg == 1 ? 1 + fstart : 1 + F + fstart;
    const int d_hi = d_lo + nf;
    const int i_lo = seg == 1 ? 1 : 1 + FPW;
#pragma unroll
    for (int jj = 0; jj < FPW / 4 + 1; ++jj) {
      const int j = (d_lo >> 2) + jj;
      float n4[4];
      philox_normal4_fast(seed, gchain, tgen, (unsigned int)j, n4);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int d = 4 * j + q;
        if (d >= d_lo && d < d_hi) dst[(i_lo + d - d_lo) * TC_WORKERS] = n4[q];
      }
    }
  }
}
template <int FPW>
__device__ __noinline__ void tcs_draw_momenta_call(uint64_t seed, uint32_t gchain, uint32_t tgen, int fstart, int nf, int F,
                                                   float* dst, int parts) {
  tcs_draw_momenta_body<FPW, true>(seed, gchain, tgen, fstart, nf, F, dst, parts);
}

// GAMMA = german_credit_gammascale (models.py:930-945): beta_log_scales is not a Normal site (never
// reparameterised); beta ~ N(0, exp(overall_log_scale + beta_log_scales)).
template <int NF, bool GAMMA, bool MULTI = false>
__global__ void
#ifdef TCS_MAXNREG
__maxnreg__(TCS_MAXNREG)
#else
__launch_bounds__(TCS_THREADS, 1)
#endif
k_german_tcs_hmc(TcsParams tp, HmcWs ws, HmcArgs p) {
  using K = Tcs<NF>;
  // multi-run launch (leapfrog-step tuning grid): this CTA's run supplies L, the transition counts, the step sizes and
  // the output buffers; rv.chain = index of the CTA's first chain inside its run
  const HmcRun rv = hmc_run_view<MULTI>(p, (int)blockIdx.x * TC_CHAINS);
  const int chain0 = rv.chain;
  constexpr int FPW = K::FPW, NLOC = K::NLOC;
  constexpr bool V_IN_REGS = (NF <= 32);
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_a = sbase + K::BAR, bar_h0 = bar_a + 8, bar_r0 = bar_a + 24, bar_g = bar_a + 40;
  const uint32_t bar_xf = bar_a + 48, bar_xe = bar_xf + 8 * TCS_NSTAGE;
  float* xch = reinterpret_cast<float*>(smem + K::XCH);
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + K::TMEM_PTR);
  {
    float* par = reinterpret_cast<float*>(smem + K::PAR);
    for (int i = tid; i < p.D; i += TCS_THREADS) {
      par[i] = p.a[i];
      par[(2 * NF + 4) + i] = p.b[i];
      par[2 * (2 * NF + 4) + i] = ARP_RUN(eps0)[i];
    }
  }
  if (warp == TCS_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)),
                 "r"((uint32_t)TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(bar_a, TC_WORKERS);
    mbar_init(bar_h0, 1); mbar_init(bar_h0 + 8, 1);
    mbar_init(bar_r0, TC_WORKERS); mbar_init(bar_r0 + 8, TC_WORKERS);
    mbar_init(bar_g, 1);
    for (int s = 0; s < TCS_NSTAGE; ++s) { mbar_init(bar_xf + 8 * s, 1); mbar_init(bar_xe + 8 * s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_s;
  const int n_lf = ARP_RUN(T) * ARP_RUN(L);
  const int NCH = tp.nchunk;

  if (warp == TCS_PROD_WARP) {
    // =========================== producer: chunk images L2 -> smem ring ===========================
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int s = 0; s < n_lf; ++s)
        for (int c = 0; c < NCH; ++c, ++cnt) {
          const uint32_t st = cnt % TCS_NSTAGE, use = cnt / TCS_NSTAGE;
          if (use > 0) mbar_wait(bar_xe + 8 * st, (use - 1) & 1);
          mbar_expect_tx(bar_xf + 8 * st, K::STAGE);
          bulk_g2s(sbase + K::RING + st * K::STAGE, tp.img + (size_t)c * K::STAGE, K::STAGE, bar_xf + 8 * st);
        }
    }
    __syncwarp();
  } else if (warp == TCS_MMA_WARP) {
    // =========================== MMA issuer (warp-uniform loop, lane 0 issues) ===========================
    {
      const uint32_t issue = lane == 0 ? 1u : 0u;   // all lanes run the loop; lane 0 issues
      const uint32_t sA[2] = {sbase + K::A1, sbase + K::A2};
      const int pa_sel[3] = {0, 0, 1}, pb_sel[3] = {0, 1, 0};
      const uint32_t tmu = __shfl_sync(0xffffffffu, tmem, 0);   // warp-uniform copy for the uniform datapath
      uint32_t cnt = 0;  // global chunk counter of the next GEMM1 to issue
      auto stage_of = [&](uint32_t k) { return sbase + K::RING + (k % TCS_NSTAGE) * K::STAGE; };
      auto issue_g1 = [&](int c, uint32_t k, uint32_t commit_bar) {
        mbar_wait(bar_xf + 8 * (k % TCS_NSTAGE), (k / TCS_NSTAGE) & 1);
        tc_fence_after();
        const uint32_t d = tmu + K::COL_H + (uint32_t)(c & 1) * TC_CHUNK;
        const uint32_t xs = __shfl_sync(0xffffffffu, stage_of(k), 0);
        // one elected lane issues the whole batch; the descriptors of a batch differ only in the 14-bit start
        // address field (16-byte units, no carry for shared-memory addresses), so each one is base + constant
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
                const uint32_t og = 4 * w + 2 * kk;  // 8-observation group inside the chunk
                if constexpr (!K::GSPLIT) {
                  mma_ts(tmu + K::COL_G, a_t, b_base + (uint64_t)((pb_sel[q] * K::XCHUNK + og * K::SG) >> 4), K::IDESC_G2,
                         (c | q | w | kk) ? 1u : 0u);
                } else if (q == 0) {
                  const uint32_t hc = K::NHEAD == 2 ? (uint32_t)(c >> 1) : (uint32_t)c;   // uses of this accumulator so far
                  mma_ts(tmu + ((K::NHEAD == 2 && (c & 1)) ? K::COL_G1 : K::COL_G), a_t,
                         b_base + (uint64_t)((pb_sel[q] * K::XCHUNK + og * K::SG) >> 4), K::IDESC_G2, (hc | w | kk) ? 1u : 0u);
                } else {
                  mma_ts(tmu + K::COL_GB, a_t, b_base + (uint64_t)((pb_sel[q] * K::XCHUNK + og * K::SG) >> 4), K::IDESC_G2,
                         (c | (q - 1) | w | kk) ? 1u : 0u);
                }
              }
          tc_commit(stage_free_bar);   // stage free once GEMM2(c) has read it
        }
        __syncwarp();
      };
      for (int s = 0; s < n_lf; ++s) {
        const uint32_t k0 = cnt;  // global index of chunk 0 of this step
        nb_sync(TCS_NB_A);
        tc_fence_after();
        issue_g1(0, k0, bar_h0);
        if (NCH > 1) issue_g1(1, k0 + 1, bar_h0 + 8);
        for (int c = 0; c < NCH; ++c) {
          const int b = c & 1;
          nb_sync(TCS_NB_R + b);
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
    // ====================== chain workers (4 per chain) ======================
    const int w = tid >> 7, r = tid & 127;
    const int row = blockIdx.x * TC_CHAINS + r;    // workspace row
    const int chain = MULTI ? chain0 + r : row;    // chain index inside the run: RNG streams, outputs
    const bool valid = chain < p.C;
    const int D = p.D, F = tp.F;
    // features are dealt to the four workers of a chain in contiguous, balanced ranges (25 -> 7, 6, 6, 6); worker w
    // owns K-slots [FPW w, FPW w + nf) of the A operand / X images and the matching columns of G
    const int nf = F / TC_NQ + (w < F % TC_NQ ? 1 : 0);
    const int fstart = w * (F / TC_NQ) + min(w, F % TC_NQ);
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const size_t co = (size_t)row * ws.sc;
    Vec Z{ws.z + co, ws.sd}, G{ws.g + co, ws.sd}, XC{ws.xc + co, ws.sd}, VG{ws.v + co, ws.sd};
    float* xs = reinterpret_cast<float*>(smem + K::XS) + tid;
    const float* pa_s = reinterpret_cast<const float*>(smem + K::PAR);
    const float* pb_s = pa_s + (2 * NF + 4);
    const float* pe_s = pb_s + (2 * NF + 4);
    float lp_cur = ws.lp[row], Hc = ws.H[row], lavg = ws.lavg[row], mult = ws.mult[row];
    int nacc = ws.nacc[row];
    const unsigned int gchain = p.chain_offset + (unsigned int)chain;
    uint32_t ph[2] = {0, 0}, pg = 0;
    const float a0 = pa_s[0], b0 = pb_s[0];
    // coordinate 0 (overall_log_scale) is replicated in all four workers of a chain.  Each keeps its own copy of
    // the current z / gradient in registers (all four take identical accept decisions), so no worker ever reads
    // what another worker of the chain writes to the global workspace.
    float z0_cur = Z(0), g0_cur = G(0), g0_prop = 0.f;
    // The gradient / centred values of the current state and of the proposal live in two buffer sets
    // (g, xc) and (gx, xcx) that swap roles on accept: no copy.  ocur = element offset of the current set.
    const uint32_t oflip = (uint32_t)(ws.gx - ws.g);   // == ws.xcx - ws.xc (checked by the host)
    uint32_t ocur = 0;
    auto Gc = [&](int d) -> float& { return ws.g[co + (size_t)d * ws.sd + ocur]; };
    auto XCc = [&](int d) -> float& { return ws.xc[co + (size_t)d * ws.sd + ocur]; };
    auto Gp = [&](int d) -> float& { return ws.g[co + (size_t)d * ws.sd + (oflip - ocur)]; };
    auto XCp = [&](int d) -> float& { return ws.xc[co + (size_t)d * ws.sd + (oflip - ocur)]; };
    uint8_t* a_row1 = smem + K::A1 + (r >> 3) * K::SG + (r & 7) * 16 + w * (FPW / 8) * K::SF;
    uint8_t* a_row2 = smem + K::A2 + (r >> 3) * K::SG + (r & 7) * 16 + w * (FPW / 8) * K::SF;
    const float LOG2E = 1.4426950408889634f;
    auto dof = [&](int i) { return i == 0 ? 0 : (i <= FPW ? fstart + i : F + fstart + i - FPW); };
    auto owned = [&](int i) { return i == 0 || (i <= FPW ? (i - 1) < nf : (i - 1 - FPW) < nf); };
    // cross-quarter exchange slots, double-buffered by the parity of the step counter so ONE named barrier per
    // leapfrog step suffices (a thread can be at most one step ahead of the slowest one)
    uint32_t par = 0;
    auto xch_at = [&](int slot, int q) -> float& { return xch[((par * 4 + slot) * TC_NQ + q) * TC_CHAINS + r]; };
    auto xch_sum = [&](int slot) { return (xch_at(slot, 0) + xch_at(slot, 1)) + (xch_at(slot, 2) + xch_at(slot, 3)); };
    // momentum: registers (NF = 32) or the global workspace (NF = 64); coordinate 0 is replicated in every
    // quarter, so its momentum always stays in a private register
    float vreg[V_IN_REGS ? NLOC : 1] = {};
    float v0r = 0.f;
    auto vget = [&](int i) -> float {
      if (i == 0) return v0r;
      if constexpr (V_IN_REGS) return vreg[i]; else return VG(dof(i));
    };
    auto vset = [&](int i, float x) {
      if (i == 0) { v0r = x; return; }
      if constexpr (V_IN_REGS) vreg[i] = x; else VG(dof(i)) = x;
    };
#if TCS_PROFILE
    uint32_t pt[16] = {}, plast = (uint32_t)clock();
#define TCS_TICK(i) { const uint32_t now_ = (uint32_t)clock(); pt[i] += now_ - plast; plast = now_; }
#else
#define TCS_TICK(i)
#endif

    float* mom_next = reinterpret_cast<float*>(smem + K::MOM) + tid;
    const bool mom_ahead = K::MOM_AHEAD && !p.ext_momenta;
    if (mom_ahead) tcs_draw_momenta_call<FPW>(p.seed, gchain, (unsigned int)p.t_begin, fstart, nf, F, mom_next, 3);

    for (int t = 0; t < ARP_RUN(T); ++t) {
      const int tg = p.t_begin + t;
      TCS_TICK(9)
      // current state: global loads issued before the Philox block so that their L2 round trip hides behind it
      float gq0[NLOC], zq0[NLOC];
#pragma unroll
      for (int i = 0; i < NLOC; ++i) {
        gq0[i] = 0.f; zq0[i] = 0.f;
        if (i == 0) { gq0[0] = g0_cur; zq0[0] = z0_cur; }
        else if (owned(i)) { const int d = dof(i); gq0[i] = Gc(d); zq0[i] = Z(d); }
      }
      if (p.ext_momenta) {
        const float* mom = p.ext_momenta + ((size_t)tg * p.C + (valid ? chain : 0)) * D;
#pragma unroll
        for (int i = 0; i < NLOC; ++i)
          if (owned(i)) xs[i * TC_WORKERS] = mom[dof(i)];
      } else if (!K::MOM_AHEAD) {
        tcs_draw_momenta_body<FPW, false>(p.seed, gchain, (unsigned int)tg, fstart, nf, F, xs, 3);
      }
      TCS_TICK(10)
      float ke0 = 0.f, ke1 = 0.f, ke0_tot = 0.f, ke1_tot = 0.f;
      {
#pragma unroll
        for (int i = 0; i < NLOC; ++i) {
          if (owned(i)) {
            const int d = dof(i);
            float vi = mom_ahead ? mom_next[i * TC_WORKERS] : xs[i * TC_WORKERS];
            if (i > 0 || w == 0) ke0 = fmaf(vi, vi, ke0);
            const float e = pe_s[d] * mult;
            vi = vi + 0.5f * e * gq0[i];
            vset(i, vi);
            xs[i * TC_WORKERS] = zq0[i] + e * vi;
          }
        }
      }
      float lpx = 0.f;
      // a coefficient that leaves the fp16 range of the A operand (only on wildly diverging trajectories) would make
      // GEMM1 return inf / NaN and the clamp below would swallow it: such a trajectory is rejected outright
      bool ovf = false;
      TCS_TICK(0)
      for (int l = 0; l < ARP_RUN(L); ++l) {
        const bool last = (l == ARP_RUN(L) - 1);
        float lp_top = 0.f;
        const Site s0 = site_fwd_fast(xs[0], 0.f, ARP_LOG_10, a0, b0, lp_top);
#pragma unroll
        for (int fc = 0; fc < FPW / 8; ++fc) {
          float be[8];
#pragma unroll
          for (int k8 = 0; k8 < 8; ++k8) {
            const int k = 8 * fc + k8;
            be[k8] = 0.f;
            if (k < nf) {
              const int f = fstart + k;
              float dummy = 0.f;
              float ls;
              if (GAMMA) ls = s0.x + xs[(1 + k) * TC_WORKERS];
              else ls = site_fwd_unit(xs[(1 + k) * TC_WORKERS], s0.x, pa_s[1 + f], dummy).x;
              const Site sb = site_fwd_fast(xs[(1 + FPW + k) * TC_WORKERS], 0.f, ls, pa_s[1 + F + f], pb_s[1 + F + f], dummy);
              be[k8] = sb.x * LOG2E;   // GEMM1 then yields log2(e) t eta: one FMUL less per likelihood element
              ovf |= !(fabsf(be[k8]) < 60000.f);   // outside the fp16 range of the A operand (or NaN)
            }
          }
          uint4 hi, lo;
          split_pack(be[0], be[1], hi.x, lo.x);
          split_pack(be[2], be[3], hi.y, lo.y);
          split_pack(be[4], be[5], hi.z, lo.z);
          split_pack(be[6], be[7], hi.w, lo.w);
          *reinterpret_cast<uint4*>(a_row1 + fc * K::SF) = hi;
          *reinterpret_cast<uint4*>(a_row2 + fc * K::SF) = lo;
        }
        fence_async_smem();
        tc_fence_before();
        nb_arrive(TCS_NB_A);
        // the momenta of the next transition do not depend on the state: draw them now, while GEMM1 of chunk 0 runs
        // and every worker of the CTA would otherwise idle (~2.3k clk per step); half of the draw in each of the last
        // two leapfrog steps (all of it in the only step when L = 1)
        {
          const int parts = (last ? 2 : 0) | ((l == ARP_RUN(L) - 2 || ARP_RUN(L) == 1) ? 1 : 0);
          if (mom_ahead && parts && t + 1 < ARP_RUN(T))
            tcs_draw_momenta_call<FPW>(p.seed, gchain, (unsigned int)(tg + 1), fstart, nf, F, mom_next, parts);
        }
        TCS_TICK(1)
        float lik = 0.f;
        for (int c = 0; c < NCH; ++c) {
          const int b = c & 1;
          mbar_wait(bar_h0 + 8 * b, ph[b]); ph[b] ^= 1;
          TCS_TICK(2)
          tc_fence_after();
          const uint32_t h_addr = tmem + lane_off + K::COL_H + b * TC_CHUNK + 32 * w;
          uint32_t hv[32], r1[16], r2[16];
          TC_LD32(h_addr, hv);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          TCS_TICK(3)
          if constexpr (TCS_PACKED && K::RCP_SHARE) {
            if (last) tcs_epilogue32_packed<true>(hv, r1, r2, lik);
            else tcs_epilogue32_packed<false>(hv, r1, r2, lik);
          } else {
            if (last) tcs_epilogue32<K::RCP_SHARE, true>(hv, r1, r2, lik);
            else tcs_epilogue32<K::RCP_SHARE, false>(hv, r1, r2, lik);
          }
          TCS_TICK(14)
          TC_ST16(tmem + lane_off + K::COL_H + b * TC_CHUNK + 32 * w, r1);
          TC_ST16(tmem + lane_off + K::COL_R2 + b * 64 + 16 * w, r2);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          TCS_TICK(15)
          tc_fence_before();
          nb_arrive(TCS_NB_R + b);
          TCS_TICK(4)
        }
        mbar_wait(bar_g, pg); pg ^= 1;
        TCS_TICK(5)
        tc_fence_after();
        uint32_t gv[FPW];
        {
          auto g_load = [&](uint32_t col, uint32_t* dst) {
            const uint32_t ga = tmem + lane_off + col + FPW * w;
            if constexpr (FPW == 8) { TC_LD8(ga, dst); } else { TC_LD16(ga, dst); }
          };
          g_load(K::COL_G, gv);
          if constexpr (K::GSPLIT) {
          // two rounds through one scratch array (head of the odd chunks, then the tails): all three accumulators in
          // flight at once were 16 more live registers at the 96-register cap (stack frame 56 -> 80 B)
          uint32_t gt[FPW];
          const bool two = K::NHEAD == 2 && NCH > 1;     // the odd-chunk accumulator has been written
          if constexpr (K::NHEAD == 2) {
            if (two) g_load(K::COL_G1, gt);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (two) {
#pragma unroll
              for (int k = 0; k < FPW; ++k) gv[k] = __float_as_uint(__uint_as_float(gv[k]) + __uint_as_float(gt[k]));
            }
          }
          g_load(K::COL_GB, gt);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int k = 0; k < FPW; ++k) gv[k] = __float_as_uint(__uint_as_float(gv[k]) + __uint_as_float(gt[k]));
          } else {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          }
        }
        float acc0 = 0.f, lps = 0.f;
        // NF = 64: fetch all momenta before the update loop (inside it every iteration would wait an L2 round trip)
        float vq[V_IN_REGS ? 1 : 2 * FPW];
        if constexpr (!V_IN_REGS) {
#pragma unroll
          for (int k = 0; k < FPW; ++k) {
            vq[k] = 0.f; vq[FPW + k] = 0.f;
            if (k < nf) { vq[k] = VG(dof(1 + k)); vq[FPW + k] = VG(dof(1 + FPW + k)); }
          }
        }
#pragma unroll
        for (int k = 0; k < FPW; ++k) {
          if (k < nf) {
            const int f = fstart + k;
            const float af = pa_s[1 + f], ab_ = pa_s[1 + F + f], bb_ = pb_s[1 + F + f];
            const float xs_s = xs[(1 + k) * TC_WORKERS], xs_b = xs[(1 + FPW + k) * TC_WORKERS];
            Site ss;
            if (GAMMA) {
              ss.x = xs_s;   // the coordinate is the centred value itself
              // log Gamma(1/2, 1/2) density of v: v/2 - e^v/2 + log(1/2)/2 - lgamma(1/2)
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
            float vs = (V_IN_REGS ? vget(1 + k) : vq[V_IN_REGS ? 0 : k]) + 0.5f * es * gs;
            float vb = (V_IN_REGS ? vget(1 + FPW + k) : vq[V_IN_REGS ? 0 : FPW + k]) + 0.5f * eb * gb;
            if (last) {
              ke1 = fmaf(vs, vs, ke1);
              ke1 = fmaf(vb, vb, ke1);
              // proposal gradient / centred values: parked in the global proposal slots until the accept decision
              Gp(1 + f) = gs; Gp(1 + F + f) = gb;
              XCp(1 + f) = ss.x; XCp(1 + F + f) = sb.x;
            } else {
              vs = vs + 0.5f * es * gs;
              vb = vb + 0.5f * eb * gb;
              xs[(1 + k) * TC_WORKERS] = xs_s + es * vs;
              xs[(1 + FPW + k) * TC_WORKERS] = xs_b + eb * vb;
            }
            vset(1 + k, vs);
            vset(1 + FPW + k, vb);
          }
        }
        if (last) {
          // padded rows have h = 0, q = 1/2: -1 each
          lik = 0.69314718055994531f * (lik + (w == 0 ? (float)(NCH * TC_CHUNK - tp.N) : 0.f));
        }
        xch_at(0, w) = acc0;
        xch_at(1, w) = ovf ? -INFINITY : lik + lps;
        xch_at(2, w) = ke0;   // partial kinetic energies ride along (meaningful on the first / last step)
        xch_at(3, w) = ke1;
        TCS_TICK(6)
        epi_bar();
        TCS_TICK(7)
        const float acc0_t = xch_sum(0);
        lpx = xch_sum(1) + lp_top;
        if (l == 0) ke0_tot = xch_sum(2);
        if (last) ke1_tot = xch_sum(3);
        {
          float g0, mb, lb, ab;
          site_rev(s0, acc0_t, 0.f, a0, b0, g0, mb, lb, ab);
          const float e = pe_s[0] * mult;
          float v0 = v0r + 0.5f * e * g0;
          if (last) {
            ke1_tot = fmaf(v0, v0, ke1_tot);   // coordinate 0 is replicated: every quarter adds it itself
            g0_prop = g0;
            if (w == 0) XCp(0) = s0.x;
          } else {
            v0 = v0 + 0.5f * e * g0;
            xs[0] = xs[0] + e * v0;
          }
          v0r = v0;
        }
        par ^= 1;
        TCS_TICK(8)
      }
      float log_alpha = lpx - lp_cur + 0.5f * ke0_tot - 0.5f * ke1_tot;
      if (!(log_alpha == log_alpha) || log_alpha == -INFINITY) log_alpha = -INFINITY;
      float log_u;
      if (p.ext_log_u) log_u = p.ext_log_u[(size_t)tg * p.C + (valid ? chain : 0)];
      else log_u = philox_log_uniform(p.seed, gchain, (unsigned int)tg);
      const bool acc = log_u < log_alpha;
      TCS_TICK(11)
      if (acc) {
        z0_cur = xs[0];
        g0_cur = g0_prop;
#pragma unroll
        for (int i = 0; i < NLOC; ++i)
          if (owned(i) && (i > 0 || w == 0)) Z(dof(i)) = xs[i * TC_WORKERS];
        ocur = oflip - ocur;   // the proposal's gradient / centred values become the current ones
        lp_cur = lpx;
        ++nacc;
      }
      TCS_TICK(12)
      const int t1 = tg + 1;
      if (t1 <= ARP_RUN(num_adapt)) {
        const float ft = (float)t1;
        Hc += p.target_accept - expf(log_alpha < 0.f ? log_alpha : 0.f);
        const float log_step = ARP_LOG_10 - Hc * sqrtf(ft) / ((ft + 10.f) * 0.05f);
        const float eta = powf(ft, -0.75f);
        lavg = eta * log_step + (1.f - eta) * lavg;
        mult = (t1 < ARP_RUN(num_adapt)) ? expf(log_step) : expf(lavg);
      }
      TCS_TICK(13)
      const int since = tg - ARP_RUN(num_burnin);
      if (since >= 0 && (since % p.stride) == 0 && valid) {
        const int s = since / p.stride;
        if (s < ARP_RUN(S)) {
          const size_t o = ((size_t)s * p.C + chain) * D;
          float xq[NLOC], zq[NLOC];
#pragma unroll
          for (int i = 0; i < NLOC; ++i) {
            xq[i] = 0.f; zq[i] = 0.f;
            if (owned(i) && (i > 0 || w == 0)) {
              const int d = dof(i);
              if (ARP_RUN(samples)) xq[i] = XCc(d);
              if (ARP_RUN(samples_orig)) zq[i] = Z(d);
            }
          }
#pragma unroll
          for (int i = 0; i < NLOC; ++i)
            if (owned(i) && (i > 0 || w == 0)) {
              const int d = dof(i);
              if (ARP_RUN(samples)) ARP_RUN(samples)[o + d] = xq[i];
              if (ARP_RUN(samples_orig)) ARP_RUN(samples_orig)[o + d] = zq[i];
            }
          if (ARP_RUN(is_accepted) && w == 0) ARP_RUN(is_accepted)[(size_t)s * p.C + chain] = acc ? 1 : 0;
        }
      }
    }
#if TCS_PROFILE
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 5 || warp == 15))
      printf("tcs warp %d: philox %u kick0 %u fwd %u waitH %u ld %u compute %u st+wait %u arrive %u waitG %u rev %u bar %u top %u | alpha+u %u accept %u adapt %u store %u\n",
             warp, pt[10], pt[0], pt[1], pt[2], pt[3], pt[14], pt[15], pt[4], pt[5], pt[6], pt[7], pt[8], pt[11], pt[12], pt[13], pt[9]);
#endif
    if (ocur != 0) {   // leave the current gradient / centred values in (g, xc), as the workspace contract says
#pragma unroll
      for (int i = 0; i < NLOC; ++i)
        if (owned(i) && (i > 0 || w == 0)) { const int d = dof(i); const float gv_ = Gc(d), xv_ = XCc(d); G(d) = gv_; XC(d) = xv_; }
    }
    if (w == 0) {
      G(0) = g0_cur;   // coordinate 0 lives in registers during the run
      ws.lp[row] = lp_cur; ws.H[row] = Hc; ws.lavg[row] = lavg; ws.mult[row] = mult; ws.nacc[row] = nacc;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TCS_MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)TC_TMEM_COLS) : "memory");
  }
}

// --------------------------------------------------------------------------- host ---
struct GermanTcs {
  DevBuf img;
  int N = 0, F = 0, nf_pad = 0, nchunk = 0;
  bool ok = false;

  // X [N, F] fp32 row-major, y [N] in {0, 1} -> per-chunk stage images (head | tail) of the rows t_n X_n,
  // t_n = 2 y_n - 1, zero padded
  bool build(const float* X, const float* y, int n, int f, std::string* err) {
    ok = false;
    if (f > 64) return true;   // SIMT engine only
    nf_pad = f <= 32 ? 32 : 64;
    nchunk = (n + TC_CHUNK - 1) / TC_CHUNK;
    const uint32_t SG = (uint32_t)(nf_pad / 8) * 128, XCHUNK = (TC_CHUNK / 8) * SG, STAGE = 2 * XCHUNK;
    std::vector<uint8_t> buf((size_t)nchunk * STAGE, 0);
    // K-slot of feature j: the kernels deal the features to the four workers of a chain in balanced contiguous
    // ranges; worker w owns slots [fpw w, fpw (w + 1))
    std::vector<int> slot(f);
    {
      const int fpw = nf_pad / TC_NQ, q = f / TC_NQ, r = f % TC_NQ;
      for (int w = 0, j = 0; w < TC_NQ; ++w)
        for (int k = 0; k < q + (w < r ? 1 : 0); ++k, ++j) slot[j] = w * fpw + k;
    }
    for (int i = 0; i < n; ++i) {
      const int c = i / TC_CHUNK, rloc = i % TC_CHUNK;
      uint8_t* st = buf.data() + (size_t)c * STAGE;
      if (y[i] != 0.f && y[i] != 1.f) return true;   // not a Bernoulli outcome: SIMT engine only
      const float t = 2.f * y[i] - 1.f;
      for (int j = 0; j < f; ++j) {
        const float x = t * X[(size_t)i * f + j];
        if (!(fabsf(x) < 60000.f)) return true;  // outside fp16 range: SIMT engine only
        const __half h1 = __float2half_rn(x);
        const __half h2 = __float2half_rn(x - __half2float(h1));
        const size_t off = (size_t)(rloc / 8) * SG + (size_t)(slot[j] / 8) * 128 + (size_t)(rloc % 8) * 16 + (size_t)(slot[j] % 8) * 2;
        memcpy(st + off, &h1, 2);
        memcpy(st + XCHUNK + off, &h2, 2);
      }
    }
    cudaError_t e = upload(img, buf);
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return false; }
    N = n; F = f; ok = true;
    return true;
  }
  bool ready() const { return ok; }
};

template <int NF, bool GAMMA, bool MULTI>
static inline cudaError_t tcs_launch_one(dim3 grid, cudaStream_t st, const TcsParams& tp, const HmcWs& ws, const HmcArgs& p) {
  cudaError_t e = cudaFuncSetAttribute(k_german_tcs_hmc<NF, GAMMA, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)Tcs<NF>::BYTES);
  if (e != cudaSuccess) return e;
  k_german_tcs_hmc<NF, GAMMA, MULTI><<<grid, TCS_THREADS, Tcs<NF>::BYTES, st>>>(tp, ws, p);
  return cudaGetLastError();
}
template <bool GAMMA>
static inline cudaError_t tcs_launch(int nf_pad, dim3 grid, cudaStream_t st, const TcsParams& tp, const HmcWs& ws, const HmcArgs& p) {
  if (p.slices) return nf_pad == 32 ? tcs_launch_one<32, GAMMA, true>(grid, st, tp, ws, p) : tcs_launch_one<64, GAMMA, true>(grid, st, tp, ws, p);
  return nf_pad == 32 ? tcs_launch_one<32, GAMMA, false>(grid, st, tp, ws, p) : tcs_launch_one<64, GAMMA, false>(grid, st, tp, ws, p);
}

static inline int german_tcs_hmc(GermanTcs& tc, const DevModel& dm, int fp_simt, const HmcArgs& p, const real* z0,
                                 cudaStream_t st, bool want_final, DevBuf* wsbuf, DevBuf* dfz, DevBuf* scal, DevBuf* nacc,
                                 std::atomic<long long>* launches, std::string* err, int n_runs = 1) {
#define TCS_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { *err = std::string(#expr) + ": " + cudaGetErrorString(_e); return 1; } } while (0)
  const long long C = p.C;
  // n_runs > 1: p.slices describes the runs, each owning p.slice_rows = round_up(C, 128) workspace rows
  const long long Cpad = n_runs * ((C + TC_CHAINS - 1) / TC_CHAINS * TC_CHAINS);
  const long long Dpad = (p.D + 7) / 8 * 8;
  const size_t vec = (size_t)Cpad * Dpad;
  TCS_CUDA(wsbuf->alloc(7 * vec * sizeof(real)));
  TCS_CUDA(cudaMemsetAsync(wsbuf->p, 0, 7 * vec * sizeof(real), st));
  TCS_CUDA(scal->alloc(4 * Cpad * sizeof(real)));
  TCS_CUDA(nacc->alloc(Cpad * sizeof(int)));
  HmcWs ws{};
  real* base = wsbuf->as<real>();
  ws.z = base; ws.g = base + vec; ws.xc = base + 2 * vec; ws.x = base + 3 * vec;
  ws.gx = base + 4 * vec; ws.xcx = base + 5 * vec; ws.v = base + 6 * vec;
  real* sb = scal->as<real>();
  ws.mult = sb; ws.lp = sb + Cpad; ws.H = sb + 2 * Cpad; ws.lavg = sb + 3 * Cpad;
  ws.nacc = nacc->as<int>();
  ws.sd = (int)Cpad; ws.sc = 1;
  if (ws.gx - ws.g != ws.xcx - ws.xc) { *err = "german_tcs_hmc: workspace layout"; return 1; }   // buffer-set flip
  const dim3 grid((unsigned)(Cpad / TC_CHAINS));
  const bool gamma = dm.kind == MODEL_GERMAN_GAMMA;
#define TCS_INIT(KIND, FP)                                                                        \
  do {                                                                                            \
    if (p.slices) k_hmc_init<KIND, 1, FP, true><<<grid, ARP_BLOCK, 0, st>>>(dm, ws, p, z0);       \
    else k_hmc_init<KIND, 1, FP, false><<<grid, ARP_BLOCK, 0, st>>>(dm, ws, p, z0);               \
  } while (0)
  if (gamma) {
    if (fp_simt == 32) TCS_INIT(MODEL_GERMAN_GAMMA, 32); else TCS_INIT(MODEL_GERMAN_GAMMA, 64);
  } else {
    if (fp_simt == 32) TCS_INIT(MODEL_GERMAN_LOGNORMAL, 32); else TCS_INIT(MODEL_GERMAN_LOGNORMAL, 64);
  }
#undef TCS_INIT
  launches->fetch_add(1);
  TCS_CUDA(cudaGetLastError());
  TcsParams tp{tc.img.as<uint8_t>(), tc.N, tc.F, tc.nchunk};
  TCS_CUDA(gamma ? tcs_launch<true>(tc.nf_pad, grid, st, tp, ws, p) : tcs_launch<false>(tc.nf_pad, grid, st, tp, ws, p));
  launches->fetch_add(1);
  if (want_final) {
    TCS_CUDA(dfz->alloc((size_t)C * p.D * sizeof(real)));
    const long long n = C * p.D;
    k_gather_ws_tc<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ws.z, ws.sd, ws.sc, (int)C, p.D, dfz->as<real>());
    launches->fetch_add(1);
    TCS_CUDA(cudaGetLastError());
  }
#undef TCS_CUDA
  return 0;
}

}  // namespace arp
