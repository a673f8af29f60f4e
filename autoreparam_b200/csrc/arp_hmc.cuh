// Generic (FP32 / FP64 SIMT) kernels: batched log-joint+gradient and the
// persistent HMC kernel.  One launch runs EVERY transition of EVERY chain:
// Philox momenta, L fused leapfrog steps, Metropolis accept, per-chain dual
// averaging, burn-in / thinning, centred-sample store.
//
// Restates [TFP 0.7] HamiltonianMonteCarlo.one_step / _leapfrog_integrator_one_step
// / MetropolisHastings.one_step / DualAveragingStepSizeAdaptation.one_step /
// sample_chain as called from reference inference.py:218-234.
#pragma once
#include "arp_models.cuh"

namespace arp {

#define ARP_BLOCK 128

// Workspace: per-chain vectors, element (d, c) at base[d * sd + c * sc].
struct HmcWs {
  real *z, *g, *xc;     // current state, its gradient, its centred values
  real *x, *gx, *xcx;   // proposal
  real *v;              // momentum
  real *lp, *H, *lavg, *mult;  // [Cpad] current log-prob and dual-averaging state
  int* nacc;            // [Cpad] accepted-transition counter
  int sd, sc;
};

struct HmcArgs {
  int C, D, L;
  int T;          // transitions in this launch
  int t_begin;    // global index of the first transition
  int num_adapt, num_burnin, stride;  // stride = 1 + num_steps_between_results
  int S;          // kept samples in total
  unsigned long long seed;
  unsigned int chain_offset;
  real target_accept;
  const real* eps0;          // [D]
  const real* a;             // [D]
  const real* b;             // [D]
  const real* ext_momenta;   // [T_total, C, D] or null
  const real* ext_log_u;     // [T_total, C] or null
  real* samples;             // [S, C, D] or null
  real* samples_orig;        // [S, C, D] or null
  unsigned char* is_accepted;  // [S, C] or null
  // several independent runs in ONE launch (the num_leapfrog_steps tuning grid, reference main.py:316-329 run once per
  // L): run i owns workspace rows [i * slice_rows, (i + 1) * slice_rows) -- whole blocks -- and takes its own
  // (L, transitions, adaptation / burn-in lengths, kept samples, step sizes, output buffers) from slices[i]
  const struct HmcSlice* slices;   // null: one run, the fields above
  int slice_rows;
  // streaming statistics of the kept (centred) samples, for runs whose [S, C, D] traces cannot be stored
  // (BASELINE configs[4]: 65 536 chains x 10 003 coordinates): per (chain, coordinate) a pivot (the first kept
  // sample), the sum of y = x - pivot, a ring of the last 2 W values of y (two blocks of W), the first W values, and
  // the W lag products sum_t y_t y_{t-k}, updated once per block of W kept samples (stream_block).  All planes share
  // the workspace layout (element (d, row) at d * sd + row * sc); k_stream_finalize adds the last, partial block and
  // turns the sums into mean / variance / ESS.  W = 0: off.
  int stream_W;
  size_t stream_plane;             // elements per plane
  real* stream_pivot;              // [plane]
  real* stream_sum;                // [plane]
  real* stream_ring;               // [2 W][plane]: the last two blocks of W kept values
  real* stream_head;               // [W][plane]
  real* stream_acc;                // [W][plane]
};

struct HmcSlice {
  int L, T, num_adapt, num_burnin, S;
  const real* eps0;            // [D]
  real* samples;               // [S, C, D] or null
  unsigned char* is_accepted;  // [S, C] or null
};

// Per-block view of the per-run launch arguments of a multi-run launch.  Kernels are templated on MULTI and read a
// field through ARP_RUN(f): the single-run instantiation takes it straight from the kernel parameters (constant-bank
// operands, no registers), the multi-run one from this struct.  (Measured on the tcgen05 kernel: overriding fields of
// the by-value parameter struct made ptxas copy it to local memory, -2 %; holding the per-run fields in registers in
// the single-run path too cost 3 %.)
struct HmcRun {
  int L, T, num_adapt, num_burnin, S;
  int chain;                   // index of workspace row `row` inside its run: keys z0, the RNG streams and the outputs
  const real* eps0;
  real* samples;
  real* samples_orig;
  unsigned char* is_accepted;
};
template <bool MULTI>
__device__ __forceinline__ HmcRun hmc_run_view(const HmcArgs& p, int row) {
  HmcRun r{};
  r.chain = row;
  if constexpr (MULTI) {
    const int sl = row / p.slice_rows;
    const HmcSlice s = p.slices[sl];
    r.L = s.L; r.T = s.T; r.num_adapt = s.num_adapt; r.num_burnin = s.num_burnin; r.S = s.S;
    r.chain = row - sl * p.slice_rows;
    r.eps0 = s.eps0; r.samples = s.samples; r.samples_orig = nullptr; r.is_accepted = s.is_accepted;
  }
  return r;
}
#define ARP_RUN(f) (MULTI ? rv.f : p.f)

// Streaming lag products, one BLOCK of kept samples at a time.  The ring holds the last two blocks of W values of
// y = x - pivot (slots [cb, cb + W) = the block that just completed, [pb, pb + W) = the one before it).  For every lag
// k < W:  acc_k += sum_{t in block} y_t y_{t-k}, with y_{t-k} taken from the previous block where t - k falls there.
// Done per block instead of per kept sample, the W ring values and W lag sums of a coordinate cross HBM once per W
// samples instead of once per sample: (3 W + 6) -> ~6 plane accesses per kept sample and coordinate (the re-reads inside
// the block hit L1).  `len` < W only for the last, partial block (k_stream_finalize).  Returns the block's sum of y.
__device__ __forceinline__ real stream_block(const HmcArgs& p, size_t e, int cb, int len, bool has_prev) {
  const int W = p.stream_W;
  const size_t plane = p.stream_plane;
  const int pb = cb == 0 ? W : 0;
  const real* ring = p.stream_ring;
  real bsum = 0;
  for (int t = 0; t < len; ++t) bsum += ring[(size_t)(cb + t) * plane + e];
  for (int k0 = 0; k0 < W; k0 += 16) {
    real a16[16];
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) a16[kk] = 0;
    for (int t = 0; t < len; ++t) {
      const real yt = ring[(size_t)(cb + t) * plane + e];
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        const int k = k0 + kk, j = t - k;
        if (k < W) {
          if (j >= 0) a16[kk] = fma(yt, ring[(size_t)(cb + j) * plane + e], a16[kk]);
          else if (has_prev) a16[kk] = fma(yt, ring[(size_t)(pb + W + j) * plane + e], a16[kk]);
        }
      }
    }
#pragma unroll
    for (int kk = 0; kk < 16; ++kk)
      if (k0 + kk < W) p.stream_acc[(size_t)(k0 + kk) * plane + e] += a16[kk];
  }
  return bsum;
}

template <int KIND, int LPC, bool WITH_A, int FP>
__global__ void __launch_bounds__(ARP_BLOCK)
k_log_joint_grad(DevModel m, const real* __restrict__ a, const real* __restrict__ b,
                 const real* z, int C, real* lp_out, real* g_scratch, real* xc_scratch, real* abar_scratch,
                 real* bbar_scratch) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int chain = gtid / LPC;
  const int sub = gtid % LPC;
  // padded chains (chain >= C) still run: scratch buffers are padded to the grid
  const size_t off = (size_t)chain * m.D;
  const int zc = chain < C ? chain : C - 1;
  Vec vz{const_cast<real*>(z) + (size_t)zc * m.D, 1};
  Vec vg_{g_scratch + off, 1}, vxc{xc_scratch + off, 1};
  Vec vab{abar_scratch ? abar_scratch + off : nullptr, 1}, vbb{bbar_scratch ? bbar_scratch + off : nullptr, 1};
  real lp = vg<KIND, LPC, WITH_A, FP>(m, a, b, vz, vg_, vxc, vab, vbb, sub, true);
  if (chain < C && sub == 0 && lp_out) lp_out[chain] = lp;
}

template <int KIND, int LPC, int FP, bool MULTI = false>
__global__ void __launch_bounds__(ARP_BLOCK)
k_hmc_init(DevModel m, HmcWs ws, HmcArgs p, const real* z0) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = gtid / LPC;
  const int sub = gtid % LPC;
  const int chain = hmc_run_view<MULTI>(p, row).chain;  // every run of a multi-run launch starts from the same z0
  const bool valid = chain < p.C;
  const size_t co = (size_t)row * ws.sc;
  Vec Z{ws.z + co, ws.sd}, G{ws.g + co, ws.sd}, XC{ws.xc + co, ws.sd};
  for (int d = sub; d < p.D; d += LPC) Z(d) = valid ? z0[(size_t)chain * p.D + d] : (real)0;
  __syncwarp();
  real lp = vg<KIND, LPC, false, FP>(m, p.a, p.b, Z, G, XC, Vec{nullptr, 1}, Vec{nullptr, 1}, sub, true);
  if (sub == 0) {
    ws.lp[row] = lp;
    ws.H[row] = 0;
    ws.lavg[row] = 0;
    ws.mult[row] = 1;
    ws.nacc[row] = 0;
  }
}

#ifndef ARP_RADON_FUSED
#define ARP_RADON_FUSED 1   // 0: the generic sweeps for the radon models too (A/B measurements)
#endif
// Resident blocks per SM ptxas is asked to make room for (register budget = 65536 / (128 threads x blocks)).  Left to
// itself ptxas gave the radon kernels 80 or 128 registers and the electric kernel 168 or 240 depending on unrelated
// details of the surrounding code (measured: radon_synth 3.5e6 vs 2.75e6 grad-evals/s, electric 6.7e8 vs 5.3e8), so the
// occupancy each model needs is stated: six blocks for the light models (HBM-resident or 32 KB of shared-memory
// state per block), three for electric / time_series.
template <int KIND, int LPC>
constexpr int hmc_min_blocks() {
  return (KIND == MODEL_8SCHOOLS || KIND == MODEL_RADON || KIND == MODEL_RADON_STDDVS) ? 6
       : (KIND == MODEL_ELECTION) ? 4
       : (KIND == MODEL_ELECTRIC || KIND == MODEL_TIME_SERIES) ? 3 : 1;
}

template <int KIND, int LPC, int FP, bool MULTI = false>
__global__ void __launch_bounds__(ARP_BLOCK, hmc_min_blocks<KIND, LPC>())
k_hmc_run(DevModel m, HmcWs ws, HmcArgs p, int oc_dpad, int oc_stride) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = gtid / LPC;
  const int sub = gtid % LPC;
  const HmcRun rv = hmc_run_view<MULTI>(p, row);
  const int chain = rv.chain;
  const bool valid = chain < p.C;
  const int D = p.D;
  const size_t co = (size_t)row * ws.sc;
  // The seven state vectors of a chain: z, g, xc (current state, its gradient and centred values), x, gx, xcx (the
  // proposal's) and v, at Bv + k VS + d sd.  The two buffer sets swap roles when a proposal is accepted (`cur` = offset
  // of the current set: 0 or 3 VS), so nothing is copied.
  real* Bv = ws.z + co;
  size_t VS = (size_t)(ws.g - ws.z);   // the host lays the seven vectors out at equal distances
  int sd = ws.sd;
  // The seven state vectors of the block's chains fit in shared memory (oc_dpad > 0): the whole run works on a
  // shared-memory copy -- north_star's "state on chip"; the global workspace is touched at the start and at the end
  // only.  LPC > 1: chain-major (stride 1 inside a vector, oc_stride floats between chains, chosen = LPC mod 32 so
  // that the chains of a warp tile the 32 banks: with an arbitrary stride 70 % of the shared-memory wavefronts of the
  // radon kernel were bank conflicts).  LPC = 1 (D small, e.g. 8schools): coordinate-major, thread t owns column t.
  extern __shared__ __align__(16) unsigned char hmc_smem_raw[];
  // (LPC = 1: compiled for 8schools only -- for time_series, D = 123, 64 on-chip chains per SM were no faster than the
  // HBM-resident layout at full occupancy, and carrying both paths slowed the latter)
  const bool on_chip = oc_dpad > 0 && (LPC > 1 || KIND == MODEL_8SCHOOLS);
  if (on_chip) {
    real* base;
    int osd, vs;   // element stride, vector stride
    if (LPC > 1) { base = reinterpret_cast<real*>(hmc_smem_raw) + (size_t)(threadIdx.x / LPC) * oc_stride; osd = 1; vs = oc_dpad; }
    else { base = reinterpret_cast<real*>(hmc_smem_raw) + threadIdx.x; osd = (int)blockDim.x; vs = oc_dpad * (int)blockDim.x; }
    for (int d = sub; d < D; d += LPC) {
      base[d * osd] = Bv[(size_t)d * sd]; base[vs + d * osd] = Bv[VS + (size_t)d * sd];
      base[2 * vs + d * osd] = Bv[2 * VS + (size_t)d * sd];
    }
    Bv = base; VS = (size_t)vs; sd = osd;
    __syncwarp();
  }
  size_t cur = 0;
  const Vec V{Bv + 6 * VS, sd};
  const real* __restrict__ eps0 = ARP_RUN(eps0);
  real lp_cur = ws.lp[row], Hc = ws.H[row], lavg = ws.lavg[row], mult = ws.mult[row];
  int nacc = ws.nacc[row];
  const unsigned int gchain = p.chain_offset + (unsigned int)chain;

  for (int t = 0; t < ARP_RUN(T); ++t) {
    const int tg = p.t_begin + t;
    const size_t oth = 3 * VS - cur;
    const Vec Z{Bv + cur, sd}, G{Bv + VS + cur, sd}, XC{Bv + 2 * VS + cur, sd};
    const Vec X{Bv + oth, sd}, GX{Bv + VS + oth, sd}, XCX{Bv + 2 * VS + oth, sd};
    // ---- momenta v0 ~ N(0, I); proposal starts at the current state
    real ke0 = 0;
    if (p.ext_momenta) {
      const real* mom = p.ext_momenta + ((size_t)tg * p.C + (valid ? chain : 0)) * D;
      for (int d = sub; d < D; d += LPC) {
        const real v = mom[d];
        V(d) = v;
        ke0 = fma(v, v, ke0);
      }
    } else {
      const int nb = (D + 3) >> 2;
      for (int j = sub; j < nb; j += LPC) {
        real n4[4];
        philox_normal4(p.seed, gchain, (unsigned int)tg, (unsigned int)j, ARP_STREAM_MOMENTUM, n4);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int d = 4 * j + q;
          if (d < D) {
            V(d) = n4[q];
            ke0 = fma(n4[q], n4[q], ke0);
          }
        }
      }
      __syncwarp();
    }
    ke0 = group_sum<LPC>(ke0);
    // ---- L leapfrog steps, two half kicks per step as TFP 0.7 does (v += e g / 2 twice, each rounded separately).
    // Sweeps over the state are what a large model pays for (state in HBM: 10^4 coordinates per chain), so they are
    // fused: the first kick reads the CURRENT state directly (no copy into the proposal buffers), the second half
    // kick of step l and the first of step l + 1 share one sweep, the last half kick only forms the kinetic energy,
    // and an accepted proposal swaps the roles of the two buffer sets instead of being copied.
    real lpx = 0, ke1 = 0;
    {
#pragma unroll 4
      for (int d = sub; d < D; d += LPC) {
        const real e = ldg(eps0 + d) * mult;
        const real v = V(d) + (real)0.5 * e * G(d);
        V(d) = v;
        X(d) = Z(d) + e * v;
      }
    }
    for (int l = 0; l < ARP_RUN(L); ++l) {
      __syncwarp();
      const bool last = (l == ARP_RUN(L) - 1);
      if constexpr (ARP_RADON_FUSED && (KIND == MODEL_RADON || KIND == MODEL_RADON_STDDVS)) {
        // radon: the gradient sweep applies the kicks itself (arp_models.cuh: vg_radon_kick)
        lpx = vg_radon_kick<LPC, KIND == MODEL_RADON_STDDVS>(m, p.a, p.b, X, GX, XCX, V, eps0, mult, sub, last, ke1);
        continue;
      }
      lpx = vg<KIND, LPC, false, FP>(m, p.a, p.b, X, GX, XCX, Vec{nullptr, 1}, Vec{nullptr, 1}, sub, last);
      __syncwarp();
      if (last) {
#pragma unroll 4
        for (int d = sub; d < D; d += LPC) {
          const real e = ldg(eps0 + d) * mult;
          const real v = V(d) + (real)0.5 * e * GX(d);
          ke1 = fma(v, v, ke1);
        }
      } else {
#pragma unroll 4
        for (int d = sub; d < D; d += LPC) {
          const real e = ldg(eps0 + d) * mult;
          const real gx = GX(d);
          real v = V(d) + (real)0.5 * e * gx;
          v = v + (real)0.5 * e * gx;
          V(d) = v;
          X(d) = X(d) + e * v;
        }
      }
    }
    ke1 = group_sum<LPC>(ke1);
    // ---- Metropolis-Hastings
    real log_alpha = lpx - lp_cur + (real)0.5 * ke0 - (real)0.5 * ke1;
    if (!(log_alpha == log_alpha) || log_alpha == -INFINITY) log_alpha = -INFINITY;  // safe_sum: nan -> reject
    real log_u;
    if (p.ext_log_u) log_u = p.ext_log_u[(size_t)tg * p.C + (valid ? chain : 0)];
    else log_u = philox_log_uniform(p.seed, gchain, (unsigned int)tg);
    const bool acc = log_u < log_alpha;
    // an accepted proposal: its buffers become the current state (every lane of the chain takes the same decision)
    const Vec Zc = acc ? X : Z, XCc = acc ? XCX : XC;
    if (acc) {
      cur = oth;
      lp_cur = lpx;
      ++nacc;
    }
    // ---- per-chain dual averaging (TFP defaults: gamma 0.05, t0 10, kappa 0.75)
    const int t1 = tg + 1;
    if (t1 <= ARP_RUN(num_adapt)) {
      const real ft = (real)t1;
      Hc += p.target_accept - r_exp(log_alpha < (real)0 ? log_alpha : (real)0);
      const real log_step = ARP_LOG_10 - Hc * r_sqrt(ft) / ((ft + (real)10) * (real)0.05);
      const real eta = r_pow(ft, (real)-0.75);
      lavg = eta * log_step + ((real)1 - eta) * lavg;
      mult = (t1 < ARP_RUN(num_adapt)) ? r_exp(log_step) : r_exp(lavg);
    }
    // ---- keep every `stride`-th state after burn-in
    const int since = tg - ARP_RUN(num_burnin);
    if (since >= 0 && (since % p.stride) == 0 && valid) {
      const int s = since / p.stride;
      if (s < ARP_RUN(S)) {
        const size_t o = ((size_t)s * p.C + chain) * D;
        if (ARP_RUN(samples)) for (int d = sub; d < D; d += LPC) ARP_RUN(samples)[o + d] = XCc(d);
        if (ARP_RUN(samples_orig)) for (int d = sub; d < D; d += LPC) ARP_RUN(samples_orig)[o + d] = Zc(d);
        if (ARP_RUN(is_accepted) && sub == 0) ARP_RUN(is_accepted)[(size_t)s * p.C + chain] = acc ? 1 : 0;
        if (p.stream_W > 0) {
          const int W = p.stream_W, slot = s % (2 * W);
          const size_t plane = p.stream_plane;
          const bool block_done = (s % W) == W - 1;
          for (int d = sub; d < D; d += LPC) {
            const size_t e = (size_t)d * ws.sd + co;
            const real xv = XCc(d);
            real piv;
            if (s == 0) { piv = xv; p.stream_pivot[e] = xv; } else piv = p.stream_pivot[e];
            const real y = xv - piv;
            p.stream_ring[(size_t)slot * plane + e] = y;
            if (s < W) p.stream_head[(size_t)s * plane + e] = y;
            if (block_done) p.stream_sum[e] += stream_block(p, e, slot - (W - 1), W, s >= W);
          }
        }
      }
    }
    __syncwarp();
  }
  if (on_chip || cur != 0) {   // final state back to (z, g, xc) of the workspace: final_z, and the contract that they describe the chain
    __syncwarp();
    for (int d = sub; d < D; d += LPC) {
      const real zv = Bv[cur + (size_t)d * sd], gv = Bv[VS + cur + (size_t)d * sd], xv = Bv[2 * VS + cur + (size_t)d * sd];
      ws.z[co + (size_t)d * ws.sd] = zv; ws.g[co + (size_t)d * ws.sd] = gv; ws.xc[co + (size_t)d * ws.sd] = xv;
    }
  }
  if (sub == 0) {
    ws.lp[row] = lp_cur;
    ws.H[row] = Hc;
    ws.lavg[row] = lavg;
    ws.mult[row] = mult;
    ws.nacc[row] = nacc;
  }
}

// Streaming statistics -> mean, biased variance and ESS per (chain, coordinate), [C][D] outputs.  With y = x - pivot,
// m = mean(y) and A_k = sum_{t >= k} y_t y_{t-k}:
//   c_k = sum_{t >= k} (y_t - m)(y_{t-k} - m) = A_k - m [(sum - head_k) + (sum - tail_k)] + (S - k) m^2,
// head_k / tail_k = sums of the first / last k kept values -- exactly the centred lag products TFP's
// effective_sample_size forms, so for a series whose first negative autocorrelation lies inside the window the
// result equals arp_ess on the stored trace (up to fp32 round-off); otherwise `truncated` is set and the ESS uses all
// W lags (an upper bound).  Same truncation rule and weights as arp_ess (reference inference.py:240).
__global__ void k_stream_finalize(HmcWs ws, HmcArgs p, int S, real* mean_cd, real* var_cd, real* ess_cd, int* trunc_cd) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)p.C * p.D) return;
  const int c = (int)(i / p.D), d = (int)(i % p.D);
  const size_t e = (size_t)d * ws.sd + (size_t)c * ws.sc, plane = p.stream_plane;
  const int W = p.stream_W;
  // the last, partial block of kept samples (S is not a multiple of W in general)
  const int rem = S % W;
  if (rem > 0) {
    const int s0 = S - rem;
    p.stream_sum[e] += stream_block(p, e, s0 % (2 * W), rem, s0 >= W);
  }
  const double sum = (double)p.stream_sum[e], m = sum / S;
  if (mean_cd) mean_cd[i] = (real)((double)p.stream_pivot[e] + m);
  const int K = W < S ? W : S;
  double head = 0, tail = 0, acov0 = 0, acc = 0;
  bool done = false;
  for (int k = 0; k < K; ++k) {
    if (k > 0) {
      head += (double)p.stream_head[(size_t)(k - 1) * plane + e];
      int sl = (S - k) % (2 * W);                // slot of kept sample S - k
      tail += (double)p.stream_ring[(size_t)sl * plane + e];
    }
    const double ck = (double)p.stream_acc[(size_t)k * plane + e] - m * ((sum - head) + (sum - tail)) + (double)(S - k) * m * m;
    if (k == 0) {
      acov0 = ck / S;
      if (var_cd) var_cd[i] = (real)acov0;
      if (!(acov0 > 0.0)) { if (ess_cd) ess_cd[i] = (real)NAN; if (trunc_cd) trunc_cd[i] = 0; return; }
    }
    const double rho = (ck / (double)(S - k)) / acov0;
    if (rho < 0.0) { done = true; break; }
    acc += (double)(S - k) / S * rho;
  }
  if (ess_cd) ess_cd[i] = (real)((double)S / (-1.0 + 2.0 * acc));
  if (trunc_cd) trunc_cd[i] = (done || K == S) ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------
// Interleaved CP / NCP sampler (--method=i): reference interleaved.py:113-155 + inference.py:258-329.
// The chain state is kept in the centred space.  One transition = for rule r in (A, B): map the centred
// state to rule r's coordinates (to_rule), re-bootstrap (log-prob, gradient) there -- one extra gradient
// evaluation, as Interleaved.one_step does with bootstrap_results -- take one HMC step of L_r leapfrog
// steps, adapt that rule's step size with [TFP 0.7] SimpleStepSizeAdaptation (x (1 + rate) if the
// acceptance probability exceeds the target, / (1 + rate) otherwise, while step < num_adaptation_steps),
// and map back to the centred space (free: vg returns the centred values).  2 L + 2 gradient evaluations.
struct IlvArgs {
  const real* a2;            // rule B parameters (rule A = p.a, p.b)
  const real* b2;
  const real* eps0_2;        // [D] base step sizes of rule B (rule A = p.eps0)
  int L2;
  real rate;                 // adaptation_rate (0.05)
  unsigned char* is_accepted2;   // [S, C] accepts of the rule-B sub-step (rule A -> p.is_accepted)
  real* mult2;               // [Cpad] rule-B step multiplier
  int* nacc2;
};

template <int KIND, int LPC, int FP>
__global__ void __launch_bounds__(ARP_BLOCK)
k_hmc_interleaved(DevModel m, HmcWs ws, HmcArgs p, IlvArgs q, const real* x0) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int chain = gtid / LPC;
  const int sub = gtid % LPC;
  const bool valid = chain < p.C;
  const int D = p.D;
  const size_t co = (size_t)chain * ws.sc;
  Vec Z{ws.z + co, ws.sd}, G{ws.g + co, ws.sd}, XC{ws.xc + co, ws.sd};
  Vec X{ws.x + co, ws.sd}, GX{ws.gx + co, ws.sd}, XCX{ws.xcx + co, ws.sd};
  Vec V{ws.v + co, ws.sd};
  // the initial state is given in the centred space (initial_states_cp, inference.py:265)
  for (int d = sub; d < D; d += LPC) XC(d) = valid ? x0[(size_t)chain * D + d] : (real)0;
  __syncwarp();
  real mult[2] = {1, 1};
  int nacc[2] = {0, 0};
  const unsigned int gchain = p.chain_offset + (unsigned int)chain;

  for (int t = 0; t < p.T; ++t) {
    const int tg = p.t_begin + t;
    bool acc_r[2] = {false, false};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const real* a = r == 0 ? p.a : q.a2;
      const real* b = r == 0 ? p.b : q.b2;
      const real* eps0 = r == 0 ? p.eps0 : q.eps0_2;
      const int L = r == 0 ? p.L : q.L2;
      const unsigned int sid = 2u * (unsigned int)tg + (unsigned int)r;   // RNG / injected-stream index
      // ---- centred -> rule coordinates, re-bootstrap
      to_rule<KIND, LPC>(m, a, b, XC, Z, sub);
      __syncwarp();
      real lp_cur = vg<KIND, LPC, false, FP>(m, a, b, Z, G, XC, Vec{nullptr, 1}, Vec{nullptr, 1}, sub, true);
      __syncwarp();
      // ---- one HMC step (same op order as k_hmc_run)
      real ke0 = 0;
      if (p.ext_momenta) {
        const real* mom = p.ext_momenta + ((size_t)sid * p.C + (valid ? chain : 0)) * D;
        for (int d = sub; d < D; d += LPC) { const real v = mom[d]; V(d) = v; ke0 = fma(v, v, ke0); }
      } else {
        const int nb = (D + 3) >> 2;
        for (int j = sub; j < nb; j += LPC) {
          real n4[4];
          philox_normal4(p.seed, gchain, sid, (unsigned int)j, ARP_STREAM_MOMENTUM, n4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int d = 4 * j + k;
            if (d < D) { V(d) = n4[k]; ke0 = fma(n4[k], n4[k], ke0); }
          }
        }
        __syncwarp();
      }
      ke0 = group_sum<LPC>(ke0);
      for (int d = sub; d < D; d += LPC) { X(d) = Z(d); GX(d) = G(d); }
      real lpx = 0, ke1 = 0;
      for (int l = 0; l < L; ++l) {
        for (int d = sub; d < D; d += LPC) {
          const real e = ldg(eps0 + d) * mult[r];
          const real v = V(d) + (real)0.5 * e * GX(d);
          V(d) = v;
          X(d) = X(d) + e * v;
        }
        __syncwarp();
        const bool last = (l == L - 1);
        lpx = vg<KIND, LPC, false, FP>(m, a, b, X, GX, XCX, Vec{nullptr, 1}, Vec{nullptr, 1}, sub, last);
        __syncwarp();
        for (int d = sub; d < D; d += LPC) {
          const real e = ldg(eps0 + d) * mult[r];
          const real v = V(d) + (real)0.5 * e * GX(d);
          V(d) = v;
          if (last) ke1 = fma(v, v, ke1);
        }
      }
      ke1 = group_sum<LPC>(ke1);
      real log_alpha = lpx - lp_cur + (real)0.5 * ke0 - (real)0.5 * ke1;
      if (!(log_alpha == log_alpha) || log_alpha == -INFINITY) log_alpha = -INFINITY;
      real log_u;
      if (p.ext_log_u) log_u = p.ext_log_u[(size_t)sid * p.C + (valid ? chain : 0)];
      else log_u = philox_log_uniform(p.seed, gchain, sid);
      const bool acc = log_u < log_alpha;
      acc_r[r] = acc;
      if (acc) {
        for (int d = sub; d < D; d += LPC) XC(d) = XCX(d);   // only the centred state is carried over
        ++nacc[r];
      }
      if (tg < p.num_adapt) {
        const real one_plus = (real)1 + q.rate;
        const real pacc = r_exp(log_alpha < (real)0 ? log_alpha : (real)0);
        mult[r] = pacc > p.target_accept ? mult[r] * one_plus : mult[r] / one_plus;
      }
      __syncwarp();
    }
    const int since = tg - p.num_burnin;
    if (since >= 0 && (since % p.stride) == 0 && valid) {
      const int s = since / p.stride;
      if (s < p.S) {
        const size_t o = ((size_t)s * p.C + chain) * D;
        if (p.samples) for (int d = sub; d < D; d += LPC) p.samples[o + d] = XC(d);
        if (sub == 0) {
          if (p.is_accepted) p.is_accepted[(size_t)s * p.C + chain] = acc_r[0] ? 1 : 0;
          if (q.is_accepted2) q.is_accepted2[(size_t)s * p.C + chain] = acc_r[1] ? 1 : 0;
        }
      }
    }
    __syncwarp();
  }
  if (sub == 0) {
    ws.mult[chain] = mult[0]; q.mult2[chain] = mult[1];
    ws.nacc[chain] = nacc[0]; q.nacc2[chain] = nacc[1];
  }
}

}  // namespace arp
