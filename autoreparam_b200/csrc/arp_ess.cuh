// Effective sample size, per (chain, coordinate) series.
//
// Restates [TFP 0.7] tfp.mcmc.effective_sample_size(states, filter_threshold=0)
// as called at reference inference.py:240,327:
//   rho_k = (sum_t x_t x_{t+k} / (S-k)) / (sum_t x_t^2 / S)   (centred x)
//   every lag from the first rho_k < 0 on is zeroed
//   ESS = S / (-1 + 2 sum_k (S-k)/S rho_k)
// TFP gets the autocovariances from an FFT of length >= 2S; because everything
// from the first negative lag on is discarded, only lags 0..K (K = first
// negative lag, typically tens) are ever used, so this kernel computes exactly
// those by direct summation, ARP_ESS_W lags per pass, and stops at the first
// negative one: O(S*K) instead of O(S log S) + a [2S] complex buffer per series.
//
// One thread per series.  The input is first transposed [S][C][D] -> [S][D][C]
// (k_ess_transpose) so that the 32 lanes of a warp own the SAME coordinate of 32
// consecutive chains: their autocorrelation lengths are similar, so the
// data-dependent number of passes barely diverges inside a warp, and every load
// xT[t][d][c] is coalesced.  The mean is accumulated in double; the lag
// products in `real` with one partial sum per 32-sample block folded into a
// double total (fp32 build: error ~1e-6 of the lag-0 term, far below the
// Monte-Carlo error of an ESS estimate; fp64 build: exact to round-off).
#pragma once
#include "arp_common.cuh"

namespace arp {

#define ARP_ESS_BLOCK 128
#define ARP_ESS_W 32
#define ARP_ESS_FOLD 8   // blocks of ARP_ESS_W samples accumulated in `real` before folding into double

// [S][n] -> [n][S] (series-major), 32 x 32 tiles through shared memory.  With the series
// contiguous, a thread walks 4*S bytes of one page instead of striding 4*n bytes (3.3 MB in the
// bench) per sample, which thrashed the TLB (ncu: 12 % issue-active, long_scoreboard 12.6).
__global__ void k_ess_transpose(const real* __restrict__ x, int S, long long n, real* __restrict__ xt) {
  __shared__ real tile[32][33];
  const long long i0 = (long long)blockIdx.x * 32;
  for (int t0 = blockIdx.y * 32; t0 < S; t0 += gridDim.y * 32) {
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      const int t = t0 + r;
      const long long i = i0 + threadIdx.x;
      if (t < S && i < n) tile[r][threadIdx.x] = x[(size_t)t * n + i];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      const long long i = i0 + r;
      const int t = t0 + threadIdx.x;
      if (t < S && i < n) xt[(size_t)i * S + t] = tile[threadIdx.x][r];
    }
    __syncthreads();
  }
}

// x is [n][S] with n = C*D series in [c][d] order.  Thread j owns series (d = j / C, c = j % C) so the
// lanes of a warp own the same coordinate of 32 consecutive chains (similar autocorrelation lengths).
// VEC: S is a multiple of 4 (series are then 16-byte aligned: fp32 build only), samples are fetched four at a time.
// Every lane walks its own series, so a warp-wide load touches 32 different sectors whatever its width: with scalar
// loads the kernel sat on the L1 wavefront rate (ncu: 17 sectors per request, 8.6e9 sectors for 836k series x 1000
// samples); 128-bit loads need a quarter of the requests for the same bytes.
template <bool VEC>
__global__ void __launch_bounds__(ARP_ESS_BLOCK)
k_ess(const real* __restrict__ x, int S, int C, int D, real* __restrict__ ess_cd, real* __restrict__ mean_cd,
      real* __restrict__ var_cd) {
  const long long ntot = (long long)C * D;
  const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (j >= ntot) return;
  const long long i = (j % C) * D + j / C;   // series / output index [c][d]
  real* ess = ess_cd + i;
  real* mean_out = mean_cd ? mean_cd + i : nullptr;
  real* var_out = var_cd ? var_cd + i : nullptr;
  const size_t n = 1;                         // stride between consecutive samples of a series
  const real* xs = x + (size_t)i * S;
  double mean = 0;
  if constexpr (VEC && !ARP_REAL_IS_DOUBLE) {
    const float4* x4 = reinterpret_cast<const float4*>(xs);
    for (int t = 0; t < S / 4; ++t) {
      const float4 v = x4[t];
      mean += (double)((v.x + v.y) + (v.z + v.w));
    }
  } else {
    for (int t = 0; t < S; ++t) mean += (double)xs[(size_t)t * n];
  }
  mean /= S;
  if (mean_out) *mean_out = (real)mean;
  // centring in `real` with a two-term mean (hi + lo) keeps the conversions off the XU pipe
  const real mean_hi = (real)mean, mean_lo = (real)(mean - (double)mean_hi);

  double sum = 0;       // sum_k (S-k)/S rho_k over the kept lags
  double acov0 = 0;
  bool done = false;
  for (int k0 = 0; k0 < S && !done; k0 += ARP_ESS_W) {
    // (double totals in shared memory / a register cap for 4-5 resident blocks per SM were measured: no gain,
    // the kernel is not occupancy-bound; 64 lags per pass to halve the re-streaming of the series: 255 registers
    // with spills, 23.4 -> 32.3 ms)
    double acc[ARP_ESS_W];
    real ring[ARP_ESS_W], part[ARP_ESS_W];
#pragma unroll
    for (int w = 0; w < ARP_ESS_W; ++w) { acc[w] = 0; ring[w] = 0; part[w] = 0; }
    const int ns = S - k0;  // pairs (s + k0, s - w), s = 0 .. ns-1
    int fold = 0;
    for (int s0 = 0; s0 < ns; s0 += ARP_ESS_W) {
      if constexpr (VEC && !ARP_REAL_IS_DOUBLE) {
        // ns = S - k0 is a multiple of 4 (k0 is a multiple of 32): whole float4 groups are in or out of range
        const float4* p4 = reinterpret_cast<const float4*>(xs + s0);
        const float4* q4 = reinterpret_cast<const float4*>(xs + s0 + k0);
#pragma unroll
        for (int u4 = 0; u4 < ARP_ESS_W / 4; ++u4) {
          float4 pv = make_float4(mean_hi, mean_hi, mean_hi, mean_hi), qv = pv;   // out of range -> centred value 0
          const bool in = s0 + 4 * u4 < ns;
          if (in) { pv = p4[u4]; qv = (k0 == 0) ? pv : q4[u4]; }
          const float lo = in ? mean_lo : 0.f;
          const float pa[4] = {(pv.x - mean_hi) - lo, (pv.y - mean_hi) - lo, (pv.z - mean_hi) - lo, (pv.w - mean_hi) - lo};
          const float pr[4] = {(qv.x - mean_hi) - lo, (qv.y - mean_hi) - lo, (qv.z - mean_hi) - lo, (qv.w - mean_hi) - lo};
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const int u = 4 * u4 + v;
            ring[u] = pa[v];
#pragma unroll
            for (int w = 0; w < ARP_ESS_W; ++w)
              part[w] = fma(pr[v], ring[(u - w + ARP_ESS_W) % ARP_ESS_W], part[w]);
          }
        }
      } else {
#pragma unroll
      for (int u = 0; u < ARP_ESS_W; ++u) {
        const int s = s0 + u;
        real past = 0, pres = 0;
        if (s < ns) {
          past = (xs[(size_t)s * n] - mean_hi) - mean_lo;
          pres = (k0 == 0) ? past : (xs[(size_t)(s + k0) * n] - mean_hi) - mean_lo;
        }
        ring[u] = past;
#pragma unroll
        for (int w = 0; w < ARP_ESS_W; ++w)
          part[w] = fma(pres, ring[(u - w + ARP_ESS_W) % ARP_ESS_W], part[w]);
      }
      }
      if (++fold == ARP_ESS_FOLD) {   // fold the `real` partial sums into the double totals
        fold = 0;
#pragma unroll
        for (int w = 0; w < ARP_ESS_W; ++w) { acc[w] += (double)part[w]; part[w] = 0; }
      }
    }
#pragma unroll
    for (int w = 0; w < ARP_ESS_W; ++w) acc[w] += (double)part[w];
    if (k0 == 0) {
      acov0 = acc[0] / S;
      if (var_out) *var_out = (real)acov0;
      if (!(acov0 > 0.0)) {  // constant (or non-finite) series: TFP yields NaN
        *ess = (real)NAN;
        return;
      }
    }
#pragma unroll
    for (int w = 0; w < ARP_ESS_W; ++w) {
      const int k = k0 + w;
      if (!done && k < S) {
        const double rho = (acc[w] / (double)(S - k)) / acov0;
        if (rho < 0.0) done = true;           // filter_threshold = 0: this lag and all later ones are zeroed
        else sum += (double)(S - k) / S * rho;
      }
    }
  }
  *ess = (real)((double)S / (-1.0 + 2.0 * sum));
}

}  // namespace arp
