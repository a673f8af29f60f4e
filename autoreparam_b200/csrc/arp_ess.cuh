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
// One thread per series; consecutive threads own consecutive (c, d) so every
// load samples[t][i] is coalesced.  The mean is accumulated in double; the lag
// products in `real` with one partial sum per 32-sample block folded into a
// double total (fp32 build: error ~1e-6 of the lag-0 term, far below the
// Monte-Carlo error of an ESS estimate; fp64 build: exact to round-off).
#pragma once
#include "arp_common.cuh"

namespace arp {

#define ARP_ESS_BLOCK 128
#define ARP_ESS_W 32

__global__ void __launch_bounds__(ARP_ESS_BLOCK)
k_ess(const real* __restrict__ x, int S, long long n, real* __restrict__ ess, real* __restrict__ mean_out,
      real* __restrict__ var_out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const real* xs = x + i;
  double mean = 0;
  for (int t = 0; t < S; ++t) mean += (double)xs[(size_t)t * n];
  mean /= S;
  if (mean_out) mean_out[i] = (real)mean;

  double sum = 0;       // sum_k (S-k)/S rho_k over the kept lags
  double acov0 = 0;
  bool done = false;
  for (int k0 = 0; k0 < S && !done; k0 += ARP_ESS_W) {
    double acc[ARP_ESS_W];
    real ring[ARP_ESS_W];
#pragma unroll
    for (int w = 0; w < ARP_ESS_W; ++w) { acc[w] = 0; ring[w] = 0; }
    real part[ARP_ESS_W];
    const int ns = S - k0;  // pairs (s + k0, s - w), s = 0 .. ns-1
    for (int s0 = 0; s0 < ns; s0 += ARP_ESS_W) {
#pragma unroll
      for (int w = 0; w < ARP_ESS_W; ++w) part[w] = 0;
#pragma unroll
      for (int u = 0; u < ARP_ESS_W; ++u) {
        const int s = s0 + u;
        real past = 0, pres = 0;
        if (s < ns) {
          past = (real)((double)xs[(size_t)s * n] - mean);
          pres = (k0 == 0) ? past : (real)((double)xs[(size_t)(s + k0) * n] - mean);
        }
        ring[u] = past;
#pragma unroll
        for (int w = 0; w < ARP_ESS_W; ++w)
          part[w] = fma(pres, ring[(u - w + ARP_ESS_W) % ARP_ESS_W], part[w]);
      }
#pragma unroll
      for (int w = 0; w < ARP_ESS_W; ++w) acc[w] += (double)part[w];
    }
    if (k0 == 0) {
      acov0 = acc[0] / S;
      if (var_out) var_out[i] = (real)acov0;
      if (!(acov0 > 0.0)) {  // constant (or non-finite) series: TFP yields NaN
        ess[i] = (real)NAN;
        return;
      }
    }
#pragma unroll
    for (int w = 0; w < ARP_ESS_W; ++w) {
      const int k = k0 + w;
      if (!done && k < S) {
        const double rho = (acc[w] / (double)(S - k)) / acov0;
        if (rho < 0.0) done = true;           // filter_threshold = 0: this lag and all later ones are zeroed
        else sum += (double)(S - k) / S * rho;  // NaN (constant series) propagates, as in TFP
      }
    }
  }
  ess[i] = (real)((double)S / (-1.0 + 2.0 * sum));
}

}  // namespace arp
