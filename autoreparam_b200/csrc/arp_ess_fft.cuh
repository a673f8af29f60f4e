// Effective sample size through a shared-memory FFT: the autocovariance of EVERY lag in O(N log N) per series.
//
// Restates [TFP 0.7] tfp.mcmc.effective_sample_size(states, filter_threshold=0) (reference inference.py:240,327)
// the way TFP computes it -- zero-pad the centred series to N >= 2S, power spectrum, inverse transform:
//   c_k = sum_t x_t x_{t+k},  rho_k = (c_k / (S - k)) / (c_0 / S),  every lag from the first rho_k < 0 on is dropped,
//   ESS = S / (-1 + 2 sum_k (S - k)/S rho_k).
// The direct method of arp_ess.cuh costs S K multiply-adds per series (K = first negative lag, ~400 on slowly mixing
// chains: 23 ms for 836k series x 1000 samples, 7 x the algorithmic HBM traffic).  Here:
//   * a CTA owns ARP_FFT_G adjacent series of the [S][C*D] sample array and reads them straight from it, row by row
//     (ARP_FFT_G floats = whole 32-byte sectors per row): the samples cross HBM exactly once and no transposed copy
//     exists;
//   * two REAL series ride in one COMPLEX transform (z = x1 + i x2; X1 = (Z_k + conj Z_{N-k}) / 2, X2 = (Z_k - conj
//     Z_{N-k}) / 2i), their power spectra are packed again as P1 + i P2, and because a power spectrum is real and even
//     its inverse transform equals its forward transform: ONE routine, c1 = Re, c2 = Im of the second pass;
//   * N = 2048 = 16 x 16 x 8: three Stockham autosort passes, one radix-16 / radix-8 butterfly per thread per pass,
//     data in shared memory between passes (padded against bank conflicts), 128 threads per transform.
// S <= 1024 only (N = 2048 covers every lag of the linear autocovariance); longer series take the direct kernel.
#pragma once
#include "arp_common.cuh"

namespace arp {

#define ARP_FFT_N 2048
#define ARP_FFT_MAXS 1024
#define ARP_FFT_G 8                       // series per CTA (4 complex transforms)
#define ARP_FFT_TPF 128                   // threads per transform
#define ARP_FFT_THREADS (ARP_FFT_TPF * ARP_FFT_G / 2)
#ifndef ARP_FFT_MINB
#define ARP_FFT_MINB 2                    // resident CTAs per SM the register allocation aims at
#endif
#define ARP_FFT_PAD(i) ((i) + ((i) >> 4)) // one padding element per 16: the stride-16 stores of the first pass spread over all banks
#define ARP_FFT_BUF (ARP_FFT_N + ARP_FFT_N / 16 + 4)   // + 4: the four transforms of a CTA start 8 banks apart

template <typename T> struct cplx { T x, y; };
template <typename T> __host__ __device__ __forceinline__ cplx<T> cmul(cplx<T> a, cplx<T> b) {
  return cplx<T>{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <typename T> __host__ __device__ __forceinline__ cplx<T> cadd(cplx<T> a, cplx<T> b) { return cplx<T>{a.x + b.x, a.y + b.y}; }
template <typename T> __host__ __device__ __forceinline__ cplx<T> csub(cplx<T> a, cplx<T> b) { return cplx<T>{a.x - b.x, a.y - b.y}; }
// multiplication by -i (forward transform: e^{-i pi/2})
template <typename T> __host__ __device__ __forceinline__ cplx<T> cmul_mi(cplx<T> a) { return cplx<T>{a.y, -a.x}; }

// (Measured and not kept: complex arithmetic on the packed FP32x2 instructions of sm_100 (FADD2 / FMUL2 / FFMA2 through
// __fadd2_rn / __fmul2_rn / __ffma2_rn).  592 packed instructions replaced 1330 scalar ones, but ptxas needs aligned
// register pairs: +190 MOV and spills at the 64-register budget; 6.50 ms against 5.38 ms for the scalar code.)

// exp(-2 pi i num / den)
template <typename T> __host__ __device__ __forceinline__ cplx<T> twiddle(int num, int den) {
#ifdef __CUDA_ARCH__
  T s, c;
  const T scale = (T)(-2.0) / (T)den;          // den is a power of two: exact, and a constant after inlining
  if constexpr (sizeof(T) == 8) sincospi((T)num * scale, &s, &c);
  else sincospif((T)num * scale, &s, &c);
  return cplx<T>{c, s};
#else
  const double ang = -2.0 * 3.14159265358979323846 * (double)num / (double)den;
  return cplx<T>{(T)cos(ang), (T)sin(ang)};
#endif
}

// in-place 4-point DFT (forward), natural order out
template <typename T> __host__ __device__ __forceinline__ void dft4(cplx<T>& a0, cplx<T>& a1, cplx<T>& a2, cplx<T>& a3) {
  const cplx<T> s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = cmul_mi(csub(a1, a3));
  a0 = cadd(s02, s13); a2 = csub(s02, s13);
  a1 = cadd(d02, d13); a3 = csub(d02, d13);
}
template <typename T> __host__ __device__ __forceinline__ void dft2(cplx<T>& a0, cplx<T>& a1) {
  const cplx<T> s = cadd(a0, a1), d = csub(a0, a1);
  a0 = s; a1 = d;
}

// R-point DFT of v[0..R), natural order in and out; R = 16 (4 x 4) or 8 (2 x 4).
// n = R1 n2 + n1 (n1 < R1), k = k1 R2' ... written out for the two radices to keep everything in registers.
template <typename T> __host__ __device__ __forceinline__ void dft16(cplx<T>* v) {
  // step 1: four 4-point DFTs over n2 (stride 4) for each n1
#pragma unroll
  for (int n1 = 0; n1 < 4; ++n1) dft4(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);
  // v[n1 + 4 k2] now holds the k2-th output of column n1; twiddle by W16^(n1 k2)
  const T c1 = (T)0.92387953251128673848, s1 = (T)0.38268343236508978178, h = (T)0.70710678118654752440;
  const cplx<T> w1{c1, -s1}, w2{h, -h}, w3{s1, -c1}, w6{-h, -h}, w9{-c1, s1};
  v[1 + 4] = cmul(v[1 + 4], w1);  v[1 + 8] = cmul(v[1 + 8], w2);  v[1 + 12] = cmul(v[1 + 12], w3);
  v[2 + 4] = cmul(v[2 + 4], w2);  v[2 + 8] = cmul_mi(v[2 + 8]);   v[2 + 12] = cmul(v[2 + 12], w6);
  v[3 + 4] = cmul(v[3 + 4], w3);  v[3 + 8] = cmul(v[3 + 8], w6);  v[3 + 12] = cmul(v[3 + 12], w9);
  // step 2: four 4-point DFTs over n1 for each k2: output k = k2 + 4 k1 lands in v[4 k2 + k1]
#pragma unroll
  for (int k2 = 0; k2 < 4; ++k2) dft4(v[4 * k2], v[4 * k2 + 1], v[4 * k2 + 2], v[4 * k2 + 3]);
  // v[4 k2 + k1] = X[k2 + 4 k1]: transpose to natural order
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i + 1; j < 4; ++j) { const cplx<T> t = v[4 * i + j]; v[4 * i + j] = v[4 * j + i]; v[4 * j + i] = t; }
}

template <typename T> __host__ __device__ __forceinline__ void dft8(cplx<T>* v) {
  // n = 2 n2 + n1 (n1 < 2, n2 < 4): 4-point DFTs over n2 for each n1
  dft4(v[0], v[2], v[4], v[6]);
  dft4(v[1], v[3], v[5], v[7]);
  // v[n1 + 2 k2]; twiddle the n1 = 1 column by W8^k2
  const T h = (T)0.70710678118654752440;
  v[3] = cmul(v[3], cplx<T>{h, -h});
  v[5] = cmul_mi(v[5]);
  v[7] = cmul(v[7], cplx<T>{-h, -h});
  // 2-point DFTs over n1: X[k2 + 4 k1]
  dft2(v[0], v[1]); dft2(v[2], v[3]); dft2(v[4], v[5]); dft2(v[6], v[7]);
  // v[2 k2 + k1] = X[k2 + 4 k1] -> natural order
  const cplx<T> x0 = v[0], x4 = v[1], x1 = v[2], x5 = v[3], x2 = v[4], x6 = v[5], x3 = v[6], x7 = v[7];
  v[0] = x0; v[1] = x1; v[2] = x2; v[3] = x3; v[4] = x4; v[5] = x5; v[6] = x6; v[7] = x7;
}

// One Stockham autosort pass of radix R over a length-N transform held in buf (padded indexing): butterfly j of N / R.
// Ns = product of the radices of the earlier passes.  load: v[r] = x[j + r N/R] w^r, w = exp(-2 pi i (j mod Ns) / (Ns R));
// store: y[(j / Ns) Ns R + (j mod Ns) + r Ns] = DFT_R(v)[r].   (Govindaraju et al., "High performance discrete Fourier
// transforms on graphics processors", SC'08 -- the published algorithm; all reads of a pass precede its writes.)
template <typename T, int R>
__host__ __device__ __forceinline__ void stockham_load(const cplx<T>* buf, int j, int Ns, cplx<T>* v) {
  const int k = j % Ns;
#pragma unroll
  for (int r = 0; r < R; ++r) v[r] = buf[ARP_FFT_PAD(j + r * (ARP_FFT_N / R))];
  if (Ns > 1) {
    const cplx<T> w = twiddle<T>(k, Ns * R);
    cplx<T> wr = w;
#pragma unroll
    for (int r = 1; r < R; ++r) {
      v[r] = cmul(v[r], wr);
      wr = cmul(wr, w);
    }
  }
}
template <typename T, int R>
__host__ __device__ __forceinline__ void stockham_store(cplx<T>* buf, int j, int Ns, const cplx<T>* v) {
  const int k = j % Ns, base = (j / Ns) * Ns * R + k;
#pragma unroll
  for (int r = 0; r < R; ++r) buf[ARP_FFT_PAD(base + r * Ns)] = v[r];
}

#ifdef __CUDACC__
// named barrier among the ARP_FFT_TPF threads of one transform (ids 1 .. G/2)
__device__ __forceinline__ void fft_bar(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(ARP_FFT_TPF) : "memory"); }

// forward transform of buf (length ARP_FFT_N, padded indexing) by the 128 threads of one group, in place.
// FIRST: the input is the raw pair of series (x1 + i x2), S <= N / 2 samples; centring (mean m1 / m2, split hi + lo),
// scaling to unit variance (s1 / s2) and the zero padding are applied while the first pass loads its operands, so the
// padding never exists in shared memory and only the first 8 of the 16 operands are read at all.
// HALF_OUT: only outputs k < N / 2 are stored by the last pass (the lags 0 .. S - 1 of the second transform).
template <typename T, bool FIRST, bool HALF_OUT>
__device__ __forceinline__ void fft2048(cplx<T>* buf, int t, int bar_id, int S, T m1h, T m1l, T s1, T m2h, T m2l, T s2) {
  cplx<T> v[16];
  // pass 1: radix 16, Ns = 1 (no twiddles)
  if constexpr (FIRST) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int idx = t + r * (ARP_FFT_N / 16);
      v[r] = cplx<T>{0, 0};
      if (r < 8 && idx < S) {                       // S <= N / 2: operands 8 .. 15 are padding
        const cplx<T> raw = buf[ARP_FFT_PAD(idx)];
        v[r] = cplx<T>{((raw.x - m1h) - m1l) * s1, ((raw.y - m2h) - m2l) * s2};
      }
    }
  } else {
    stockham_load<T, 16>(buf, t, 1, v);
  }
  dft16(v);
  fft_bar(bar_id);
  stockham_store<T, 16>(buf, t, 1, v);
  fft_bar(bar_id);
  // pass 2: radix 16, Ns = 16
  stockham_load<T, 16>(buf, t, 16, v);
  dft16(v);
  fft_bar(bar_id);
  stockham_store<T, 16>(buf, t, 16, v);
  fft_bar(bar_id);
  // pass 3: radix 8, Ns = 256: 256 butterflies, two per thread
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    stockham_load<T, 8>(buf, t + h * ARP_FFT_TPF, 256, v + 8 * h);
    dft8(v + 8 * h);
  }
  fft_bar(bar_id);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = t + h * ARP_FFT_TPF;               // Ns = 256 = N / 8: output index j + 256 r
#pragma unroll
    for (int r = 0; r < (HALF_OUT ? 4 : 8); ++r) buf[ARP_FFT_PAD(j + r * 256)] = v[8 * h + r];
  }
  fft_bar(bar_id);
}

// x: [S][n] samples (n = C * D series); ess / mean / var: [n].
template <typename T>
__global__ void __launch_bounds__(ARP_FFT_THREADS, ARP_FFT_MINB)
k_ess_fft(const T* __restrict__ x, int S, long long n, T* __restrict__ ess, T* __restrict__ mean_out, T* __restrict__ var_out) {
  extern __shared__ __align__(16) unsigned char ess_smem[];
  cplx<T>* bufs = reinterpret_cast<cplx<T>*>(ess_smem);                                   // [G/2][ARP_FFT_BUF]
  double* red = reinterpret_cast<double*>(ess_smem + sizeof(cplx<T>) * (ARP_FFT_G / 2) * ARP_FFT_BUF);   // [2][THREADS / 32][G] partial sums
  __shared__ double sum_s[ARP_FFT_G];
  __shared__ T mh_s[ARP_FFT_G], ml_s[ARP_FFT_G], sc_s[ARP_FFT_G];
  __shared__ int kneg_s[ARP_FFT_G], live_s[ARP_FFT_G];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = ARP_FFT_THREADS / 32;
  const long long i0 = (long long)blockIdx.x * ARP_FFT_G;
  // ---- load: thread tid reads column j = tid % G of rows tid / G, tid / G + THREADS / G, ... (whole 32-byte sectors per
  // row); raw values go to shared memory.  First and second moment are accumulated about a pivot (the series' first
  // sample) so that they can stay in `T` without cancellation; the cross-thread reduction is in double.
  {
    const int j = tid % ARP_FFT_G;
    const bool col_ok = i0 + j < n;
    T* dst = reinterpret_cast<T*>(bufs + (size_t)(j >> 1) * ARP_FFT_BUF) + (j & 1);
    const T pivot = col_ok ? x[i0 + j] : (T)0;
    T q1 = 0, q2 = 0;
    for (int t = tid / ARP_FFT_G; t < S; t += ARP_FFT_THREADS / ARP_FFT_G) {
      const T v = col_ok ? x[(size_t)t * n + i0 + j] : (T)0;
      const T dv = v - pivot;
      q1 += dv;
      q2 = fma(dv, dv, q2);
      dst[2 * ARP_FFT_PAD(t)] = v;
    }
    double p1 = (double)q1, p2 = (double)q2;
    // lanes with the same j: lane, lane ^ 8, ^ 16 (G = 8 divides 32)
    p1 += __shfl_xor_sync(0xffffffffu, p1, 8);  p2 += __shfl_xor_sync(0xffffffffu, p2, 8);
    p1 += __shfl_xor_sync(0xffffffffu, p1, 16); p2 += __shfl_xor_sync(0xffffffffu, p2, 16);
    if (lane < ARP_FFT_G) { red[warp * ARP_FFT_G + lane] = p1; red[(NW + warp) * ARP_FFT_G + lane] = p2; }
  }
  __syncthreads();
  if (tid < ARP_FFT_G) {
    double s1 = 0, s2 = 0;
    for (int w = 0; w < NW; ++w) { s1 += red[w * ARP_FFT_G + tid]; s2 += red[(NW + w) * ARP_FFT_G + tid]; }
    const double pivot = (i0 + tid < n) ? (double)x[i0 + tid] : 0.0;
    const double dm = s1 / S;                        // mean - pivot
    double v = s2 / S - dm * dm;                     // biased variance (= lag-0 autocovariance)
    // a constant series gives exactly 0 here (every dv is 0); a non-finite one gives NaN / inf
    const bool okv = v > 0.0 && v < (double)INFINITY;
    const double m = pivot + dm;
    // centring with a two-term mean (hi + lo); scaling to unit variance: ESS does not depend on the scale of a series,
    // but the two series that share one complex transform see each other's rounding noise -- without the scaling a
    // series 1e4 x smaller than its partner would inherit a relative error of 1e4 x 2^-24.  A constant or non-finite
    // series is zeroed (scale 0) so that it cannot contaminate its partner.
    const T mh = okv ? (T)m : (T)0;
    mh_s[tid] = mh;
    ml_s[tid] = okv ? (T)(m - (double)mh) : (T)0;
    sc_s[tid] = okv ? (T)(1.0 / sqrt(v)) : (T)0;
    live_s[tid] = okv ? 1 : 0;
    kneg_s[tid] = S;     // first lag with a negative autocorrelation (S = none)
    sum_s[tid] = 0;
    if (i0 + tid < n) {
      if (mean_out) mean_out[i0 + tid] = (T)m;
      if (var_out) var_out[i0 + tid] = (T)v;
    }
  }
  __syncthreads();
  // ---- per pair of series: forward transform, power spectra, forward transform again
  const int grp = tid / ARP_FFT_TPF, t = tid % ARP_FFT_TPF;
  cplx<T>* buf = bufs + (size_t)grp * ARP_FFT_BUF;
  fft2048<T, true, false>(buf, t, 1 + grp, S, mh_s[2 * grp], ml_s[2 * grp], sc_s[2 * grp], mh_s[2 * grp + 1],
                          ml_s[2 * grp + 1], sc_s[2 * grp + 1]);
  // W_k = |X1_k|^2 + i |X2_k|^2 with X1 = (Z_k + conj Z_{N-k}) / 2, X2 = (Z_k - conj Z_{N-k}) / (2 i); W_{N-k} = W_k
  for (int k = t; k <= ARP_FFT_N / 2; k += ARP_FFT_TPF) {
    const int km = (ARP_FFT_N - k) & (ARP_FFT_N - 1);
    const cplx<T> zk = buf[ARP_FFT_PAD(k)], zm = buf[ARP_FFT_PAD(km)];
    const T ar = (T)0.5 * (zk.x + zm.x), ai = (T)0.5 * (zk.y - zm.y);    // X1
    const T br = (T)0.5 * (zk.y + zm.y), bi = (T)0.5 * (zm.x - zk.x);    // X2 = (Z_k - conj Z_m) / (2 i)
    const cplx<T> w{ar * ar + ai * ai, br * br + bi * bi};
    buf[ARP_FFT_PAD(k)] = w;
    buf[ARP_FFT_PAD(km)] = w;
  }
  fft_bar(1 + grp);
  fft2048<T, false, true>(buf, t, 1 + grp, S, 0, 0, 0, 0, 0, 0);
  // buf[k] = N (c1_k + i c2_k) for k < N / 2 (the factor N cancels in rho)
  // ---- ESS per series: 64 threads each.  With rho_k = (c_k / (S - k)) / (c_0 / S) the weights (S - k) / S cancel:
  //   ESS = S / (-1 + 2 sum_{k < K} c_k / c_0),   K = first lag with c_k < 0  (rho_k < 0 <=> c_k < 0).
  {
    const int sidx = 2 * grp + (t >> 6);          // series within the CTA
    const int u = t & 63;
    const T* c = reinterpret_cast<const T*>(buf) + (t >> 6);
    const T c0 = c[0];
    const bool live = live_s[sidx] && c0 > (T)0;  // constant / non-finite series: NaN, as TFP
    int kneg = S;
    for (int k = u; k < S; k += 64)
      if (c[2 * ARP_FFT_PAD(k)] < (T)0) { kneg = k; break; }          // k increases: the first negative lag this thread sees
    if (live) atomicMin(&kneg_s[sidx], kneg);
    fft_bar(1 + grp);
    const int kn = kneg_s[sidx];
    T acc = 0;
    for (int k = u; k < kn; k += 64) acc += c[2 * ARP_FFT_PAD(k)];
    double part = (double)acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((u & 31) == 0) atomicAdd(&sum_s[sidx], part);
    fft_bar(1 + grp);
    if (u == 0 && i0 + sidx < n)
      ess[i0 + sidx] = live ? (T)((double)S / (-1.0 + 2.0 * sum_s[sidx] / (double)c0)) : (T)NAN;
  }
}
#endif  // __CUDACC__

}  // namespace arp
