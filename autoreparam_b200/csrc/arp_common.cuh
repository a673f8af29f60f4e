// Common device utilities for the autoreparam B200 library (sm_100a only).
//
//  * `real`            -- float (product build) or double (-DARP_FP64 check build)
//  * site rule         -- the (a, b) partial-centring rule of the reference's
//                         `recenter` interceptor (program_transformations.py:555-600,
//                         NCP special case :262-279) with its hand-derived adjoint
//                         
//  * Philox4x32-10     -- counter-based RNG; counter = (global chain id, step,
//                         block, stream) so results do not depend on the GPU count
//  * sub-warp groups   -- LPC lanes cooperate on one chain; butterfly reductions
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#ifdef ARP_FP64
typedef double real;
#define ARP_REAL_IS_DOUBLE 1
#else
typedef float real;
#define ARP_REAL_IS_DOUBLE 0
#endif

#define ARP_HALF_LOG_2PI ((real)0.91893853320467274178)
#define ARP_LOG_10 ((real)2.30258509299404568402)
#define ARP_LOG_5 ((real)1.60943791243410037460)
#define ARP_LOG_100 ((real)4.60517018598809136804)

namespace arp {

// ------------------------------------------------------------------ math ---
__device__ __forceinline__ float r_exp(float x) { return expf(x); }
__device__ __forceinline__ double r_exp(double x) { return exp(x); }
__device__ __forceinline__ float r_log(float x) { return logf(x); }
__device__ __forceinline__ double r_log(double x) { return log(x); }
__device__ __forceinline__ float r_log1p(float x) { return log1pf(x); }
__device__ __forceinline__ double r_log1p(double x) { return log1p(x); }
__device__ __forceinline__ float r_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double r_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float r_pow(float x, float y) { return powf(x, y); }
__device__ __forceinline__ double r_pow(double x, double y) { return pow(x, y); }
__device__ __forceinline__ float r_abs(float x) { return fabsf(x); }
__device__ __forceinline__ double r_abs(double x) { return fabs(x); }

// sigmoid used inside the likelihood hot loops.  float: ex2.approx + rcp.approx
// (2 MUFU ops, ~2 ulp); the residual y - sigmoid(eta) carries an absolute error
// of ~1e-7 which is far inside the 1e-5 gradient tolerance.
__device__ __forceinline__ float r_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ double r_sigmoid(double x) { return 1.0 / (1.0 + exp(-x)); }
// softplus(x) = max(x,0) + log1p(exp(-|x|)): the Bernoulli log-likelihood term
template <typename T>
__device__ __forceinline__ T r_softplus(T x) {
  return (x > (T)0 ? x : (T)0) + r_log1p(r_exp(-r_abs(x)));
}

// --------------------------------------------------------------- groups ---
// LPC (lanes per chain) consecutive lanes of a warp cooperate on one chain.
template <int LPC, typename T>
__device__ __forceinline__ T group_sum(T v) {
#pragma unroll
  for (int o = LPC / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Strided per-chain vector: element d lives at p[d * sd].
struct Vec {
  real* p;
  int sd;
  __device__ __forceinline__ real& operator()(int d) const { return p[(size_t)d * sd]; }
};

// ------------------------------------------------------------ site rule ---
// One `ed.Normal(loc=mu, scale=sigma)` latent site under rule (a, b):
//   prior term   N(z ; a*mu, sigma^b)           (program_transformations.py:569-572)
//   centred      x = mu + sigma^(1-b) (z - a mu) (:574-576,600)
// `ls` is log(sigma).  CP is a=b=1, NCP a=b=0.
// The scalar type is a template parameter: every model works in `real`; the time-series scans run their site
// arithmetic in double even in the fp32 build (see vg_time_series).
template <typename T>
struct SiteT {
  T x;    // centred value
  T r;    // sigma^(1-b)
  T usb;  // u / sigma^b  with u = (z - a mu) / sigma^b
  T dz;   // z - a mu
};
typedef SiteT<real> Site;

#define ARP_HALF_LOG_2PI_D 0.91893853320467274178

template <typename T>
__device__ __forceinline__ SiteT<T> site_fwd(T z, T mu, T ls, T a, T b, T& lp) {
  SiteT<T> s;
  T sb_inv;
  if (b == (T)1) {
    sb_inv = r_exp(-ls);
    s.r = (T)1;
  } else if (b == (T)0) {
    sb_inv = (T)1;
    s.r = r_exp(ls);
  } else {
    sb_inv = r_exp(-b * ls);
    s.r = r_exp(((T)1 - b) * ls);
  }
  s.dz = z - a * mu;
  T u = s.dz * sb_inv;
  s.usb = u * sb_inv;
  s.x = mu + s.r * s.dz;
  lp += (T)-0.5 * u * u - b * ls - (T)ARP_HALF_LOG_2PI_D;
  return s;
}

// sigma == 1 (ls == 0): sigma^b = 1 for every b.
template <typename T>
__device__ __forceinline__ SiteT<T> site_fwd_unit(T z, T mu, T a, T& lp) {
  SiteT<T> s;
  s.r = (T)1;
  s.dz = z - a * mu;
  s.usb = s.dz;
  s.x = mu + s.dz;
  lp += (T)-0.5 * s.dz * s.dz - (T)ARP_HALF_LOG_2PI_D;
  return s;
}

// Reverse sweep of one site, with d/d(log sigma)
// instead of d/d(sigma):  lsbar = xbar dz (1-b) r + (u^2 - 1) b.
template <typename T>
__device__ __forceinline__ void site_rev(const SiteT<T>& s, T xbar, T mu, T a, T b,
                                         T& zbar, T& mubar, T& lsbar, T& abar) {
  zbar = xbar * s.r - s.usb;
  mubar = xbar * ((T)1 - s.r * a) + a * s.usb;
  lsbar = xbar * s.dz * ((T)1 - b) * s.r + (s.usb * s.dz - (T)1) * b;
  abar = mu * (s.usb - xbar * s.r);
}

// d log_joint / d b of one site: ln(sigma) [(u^2 - 1) - xbar (z - a mu) sigma^(1-b)].
// Needed only when b is learned (untied VIP, or the paper's tied b = a): program_transformations.py:512-523.
template <typename T>
__device__ __forceinline__ T site_bbar(const SiteT<T>& s, T xbar, T ls) {
  return ls * ((s.usb * s.dz - (T)1) - xbar * s.dz * s.r);
}

// --------------------------------------------------------------- Philox ---
#define ARP_STREAM_MOMENTUM 0u
#define ARP_STREAM_ACCEPT 1u
#define ARP_STREAM_VI 2u

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

// uint32 -> uniform in (0,1): (top 24 bits + 0.5) * 2^-24
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * 5.9604644775390625e-08f; }

// 4 standard normals for coordinates 4j .. 4j+3 of one chain at one step
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint32_t chain, uint32_t step, uint32_t j,
                                               uint32_t stream, real out[4]) {
  uint4 r = philox4x32_10(make_uint4(chain, step, j, stream), (uint32_t)seed, (uint32_t)(seed >> 32));
#if ARP_REAL_IS_DOUBLE
  double u0 = (double)u01(r.x), u1 = (double)u01(r.y), u2 = (double)u01(r.z), u3 = (double)u01(r.w);
  double rad0 = sqrt(-2.0 * log(u0)), rad1 = sqrt(-2.0 * log(u2));
  double s0, c0, s1, c1;
  sincospi(2.0 * u1, &s0, &c0);
  sincospi(2.0 * u3, &s1, &c1);
#else
  float u0 = u01(r.x), u1 = u01(r.y), u2 = u01(r.z), u3 = u01(r.w);
  float rad0 = sqrtf(-2.0f * logf(u0)), rad1 = sqrtf(-2.0f * logf(u2));
  float s0, c0, s1, c1;
  sincospif(2.0f * u1, &s0, &c0);
  sincospif(2.0f * u3, &s1, &c1);
#endif
  out[0] = rad0 * c0;
  out[1] = rad0 * s0;
  out[2] = rad1 * c1;
  out[3] = rad1 * s1;
}

__device__ __forceinline__ real philox_log_uniform(uint64_t seed, uint32_t chain, uint32_t step) {
  uint4 r = philox4x32_10(make_uint4(chain, step, 0u, ARP_STREAM_ACCEPT), (uint32_t)seed, (uint32_t)(seed >> 32));
  return r_log((real)u01(r.x));
}

}  // namespace arp
