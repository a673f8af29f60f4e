// tcgen05 primitives shared by the German-credit tensor-core engine (arp_german_tcs.cuh): PTX wrappers
// (mbarrier, tcgen05.mma / ld / st / commit, shared-memory matrix descriptors), the fp16 head / tail
// split, MUFU-based site math and Box-Muller.
//
//   eta = X beta   (reference models.py:903 einsum)  ->  GEMM1  H[chain, obs] = B[chain, f] X[obs, f]^T
//   gbeta = X^T r  (autodiff of :903-904)            ->  GEMM2  G[chain, f]  = R[chain, obs] X[obs, f]
//
// Precision: fp32 operands are split into an fp16 head and tail (x = x1 + x2, 22 significant bits) and
// each contraction is three kind::f16 MMAs with fp32 accumulation (x1 b1 + x2 b1 + x1 b2) -- the
// split-precision scheme of 3xTF32, done in fp16 because the f16 MMA rate is twice the tf32 rate and
// the images are half the size.  Error per product ~2^-22, which holds the 1e-5 fp32 tolerance of the
// log-joint gradient (tests/test_gpu_tc.py).
//
// (Round 1 also had a kernel with X resident in shared memory and a dual-tile M = 64 variant; both were
// superseded by the streaming kernel -- 2 % / 9 % slower -- and were removed; see git history.)
#pragma once
#include <atomic>
#include <cuda_fp16.h>
#include <string>
#include <vector>
#include "arp_host.cuh"
#include "arp_hmc.cuh"

namespace arp {

#ifndef TC_FAST_SITE
#define TC_FAST_SITE 1    // MUFU-based exp / log / sincos in the site math and the Box-Muller transform
#endif
#ifndef TC_NEWTON_MIX
#define TC_NEWTON_MIX 0   // 1: odd likelihood elements take their reciprocal on the FMA pipe (measured 3.6 % slower: issue-bound)
#endif
#define TC_CHAINS 128
#define TC_CHUNK 128      // observations per GEMM1 tile
#define TC_NQ 4            // worker threads per chain
#define TC_WORKERS (TC_NQ * TC_CHAINS)   // 512
#define TC_TMEM_COLS 512


// ---------------------------------------------------------------- PTX wrappers ---
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (sm_100)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         ((uint64_t)1 << 46);
}

__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}

// Warp-uniform issue: the whole issuer warp runs the issue loop convergently and each tcgen05.mma / commit sits
// under elect.sync, which tells ptxas the region is single-threaded, so descriptors and TMEM addresses stay in
// uniform registers and every MMA is one UTCHMMA.  (Inside an `if (lane == 0)` region nvcc wraps each UTCHMMA in
// an ELECT / R2UR.BROADCAST / BRA.U.ANY loop: ~45 clk of issue latency per MMA, which was on the critical path.)
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void mma_ss_if(uint32_t, uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (elect_one()) mma_ss(d, a, b, idesc, acc);
}
__device__ __forceinline__ void mma_ts_if(uint32_t, uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (elect_one()) mma_ts(d, a_tmem, b, idesc, acc);
}
__device__ __forceinline__ void tc_commit_if(uint32_t, uint32_t bar) {
  if (elect_one()) tc_commit(bar);
}

#define TC_LD32(taddr, v)                                                                                  \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                   \
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"                                    \
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                   \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) \
               : "r"(taddr) : "memory")
#define TC_LD16(taddr, v)                                                                                  \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                   \
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                            \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) \
               : "r"(taddr) : "memory")
#define TC_LD8(taddr, v)                                                                                   \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                     \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) \
               : "r"(taddr) : "memory")
#define TC_ST8(taddr, v)                                                                                   \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"                     \
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory")
#define TC_ST16(taddr, v)                                                                                  \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "                                             \
               "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"                                  \
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), \
                 "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory")

// fp32 pair -> packed fp16x2 head (element 0 in the low half) and the packed tail of the remainders
__device__ __forceinline__ void split_pack(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp for the site scales (sigma = exp(log-scale), once per coefficient per sweep).
//   TC_EXP_MODE 0: one MUFU.EX2 of the rounded product x log2(e): relative error |x| 2^-24 + 2^-22 (6e-7 at x = 9)
//               1: the product carried as hi + lo (two FFMA), the low part applied to first order: 2^-22 (MUFU only)
//               2: expf (2 ulp)
// The coefficient scales feed beta = sigma z into GEMM1; at states with |eta| ~ 1e3 .. 1e4 (far outside the typical
// set, but the parity tests go there) a relative error of 6e-7 in sigma moves eta by several 1e-3 and the tensor-core
// gradient drifted to 1e-5 of the fp64 oracle where the SIMT engine (expf) holds 1e-7 .. 1e-6.
#ifndef TC_EXP_MODE
#define TC_EXP_MODE 1
#endif
__device__ __forceinline__ float exp_fast(float x) {
#if TC_EXP_MODE == 0
  return ex2_approx(x * 1.4426950408889634f);
#elif TC_EXP_MODE == 1
  const float t = x * 1.4426950216293335f;                                      // fp32(log2 e)
  const float r = fmaf(x, 1.925963033500011e-08f, fmaf(x, 1.4426950216293335f, -t));   // what the product lost
  const float e = ex2_approx(t);
  return fmaf(e, r * 0.6931471805599453f, e);
#else
  return expf(x);
#endif
}
// reciprocal of d in [1, 2^60] on the FMA pipe: bit-trick seed + 3 Newton steps (error ~5e-8);
// takes MUFU pressure off the epilogue, where the XU pipe is the binding resource
__device__ __forceinline__ float rcp_newton(float d) {
  float r = __int_as_float(0x7EF311C3 - __float_as_int(d));
  r = r * fmaf(-d, r, 2.0f);
  r = r * fmaf(-d, r, 2.0f);
  r = r * fmaf(-d, r, 2.0f);
  return r;
}
// site rule with fast exponentials (same algebra as arp_common.cuh: site_fwd)
__device__ __forceinline__ Site site_fwd_fast(float z, float mu, float ls, float a, float b, float& lp) {
#if !TC_FAST_SITE
  return site_fwd(z, mu, ls, a, b, lp);
#endif
  Site s;
  float sb_inv;
  if (b == 1.f) { sb_inv = exp_fast(-ls); s.r = 1.f; }
  else if (b == 0.f) { sb_inv = 1.f; s.r = exp_fast(ls); }
  else { sb_inv = exp_fast(-b * ls); s.r = exp_fast((1.f - b) * ls); }
  s.dz = z - a * mu;
  const float u = s.dz * sb_inv;
  s.usb = u * sb_inv;
  s.x = mu + s.r * s.dz;
  lp += -0.5f * u * u - b * ls - ARP_HALF_LOG_2PI;
  return s;
}
// 4 standard normals from one Philox block, MUFU log / sincos (the XU pipe idles outside the epilogue)
__device__ __forceinline__ void philox_normal4_fast(uint64_t seed, uint32_t chain, uint32_t step, uint32_t j, float out[4]) {
#if !TC_FAST_SITE
  philox_normal4(seed, chain, step, j, ARP_STREAM_MOMENTUM, out);
  return;
#endif
  const uint4 r = philox4x32_10(make_uint4(chain, step, j, ARP_STREAM_MOMENTUM), (uint32_t)seed, (uint32_t)(seed >> 32));
  const float rad0 = sqrtf(-2.0f * __logf(u01(r.x))), rad1 = sqrtf(-2.0f * __logf(u01(r.z)));
  float s0, c0, s1, c1;   // sin/cos(2 pi u) = -sin/cos(2 pi (u - 1/2)), argument in (-pi, pi)
  __sincosf(6.283185307179586f * (u01(r.y) - 0.5f), &s0, &c0);
  __sincosf(6.283185307179586f * (u01(r.w) - 0.5f), &s1, &c1);
  out[0] = -rad0 * c0; out[1] = -rad0 * s0; out[2] = -rad1 * c1; out[3] = -rad1 * s1;
}

static inline bool german_tc_auto(long long C) { return C >= 2 * TC_CHAINS; }

__global__ void k_gather_ws_tc(const real* ws, int sd, int sc, int C, int D, real* out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)C * D) return;
  const int c = (int)(i / D), d = (int)(i % D);
  out[i] = ws[(size_t)d * sd + (size_t)c * sc];
}

}  // namespace arp
