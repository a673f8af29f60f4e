// Fused mean-field VI: reparameterised sample -> ELBO gradient -> Adam, every
// optimisation step of every learning rate in ONE persistent kernel launch.
//
// Replaces util.get_mean_field_elbo (reference util.py:232-268), the mean-field
// guide of program_transformations.make_variational_model (:192-241: q = prod
// N(loc, softplus(rho))) and the Adam loop of inference.find_best_learning_rate
// (inference.py:26-154: TF1 Adam, lr/5 after 1/3 and lr/20 after 2/3 of the
// steps, NaN gradients zeroed).  The reference crosses the host<->runtime
// boundary once per step (sess.run, inference.py:93); here a cluster of 8 CTAs
// owns one learning rate, 8 lanes own one Monte-Carlo sample, parameters, Adam
// moments and per-sample vectors stay in shared memory, and nothing leaves the
// GPU until the end.
//
// ELBO = mean_s [ log_joint(z_s) + sum_d (eps^2/2 + log scale_d + log(2 pi)/2) ],
// z_s = loc + scale * eps_s; gradients of -ELBO:
//   d loc = -mean_s g_s          d scale = -mean_s g_s eps_s - 1/scale
//   d rho = d scale * sigmoid(rho)
// Learnable reparameterisation (cVIP, program_transformations.py:486-533): P unconstrained parameters u_p, value
// sigmoid(u_p); coordinate d reads its `a` from slot ia[d] and its `b` from slot ib[d] (-1 = fixed).  This covers
//   tied as written   ia[d] = d, ib = -1 (b = 1)              [the reference's default, SURVEY.md section 0 item 3]
//   tied, b = a       ia[d] = ib[d] = d                        [the paper's intent]
//   untied            ia = slot by loc shape, ib = slot by scale shape (a scalar loc / scale shares one slot per site)
//   d u_p = -[ mean_s (sum_{ia[d]=p} abar_{s,d} + sum_{ib[d]=p} bbar_{s,d}) + d log prior / d p ] * p (1 - p)
#pragma once
#include <atomic>
#include <string>
#include "arp_host.cuh"
#include "arp_models.cuh"

namespace arp {

#define ARP_VI_MAX_S 4096
#define ARP_VI_BLOCK 256
#ifndef ARP_VI_MAX_RUNS
#define ARP_VI_MAX_RUNS 16
#endif

struct ViArgs {
  int D, S, steps, R;
  int P;                // learnable reparameterisation parameters (0 = fixed (a, b): CP / NCP / dVIP)
  int discrete_prior;   // 1: mixture-of-Laplace prior on the learnable parameters (main.py:244-253)
  real lrs[ARP_VI_MAX_RUNS];
  unsigned long long seed;
  real* loc;            // [R, D] in/out
  real* rho;            // [R, D] in/out
  real* u;              // [R, P] in/out  unconstrained parameters, value = sigmoid(u) (program_transformations.py:507-523)
  const int* ia;        // [D] parameter slot that coordinate d's `a` reads (-1: a_in[d] is used, not learned)
  const int* ib;        // [D] same for `b`
  const real* ext_eps;  // [steps, S, D] or null (shared by all runs)
  real* elbo;           // [R, steps]  ELBO (+ prior log-prob of the parameters when discrete_prior)
  real* prior_logp;     // [R, steps] or null
  const real* a_in;     // [D]
  const real* b_in;     // [D]
};

__device__ __forceinline__ void adam_update(real& theta, real grad, real& m1, real& m2, real lr_t) {
  if (grad != grad) grad = 0;  // inference.py:62 remove_nans
  m1 = (real)0.9 * m1 + (real)0.1 * grad;
  m2 = (real)0.999 * m2 + (real)0.001 * grad * grad;
  theta -= lr_t * m1 / (r_sqrt(m2) + (real)1e-8);
}

// log density and its derivative of the reference's prior on a learnable parameter p in (0, 1) (main.py:244-253):
// Mixture(Categorical(logits = [0, 5, 0]), [Laplace(0, 0.1), Uniform(0, 1), Laplace(1, 0.1)]).
__device__ __forceinline__ real discrete_prior_logp(real p, real& dlogp) {
  const real e5 = (real)148.41315910257660342;           // exp(5)
  const real w0 = (real)1 / ((real)2 + e5), w1 = e5 * w0;
  const real l0 = (real)5 * r_exp((real)-10 * p), l2 = (real)5 * r_exp((real)-10 * ((real)1 - p));
  const real mix = w0 * l0 + w1 + w0 * l2;
  dlogp = (real)10 * w0 * (l2 - l0) / mix;
  return r_log(mix);
}

// One learning rate = one thread-block CLUSTER of ARP_VI_NCTA CTAs (8 SMs): the S Monte-Carlo samples are dealt to the
// CTAs, ARP_VI_LPC lanes cooperate on one sample (S = 256: 32 samples per CTA x 8 lanes = 256 threads), and the
// per-sample state vectors live in shared memory.  Every CTA keeps its OWN copy of the parameters and Adam moments; per
// step each CTA reduces the gradient over its samples into a partial block in its shared memory, one cluster barrier
// later every CTA reads all partial blocks through distributed shared memory in the same (rank) order -- so all copies
// take bit-identical Adam steps and nothing has to be broadcast.  The partial blocks are double-buffered by step
// parity, which makes ONE cluster barrier per optimisation step sufficient.
// (Round 1 ran one CTA per learning rate with one thread per sample: 5 of 148 SMs, every log-joint serial in a thread.)
#define ARP_VI_NCTA 8
#define ARP_VI_LPC 8
#define ARP_VI_SPB (ARP_VI_BLOCK / ARP_VI_LPC)   // samples per CTA per pass

__device__ __forceinline__ unsigned int vi_cluster_rank() {
  unsigned int r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void vi_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// generic address of `p` (a shared-memory address of this CTA) in CTA `rank` of the cluster
template <typename T>
__device__ __forceinline__ const T* vi_map_rank(const T* p, unsigned int rank) {
  uint64_t out;
  asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"(reinterpret_cast<uint64_t>(p)), "r"(rank));
  return reinterpret_cast<const T*>(out);
}

template <int KIND, bool LEARN, int FP>
__global__ void __cluster_dims__(ARP_VI_NCTA, 1, 1) __launch_bounds__(ARP_VI_BLOCK)
k_vi(DevModel m, ViArgs v, real* ws_all, int spc, int Dpad, int ws_in_smem) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  real* sm = reinterpret_cast<real*>(smem_raw);
  constexpr int LPC = ARP_VI_LPC, NCTA = ARP_VI_NCTA, nthr = ARP_VI_BLOCK;
  const int D = v.D, S = v.S, P = v.P;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
  const int sub = tid % LPC, slot = tid / LPC;
  const unsigned int crank = vi_cluster_rank();
  const int run = blockIdx.x / NCTA;
  // ---- shared memory: parameters, Adam moments, partial gradient blocks, (per-sample vectors)
  real* loc = sm;          real* rho = loc + D;    real* scale = rho + D;
  real* a_s = scale + D;   real* b_s = a_s + D;
  real* gl = b_s + D;      real* gs = gl + D;      real* ga = gs + D;   real* gb = ga + D;
  real* mom = gb + D;      // [4][D] Adam first/second moments of loc, rho
  real* up = mom + 4 * D;  // [P] unconstrained parameters; pv [P] values; gu [P] gradient sums; pm [2][P] Adam moments
  real* pv = up + P;       real* gu = pv + P;      real* pm = gu + P;
  real* red = pm + 2 * P;  // [32] block-reduction scratch
  real* part = red + 32;   // [2][4 D + 4] this CTA's partial sums (loc, scale, a, b gradients; ELBO), by step parity
  const int PB = 4 * D + 4;
  int* ia = reinterpret_cast<int*>(part + 2 * PB);
  int* ib = ia + D;
  real* wsm = reinterpret_cast<real*>(ib + D);   // [ARP_VI_SPB][6][Dpad] per-sample vectors of the current pass
  real* ws_g = ws_all + (size_t)blockIdx.x * ARP_VI_SPB * 6 * Dpad;
  real* wbase = (ws_in_smem ? wsm : ws_g) + (size_t)slot * 6 * Dpad;
  Vec Z{wbase, 1}, G{wbase + Dpad, 1}, XC{wbase + 2 * Dpad, 1}, AB{wbase + 3 * Dpad, 1}, E{wbase + 4 * Dpad, 1},
      BB{wbase + 5 * Dpad, 1};
  for (int d = tid; d < D; d += nthr) {
    loc[d] = v.loc[(size_t)run * D + d];
    rho[d] = v.rho[(size_t)run * D + d];
    a_s[d] = v.a_in[d];
    b_s[d] = v.b_in[d];
    ia[d] = LEARN ? v.ia[d] : -1;
    ib[d] = LEARN ? v.ib[d] : -1;
#pragma unroll
    for (int q = 0; q < 4; ++q) mom[q * D + d] = 0;
  }
  if (LEARN)
    for (int p = tid; p < P; p += nthr) { up[p] = v.u[(size_t)run * P + p]; pm[p] = 0; pm[P + p] = 0; }
  const real base_lr = v.lrs[run];
  __syncthreads();
  for (int step = 0; step < v.steps; ++step) {
    if (LEARN) {
      for (int p = tid; p < P; p += nthr) pv[p] = (real)1 / ((real)1 + r_exp(-up[p]));
      __syncthreads();
    }
    for (int d = tid; d < D; d += nthr) {
      scale[d] = r_softplus(rho[d]);
      if (LEARN) {
        if (ia[d] >= 0) a_s[d] = pv[ia[d]];
        if (ib[d] >= 0) b_s[d] = pv[ib[d]];
      }
      gl[d] = 0; gs[d] = 0; ga[d] = 0; gb[d] = 0;
    }
    __syncthreads();
    // ---- this CTA's samples crank * spc .. + spc, ARP_VI_SPB per pass, LPC lanes each
    real el = 0;
    for (int s0 = 0; s0 < spc; s0 += ARP_VI_SPB) {
      const int sl = s0 + slot;                       // sample index inside this CTA's share
      const int smp = (int)crank * spc + sl;          // global sample index: keys the Philox stream
      const bool live = sl < spc && smp < S;
      real ent = 0;
      if (v.ext_eps) {
        const real* e = v.ext_eps + ((size_t)step * S + (live ? smp : 0)) * D;
        for (int d = sub; d < D; d += LPC) {
          const real ee = live ? e[d] : (real)0;
          E(d) = ee;
          Z(d) = loc[d] + scale[d] * ee;
          ent += (real)0.5 * ee * ee + r_log(scale[d]) + ARP_HALF_LOG_2PI;
        }
      } else {
        const int nb = (D + 3) >> 2;
        for (int j = sub; j < nb; j += LPC) {
          real n4[4];
          philox_normal4(v.seed, (unsigned int)smp, (unsigned int)step, (unsigned int)j, ARP_STREAM_VI, n4);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int d = 4 * j + q;
            if (d < D) {
              E(d) = n4[q];
              Z(d) = loc[d] + scale[d] * n4[q];
              ent += (real)0.5 * n4[q] * n4[q] + r_log(scale[d]) + ARP_HALF_LOG_2PI;
            }
          }
        }
      }
      ent = group_sum<LPC>(ent);
      __syncwarp();
      const real lp = vg<KIND, LPC, LEARN, FP>(m, a_s, b_s, Z, G, XC, AB, BB, sub, true);
      if (live && sub == 0) el += lp + ent;
      __syncthreads();   // G / E / AB / BB of the pass are visible to the block
      // ---- gradient sums over the samples of this pass: one thread per coordinate, samples in slot order
      for (int d = tid; d < D; d += nthr) {
        const real* base = ws_in_smem ? wsm : ws_g;
        const bool la = LEARN && ia[d] >= 0, lb = LEARN && ib[d] >= 0;
        real s_l = 0, s_s = 0, s_a = 0, s_b = 0;
        const int nlive = min(ARP_VI_SPB, min(spc - s0, S - ((int)crank * spc + s0)));
        for (int q = 0; q < nlive; ++q) {
          const real* w = base + (size_t)q * 6 * Dpad;
          const real gg = w[Dpad + d];
          s_l += gg;
          s_s = fma(gg, w[4 * Dpad + d], s_s);
          if (la) s_a += w[3 * Dpad + d];
          if (lb) s_b += w[5 * Dpad + d];
        }
        gl[d] += s_l; gs[d] += s_s; ga[d] += s_a; gb[d] += s_b;
      }
      __syncthreads();
    }
    // ---- this CTA's partial block (step parity), then one cluster barrier
    el = group_sum<32>(el);
    if (lane == 0) red[warp] = el;
    real* mine = part + (step & 1) * PB;
    for (int d = tid; d < D; d += nthr) { mine[d] = gl[d]; mine[D + d] = gs[d]; mine[2 * D + d] = ga[d]; mine[3 * D + d] = gb[d]; }
    __syncthreads();
    if (tid == 0) {
      real tot = 0;
      for (int w = 0; w < nwarp; ++w) tot += red[w];
      mine[4 * D] = tot;
    }
    vi_cluster_sync();
    // ---- totals: every CTA sums the NCTA partial blocks in rank order (identical results in every CTA)
    for (int i = tid; i < 4 * D + 1; i += nthr) {
      real tot = 0;
#pragma unroll
      for (unsigned int r = 0; r < NCTA; ++r) tot += vi_map_rank(mine, r)[i];
      if (i < D) gl[i] = tot;
      else if (i < 2 * D) gs[i - D] = tot;
      else if (i < 3 * D) ga[i - 2 * D] = tot;
      else if (i < 4 * D) gb[i - 3 * D] = tot;
      else red[0] = tot;
    }
    __syncthreads();
    if (tid == 0 && crank == 0) {
      real plp = 0;
      if (LEARN && v.discrete_prior)
        for (int p = 0; p < P; ++p) { real dl; plp += discrete_prior_logp(pv[p], dl); }
      v.elbo[(size_t)run * v.steps + step] = red[0] / (real)S + plp;   // elbo_with_prior (inference.py:54)
      if (v.prior_logp) v.prior_logp[(size_t)run * v.steps + step] = plp;
    }
    if (LEARN) {
      // a parameter may be shared by several coordinates (untied: `a` has the shape of the site's loc, `b` of its
      // scale) and by a and b (tied b = a): sum the coordinate adjoints per slot, in coordinate order (deterministic)
      for (int p = tid; p < P; p += nthr) {
        real acc = 0;
        for (int d = 0; d < D; ++d) {
          if (ia[d] == p) acc += ga[d];
          if (ib[d] == p) acc += gb[d];
        }
        gu[p] = acc;
      }
      __syncthreads();
    }
    // ---- Adam (TF1 formulation, inference.py:47) with the reference's lr schedule (:69-75)
    real lr = base_lr;
    if (3LL * step > 2LL * v.steps) lr = base_lr / (real)20;
    else if (3LL * step > (long long)v.steps) lr = base_lr / (real)5;
    const double t = (double)(step + 1);
    const real lr_t = (real)((double)lr * sqrt(1.0 - pow(0.999, t)) / (1.0 - pow(0.9, t)));
    const real inv_S = (real)1 / (real)S;
    for (int d = tid; d < D; d += nthr) {
      const real g_loc = -gl[d] * inv_S;
      const real g_scale = -gs[d] * inv_S - (real)1 / scale[d];
      const real sig_rho = (real)1 / ((real)1 + r_exp(-rho[d]));
      adam_update(loc[d], g_loc, mom[d], mom[D + d], lr_t);
      adam_update(rho[d], g_scale * sig_rho, mom[2 * D + d], mom[3 * D + d], lr_t);
    }
    if (LEARN)
      for (int p = tid; p < P; p += nthr) {
        const real pp = pv[p];
        real dl = 0;
        if (v.discrete_prior) discrete_prior_logp(pp, dl);
        adam_update(up[p], -(gu[p] * inv_S + dl) * pp * ((real)1 - pp), pm[p], pm[P + p], lr_t);
      }
    __syncthreads();
  }
  vi_cluster_sync();   // no CTA may exit while another still reads its partial block
  if (crank == 0) {
    for (int d = tid; d < D; d += nthr) {
      v.loc[(size_t)run * D + d] = loc[d];
      v.rho[(size_t)run * D + d] = rho[d];
    }
    if (LEARN)
      for (int p = tid; p < P; p += nthr) v.u[(size_t)run * P + p] = up[p];
  }
}

static inline int vi_launch(const DevModel& dm, int fp, const ViArgs& v, cudaStream_t st, DevBuf* ws,
                            std::atomic<long long>* launches, std::string* err) {
  const int spc = (v.S + ARP_VI_NCTA - 1) / ARP_VI_NCTA;             // samples per CTA
  const int Dpad = (v.D + 3) / 4 * 4 + 1;                             // + 1: the samples of a warp start in different banks
  const size_t fixed = (size_t)(13 * v.D + 5 * v.P + 32 + 2 * (4 * v.D + 4)) * sizeof(real) + (size_t)2 * v.D * sizeof(int);
  const size_t wsb = (size_t)ARP_VI_SPB * 6 * Dpad * sizeof(real);
  const int ws_in_smem = fixed + wsb <= 200 * 1024;
  const size_t smem = fixed + (ws_in_smem ? wsb : 0);
  if (smem > 200 * 1024) { *err = "vi: model too large for the shared-memory parameter block"; return 1; }
  const size_t ws_bytes = ws_in_smem ? 256 : (size_t)v.R * ARP_VI_NCTA * wsb;
  cudaError_t e = ws->alloc(ws_bytes);
  if (e != cudaSuccess) { *err = std::string("vi workspace: ") + cudaGetErrorString(e); return 1; }
  cudaMemsetAsync(ws->p, 0, ws_bytes, st);
  const dim3 grid((unsigned)(v.R * ARP_VI_NCTA));
#define ARP_VI_GO(KIND, FP)                                                                              \
  do {                                                                                                   \
    if (v.P > 0) {                                                                                       \
      cudaFuncSetAttribute(k_vi<KIND, true, FP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      k_vi<KIND, true, FP><<<grid, ARP_VI_BLOCK, smem, st>>>(dm, v, ws->as<real>(), spc, Dpad, ws_in_smem); \
    } else {                                                                                             \
      cudaFuncSetAttribute(k_vi<KIND, false, FP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      k_vi<KIND, false, FP><<<grid, ARP_VI_BLOCK, smem, st>>>(dm, v, ws->as<real>(), spc, Dpad, ws_in_smem); \
    }                                                                                                    \
  } while (0)
  switch (dm.kind) {
    case MODEL_8SCHOOLS: ARP_VI_GO(MODEL_8SCHOOLS, 32); break;
    case MODEL_GERMAN_LOGNORMAL: if (fp == 32) ARP_VI_GO(MODEL_GERMAN_LOGNORMAL, 32); else ARP_VI_GO(MODEL_GERMAN_LOGNORMAL, 64); break;
    case MODEL_GERMAN_GAMMA: if (fp == 32) ARP_VI_GO(MODEL_GERMAN_GAMMA, 32); else ARP_VI_GO(MODEL_GERMAN_GAMMA, 64); break;
    case MODEL_RADON: ARP_VI_GO(MODEL_RADON, 32); break;
    case MODEL_RADON_STDDVS: ARP_VI_GO(MODEL_RADON_STDDVS, 32); break;
    case MODEL_ELECTION: ARP_VI_GO(MODEL_ELECTION, 32); break;
    case MODEL_ELECTRIC: ARP_VI_GO(MODEL_ELECTRIC, 32); break;
    default: ARP_VI_GO(MODEL_TIME_SERIES, 32); break;
  }
#undef ARP_VI_GO
  launches->fetch_add(1);
  e = cudaGetLastError();
  if (e != cudaSuccess) { *err = std::string("k_vi launch: ") + cudaGetErrorString(e); return 1; }
  return 0;
}

}  // namespace arp
