// Per-model log-joint + hand-derived adjoint under the (a, b) site rule.
//
// Every function evaluates ONE chain cooperatively on LPC consecutive lanes of a
// warp (LPC = 1 .. 32): per-site / per-observation loops are strided over the
// lanes, cross-lane sums are xor-butterflies (so every lane of the group ends up
// with bit-identical scalars), and the state vectors live in (L1/L2-cached)
// global memory behind a strided accessor.
//
//   lp = vg<KIND, LPC, WITH_A, FP>(m, a, b, z, g, xc, abar, bbar, sub, want_lp)
//     z     in   state coordinates (trace order)
//     g     out  d log_joint / d z
//     xc    out  centred value of every coordinate (= make_to_centered(z),
//                reference models.py:59-81); also used as scratch between lanes
//     abar  out  d log_joint / d a per coordinate (only if WITH_A; cVIP VI)
//     bbar  out  d log_joint / d b per coordinate (only if WITH_A; VI with a learned b)
//     returns log_joint (valid only if want_lp; the gradient never needs it)
//
// Model bodies follow reference models.py (line ranges per function) with the
// group look-ups done as gathers / segmented sums instead of the reference's
// dense one-hot matmuls.
#pragma once
#include "arp_common.cuh"

namespace arp {

enum ModelKind {
  MODEL_8SCHOOLS = 0,
  MODEL_GERMAN_LOGNORMAL = 1,
  MODEL_GERMAN_GAMMA = 2,
  MODEL_RADON = 3,
  MODEL_RADON_STDDVS = 4,
  MODEL_ELECTION = 5,
  MODEL_ELECTRIC = 6,
  MODEL_TIME_SERIES = 7,
  MODEL_COUNT = 8
};

// Device-resident model data (all pointers are device pointers owned by arp_model).
struct DevModel {
  int kind;
  int D;     // number of state coordinates
  int N;     // observations (election: weighted cells)
  int F;     // german: features
  int Fpad;  // german: row stride of X (multiple of 4, zero padded)
  int J;     // groups: radon counties | election states (+1 NONE group) | electric pairs (+1 NONE)
  int K;     // election: n_state | electric: n_pair | time series: T
  const real* X;    // german [N][Fpad]
  const real* y;    // observations (election: sum of y per cell)
  const real* x1;   // radon floor | election female | electric treatment | time_series year | 8schools sigma
  const real* x2;   // election black
  const real* w;    // election: cell count | radon: per-county sufficient statistics [J][6]
  const real* u;    // radon: log uranium [J]
  const int* offs;  // CSR offsets of the groups into the (sorted) observations [J+1]
  const int* gidx;  // electric: grade index per observation (-1 = one-hot out of range)
  const int* pidx;  // electric: grade_pair index per pair (-1 = out of range)
};

template <typename T>
__device__ __forceinline__ T ldg(const T* p) { return __ldg(p); }

// ------------------------------------------------------------- 8 schools ---
// reference models.py:139-147.  z = [mu, log_tau, theta[8]]
template <int LPC, bool WITH_A>
__device__ real vg_8schools(const DevModel& m, const real* a, const real* b,
                            Vec z, Vec g, Vec xc, Vec abar, Vec bbar, int sub, bool want_lp) {
  real lp_top = 0;
  const real a0 = (*(a + 0)), b0 = (*(b + 0)), a1 = (*(a + 1)), b1 = (*(b + 1));
  Site smu = site_fwd(z(0), (real)0, ARP_LOG_5, a0, b0, lp_top);
  Site slt = site_fwd(z(1), (real)0, ARP_LOG_5, a1, b1, lp_top);
  const real mu = smu.x, lt = slt.x;
  real lp = 0, acc_mu = 0, acc_lt = 0;
  for (int i = sub; i < 8; i += LPC) {
    const real ai = (*(a + 2 + i)), bi = (*(b + 2 + i));
    Site st = site_fwd(z(2 + i), mu, lt, ai, bi, lp);
    const real sig = ldg(m.x1 + i);
    const real e = (ldg(m.y + i) - st.x) / sig;
    lp += (real)-0.5 * e * e - r_log(sig) - ARP_HALF_LOG_2PI;
    real zb, mb, lb, ab;
    site_rev(st, e / sig, mu, ai, bi, zb, mb, lb, ab);
    g(2 + i) = zb;
    xc(2 + i) = st.x;
    if (WITH_A) { abar(2 + i) = ab; bbar(2 + i) = site_bbar(st, e / sig, lt); }
    acc_mu += mb;
    acc_lt += lb;
  }
  acc_mu = group_sum<LPC>(acc_mu);
  acc_lt = group_sum<LPC>(acc_lt);
  lp = group_sum<LPC>(lp) + lp_top;
  if (sub == 0) {
    real zb, mb, lb, ab;
    site_rev(smu, acc_mu, (real)0, a0, b0, zb, mb, lb, ab);
    g(0) = zb;
    xc(0) = mu;
    site_rev(slt, acc_lt, (real)0, a1, b1, zb, mb, lb, ab);
    g(1) = zb;
    xc(1) = lt;
    if (WITH_A) {
      abar(0) = 0; abar(1) = 0;
      bbar(0) = site_bbar(smu, acc_mu, ARP_LOG_5); bbar(1) = site_bbar(slt, acc_lt, ARP_LOG_5);
    }
  }
  return lp;
}

// ---------------------------------------------------------- German credit ---
// Bernoulli-logit likelihood over the design matrix: eta = X beta,
// r = y - sigmoid(eta), gbeta = X^T r  (reference models.py:903-904 einsum +
// ed.Bernoulli; autodiff gives X^T r).  beta is read from xc(beta_off + f);
// the reduced gradient is left in g(beta_off + f).  FP = padded feature count
// held in registers.
template <int LPC, int FP>
__device__ real german_likelihood(const DevModel& m, Vec xc, Vec g, int beta_off, int sub, bool want_lp) {
  real be[FP], gb[FP];
#pragma unroll
  for (int f = 0; f < FP; ++f) {
    be[f] = (f < m.F) ? xc(beta_off + f) : (real)0;
    gb[f] = 0;
  }
  real lp = 0;
  const real* __restrict__ X = m.X;
  const real* __restrict__ y = m.y;
#pragma unroll 2
  for (int n = sub; n < m.N; n += LPC) {
    const real* row = X + (size_t)n * FP;
    real xr[FP];
#if ARP_REAL_IS_DOUBLE
#pragma unroll
    for (int f = 0; f < FP; f += 2) {
      double2 v = __ldg(reinterpret_cast<const double2*>(row + f));
      xr[f] = v.x; xr[f + 1] = v.y;
    }
#else
#pragma unroll
    for (int f = 0; f < FP; f += 4) {
      float4 v = __ldg(reinterpret_cast<const float4*>(row + f));
      xr[f] = v.x; xr[f + 1] = v.y; xr[f + 2] = v.z; xr[f + 3] = v.w;
    }
#endif
    real e0 = 0, e1 = 0;
#pragma unroll
    for (int f = 0; f < FP; f += 2) {
      e0 = fma(xr[f], be[f], e0);
      e1 = fma(xr[f + 1], be[f + 1], e1);
    }
    const real eta = e0 + e1;
    const real yn = ldg(y + n);
    const real r = yn - r_sigmoid(eta);
    if (want_lp) lp += yn * eta - r_softplus(eta);
#pragma unroll
    for (int f = 0; f < FP; ++f) gb[f] = fma(xr[f], r, gb[f]);
  }
#pragma unroll
  for (int f = 0; f < FP; ++f) {
    real v = group_sum<LPC>(gb[f]);
    if ((f % LPC) == sub && f < m.F) g(beta_off + f) = v;
  }
  return group_sum<LPC>(lp);
}

// reference models.py:888-904 (lognormalcentered) and :930-945 (gammascale).
// z = [overall_log_scale, beta_log_scales[F], beta[F]]
template <int LPC, bool WITH_A, int FP, bool GAMMA>
__device__ real vg_german(const DevModel& m, const real* a, const real* b,
                          Vec z, Vec g, Vec xc, Vec abar, Vec bbar, int sub, bool want_lp) {
  const int F = m.F;
  real lp_top = 0;
  const real a0 = (*a), b0 = (*b);
  Site s0 = site_fwd(z(0), (real)0, ARP_LOG_10, a0, b0, lp_top);
  const real s0x = s0.x;
  real lp = 0;
  // forward: centred log-scales and coefficients
  for (int f = sub; f < F; f += LPC) {
    real dummy = 0;
    real ls;
    if (GAMMA) {
      const real v = z(1 + f);
      ls = s0x + v;
      xc(1 + f) = v;
    } else {
      Site ss = site_fwd_unit(z(1 + f), s0x, (*(a + 1 + f)), dummy);
      ls = ss.x;
      xc(1 + f) = ls;
    }
    Site sb = site_fwd(z(1 + F + f), (real)0, ls, (*(a + 1 + F + f)), (*(b + 1 + F + f)), dummy);
    xc(1 + F + f) = sb.x;
  }
  __syncwarp();
  lp = german_likelihood<LPC, FP>(m, xc, g, 1 + F, sub, want_lp);
  // reverse through beta -> log-scales -> overall scale
  real acc0 = 0, lp_sites = 0;
  for (int f = sub; f < F; f += LPC) {
    const real af = (*(a + 1 + f)), ab_ = (*(a + 1 + F + f)), bb_ = (*(b + 1 + F + f));
    real zb, mb, lb, ab;
    if (GAMMA) {
      const real v = z(1 + f);
      const real ls = s0x + v;
      Site sb = site_fwd(z(1 + F + f), (real)0, ls, ab_, bb_, lp_sites);
      const real xb = g(1 + F + f);
      site_rev(sb, xb, (real)0, ab_, bb_, zb, mb, lb, ab);
      g(1 + F + f) = zb;
      if (WITH_A) { abar(1 + F + f) = 0; abar(1 + f) = 0; bbar(1 + F + f) = site_bbar(sb, xb, ls); bbar(1 + f) = 0; }
      // log Gamma(1/2, 1/2) density of v: 0.5 v - 0.5 e^v + 0.5 log 0.5 - lgamma(0.5)
      const real ev = r_exp(v);
      lp_sites += (real)0.5 * v - (real)0.5 * ev + (real)(-0.34657359027997264 - 0.57236494292470008);
      g(1 + f) = (real)0.5 - (real)0.5 * ev + lb;
      acc0 += lb;
    } else {
      Site ss = site_fwd_unit(z(1 + f), s0x, af, lp_sites);
      Site sb = site_fwd(z(1 + F + f), (real)0, ss.x, ab_, bb_, lp_sites);
      const real xb = g(1 + F + f);
      site_rev(sb, xb, (real)0, ab_, bb_, zb, mb, lb, ab);
      g(1 + F + f) = zb;
      if (WITH_A) { abar(1 + F + f) = 0; bbar(1 + F + f) = site_bbar(sb, xb, ss.x); }
      real zb2, mb2, lb2, ab2;
      site_rev(ss, lb, s0x, af, (real)1, zb2, mb2, lb2, ab2);
      g(1 + f) = zb2;
      if (WITH_A) { abar(1 + f) = ab2; bbar(1 + f) = 0; }   // unit prior scale: no dependence on b
      acc0 += mb2;
    }
  }
  acc0 = group_sum<LPC>(acc0);
  lp += group_sum<LPC>(lp_sites) + lp_top;
  if (sub == 0) {
    real zb, mb, lb, ab;
    site_rev(s0, acc0, (real)0, a0, b0, zb, mb, lb, ab);
    g(0) = zb;
    xc(0) = s0x;
    if (WITH_A) { abar(0) = 0; bbar(0) = site_bbar(s0, acc0, ARP_LOG_10); }
  }
  return lp;
}

// ------------------------------------------------------------------ radon ---
// reference models.py:826-837 (sigma_y = 1) and :772-788 (per-county scales).
// z = [mua, b1, b2, m[J]] (+ log_m_stddv[J]); observations sorted by county, CSR.
template <int LPC, bool WITH_A, bool STDDVS>
__device__ real vg_radon(const DevModel& m, const real* a, const real* b,
                         Vec z, Vec g, Vec xc, Vec abar, Vec bbar, int sub, bool want_lp) {
  const int J = m.J;
  real lp_top = 0;
  const real a0 = (*a), a1 = (*(a + 1)), a2 = (*(a + 2));
  Site smua = site_fwd_unit(z(0), (real)0, a0, lp_top);
  Site sb1 = site_fwd_unit(z(1), (real)0, a1, lp_top);
  Site sb2 = site_fwd_unit(z(2), (real)0, a2, lp_top);
  const real mua = smua.x, b1 = sb1.x, b2 = sb2.x;
  real lp = 0, acc_mua = 0, acc_b1 = 0, acc_b2 = 0;
  double lik_d = 0;   // likelihood total in double: log-prob differences of O(0.1) matter to the accept test
  for (int j = sub; j < J; j += LPC) {
    const real uj = ldg(m.u + j);
    const real mu_j = mua + uj * b1;
    const real aj = (*(a + 3 + j));
    Site sm = site_fwd_unit(z(3 + j), mu_j, aj, lp);
    const real mj = sm.x;
    real lsj = 0, inv = 1, inv2 = 1;
    Site sl;
    if (STDDVS) {
      sl = site_fwd_unit(z(3 + J + j), (real)0, (*(a + 3 + J + j)), lp);
      lsj = sl.x;
      inv = r_exp(-lsj);
      inv2 = inv * inv;
    }
    // Gaussian likelihood of county j through its sufficient statistics (exact algebra, precomputed in
    // double at model-create time): with d = ybar - m_j - b2 xbar and within-county centred moments C..,
    //   sum e = n d,  sum e x = Cxy - b2 Cxx + n xbar d,  sum e^2 = Cyy - 2 b2 Cxy + b2^2 Cxx + n d^2.
    // The reference evaluates every observation through a dense one-hot matmul (models.py:834-837); the
    // roofline numerator stays the naive 6N + 8J flop of SURVEY.md 8d.
    const real* st = m.w + (size_t)6 * j;
    const real cnt = ldg(st), ybar = ldg(st + 1), xbar = ldg(st + 2), Cyy = ldg(st + 3), Cxy = ldg(st + 4), Cxx = ldg(st + 5);
    const real dj = ybar - mj - xbar * b2;
    const real se = cnt * dj;
    const real sex = Cxy - b2 * Cxx + cnt * xbar * dj;
    const real see = Cyy - (real)2 * b2 * Cxy + b2 * b2 * Cxx + cnt * dj * dj;
    lik_d += (double)((real)-0.5 * see * inv2 - cnt * (lsj + ARP_HALF_LOG_2PI));   // up to 1e4 counties of O(100) each
    acc_b2 += sex * inv2;
    real zb, mb, lb, ab;
    site_rev(sm, se * inv2, mu_j, aj, (real)1, zb, mb, lb, ab);
    g(3 + j) = zb;
    xc(3 + j) = mj;
    if (WITH_A) { abar(3 + j) = ab; bbar(3 + j) = 0; }
    acc_mua += mb;
    acc_b1 += uj * mb;
    if (STDDVS) {
      site_rev(sl, see * inv2 - cnt, (real)0, (real)0, (real)1, zb, mb, lb, ab);
      g(3 + J + j) = zb;
      xc(3 + J + j) = lsj;
      if (WITH_A) { abar(3 + J + j) = 0; bbar(3 + J + j) = 0; }
    }
  }
  acc_mua = group_sum<LPC>(acc_mua);
  acc_b1 = group_sum<LPC>(acc_b1);
  acc_b2 = group_sum<LPC>(acc_b2);
  lp = (real)(group_sum<LPC>((double)lp + lik_d)) + lp_top;
  if (sub == 0) {
    real zb, mb, lb, ab;
    site_rev(smua, acc_mua, (real)0, a0, (real)1, zb, mb, lb, ab);
    g(0) = zb; xc(0) = mua;
    site_rev(sb1, acc_b1, (real)0, a1, (real)1, zb, mb, lb, ab);
    g(1) = zb; xc(1) = b1;
    site_rev(sb2, acc_b2, (real)0, a2, (real)1, zb, mb, lb, ab);
    g(2) = zb; xc(2) = b2;
    if (WITH_A) { abar(0) = 0; abar(1) = 0; abar(2) = 0; bbar(0) = 0; bbar(1) = 0; bbar(2) = 0; }
  }
  return lp;
}

// Gradient sweep FUSED with the leapfrog kicks (SIMT HMC kernel, radon models).  The gradient of a per-county
// coordinate is local to its county, so the sweep that computes it can apply the kick on the spot instead of storing
// it for a second sweep: on a non-final leapfrog step `v += e g / 2` (end of this step), `v += e g / 2; x += e v` (start
// of the next) with the SAME two separately rounded half kicks as the unfused path, and neither the gradient nor the
// centred value is stored; on the final step the half kick only feeds the kinetic energy and (g, xc) are stored for
// the accept decision.  Per coordinate and leapfrog step: read x, v + write x, v instead of read x, v, g, x + write
// g, xc, v, x -- the synthetic 10^6 x 10^4 radon streams its state through HBM, so this is what it pays for.
// `z` is the proposal position (updated in place), `ke` accumulates this lane's share of |v|^2 on the final step.
template <int LPC, bool STDDVS>
__device__ real vg_radon_kick(const DevModel& m, const real* a, const real* b, Vec z, Vec g, Vec xc, Vec v,
                              const real* __restrict__ eps0, real mult, int sub, bool last, real& ke) {
  const int J = m.J;
  real lp_top = 0;
  const real a0 = (*a), a1 = (*(a + 1)), a2 = (*(a + 2));
  Site smua = site_fwd_unit(z(0), (real)0, a0, lp_top);
  Site sb1 = site_fwd_unit(z(1), (real)0, a1, lp_top);
  Site sb2 = site_fwd_unit(z(2), (real)0, a2, lp_top);
  const real mua = smua.x, b1 = sb1.x, b2 = sb2.x;
  auto kick = [&](int d, real grad, real centred) {
    const real e = ldg(eps0 + d) * mult;
    real vv = v(d) + (real)0.5 * e * grad;
    if (last) {
      ke = fma(vv, vv, ke);
      g(d) = grad;
      xc(d) = centred;
    } else {
      vv = vv + (real)0.5 * e * grad;
      v(d) = vv;
      z(d) = z(d) + e * vv;
    }
  };
  __syncwarp();   // every lane has read the three top-level coordinates before lane 0 moves them (below)
  real lp = 0, acc_mua = 0, acc_b1 = 0, acc_b2 = 0;
  double lik_d = 0;
  for (int j = sub; j < J; j += LPC) {
    const real uj = ldg(m.u + j);
    const real mu_j = mua + uj * b1;
    const real aj = (*(a + 3 + j));
    Site sm = site_fwd_unit(z(3 + j), mu_j, aj, lp);
    const real mj = sm.x;
    real lsj = 0, inv = 1, inv2 = 1;
    Site sl;
    if (STDDVS) {
      sl = site_fwd_unit(z(3 + J + j), (real)0, (*(a + 3 + J + j)), lp);
      lsj = sl.x;
      inv = r_exp(-lsj);
      inv2 = inv * inv;
    }
    const real* st = m.w + (size_t)6 * j;
    const real cnt = ldg(st), ybar = ldg(st + 1), xbar = ldg(st + 2), Cyy = ldg(st + 3), Cxy = ldg(st + 4), Cxx = ldg(st + 5);
    const real dj = ybar - mj - xbar * b2;
    const real se = cnt * dj;
    const real sex = Cxy - b2 * Cxx + cnt * xbar * dj;
    const real see = Cyy - (real)2 * b2 * Cxy + b2 * b2 * Cxx + cnt * dj * dj;
    lik_d += (double)((real)-0.5 * see * inv2 - cnt * (lsj + ARP_HALF_LOG_2PI));
    acc_b2 += sex * inv2;
    real zb, mb, lb, ab;
    site_rev(sm, se * inv2, mu_j, aj, (real)1, zb, mb, lb, ab);
    kick(3 + j, zb, mj);
    acc_mua += mb;
    acc_b1 += uj * mb;
    if (STDDVS) {
      site_rev(sl, see * inv2 - cnt, (real)0, (real)0, (real)1, zb, mb, lb, ab);
      kick(3 + J + j, zb, lsj);
    }
  }
  acc_mua = group_sum<LPC>(acc_mua);
  acc_b1 = group_sum<LPC>(acc_b1);
  acc_b2 = group_sum<LPC>(acc_b2);
  lp = (real)(group_sum<LPC>((double)lp + lik_d)) + lp_top;
  if (sub == 0) {
    real zb, mb, lb, ab;
    site_rev(smua, acc_mua, (real)0, a0, (real)1, zb, mb, lb, ab);
    kick(0, zb, mua);
    site_rev(sb1, acc_b1, (real)0, a1, (real)1, zb, mb, lb, ab);
    kick(1, zb, b1);
    site_rev(sb2, acc_b2, (real)0, a2, (real)1, zb, mb, lb, ab);
    kick(2, zb, b2);
  }
  return lp;
}

// --------------------------------------------------------------- election ---
// reference models.py:969-982.  z = [mua, log_sigma_a, a[K], b1, b2].
// Observations are grouped by the one-hot index of `state` (group K = "no
// column hit", reference feeds 1-based indices to tf.one_hot(depth=K)) and
// identical (group, female, black) rows are merged into weighted cells
// (w = count, y = sum of y): sum over equal-eta observations, exact.
template <int LPC, bool WITH_A>
__device__ real vg_election(const DevModel& m, const real* a, const real* b,
                            Vec z, Vec g, Vec xc, Vec abar, Vec bbar, int sub, bool want_lp) {
  const int K = m.K;
  real lp_top = 0;
  const real a0 = (*a), b0 = (*b), a1 = (*(a + 1)), bb1 = (*(b + 1));
  const real a3 = (*(a + 2 + K)), b3 = (*(b + 2 + K)), a4 = (*(a + 3 + K)), b4 = (*(b + 3 + K));
  Site smua = site_fwd(z(0), (real)0, ARP_LOG_100, a0, b0, lp_top);
  Site slsa = site_fwd(z(1), (real)0, ARP_LOG_10, a1, bb1, lp_top);
  Site sb1 = site_fwd(z(2 + K), (real)0, ARP_LOG_100, a3, b3, lp_top);
  Site sb2 = site_fwd(z(3 + K), (real)0, ARP_LOG_100, a4, b4, lp_top);
  const real mua = smua.x, lsa = slsa.x, b1 = sb1.x, b2 = sb2.x;
  real lp = 0, acc_mua = 0, acc_lsa = 0, acc_b1 = 0, acc_b2 = 0;
  for (int k = sub; k < m.J; k += LPC) {
    real ak = 0, aa = 0, ba = 0;
    Site sa;
    if (k < K) {
      aa = (*(a + 2 + k));
      ba = (*(b + 2 + k));
      sa = site_fwd(z(2 + k), mua, lsa, aa, ba, lp);
      ak = sa.x;
    }
    const int c0 = ldg(m.offs + k), c1 = ldg(m.offs + k + 1);
    real abar_k = 0;
    for (int c = c0; c < c1; ++c) {
      const real fe = ldg(m.x1 + c), bl = ldg(m.x2 + c), w = ldg(m.w + c), ys = ldg(m.y + c);
      const real eta = ak + fe * b2 + bl * b1;
      const real r = ys - w * r_sigmoid(eta);
      if (want_lp) lp += ys * eta - w * r_softplus(eta);
      abar_k += r;
      acc_b2 = fma(fe, r, acc_b2);
      acc_b1 = fma(bl, r, acc_b1);
    }
    if (k < K) {
      real zb, mb, lb, ab;
      site_rev(sa, abar_k, mua, aa, ba, zb, mb, lb, ab);
      g(2 + k) = zb;
      xc(2 + k) = ak;
      if (WITH_A) { abar(2 + k) = ab; bbar(2 + k) = site_bbar(sa, abar_k, lsa); }
      acc_mua += mb;
      acc_lsa += lb;
    }
  }
  acc_mua = group_sum<LPC>(acc_mua);
  acc_lsa = group_sum<LPC>(acc_lsa);
  acc_b1 = group_sum<LPC>(acc_b1);
  acc_b2 = group_sum<LPC>(acc_b2);
  lp = group_sum<LPC>(lp) + lp_top;
  if (sub == 0) {
    real zb, mb, lb, ab;
    site_rev(smua, acc_mua, (real)0, a0, b0, zb, mb, lb, ab);
    g(0) = zb; xc(0) = mua;
    site_rev(slsa, acc_lsa, (real)0, a1, bb1, zb, mb, lb, ab);
    g(1) = zb; xc(1) = lsa;
    site_rev(sb1, acc_b1, (real)0, a3, b3, zb, mb, lb, ab);
    g(2 + K) = zb; xc(2 + K) = b1;
    site_rev(sb2, acc_b2, (real)0, a4, b4, zb, mb, lb, ab);
    g(3 + K) = zb; xc(3 + K) = b2;
    if (WITH_A) {
      abar(0) = 0; abar(1) = 0; abar(2 + K) = 0; abar(3 + K) = 0;
      bbar(0) = site_bbar(smua, acc_mua, ARP_LOG_100); bbar(1) = site_bbar(slsa, acc_lsa, ARP_LOG_10);
      bbar(2 + K) = site_bbar(sb1, acc_b1, ARP_LOG_100); bbar(3 + K) = site_bbar(sb2, acc_b2, ARP_LOG_100);
    }
  }
  return lp;
}

// --------------------------------------------------------------- electric ---
// reference models.py:1013-1035.  z = [mua[4], sigma_y[4], a[K=96], b[4]];
// observations grouped by pair index (group K = out of range); grade /
// grade_pair indices are -1 where the reference's one-hot row is all zero.
template <int LPC, bool WITH_A>
__device__ real vg_electric(const DevModel& m, const real* a, const real* b,
                            Vec z, Vec g, Vec xc, Vec abar, Vec bbar, int sub, bool want_lp) {
  const int K = m.K;
  const int oB = 8 + K;
  real lp_top = 0;
  real mua[4], sy[4], bx[4];
  Site s_mua[4], s_sy[4], s_b[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    s_mua[q] = site_fwd_unit(z(q), (real)0, (*(a + q)), lp_top);
    s_sy[q] = site_fwd_unit(z(4 + q), (real)0, (*(a + 4 + q)), lp_top);
    s_b[q] = site_fwd(z(oB + q), (real)0, ARP_LOG_100, (*(a + oB + q)), (*(b + oB + q)), lp_top);
    mua[q] = s_mua[q].x; sy[q] = s_sy[q].x; bx[q] = s_b[q].x;
  }
  real lp = 0;
  real acc_mua[4] = {0, 0, 0, 0}, acc_sy[4] = {0, 0, 0, 0}, acc_b[4] = {0, 0, 0, 0};
  for (int p = sub; p < m.J; p += LPC) {
    real ap = 0, ap_a = 0, mu_p = 0;
    int gp = -1;
    Site sa;
    if (p < K) {
      gp = ldg(m.pidx + p);
#pragma unroll
      for (int q = 0; q < 4; ++q) if (gp == q) mu_p = (real)100 * mua[q];
      ap_a = (*(a + 8 + p));
      sa = site_fwd_unit(z(8 + p), mu_p, ap_a, lp);
      ap = sa.x;
    }
    const int n0 = ldg(m.offs + p), n1 = ldg(m.offs + p + 1);
    real abar_p = 0;
    for (int n = n0; n < n1; ++n) {
      const int gi = ldg(m.gidx + n);
      real bb = 0, lsy = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) if (gi == q) { bb = bx[q]; lsy = sy[q]; }
      const real tr = ldg(m.x1 + n);
      const real inv = r_exp(-lsy);
      const real e = (ldg(m.y + n) - ap - bb * tr) * inv;
      lp += (real)-0.5 * e * e - lsy - ARP_HALF_LOG_2PI;
      const real eb = e * inv;
      abar_p += eb;
#pragma unroll
      for (int q = 0; q < 4; ++q) if (gi == q) { acc_b[q] = fma(eb, tr, acc_b[q]); acc_sy[q] += e * e - (real)1; }
    }
    if (p < K) {
      real zb, mb, lb, ab;
      site_rev(sa, abar_p, mu_p, ap_a, (real)1, zb, mb, lb, ab);
      g(8 + p) = zb;
      xc(8 + p) = ap;
      if (WITH_A) { abar(8 + p) = ab; bbar(8 + p) = 0; }
#pragma unroll
      for (int q = 0; q < 4; ++q) if (gp == q) acc_mua[q] += (real)100 * mb;
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    acc_mua[q] = group_sum<LPC>(acc_mua[q]);
    acc_sy[q] = group_sum<LPC>(acc_sy[q]);
    acc_b[q] = group_sum<LPC>(acc_b[q]);
  }
  lp = group_sum<LPC>(lp) + lp_top;
  if (sub == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      real zb, mb, lb, ab;
      site_rev(s_mua[q], acc_mua[q], (real)0, (*(a + q)), (real)1, zb, mb, lb, ab);
      g(q) = zb; xc(q) = mua[q];
      site_rev(s_sy[q], acc_sy[q], (real)0, (*(a + 4 + q)), (real)1, zb, mb, lb, ab);
      g(4 + q) = zb; xc(4 + q) = sy[q];
      site_rev(s_b[q], acc_b[q], (real)0, (*(a + oB + q)), (*(b + oB + q)), zb, mb, lb, ab);
      g(oB + q) = zb; xc(oB + q) = bx[q];
      if (WITH_A) {
        abar(q) = 0; abar(4 + q) = 0; abar(oB + q) = 0;
        bbar(q) = 0; bbar(4 + q) = 0; bbar(oB + q) = site_bbar(s_b[q], acc_b[q], ARP_LOG_100);
      }
    }
  }
  return lp;
}

// ------------------------------------------------------------ time series ---
// reference models.py:1071-1096.  z = [sigma_alpha, sigma_mu, alpha0, mu0,
// alpha1, mu1, ..., alpha_{T-1}, mu_{T-1}, beta]; local linear trend, forward
// scan for the centred values and reverse scan for the adjoints.  The scan is
// sequential per chain: with LPC > 1 every lane runs it redundantly and lane 0
// writes.
// The scans run in DOUBLE in the fp32 build as well (state, gradient and centred values stay `real`): the level alpha_t
// absorbs -beta * year (|beta * year| ~ 2e3 for the raw years 1959 .. 2018, models.py:1098-1113) against an observation
// scale of 0.12, so one fp32 ulp of alpha_t is 1e-3 standard deviations of the residual and 60 accumulated roundings of
// the level put the fp32 log-joint / gradient at 2e-4 of the fp64 oracle (round 1 waived the tolerance); with the
// recurrences and the residual in double the fp32 build meets 1e-5.  ~2e3 flop per evaluation: the double rate is
// irrelevant next to the state traffic.
template <int LPC, bool WITH_A>
__device__ real vg_time_series_seq(const DevModel& m, const real* a, const real* b,
                                   Vec z, Vec g, Vec xc, Vec abar, Vec bbar, int sub, bool want_lp) {
  typedef double acc_t;
  typedef SiteT<acc_t> SiteA;
  const int T = m.K;
  const int oBeta = 2 + 2 * T;
  const bool wr = (sub == 0);
  acc_t lp = 0;
  auto A = [&](int i) { return (acc_t)(*(a + i)); };
  auto B = [&](int i) { return (acc_t)(*(b + i)); };
  SiteA s_sa = site_fwd_unit((acc_t)z(0), (acc_t)0, A(0), lp);
  SiteA s_sm = site_fwd_unit((acc_t)z(1), (acc_t)0, A(1), lp);
  SiteA s_be = site_fwd_unit((acc_t)z(oBeta), (acc_t)0, A(oBeta), lp);
  const acc_t sa = s_sa.x, sm = s_sm.x, be = s_be.x;
  const acc_t sig_a = r_softplus(sa), sig_m = r_softplus(sm);
  const acc_t lsa = r_log(sig_a), lsm = r_log(sig_m);
  const acc_t obs_sd = (acc_t)0.12f;  // the reference's scale=0.12 is a float32 graph constant
  const acc_t inv_obs = (acc_t)1 / obs_sd;
  const acc_t log_obs = r_log(obs_sd);
  // forward scan
  acc_t al_prev = 0, mu_prev = 0;
  for (int t = 0; t < T; ++t) {
    const int ia = 2 + 2 * t, im = 3 + 2 * t;
    SiteA s_al = site_fwd((acc_t)z(ia), al_prev + mu_prev, lsa, A(ia), B(ia), lp);
    SiteA s_mu = site_fwd((acc_t)z(im), mu_prev, lsm, A(im), B(im), lp);
    if (wr) {
      // centred values as a two-term sum: the head is the `real` output, the tail is parked in the gradient slot of the
      // same coordinate (free until the reverse scan reaches it) so that the reverse scan sees them in full precision
      const real ah = (real)s_al.x, mh = (real)s_mu.x;
      xc(ia) = ah; xc(im) = mh;
      g(ia) = (real)(s_al.x - (acc_t)ah); g(im) = (real)(s_mu.x - (acc_t)mh);
    }
    al_prev = s_al.x;
    mu_prev = s_mu.x;
    const acc_t e = ((acc_t)ldg(m.y + t) - s_al.x - be * (acc_t)ldg(m.x1 + t)) * inv_obs;
    lp += (acc_t)-0.5 * e * e - log_obs - (acc_t)ARP_HALF_LOG_2PI_D;
  }
  __syncwarp();
  // reverse scan
  acc_t carry_al = 0, carry_mu = 0, acc_lsa = 0, acc_lsm = 0, acc_be = 0;
  for (int t = T - 1; t >= 0; --t) {
    const int ia = 2 + 2 * t, im = 3 + 2 * t;
    const acc_t aa = A(ia), ba = B(ia), am = A(im), bm = B(im);
    const acc_t alp = (t > 0) ? (acc_t)xc(ia - 2) + (acc_t)g(ia - 2) : (acc_t)0;
    const acc_t mup = (t > 0) ? (acc_t)xc(im - 2) + (acc_t)g(im - 2) : (acc_t)0;
    __syncwarp();   // LPC > 1: every lane has read these slots before lane 0 overwrites them in the next iteration
    acc_t dummy = 0;
    SiteA s_al = site_fwd((acc_t)z(ia), alp + mup, lsa, aa, ba, dummy);
    SiteA s_mu = site_fwd((acc_t)z(im), mup, lsm, am, bm, dummy);
    const acc_t xt = (acc_t)ldg(m.x1 + t);
    const acc_t e = ((acc_t)ldg(m.y + t) - s_al.x - be * xt) * inv_obs;
    const acc_t lik = e * inv_obs;
    acc_be = fma(lik, xt, acc_be);
    acc_t zb, mb_al, lb, ab;
    site_rev(s_al, lik + carry_al, alp + mup, aa, ba, zb, mb_al, lb, ab);
    if (wr) { g(ia) = (real)zb; if (WITH_A) { abar(ia) = (real)ab; bbar(ia) = (real)site_bbar(s_al, lik + carry_al, lsa); } }
    acc_lsa += lb;
    acc_t mb_mu;
    site_rev(s_mu, carry_mu, mup, am, bm, zb, mb_mu, lb, ab);
    if (wr) { g(im) = (real)zb; if (WITH_A) { abar(im) = (real)ab; bbar(im) = (real)site_bbar(s_mu, carry_mu, lsm); } }
    acc_lsm += lb;
    carry_al = mb_al;
    carry_mu = mb_al + mb_mu;
  }
  if (wr) {
    acc_t zb, mb, lb, ab;
    // d log softplus(s) / d s = sigmoid(s) / softplus(s)
    const acc_t dsa = ((acc_t)1 / ((acc_t)1 + r_exp(-sa))) / sig_a;
    const acc_t dsm = ((acc_t)1 / ((acc_t)1 + r_exp(-sm))) / sig_m;
    site_rev(s_sa, acc_lsa * dsa, (acc_t)0, A(0), (acc_t)1, zb, mb, lb, ab);
    g(0) = (real)zb; xc(0) = (real)sa;
    site_rev(s_sm, acc_lsm * dsm, (acc_t)0, A(1), (acc_t)1, zb, mb, lb, ab);
    g(1) = (real)zb; xc(1) = (real)sm;
    site_rev(s_be, acc_be, (acc_t)0, A(oBeta), (acc_t)1, zb, mb, lb, ab);
    g(oBeta) = (real)zb; xc(oBeta) = (real)be;
    if (WITH_A) { abar(0) = 0; abar(1) = 0; abar(oBeta) = 0; bbar(0) = 0; bbar(1) = 0; bbar(oBeta) = 0; }
  }
  return (real)lp;
}

// Lane-parallel evaluation (LPC > 1).  Under ANY rule (a, b) the centred values obey an AFFINE recurrence in the previous
// centred values, because the scales do not depend on them: with r = sigma^(1 - b), c = 1 - r a, d = r z,
//   alpha_t = c_a (alpha_{t-1} + mu_{t-1}) + d_a,      mu_t = c_m mu_{t-1} + d_m,
// i.e. a map (alpha, mu) -> (p alpha + q mu + u, s mu + w), and such maps compose.  Lane `sub` owns K = ceil(T / LPC)
// consecutive time steps: it composes its steps, an inclusive Hillis-Steele scan over the LPC lanes of the chain gives the
// state entering every segment, and each lane then walks its own steps.  The adjoint carries obey the transposed
// recurrence (ca' = c_a (lik + ca) + a_a usb_a,  cm' = ca' + c_m cm + a_m usb_m), scanned from the right the same way.
// All of it in double (see above): 60 dependent steps become 2 K local steps + 2 log2(LPC) shuffle rounds.
template <int LPC, bool WITH_A>
__device__ real vg_time_series_par(const DevModel& m, const real* a, const real* b,
                                   Vec z, Vec g, Vec xc, Vec abar, Vec bbar, int sub, bool want_lp) {
  typedef double acc_t;
  typedef SiteT<acc_t> SiteA;
  constexpr int KMAX = 64 / LPC;            // the caller guarantees T <= 64
  constexpr unsigned FULL = 0xffffffffu;
  const int T = m.K;
  const int oBeta = 2 + 2 * T;
  const int K = (T + LPC - 1) / LPC;
  const int t0 = sub * K, t1 = (t0 + K < T) ? t0 + K : T;
  auto A = [&](int i) { return (acc_t)(*(a + i)); };
  auto B = [&](int i) { return (acc_t)(*(b + i)); };
  acc_t lp_top = 0;
  SiteA s_sa = site_fwd_unit((acc_t)z(0), (acc_t)0, A(0), lp_top);
  SiteA s_sm = site_fwd_unit((acc_t)z(1), (acc_t)0, A(1), lp_top);
  SiteA s_be = site_fwd_unit((acc_t)z(oBeta), (acc_t)0, A(oBeta), lp_top);
  const acc_t sa = s_sa.x, sm = s_sm.x, be = s_be.x;
  const acc_t sig_a = r_softplus(sa), sig_m = r_softplus(sm);
  const acc_t lsa = r_log(sig_a), lsm = r_log(sig_m);
  const acc_t obs_sd = (acc_t)0.12f;
  const acc_t inv_obs = (acc_t)1 / obs_sd;
  const acc_t log_obs = r_log(obs_sd);
  auto rfac = [&](acc_t bb, acc_t ls, acc_t sig) {
    return bb == (acc_t)1 ? (acc_t)1 : (bb == (acc_t)0 ? sig : r_exp(((acc_t)1 - bb) * ls));
  };
  // ---- 1. this lane's steps as one affine map (alpha, mu) -> (p alpha + q mu + u, s mu + w)
  acc_t p = 1, q = 0, s = 1, u = 0, w = 0;
  for (int t = t0; t < t1; ++t) {
    const int ia = 2 + 2 * t, im = 3 + 2 * t;
    const acc_t ra = rfac(B(ia), lsa, sig_a), rm = rfac(B(im), lsm, sig_m);
    const acc_t ca = (acc_t)1 - ra * A(ia), cm = (acc_t)1 - rm * A(im);
    const acc_t da = ra * (acc_t)z(ia), dm = rm * (acc_t)z(im);
    u = ca * (u + w) + da; q = ca * (q + s); p = ca * p;
    w = cm * w + dm; s = cm * s;
  }
  // ---- 2. inclusive scan over the lanes of the chain: (mine) o (everything to my left)
#pragma unroll
  for (int o = 1; o < LPC; o <<= 1) {
    const acc_t pE = __shfl_up_sync(FULL, p, o, LPC), qE = __shfl_up_sync(FULL, q, o, LPC);
    const acc_t sE = __shfl_up_sync(FULL, s, o, LPC), uE = __shfl_up_sync(FULL, u, o, LPC);
    const acc_t wE = __shfl_up_sync(FULL, w, o, LPC);
    if (sub >= o) {
      u = p * uE + q * wE + u; q = p * qE + q * sE; p = p * pE;
      w = s * wE + w; s = s * sE;
    }
  }
  acc_t al = __shfl_up_sync(FULL, u, 1, LPC), mu = __shfl_up_sync(FULL, w, 1, LPC);   // state entering my segment
  if (sub == 0) { al = 0; mu = 0; }
  // ---- 3. walk my steps: centred values, log-density terms, and the adjoint map of the segment
  //         (ca, cm) -> (P ca + U, R ca + S cm + W), built left to right: F <- F o B_t (B_t is applied first).
  // MIXED PRECISION from here on: only what lives on the scale of the level (|alpha| ~ 2e3: the means m, the site
  // offsets dz = z - a m where z ~ m cancels, the centred values and the residual) is formed in double; everything
  // derived from dz and the residual (standardised offsets, log-density terms, adjoints and their scan) is `real`.
  typedef real lo_t;
  typedef SiteT<lo_t> SiteL;
  const lo_t lsa_l = (lo_t)lsa, lsm_l = (lo_t)lsm, sig_a_l = (lo_t)sig_a, sig_m_l = (lo_t)sig_m;
  const lo_t inv_obs_l = (lo_t)inv_obs;
  // one site from its (double) offset: scales in `real`
  auto site_lo = [&](acc_t dz_acc, lo_t ls, lo_t sig, lo_t bb, lo_t& lpl) {
    SiteL sl;
    lo_t sb_inv;
    if (bb == (lo_t)1) { sb_inv = (lo_t)1 / sig; sl.r = (lo_t)1; }
    else if (bb == (lo_t)0) { sb_inv = (lo_t)1; sl.r = sig; }
    else { sb_inv = r_exp(-bb * ls); sl.r = r_exp(((lo_t)1 - bb) * ls); }
    sl.dz = (lo_t)dz_acc;
    const lo_t uu = sl.dz * sb_inv;
    sl.usb = uu * sb_inv;
    sl.x = 0;   // the centred value is kept in double by the caller
    lpl += (lo_t)-0.5 * uu * uu - bb * ls - ARP_HALF_LOG_2PI;
    return sl;
  };
  lo_t dz_al[KMAX], dz_mu[KMAX], lik_k[KMAX], mm_al[KMAX], mm_mu[KMAX];
  lo_t P = 1, U = 0, R = 0, S = 1, W = 0;
  lo_t lp_l = 0;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int t = t0 + k;
    dz_al[k] = 0; dz_mu[k] = 0; lik_k[k] = 0; mm_al[k] = 0; mm_mu[k] = 0;
    if (t < t1) {
      const int ia = 2 + 2 * t, im = 3 + 2 * t;
      const lo_t aa = (*(a + ia)), am = (*(a + im));
      const acc_t m_a = al + mu, m_m = mu;
      const acc_t dza = (acc_t)z(ia) - (acc_t)aa * m_a, dzm = (acc_t)z(im) - (acc_t)am * m_m;
      SiteL s_al = site_lo(dza, lsa_l, sig_a_l, (*(b + ia)), lp_l);
      SiteL s_mu = site_lo(dzm, lsm_l, sig_m_l, (*(b + im)), lp_l);
      al = m_a + (acc_t)s_al.r * dza;
      mu = m_m + (acc_t)s_mu.r * dzm;
      xc(ia) = (real)al; xc(im) = (real)mu;
      const lo_t e = (lo_t)(((acc_t)ldg(m.y + t) - al - be * (acc_t)ldg(m.x1 + t)) * inv_obs);
      lp_l += (lo_t)-0.5 * e * e - (lo_t)log_obs - ARP_HALF_LOG_2PI;
      const lo_t lik = e * inv_obs_l;
      dz_al[k] = s_al.dz; dz_mu[k] = s_mu.dz; lik_k[k] = lik;
      if (WITH_A) { mm_al[k] = (lo_t)m_a; mm_mu[k] = (lo_t)m_m; }
      const lo_t c_al = (lo_t)1 - s_al.r * aa, c_mu = (lo_t)1 - s_mu.r * am;
      const lo_t h_al = c_al * lik + aa * s_al.usb, h_mu = am * s_mu.usb;
      W = R * h_al + S * (h_al + h_mu) + W; R = (R + S) * c_al; S = S * c_mu;
      U = P * h_al + U; P = P * c_al;
    }
  }
  acc_t lp = (acc_t)lp_l;
  // ---- 4. inclusive scan from the right: (mine) o (everything to my right, applied first)
#pragma unroll
  for (int o = 1; o < LPC; o <<= 1) {
    const lo_t PE = __shfl_down_sync(FULL, P, o, LPC), UE = __shfl_down_sync(FULL, U, o, LPC);
    const lo_t RE = __shfl_down_sync(FULL, R, o, LPC), SE = __shfl_down_sync(FULL, S, o, LPC);
    const lo_t WE = __shfl_down_sync(FULL, W, o, LPC);
    if (sub + o < LPC) {
      W = R * UE + S * WE + W; R = R * PE + S * RE; S = S * SE;
      U = P * UE + U; P = P * PE;
    }
  }
  lo_t ca = __shfl_down_sync(FULL, U, 1, LPC), cm = __shfl_down_sync(FULL, W, 1, LPC);   // carries entering from the right
  if (sub == LPC - 1) { ca = 0; cm = 0; }
  // ---- 5. walk my steps backwards
  lo_t acc_lsa_l = 0, acc_lsm_l = 0, acc_be_l = 0;
#pragma unroll
  for (int k = KMAX - 1; k >= 0; --k) {
    const int t = t0 + k;
    if (t < t1) {
      const int ia = 2 + 2 * t, im = 3 + 2 * t;
      const lo_t aa = (*(a + ia)), ba = (*(b + ia)), am = (*(a + im)), bm = (*(b + im));
      lo_t dummy = 0;
      SiteL s_al = site_lo((acc_t)dz_al[k], lsa_l, sig_a_l, ba, dummy);
      SiteL s_mu = site_lo((acc_t)dz_mu[k], lsm_l, sig_m_l, bm, dummy);
      const lo_t lik = lik_k[k];
      acc_be_l = fma(lik, (lo_t)ldg(m.x1 + t), acc_be_l);
      lo_t zb, mb_al, lb, ab;
      site_rev(s_al, lik + ca, mm_al[k], aa, ba, zb, mb_al, lb, ab);
      g(ia) = (real)zb;
      if (WITH_A) { abar(ia) = (real)ab; bbar(ia) = (real)site_bbar(s_al, lik + ca, lsa_l); }
      acc_lsa_l += lb;
      lo_t mb_mu;
      site_rev(s_mu, cm, mm_mu[k], am, bm, zb, mb_mu, lb, ab);
      g(im) = (real)zb;
      if (WITH_A) { abar(im) = (real)ab; bbar(im) = (real)site_bbar(s_mu, cm, lsm_l); }
      acc_lsm_l += lb;
      ca = mb_al;
      cm = mb_al + mb_mu;
    }
  }
  acc_t acc_lsa = (acc_t)acc_lsa_l, acc_lsm = (acc_t)acc_lsm_l, acc_be = (acc_t)acc_be_l;
  acc_lsa = group_sum<LPC>(acc_lsa);
  acc_lsm = group_sum<LPC>(acc_lsm);
  acc_be = group_sum<LPC>(acc_be);
  lp = group_sum<LPC>(lp) + lp_top;
  if (sub == 0) {
    acc_t zb, mb, lb, ab;
    const acc_t dsa = ((acc_t)1 / ((acc_t)1 + r_exp(-sa))) / sig_a;
    const acc_t dsm = ((acc_t)1 / ((acc_t)1 + r_exp(-sm))) / sig_m;
    site_rev(s_sa, acc_lsa * dsa, (acc_t)0, A(0), (acc_t)1, zb, mb, lb, ab);
    g(0) = (real)zb; xc(0) = (real)sa;
    site_rev(s_sm, acc_lsm * dsm, (acc_t)0, A(1), (acc_t)1, zb, mb, lb, ab);
    g(1) = (real)zb; xc(1) = (real)sm;
    site_rev(s_be, acc_be, (acc_t)0, A(oBeta), (acc_t)1, zb, mb, lb, ab);
    g(oBeta) = (real)zb; xc(oBeta) = (real)be;
    if (WITH_A) { abar(0) = 0; abar(1) = 0; abar(oBeta) = 0; bbar(0) = 0; bbar(1) = 0; bbar(oBeta) = 0; }
  }
  return (real)lp;
}

template <int LPC, bool WITH_A>
__device__ __forceinline__ real vg_time_series(const DevModel& m, const real* a, const real* b,
                                               Vec z, Vec g, Vec xc, Vec abar, Vec bbar, int sub, bool want_lp) {
  if constexpr (LPC > 1) {
    if (m.K <= 64) return vg_time_series_par<LPC, WITH_A>(m, a, b, z, g, xc, abar, bbar, sub, want_lp);
  }
  return vg_time_series_seq<LPC, WITH_A>(m, a, b, z, g, xc, abar, bbar, sub, want_lp);
}

// ------------------------------------------------- centred -> rule (a, b) ---
// Inverse of the site rule: given the CENTRED values x of every coordinate, the state coordinate under
// rule (a, b) is  z = a mu + (x - mu) sigma^(b-1)  with (mu, sigma) evaluated at the centred values of the
// parents.  a = b = 0 is the reference's make_to_noncentered (models.py:84-102), general (a, b) its
// build_make_to_partially_noncentered (models.py:105-128).  Used by the interleaved CP/NCP sampler.
__device__ __forceinline__ real site_inv(real x, real mu, real ls, real a, real b) {
  if (a == (real)1 && b == (real)1) return x;   // CP: the state IS the centred value (exactly)
  return a * mu + (x - mu) * r_exp((b - (real)1) * ls);
}

template <int KIND, int LPC>
__device__ void to_rule(const DevModel& m, const real* a, const real* b, Vec x, Vec z, int sub) {
  if (KIND == MODEL_8SCHOOLS) {
    const real mu = x(0), lt = x(1);
    if (sub == 0) { z(0) = site_inv(mu, 0, ARP_LOG_5, a[0], b[0]); z(1) = site_inv(lt, 0, ARP_LOG_5, a[1], b[1]); }
    for (int i = sub; i < 8; i += LPC) z(2 + i) = site_inv(x(2 + i), mu, lt, a[2 + i], b[2 + i]);
  } else if (KIND == MODEL_GERMAN_LOGNORMAL || KIND == MODEL_GERMAN_GAMMA) {
    const int F = m.F;
    const real s0 = x(0);
    if (sub == 0) z(0) = site_inv(s0, 0, ARP_LOG_10, a[0], b[0]);
    for (int f = sub; f < F; f += LPC) {
      const real sf = x(1 + f);
      real ls;
      if (KIND == MODEL_GERMAN_GAMMA) { z(1 + f) = sf; ls = s0 + sf; }
      else { z(1 + f) = site_inv(sf, s0, 0, a[1 + f], b[1 + f]); ls = sf; }
      z(1 + F + f) = site_inv(x(1 + F + f), 0, ls, a[1 + F + f], b[1 + F + f]);
    }
  } else if (KIND == MODEL_RADON || KIND == MODEL_RADON_STDDVS) {
    const int J = m.J;
    const real mua = x(0), b1 = x(1);
    if (sub == 0) { z(0) = x(0); z(1) = x(1); z(2) = x(2); }   // mu = 0, sigma = 1
    for (int j = sub; j < J; j += LPC) {
      z(3 + j) = site_inv(x(3 + j), mua + ldg(m.u + j) * b1, 0, a[3 + j], b[3 + j]);
      if (KIND == MODEL_RADON_STDDVS) z(3 + J + j) = x(3 + J + j);
    }
  } else if (KIND == MODEL_ELECTION) {
    const int K = m.K;
    const real mua = x(0), lsa = x(1);
    if (sub == 0) {
      z(0) = site_inv(mua, 0, ARP_LOG_100, a[0], b[0]);
      z(1) = site_inv(lsa, 0, ARP_LOG_10, a[1], b[1]);
      z(2 + K) = site_inv(x(2 + K), 0, ARP_LOG_100, a[2 + K], b[2 + K]);
      z(3 + K) = site_inv(x(3 + K), 0, ARP_LOG_100, a[3 + K], b[3 + K]);
    }
    for (int k = sub; k < K; k += LPC) z(2 + k) = site_inv(x(2 + k), mua, lsa, a[2 + k], b[2 + k]);
  } else if (KIND == MODEL_ELECTRIC) {
    const int K = m.K, oB = 8 + K;
    if (sub == 0)
      for (int q = 0; q < 4; ++q) {
        z(q) = x(q); z(4 + q) = x(4 + q);
        z(oB + q) = site_inv(x(oB + q), 0, ARP_LOG_100, a[oB + q], b[oB + q]);
      }
    for (int p = sub; p < K; p += LPC) {
      const int gp = ldg(m.pidx + p);
      const real mu_p = gp >= 0 ? (real)100 * x(gp) : (real)0;
      z(8 + p) = site_inv(x(8 + p), mu_p, 0, a[8 + p], b[8 + p]);
    }
  } else {  // time series
    const int T = m.K, oBeta = 2 + 2 * T;
    const real lsa = r_log(r_softplus(x(0))), lsm = r_log(r_softplus(x(1)));
    if (sub == 0) { z(0) = x(0); z(1) = x(1); z(oBeta) = x(oBeta); }
    for (int t = sub; t < T; t += LPC) {
      const int ia = 2 + 2 * t, im = 3 + 2 * t;
      const real alp = t > 0 ? x(ia - 2) : (real)0, mup = t > 0 ? x(im - 2) : (real)0;
      z(ia) = site_inv(x(ia), alp + mup, lsa, a[ia], b[ia]);
      z(im) = site_inv(x(im), mup, lsm, a[im], b[im]);
    }
  }
}

// --------------------------------------------------------------- dispatch ---
template <int KIND, int LPC, bool WITH_A, int FP>
__device__ __forceinline__ real vg(const DevModel& m, const real* a, const real* b,
                                   Vec z, Vec g, Vec xc, Vec abar, Vec bbar, int sub, bool want_lp) {
  if (KIND == MODEL_8SCHOOLS) return vg_8schools<LPC, WITH_A>(m, a, b, z, g, xc, abar, bbar, sub, want_lp);
  if (KIND == MODEL_GERMAN_LOGNORMAL) return vg_german<LPC, WITH_A, FP, false>(m, a, b, z, g, xc, abar, bbar, sub, want_lp);
  if (KIND == MODEL_GERMAN_GAMMA) return vg_german<LPC, WITH_A, FP, true>(m, a, b, z, g, xc, abar, bbar, sub, want_lp);
  if (KIND == MODEL_RADON) return vg_radon<LPC, WITH_A, false>(m, a, b, z, g, xc, abar, bbar, sub, want_lp);
  if (KIND == MODEL_RADON_STDDVS) return vg_radon<LPC, WITH_A, true>(m, a, b, z, g, xc, abar, bbar, sub, want_lp);
  if (KIND == MODEL_ELECTION) return vg_election<LPC, WITH_A>(m, a, b, z, g, xc, abar, bbar, sub, want_lp);
  if (KIND == MODEL_ELECTRIC) return vg_electric<LPC, WITH_A>(m, a, b, z, g, xc, abar, bbar, sub, want_lp);
  return vg_time_series<LPC, WITH_A>(m, a, b, z, g, xc, abar, bbar, sub, want_lp);
}

}  // namespace arp
