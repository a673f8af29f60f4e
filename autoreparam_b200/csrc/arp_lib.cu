// C ABI of the autoreparam B200 library (see include/autoreparam_b200.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared
//        -Xcompiler -fPIC [-DARP_FP64] arp_lib.cu -o libarp_f32.so | libarp_f64.so
#include "../../include/autoreparam_b200.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "arp_host.cuh"
#include "arp_hmc.cuh"
#include "arp_ess.cuh"
#include "arp_ess_fft.cuh"
#include "arp_vi.cuh"
#ifndef ARP_FP64
#include "arp_german_tcs.cuh"
#endif

using namespace arp;

static_assert(sizeof(arp_real) == sizeof(real), "arp_real / real mismatch");

// ------------------------------------------------------------------ errors ---
static thread_local std::string g_last_error;
static std::atomic<long long> g_launches{0};

static int fail(const std::string& msg) {
  g_last_error = msg;
  return 1;
}
#define ARP_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return fail(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                  std::to_string(__LINE__) + ")");                                           \
  } while (0)
#define ARP_LAUNCH_CHECK()                  \
  do {                                      \
    g_launches.fetch_add(1);                \
    ARP_CUDA(cudaGetLastError());           \
  } while (0)

extern "C" const char* arp_last_error(void) { return g_last_error.c_str(); }
extern "C" int64_t arp_kernel_launch_count(void) { return g_launches.load(); }
extern "C" const char* arp_precision(void) { return ARP_REAL_IS_DOUBLE ? "f64" : "f32"; }
extern "C" void arp_release_cached_memory(void) { DevPool::get().release(); }

struct arp_model {
  DevModel dev{};
  std::string name;
  DevBuf X, y, x1, x2, w, u, offs, gidx, pidx;
  int fp = 32;  // german: register-resident padded feature count of the SIMT engine
#ifndef ARP_FP64
  GermanTcs tcs;  // german: operands of the tcgen05 engine (chunk images of X streamed from L2; F <= 64, any N)
#endif
};

static std::vector<real> to_real(const float* p, size_t n) {
  std::vector<real> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = (real)p[i];
  return v;
}

// --------------------------------------------------------------- model create ---
extern "C" int arp_model_create(const char* model_name, const arp_model_data* d, arp_model** out) {
  if (!model_name || !d || !out) return fail("arp_model_create: null argument");
  const std::string name(model_name);
  arp_model* m = new arp_model();
  m->name = name;
  DevModel& dm = m->dev;
  auto bail = [&](const std::string& msg) { delete m; return fail("arp_model_create(" + name + "): " + msg); };
#define UP(buf, vec) do { cudaError_t _e = upload(m->buf, vec); if (_e != cudaSuccess) return bail(cudaGetErrorString(_e)); } while (0)

  if (name == "8schools") {
    if (d->n != 8 || !d->y || !d->x1) return bail("needs n=8, y, x1 (stddevs)");
    dm.kind = MODEL_8SCHOOLS; dm.N = 8; dm.D = 10;
    UP(y, to_real(d->y, 8)); UP(x1, to_real(d->x1, 8));
  } else if (name == "german_credit_lognormalcentered" || name == "german_credit_gammascale") {
    if (d->n <= 0 || d->f <= 0 || !d->X || !d->y) return bail("needs n, f, X, y");
    dm.kind = (name == "german_credit_gammascale") ? MODEL_GERMAN_GAMMA : MODEL_GERMAN_LOGNORMAL;
    dm.N = (int)d->n; dm.F = (int)d->f; dm.D = 1 + 2 * dm.F;
    if (dm.F > 64) return bail("more than 64 features is not supported");
    m->fp = dm.F <= 32 ? 32 : 64;
    dm.Fpad = m->fp;
    std::vector<real> X((size_t)dm.N * dm.Fpad, (real)0);
    for (int n = 0; n < dm.N; ++n)
      for (int f = 0; f < dm.F; ++f) X[(size_t)n * dm.Fpad + f] = (real)d->X[(size_t)n * dm.F + f];
    UP(X, X); UP(y, to_real(d->y, dm.N));
#ifndef ARP_FP64
    {
      std::string err;
      if (!m->tcs.build(d->X, d->y, dm.N, dm.F, &err)) return bail("tcgen05 operand build: " + err);
    }
#endif
  } else if (name == "radon" || name == "radon_stddvs") {
    if (d->n <= 0 || d->j <= 0 || !d->idx0 || !d->u || !d->x1 || !d->y) return bail("needs n, j, idx0 (county), u, x1, y");
    const bool sd = (name == "radon_stddvs");
    dm.kind = sd ? MODEL_RADON_STDDVS : MODEL_RADON;
    dm.N = (int)d->n; dm.J = (int)d->j; dm.D = 3 + (sd ? 2 : 1) * dm.J;
    std::vector<int> offs(dm.J + 1, 0);
    for (int n = 0; n < dm.N; ++n) {
      if (d->idx0[n] < 0 || d->idx0[n] >= dm.J) return bail("county index out of range");
      offs[d->idx0[n] + 1]++;
    }
    for (int j = 0; j < dm.J; ++j) offs[j + 1] += offs[j];
    std::vector<int> cur(offs.begin(), offs.end() - 1);
    std::vector<real> xs(dm.N), ys(dm.N);
    for (int n = 0; n < dm.N; ++n) {  // stable counting sort by county
      const int pos = cur[d->idx0[n]]++;
      xs[pos] = (real)d->x1[n]; ys[pos] = (real)d->y[n];
    }
    // per-county sufficient statistics of the Gaussian likelihood, accumulated in double
    std::vector<real> stats((size_t)6 * dm.J, (real)0);
    for (int j = 0; j < dm.J; ++j) {
      const int n0 = offs[j], n1 = offs[j + 1];
      const double cnt = n1 - n0;
      double sy = 0, sx = 0;
      for (int n = n0; n < n1; ++n) { sy += (double)ys[n]; sx += (double)xs[n]; }
      const double yb = cnt > 0 ? sy / cnt : 0, xb = cnt > 0 ? sx / cnt : 0;
      double cyy = 0, cxy = 0, cxx = 0;
      for (int n = n0; n < n1; ++n) {
        const double dy = (double)ys[n] - yb, dx = (double)xs[n] - xb;
        cyy += dy * dy; cxy += dx * dy; cxx += dx * dx;
      }
      real* st = &stats[(size_t)6 * j];
      st[0] = (real)cnt; st[1] = (real)yb; st[2] = (real)xb; st[3] = (real)cyy; st[4] = (real)cxy; st[5] = (real)cxx;
    }
    UP(offs, offs); UP(x1, xs); UP(y, ys); UP(u, to_real(d->u, dm.J)); UP(w, stats);
  } else if (name == "election") {
    if (d->n <= 0 || d->j <= 0 || !d->idx0 || !d->x1 || !d->x2 || !d->y) return bail("needs n, j (n_state), idx0 (state), x1 (female), x2 (black), y");
    dm.kind = MODEL_ELECTION; dm.K = (int)d->j; dm.J = dm.K + 1; dm.D = dm.K + 4;
    // tf.one_hot(state, K): only 0 <= state < K hits a column (reference models.py:978)
    std::map<std::tuple<int, float, float>, std::pair<double, double>> cells;
    for (int64_t n = 0; n < d->n; ++n) {
      const int s = d->idx0[n];
      const int k = (s >= 0 && s < dm.K) ? s : dm.K;
      auto& c = cells[std::make_tuple(k, d->x1[n], d->x2[n])];
      c.first += 1.0; c.second += (double)d->y[n];
    }
    std::vector<int> offs(dm.J + 1, 0);
    std::vector<real> fe, bl, w, ys;
    for (auto& kv : cells) {  // std::map iterates sorted by group first
      offs[std::get<0>(kv.first) + 1]++;
      fe.push_back((real)std::get<1>(kv.first)); bl.push_back((real)std::get<2>(kv.first));
      w.push_back((real)kv.second.first); ys.push_back((real)kv.second.second);
    }
    for (int j = 0; j < dm.J; ++j) offs[j + 1] += offs[j];
    dm.N = (int)w.size();
    UP(offs, offs); UP(x1, fe); UP(x2, bl); UP(w, w); UP(y, ys);
  } else if (name == "electric") {
    if (d->n <= 0 || d->j <= 0 || d->k != 4 || d->k2 != 4 || !d->idx0 || !d->idx1 || !d->idx2 || !d->x1 || !d->y)
      return bail("needs n, j (n_pair), k = k2 = 4, idx0 (pair), idx1 (grade), idx2 (grade_pair), x1 (treatment), y");
    dm.kind = MODEL_ELECTRIC; dm.N = (int)d->n; dm.K = (int)d->j; dm.J = dm.K + 1; dm.D = 12 + dm.K;
    std::vector<int> offs(dm.J + 1, 0), grp(dm.N);
    for (int n = 0; n < dm.N; ++n) {
      const int p = d->idx0[n];
      grp[n] = (p >= 0 && p < dm.K) ? p : dm.K;
      offs[grp[n] + 1]++;
    }
    for (int j = 0; j < dm.J; ++j) offs[j + 1] += offs[j];
    std::vector<int> cur(offs.begin(), offs.end() - 1), gidx(dm.N), pidx(dm.K);
    std::vector<real> tr(dm.N), ys(dm.N);
    for (int n = 0; n < dm.N; ++n) {
      const int pos = cur[grp[n]]++;
      const int g = d->idx1[n];
      gidx[pos] = (g >= 0 && g < 4) ? g : -1;
      tr[pos] = (real)d->x1[n]; ys[pos] = (real)d->y[n];
    }
    for (int p = 0; p < dm.K; ++p) {
      const int g = d->idx2[p];
      pidx[p] = (g >= 0 && g < 4) ? g : -1;
    }
    UP(offs, offs); UP(gidx, gidx); UP(pidx, pidx); UP(x1, tr); UP(y, ys);
  } else if (name == "time_series") {
    if (d->n <= 0 || !d->x1 || !d->y) return bail("needs n (T), x1 (x), y");
    dm.kind = MODEL_TIME_SERIES; dm.N = (int)d->n; dm.K = dm.N; dm.D = 3 + 2 * dm.N;
    UP(x1, to_real(d->x1, dm.N)); UP(y, to_real(d->y, dm.N));
  } else {
    delete m;
    return fail("unknown model " + name);  // mirrors models.py:1173-1174
  }
#undef UP
  dm.X = m->X.as<real>(); dm.y = m->y.as<real>(); dm.x1 = m->x1.as<real>(); dm.x2 = m->x2.as<real>();
  dm.w = m->w.as<real>(); dm.u = m->u.as<real>(); dm.offs = m->offs.as<int>();
  dm.gidx = m->gidx.as<int>(); dm.pidx = m->pidx.as<int>();
  *out = m;
  return 0;
}

extern "C" void arp_model_destroy(arp_model* m) { delete m; }
extern "C" int arp_model_num_coords(const arp_model* m) { return m ? m->dev.D : -1; }

// ------------------------------------------------------------------ dispatch ---
static int hmc_onchip_dpad(int kind, int lpc, int D, size_t* bytes, int* stride, int* block);

// lanes per chain of the SIMT engine.  Measured on B200 (profiles/r02_simt.md): with the chain state in shared
// memory, 8 lanes per chain beat one lane per chain at every chain count for the hierarchical models (radon 2.2e9 vs
// 7.0e8 grad-evals/s at 131 072 chains); 8schools (D = 10) is fastest with one lane per chain once the chip is full.
static int pick_lpc(const arp_model* m, long long C, int forced) {
#ifdef ARP_DEV_GERMAN_ONLY
  return 1;
#endif
  if (forced == 1 || forced == 8 || forced == 32) return forced;
  if (m->dev.kind == MODEL_TIME_SERIES)   // lane-parallel scan with the state on chip: measured 2.0e8 (C = 4096) .. 3.3e8
    return C >= 2048 ? 8 : 32;            // (131 072) grad-evals/s with 8 lanes; 32 lanes for a few hundred chains
  const long long target = 148LL * 4 * 32 * 4;      // ~4 warps per SM sub-partition
  size_t b; int st, blk;
  const bool onchip8 = hmc_onchip_dpad(m->dev.kind, 8, m->dev.D, &b, &st, &blk) > 0;
  if (C >= target && !(onchip8 && m->dev.kind != MODEL_8SCHOOLS)) return 1;
  if (C * 8 >= target) return 8;
  return 32;
}

// expands BODY(KIND, LPC, FP) for the runtime (kind, lpc, fp)
#define ARP_DISPATCH_LPC(KIND, FP, lpc, BODY)                   \
  switch (lpc) {                                                \
    case 1: { BODY(KIND, 1, FP); } break;                       \
    case 8: { BODY(KIND, 8, FP); } break;                       \
    default: { BODY(KIND, 32, FP); } break;                     \
  }
#ifdef ARP_DEV_GERMAN_ONLY
// development build (kernel experiments only, never shipped): instantiate just the German-credit
// lognormal SIMT kernels so that the library compiles in well under a minute
#define ARP_DISPATCH(kind, lpc, fp, BODY) { BODY(MODEL_GERMAN_LOGNORMAL, 1, 32); }
#else
#define ARP_DISPATCH(kind, lpc, fp, BODY)                                                        \
  switch (kind) {                                                                                \
    case MODEL_8SCHOOLS: ARP_DISPATCH_LPC(MODEL_8SCHOOLS, 32, lpc, BODY) break;                  \
    case MODEL_GERMAN_LOGNORMAL:                                                                 \
      if (fp == 32) { ARP_DISPATCH_LPC(MODEL_GERMAN_LOGNORMAL, 32, lpc, BODY) }                  \
      else { ARP_DISPATCH_LPC(MODEL_GERMAN_LOGNORMAL, 64, lpc, BODY) } break;                    \
    case MODEL_GERMAN_GAMMA:                                                                     \
      if (fp == 32) { ARP_DISPATCH_LPC(MODEL_GERMAN_GAMMA, 32, lpc, BODY) }                      \
      else { ARP_DISPATCH_LPC(MODEL_GERMAN_GAMMA, 64, lpc, BODY) } break;                        \
    case MODEL_RADON: ARP_DISPATCH_LPC(MODEL_RADON, 32, lpc, BODY) break;                        \
    case MODEL_RADON_STDDVS: ARP_DISPATCH_LPC(MODEL_RADON_STDDVS, 32, lpc, BODY) break;          \
    case MODEL_ELECTION: ARP_DISPATCH_LPC(MODEL_ELECTION, 32, lpc, BODY) break;                  \
    case MODEL_ELECTRIC: ARP_DISPATCH_LPC(MODEL_ELECTRIC, 32, lpc, BODY) break;                  \
    default: ARP_DISPATCH_LPC(MODEL_TIME_SERIES, 32, lpc, BODY) break;                           \
  }
#endif

static inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

// On-chip state of the SIMT HMC kernel: the block's chains keep their seven state vectors in shared memory when that
// fits in 96 KB (two blocks per SM).  Returns the padded vector length (0 = use the global workspace), the stride
// between chains (LPC > 1), the dynamic shared-memory size and the block size.  LPC = 1 (one thread per chain,
// coordinate-major [7][D][block]) is compiled for 8schools only.  (Measured and not kept: time_series, D = 123, with
// 32-thread blocks = 64 on-chip chains per SM: 2.67e8 grad-evals/s, the same as the HBM-resident layout at full
// occupancy -- the sequential scan is latency-bound with two warps per SM.)  ARP_HMC_ONCHIP=0 disables it (A/B).
static int hmc_onchip_dpad(int kind, int lpc, int D, size_t* bytes, int* stride, int* block) {
  *bytes = 0; *stride = 0; *block = ARP_BLOCK;
  static const bool off = [] { const char* e = getenv("ARP_HMC_ONCHIP"); return e && e[0] == '0'; }();
  if (off) return 0;
  if (lpc > 1) {
    const int dpad = (int)round_up(D, 4);
    // chain stride = LPC (mod 32): the chains of a warp (32 / LPC of them) start LPC banks apart
    *stride = (int)round_up(7 * dpad, 32) + (lpc < 32 ? lpc : 0);
    const size_t b = (size_t)(ARP_BLOCK / lpc) * *stride * sizeof(real);
    if (b > 96 * 1024) return 0;
    *bytes = b;
    return dpad;
  }
  if (kind != MODEL_8SCHOOLS) return 0;
  const size_t b = (size_t)7 * D * ARP_BLOCK * sizeof(real);
  if (b > 96 * 1024) return 0;
  *bytes = b;
  return D;
}

// copies `n` reals from a caller buffer (host or device) into a fresh device buffer
static int stage_in(DevBuf& buf, const void* src, size_t bytes, int mem, cudaStream_t st) {
  ARP_CUDA(buf.alloc(bytes));
  ARP_CUDA(cudaMemcpyAsync(buf.p, src, bytes, mem == ARP_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
  return 0;
}

// -------------------------------------------------------- log joint + gradient ---
static int log_joint_grad_impl(arp_model* m, const arp_real* a, const arp_real* b, const arp_real* z, int64_t C,
                               arp_real* lp, arp_real* grad, arp_real* centered, arp_real* abar, arp_real* bbar, int mem,
                               void* stream) {
  if (!m || !a || !b || !z || C <= 0) return fail("arp_log_joint_grad: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int D = m->dev.D;
  const int lpc = pick_lpc(m, C, 0);
  const int cpb = ARP_BLOCK / lpc;
  const long long Cpad = round_up(C, cpb);
  DevBuf da, db, dz, dlp, dg, dxc, dab, dbb;
  if (stage_in(da, a, D * sizeof(real), ARP_MEM_HOST, st)) return 1;
  if (stage_in(db, b, D * sizeof(real), ARP_MEM_HOST, st)) return 1;
  const real* zdev = z;
  if (mem == ARP_MEM_HOST) {
    if (stage_in(dz, z, (size_t)C * D * sizeof(real), mem, st)) return 1;
    zdev = dz.as<real>();
  }
  const bool with_a = abar != nullptr || bbar != nullptr;
  ARP_CUDA(dlp.alloc(Cpad * sizeof(real)));
  ARP_CUDA(dg.alloc((size_t)Cpad * D * sizeof(real)));
  ARP_CUDA(dxc.alloc((size_t)Cpad * D * sizeof(real)));
  if (with_a) {
    ARP_CUDA(dab.alloc((size_t)Cpad * D * sizeof(real)));
    ARP_CUDA(dbb.alloc((size_t)Cpad * D * sizeof(real)));
  }
  const dim3 grid((unsigned)(Cpad / cpb)), block(ARP_BLOCK);
  const DevModel dm = m->dev;
  const int fp = m->fp;
#define BODY(KIND, LPC, FP)                                                                              \
  if (with_a) k_log_joint_grad<KIND, LPC, true, FP><<<grid, block, 0, st>>>(                              \
      dm, da.as<real>(), db.as<real>(), zdev, (int)C, dlp.as<real>(), dg.as<real>(), dxc.as<real>(), dab.as<real>(), dbb.as<real>()); \
  else k_log_joint_grad<KIND, LPC, false, FP><<<grid, block, 0, st>>>(                                    \
      dm, da.as<real>(), db.as<real>(), zdev, (int)C, dlp.as<real>(), dg.as<real>(), dxc.as<real>(), nullptr, nullptr);
  ARP_DISPATCH(dm.kind, lpc, fp, BODY)
#undef BODY
  ARP_LAUNCH_CHECK();
  const cudaMemcpyKind kd = mem == ARP_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  if (lp) ARP_CUDA(cudaMemcpyAsync(lp, dlp.p, C * sizeof(real), kd, st));
  if (grad) ARP_CUDA(cudaMemcpyAsync(grad, dg.p, (size_t)C * D * sizeof(real), kd, st));
  if (centered) ARP_CUDA(cudaMemcpyAsync(centered, dxc.p, (size_t)C * D * sizeof(real), kd, st));
  if (abar) ARP_CUDA(cudaMemcpyAsync(abar, dab.p, (size_t)C * D * sizeof(real), kd, st));
  if (bbar) ARP_CUDA(cudaMemcpyAsync(bbar, dbb.p, (size_t)C * D * sizeof(real), kd, st));
  ARP_CUDA(cudaStreamSynchronize(st));  // temporaries are freed on return
  return 0;
}

extern "C" int arp_log_joint_grad(arp_model* m, const arp_real* a, const arp_real* b, const arp_real* z, int64_t C,
                                  arp_real* lp, arp_real* grad, arp_real* centered, arp_real* abar, int mem,
                                  void* stream) {
  return log_joint_grad_impl(m, a, b, z, C, lp, grad, centered, abar, nullptr, mem, stream);
}

extern "C" int arp_log_joint_param_grad(arp_model* m, const arp_real* a, const arp_real* b, const arp_real* z, int64_t C,
                                        arp_real* abar, arp_real* bbar, int mem, void* stream) {
  return log_joint_grad_impl(m, a, b, z, C, nullptr, nullptr, nullptr, abar, bbar, mem, stream);
}

// Same contract, chosen engine.  engine 0 / 1: the SIMT kernel above.  engine 2 / 3: the gradient as the tcgen05 HMC
// engine computes it (german_credit models): one transition of one leapfrog step with step size 0, momentum 0 and
// log u = -inf is run through k_german_tcs_hmc, so the proposal IS the input state, it is always accepted, and the
// kernel's own epilogue / GEMM2 / site-reverse code leaves (log-joint, gradient, centred values) in the workspace.
// A chain whose coefficients leave the fp16 range of the A operand is rejected by that engine: lp = -inf, grad = NaN.
#ifndef ARP_FP64
__global__ void k_tc_grad_out(HmcWs ws, int C, int D, real* lp, real* grad, real* centered) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)C * D) return;
  const int c = (int)(i / D), d = (int)(i % D);
  const bool ok = ws.nacc[c] == 1;
  if (d == 0 && lp) lp[c] = ok ? ws.lp[c] : -INFINITY;
  if (grad) grad[i] = ok ? ws.g[(size_t)d * ws.sd + c] : NAN;
  if (centered) centered[i] = ok ? ws.xc[(size_t)d * ws.sd + c] : NAN;
}
#endif

extern "C" int arp_log_joint_grad_engine(arp_model* m, const arp_real* a, const arp_real* b, const arp_real* z, int64_t C,
                                         arp_real* lp, arp_real* grad, arp_real* centered, int engine, int mem,
                                         void* stream) {
  if (engine == 0 || engine == 1) return arp_log_joint_grad(m, a, b, z, C, lp, grad, centered, nullptr, mem, stream);
#ifdef ARP_FP64
  return fail("arp_log_joint_grad_engine: the fp64 check build has no tcgen05 engine");
#else
  if (!m || !a || !b || !z || C <= 0) return fail("arp_log_joint_grad_engine: bad argument");
  if (engine != 2 && engine != 3) return fail("arp_log_joint_grad_engine: unknown engine");
  if (!((m->dev.kind == MODEL_GERMAN_LOGNORMAL || m->dev.kind == MODEL_GERMAN_GAMMA) && m->tcs.ready()))
    return fail("arp_log_joint_grad_engine: the tcgen05 engine needs a german_credit model with 0/1 outcomes and at most 64 features");
  cudaStream_t st = (cudaStream_t)stream;
  const int D = m->dev.D;
  const bool host = mem == ARP_MEM_HOST;
  DevBuf da, db, dz, deps, dmom, dlu, dout;
  if (stage_in(da, a, D * sizeof(real), ARP_MEM_HOST, st)) return 1;
  if (stage_in(db, b, D * sizeof(real), ARP_MEM_HOST, st)) return 1;
  const real* zdev = z;
  if (host) {
    if (stage_in(dz, z, (size_t)C * D * sizeof(real), mem, st)) return 1;
    zdev = dz.as<real>();
  }
  ARP_CUDA(deps.alloc(D * sizeof(real)));
  ARP_CUDA(cudaMemsetAsync(deps.p, 0, D * sizeof(real), st));
  ARP_CUDA(dmom.alloc((size_t)C * D * sizeof(real)));
  ARP_CUDA(cudaMemsetAsync(dmom.p, 0, (size_t)C * D * sizeof(real), st));
  {
    std::vector<real> neg((size_t)C, -INFINITY);
    if (stage_in(dlu, neg.data(), (size_t)C * sizeof(real), ARP_MEM_HOST, st)) return 1;
    ARP_CUDA(cudaStreamSynchronize(st));   // `neg` dies at the end of this scope
  }
  HmcArgs p{};
  p.C = (int)C; p.D = D; p.L = 1; p.T = 1; p.t_begin = 0; p.num_adapt = 0; p.num_burnin = 0; p.stride = 2; p.S = 1;
  p.seed = 0; p.chain_offset = 0; p.target_accept = (real)0.75;
  p.eps0 = deps.as<real>(); p.a = da.as<real>(); p.b = db.as<real>();
  p.ext_momenta = dmom.as<real>(); p.ext_log_u = dlu.as<real>();
  DevBuf wsbuf, dfz, scal, nacc;
  if (german_tcs_hmc(m->tcs, m->dev, m->fp, p, zdev, st, false, &wsbuf, &dfz, &scal, &nacc, &g_launches, &g_last_error)) return 1;
  const long long Cpad = round_up(C, TC_CHAINS), Dpad = round_up(D, 8);
  const size_t vec = (size_t)Cpad * Dpad;
  HmcWs ws{};
  ws.g = wsbuf.as<real>() + vec; ws.xc = wsbuf.as<real>() + 2 * vec;
  ws.lp = scal.as<real>() + Cpad; ws.nacc = nacc.as<int>(); ws.sd = (int)Cpad; ws.sc = 1;
  real *olp = lp, *og = grad, *oxc = centered;
  const size_t n = (size_t)C * D;
  if (host) {
    ARP_CUDA(dout.alloc((2 * n + (size_t)C) * sizeof(real)));
    og = dout.as<real>(); oxc = og + n; olp = oxc + n;
  }
  k_tc_grad_out<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ws, (int)C, D, olp, og, oxc);
  ARP_LAUNCH_CHECK();
  if (host) {
    if (lp) ARP_CUDA(cudaMemcpyAsync(lp, olp, C * sizeof(real), cudaMemcpyDeviceToHost, st));
    if (grad) ARP_CUDA(cudaMemcpyAsync(grad, og, n * sizeof(real), cudaMemcpyDeviceToHost, st));
    if (centered) ARP_CUDA(cudaMemcpyAsync(centered, oxc, n * sizeof(real), cudaMemcpyDeviceToHost, st));
  }
  ARP_CUDA(cudaStreamSynchronize(st));
  return 0;
#endif
}

// ------------------------------------------------------------------------ HMC ---
extern "C" int64_t arp_hmc_num_transitions(const arp_hmc_config* cfg) {
  if (!cfg || cfg->num_results <= 0) return 0;
  return 1 + (int64_t)cfg->num_burnin_steps + (1 + (int64_t)cfg->num_steps_between_results) * (cfg->num_results - 1);
}

__global__ void k_gather_ws(const real* ws, int sd, int sc, int C, int D, real* out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)C * D) return;
  const int c = (int)(i / D), d = (int)(i % D);
  out[i] = ws[(size_t)d * sd + (size_t)c * sc];
}

extern "C" int arp_hmc_run(arp_model* m, const arp_hmc_config* cfg, const arp_real* a, const arp_real* b, int64_t C,
                           const arp_hmc_buffers* buf, int mem, void* stream) {
  if (!m || !cfg || !a || !b || !buf || C <= 0) return fail("arp_hmc_run: bad argument");
  if (!buf->z0 || !buf->eps0) return fail("arp_hmc_run: z0 and eps0 are required");
  if (cfg->num_leapfrog_steps < 1 || cfg->num_results < 1 || cfg->num_burnin_steps < 0 ||
      cfg->num_steps_between_results < 0)
    return fail("arp_hmc_run: bad configuration");
  cudaStream_t st = (cudaStream_t)stream;
  const int D = m->dev.D;
  const long long T = arp_hmc_num_transitions(cfg);
  const long long S = cfg->num_results;
  const bool host = mem == ARP_MEM_HOST;

#ifndef ARP_FP64
  const bool tc_ok = (m->dev.kind == MODEL_GERMAN_LOGNORMAL || m->dev.kind == MODEL_GERMAN_GAMMA) && m->tcs.ready();
  if (cfg->engine < 0 || cfg->engine > 3) return fail("arp_hmc_run: unknown engine");
  if (cfg->engine >= 2 && !tc_ok)
    return fail("arp_hmc_run: the tcgen05 engine needs a german_credit model with 0/1 outcomes and at most 64 features");
  if (cfg->engine >= 2 && cfg->stream_window > 0) return fail("arp_hmc_run: streaming statistics need the SIMT engine");
  const bool use_tc = tc_ok && cfg->stream_window <= 0 && (cfg->engine >= 2 || (cfg->engine == 0 && german_tc_auto(C)));
#else
  if (cfg->engine >= 2) return fail("arp_hmc_run: the fp64 check build has no tcgen05 engine");
  const bool use_tc = false;
#endif
  if (cfg->stream_window < 0 || cfg->stream_window > 1024) return fail("arp_hmc_run: stream_window must be in [0, 1024]");

  // ---- stage inputs
  DevBuf da, db, dz0, deps, dmom, dlu, dsamp, dorig, dacc;
  if (stage_in(da, a, D * sizeof(real), ARP_MEM_HOST, st)) return 1;
  if (stage_in(db, b, D * sizeof(real), ARP_MEM_HOST, st)) return 1;
  const real *z0 = buf->z0, *eps0 = buf->eps0, *mom = buf->ext_momenta, *lu = buf->ext_log_u;
  real *samples = buf->samples, *samples_orig = buf->samples_orig;
  unsigned char* is_acc = buf->is_accepted;
  if (host) {
    if (stage_in(dz0, z0, (size_t)C * D * sizeof(real), mem, st)) return 1;
    if (stage_in(deps, eps0, D * sizeof(real), mem, st)) return 1;
    z0 = dz0.as<real>(); eps0 = deps.as<real>();
    if (mom) { if (stage_in(dmom, mom, (size_t)T * C * D * sizeof(real), mem, st)) return 1; mom = dmom.as<real>(); }
    if (lu) { if (stage_in(dlu, lu, (size_t)T * C * sizeof(real), mem, st)) return 1; lu = dlu.as<real>(); }
    if (samples) { ARP_CUDA(dsamp.alloc((size_t)S * C * D * sizeof(real))); samples = dsamp.as<real>(); }
    if (samples_orig) { ARP_CUDA(dorig.alloc((size_t)S * C * D * sizeof(real))); samples_orig = dorig.as<real>(); }
    if (is_acc) { ARP_CUDA(dacc.alloc((size_t)S * C)); is_acc = dacc.as<unsigned char>(); }
  }

  HmcArgs p{};
  p.C = (int)C; p.D = D; p.L = cfg->num_leapfrog_steps; p.T = (int)T; p.t_begin = 0;
  p.num_adapt = cfg->num_adaptation_steps; p.num_burnin = cfg->num_burnin_steps;
  p.stride = 1 + cfg->num_steps_between_results; p.S = (int)S;
  p.seed = cfg->seed; p.chain_offset = (unsigned int)cfg->chain_offset;
  p.target_accept = (real)(cfg->target_accept_prob > 0 ? cfg->target_accept_prob : 0.75);
  p.eps0 = eps0; p.a = da.as<real>(); p.b = db.as<real>();
  p.ext_momenta = mom; p.ext_log_u = lu;
  p.samples = samples; p.samples_orig = samples_orig; p.is_accepted = is_acc;

  DevBuf wsbuf, scal, nacc;
  DevBuf dfz, dstream, dsout;
  real* out_mult_dev = nullptr;
  int* out_nacc_dev = nullptr;
  real* final_z_dev = nullptr;

#ifndef ARP_FP64
  if (use_tc) {
    int rc = german_tcs_hmc(m->tcs, m->dev, m->fp, p, z0, st, buf->final_z != nullptr, &wsbuf, &dfz, &scal, &nacc,
                            &g_launches, &g_last_error);
    if (rc) return rc;
    final_z_dev = dfz.as<real>();
    out_mult_dev = scal.as<real>();
    out_nacc_dev = nacc.as<int>();
  } else
#endif
  {
    const int lpc = pick_lpc(m, C, cfg->lanes_per_chain);
    const int cpb = ARP_BLOCK / lpc;
    const long long Cpad = round_up(C, cpb);
    const long long Dpad = round_up(D, 8);
    HmcWs ws{};
    const size_t vec = (size_t)Cpad * Dpad;
    ARP_CUDA(wsbuf.alloc(7 * vec * sizeof(real)));
    ARP_CUDA(cudaMemsetAsync(wsbuf.p, 0, 7 * vec * sizeof(real), st));
    real* base = wsbuf.as<real>();
    ws.z = base; ws.g = base + vec; ws.xc = base + 2 * vec; ws.x = base + 3 * vec;
    ws.gx = base + 4 * vec; ws.xcx = base + 5 * vec; ws.v = base + 6 * vec;
    ARP_CUDA(scal.alloc(4 * Cpad * sizeof(real)));
    ARP_CUDA(nacc.alloc(Cpad * sizeof(int)));
    real* sb = scal.as<real>();
    ws.mult = sb; ws.lp = sb + Cpad; ws.H = sb + 2 * Cpad; ws.lavg = sb + 3 * Cpad;
    ws.nacc = nacc.as<int>();
    if (lpc == 1) { ws.sd = (int)Cpad; ws.sc = 1; } else { ws.sd = 1; ws.sc = (int)Dpad; }
    const int W = cfg->stream_window;
    if (W > 0) {   // streaming statistics: (4 W + 2) planes in the workspace layout, zero-initialised
      ARP_CUDA(dstream.alloc((size_t)(4 * W + 2) * vec * sizeof(real)));
      ARP_CUDA(cudaMemsetAsync(dstream.p, 0, (size_t)(4 * W + 2) * vec * sizeof(real), st));
      real* sbase = dstream.as<real>();
      p.stream_W = W; p.stream_plane = vec;
      p.stream_pivot = sbase; p.stream_sum = sbase + vec; p.stream_ring = sbase + 2 * vec;      // ring: 2 W planes
      p.stream_head = sbase + (size_t)(2 + 2 * W) * vec; p.stream_acc = sbase + (size_t)(2 + 3 * W) * vec;
    }
    const dim3 grid((unsigned)(Cpad / cpb)), block(ARP_BLOCK);
    const DevModel dm = m->dev;
    const int fp = m->fp;
    size_t oc_bytes = 0;
    int oc_stride = 0, oc_block = ARP_BLOCK;
    const int oc_dpad = hmc_onchip_dpad(m->dev.kind, lpc, D, &oc_bytes, &oc_stride, &oc_block);
    const dim3 grid_run((unsigned)(Cpad * lpc / oc_block)), block_run(oc_block);
#define BODY(KIND, LPC, FP)                                                                                           \
    k_hmc_init<KIND, LPC, FP><<<grid, block, 0, st>>>(dm, ws, p, z0);                                                  \
    if (oc_bytes > 48 * 1024)                                                                                          \
      ARP_CUDA(cudaFuncSetAttribute(k_hmc_run<KIND, LPC, FP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)oc_bytes)); \
    k_hmc_run<KIND, LPC, FP><<<grid_run, block_run, oc_bytes, st>>>(dm, ws, p, oc_dpad, oc_stride);
    ARP_DISPATCH(dm.kind, lpc, fp, BODY)
#undef BODY
    g_launches.fetch_add(1);
    ARP_LAUNCH_CHECK();
    out_mult_dev = ws.mult;
    out_nacc_dev = ws.nacc;
    if (W > 0 && (buf->stream_mean || buf->stream_var || buf->stream_ess || buf->stream_truncated)) {
      const size_t n = (size_t)C * D;
      real *om = buf->stream_mean, *ov = buf->stream_var, *oe = buf->stream_ess;
      int* ot = buf->stream_truncated;
      if (host) {
        ARP_CUDA(dsout.alloc(3 * n * sizeof(real) + n * sizeof(int)));
        om = dsout.as<real>(); ov = om + n; oe = ov + n; ot = reinterpret_cast<int*>(oe + n);
      }
      k_stream_finalize<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ws, p, (int)S, om, ov, oe, ot);
      ARP_LAUNCH_CHECK();
      if (host) {
        if (buf->stream_mean) ARP_CUDA(cudaMemcpyAsync(buf->stream_mean, om, n * sizeof(real), cudaMemcpyDeviceToHost, st));
        if (buf->stream_var) ARP_CUDA(cudaMemcpyAsync(buf->stream_var, ov, n * sizeof(real), cudaMemcpyDeviceToHost, st));
        if (buf->stream_ess) ARP_CUDA(cudaMemcpyAsync(buf->stream_ess, oe, n * sizeof(real), cudaMemcpyDeviceToHost, st));
        if (buf->stream_truncated) ARP_CUDA(cudaMemcpyAsync(buf->stream_truncated, ot, n * sizeof(int), cudaMemcpyDeviceToHost, st));
      }
    }
    if (buf->final_z) {
      ARP_CUDA(dfz.alloc((size_t)C * D * sizeof(real)));
      const long long n = (long long)C * D;
      k_gather_ws<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ws.z, ws.sd, ws.sc, (int)C, D, dfz.as<real>());
      ARP_LAUNCH_CHECK();
      final_z_dev = dfz.as<real>();
    }
  }

  // ---- outputs
  const cudaMemcpyKind kd = host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  if (host) {
    if (buf->samples) ARP_CUDA(cudaMemcpyAsync(buf->samples, samples, (size_t)S * C * D * sizeof(real), kd, st));
    if (buf->samples_orig) ARP_CUDA(cudaMemcpyAsync(buf->samples_orig, samples_orig, (size_t)S * C * D * sizeof(real), kd, st));
    if (buf->is_accepted) ARP_CUDA(cudaMemcpyAsync(buf->is_accepted, is_acc, (size_t)S * C, kd, st));
  }
  if (buf->final_z && final_z_dev) ARP_CUDA(cudaMemcpyAsync(buf->final_z, final_z_dev, (size_t)C * D * sizeof(real), kd, st));
  if (buf->step_mult) ARP_CUDA(cudaMemcpyAsync(buf->step_mult, out_mult_dev, C * sizeof(real), kd, st));
  if (buf->accept_count) ARP_CUDA(cudaMemcpyAsync(buf->accept_count, out_nacc_dev, C * sizeof(int), kd, st));
  ARP_CUDA(cudaStreamSynchronize(st));  // workspace is freed on return
  return 0;
}

// ------------------------------------------------------------- several runs, one launch ---
extern "C" int arp_hmc_run_many(arp_model* m, const arp_hmc_config* cfgs, int32_t num_runs, const arp_real* a,
                                const arp_real* b, int64_t C, const arp_hmc_buffers* bufs, int mem, void* stream) {
  if (!m || !cfgs || !a || !b || !bufs || C <= 0 || num_runs < 1 || num_runs > 64) return fail("arp_hmc_run_many: bad argument");
  if (!bufs[0].z0) return fail("arp_hmc_run_many: bufs[0].z0 is required (every run starts from it)");
  for (int i = 0; i < num_runs; ++i) {
    const arp_hmc_config& c = cfgs[i];
    if (c.num_leapfrog_steps < 1 || c.num_results < 1 || c.num_burnin_steps < 0 || c.num_steps_between_results < 0)
      return fail("arp_hmc_run_many: bad configuration");
    if (c.num_steps_between_results != cfgs[0].num_steps_between_results || c.seed != cfgs[0].seed ||
        c.chain_offset != cfgs[0].chain_offset || c.engine != cfgs[0].engine || c.lanes_per_chain != cfgs[0].lanes_per_chain)
      return fail("arp_hmc_run_many: the runs may differ in num_leapfrog_steps, num_results, num_burnin_steps and "
                  "num_adaptation_steps only");
    if (!bufs[i].eps0) return fail("arp_hmc_run_many: eps0 is required for every run");
    if (bufs[i].ext_momenta || bufs[i].ext_log_u || bufs[i].samples_orig || bufs[i].final_z)
      return fail("arp_hmc_run_many: injected streams, raw samples and final states are single-run features");
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int D = m->dev.D;
  const bool host = mem == ARP_MEM_HOST;
  const arp_hmc_config& c0 = cfgs[0];
#ifndef ARP_FP64
  const bool tc_ok = (m->dev.kind == MODEL_GERMAN_LOGNORMAL || m->dev.kind == MODEL_GERMAN_GAMMA) && m->tcs.ready();
  if (c0.engine < 0 || c0.engine > 3) return fail("arp_hmc_run_many: unknown engine");
  if (c0.engine >= 2 && !tc_ok) return fail("arp_hmc_run_many: the tcgen05 engine needs a german_credit model");
  // engine and lane layout are chosen as a SEPARATE run of C chains would choose them: identical arithmetic
  const bool use_tc = tc_ok && (c0.engine >= 2 || (c0.engine == 0 && german_tc_auto(C)));
#else
  if (c0.engine >= 2) return fail("arp_hmc_run_many: the fp64 check build has no tcgen05 engine");
  const bool use_tc = false;
#endif
  const int lpc = use_tc ? 1 : pick_lpc(m, C, c0.lanes_per_chain);
  const int cpb = use_tc ? 128 : ARP_BLOCK / lpc;
  const long long slice_rows = round_up(C, cpb);

  DevBuf da, db, dz0, deps, dslices;
  std::vector<DevBuf> dsamp(num_runs), dacc(num_runs);
  if (stage_in(da, a, D * sizeof(real), ARP_MEM_HOST, st)) return 1;
  if (stage_in(db, b, D * sizeof(real), ARP_MEM_HOST, st)) return 1;
  const real* z0 = bufs[0].z0;
  if (host) {
    if (stage_in(dz0, z0, (size_t)C * D * sizeof(real), mem, st)) return 1;
    z0 = dz0.as<real>();
  }
  // per-run step sizes: one [num_runs, D] device table
  ARP_CUDA(deps.alloc((size_t)num_runs * D * sizeof(real)));
  std::vector<HmcSlice> slices(num_runs);
  for (int i = 0; i < num_runs; ++i) {
    ARP_CUDA(cudaMemcpyAsync(deps.as<real>() + (size_t)i * D, bufs[i].eps0, D * sizeof(real),
                             host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
    const long long S = cfgs[i].num_results;
    HmcSlice& s = slices[i];
    s.L = cfgs[i].num_leapfrog_steps; s.T = (int)arp_hmc_num_transitions(&cfgs[i]);
    s.num_adapt = cfgs[i].num_adaptation_steps; s.num_burnin = cfgs[i].num_burnin_steps; s.S = (int)S;
    s.eps0 = deps.as<real>() + (size_t)i * D;
    s.samples = bufs[i].samples; s.is_accepted = bufs[i].is_accepted;
    if (host) {
      if (bufs[i].samples) { ARP_CUDA(dsamp[i].alloc((size_t)S * C * D * sizeof(real))); s.samples = dsamp[i].as<real>(); }
      if (bufs[i].is_accepted) { ARP_CUDA(dacc[i].alloc((size_t)S * C)); s.is_accepted = dacc[i].as<unsigned char>(); }
    }
  }
  if (stage_in(dslices, slices.data(), slices.size() * sizeof(HmcSlice), ARP_MEM_HOST, st)) return 1;
  ARP_CUDA(cudaStreamSynchronize(st));   // `slices` (host vector) has been consumed

  HmcArgs p{};
  p.C = (int)C; p.D = D; p.L = slices[0].L; p.T = slices[0].T; p.t_begin = 0;
  p.num_adapt = slices[0].num_adapt; p.num_burnin = slices[0].num_burnin;
  p.stride = 1 + c0.num_steps_between_results; p.S = slices[0].S;
  p.seed = c0.seed; p.chain_offset = (unsigned int)c0.chain_offset;
  p.target_accept = (real)(c0.target_accept_prob > 0 ? c0.target_accept_prob : 0.75);
  p.eps0 = slices[0].eps0; p.a = da.as<real>(); p.b = db.as<real>();
  p.slices = dslices.as<HmcSlice>(); p.slice_rows = (int)slice_rows;

  DevBuf wsbuf, scal, nacc, dfz;
  real* mult_dev = nullptr;
  int* nacc_dev = nullptr;
#ifndef ARP_FP64
  if (use_tc) {
    if (german_tcs_hmc(m->tcs, m->dev, m->fp, p, z0, st, false, &wsbuf, &dfz, &scal, &nacc, &g_launches, &g_last_error, num_runs))
      return 1;
    mult_dev = scal.as<real>(); nacc_dev = nacc.as<int>();
  } else
#endif
  {
    const long long Cpad = slice_rows * num_runs, Dpad = round_up(D, 8);
    const size_t vec = (size_t)Cpad * Dpad;
    ARP_CUDA(wsbuf.alloc(7 * vec * sizeof(real)));
    ARP_CUDA(cudaMemsetAsync(wsbuf.p, 0, 7 * vec * sizeof(real), st));
    ARP_CUDA(scal.alloc(4 * Cpad * sizeof(real)));
    ARP_CUDA(nacc.alloc(Cpad * sizeof(int)));
    HmcWs ws{};
    real* base = wsbuf.as<real>();
    ws.z = base; ws.g = base + vec; ws.xc = base + 2 * vec; ws.x = base + 3 * vec;
    ws.gx = base + 4 * vec; ws.xcx = base + 5 * vec; ws.v = base + 6 * vec;
    real* sb = scal.as<real>();
    ws.mult = sb; ws.lp = sb + Cpad; ws.H = sb + 2 * Cpad; ws.lavg = sb + 3 * Cpad;
    ws.nacc = nacc.as<int>();
    if (lpc == 1) { ws.sd = (int)Cpad; ws.sc = 1; } else { ws.sd = 1; ws.sc = (int)Dpad; }
    const dim3 grid((unsigned)(Cpad / cpb)), block(ARP_BLOCK);
    const DevModel dm = m->dev;
    const int fp = m->fp;
    size_t oc_bytes = 0;
    int oc_stride = 0, oc_block = ARP_BLOCK;
    const int oc_dpad = hmc_onchip_dpad(m->dev.kind, lpc, D, &oc_bytes, &oc_stride, &oc_block);
    const dim3 grid_run((unsigned)(Cpad * lpc / oc_block)), block_run(oc_block);
#define BODY(KIND, LPC, FP)                                                                                           \
    k_hmc_init<KIND, LPC, FP, true><<<grid, block, 0, st>>>(dm, ws, p, z0);                                            \
    if (oc_bytes > 48 * 1024)                                                                                          \
      ARP_CUDA(cudaFuncSetAttribute(k_hmc_run<KIND, LPC, FP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)oc_bytes)); \
    k_hmc_run<KIND, LPC, FP, true><<<grid_run, block_run, oc_bytes, st>>>(dm, ws, p, oc_dpad, oc_stride);
    ARP_DISPATCH(dm.kind, lpc, fp, BODY)
#undef BODY
    g_launches.fetch_add(1);
    ARP_LAUNCH_CHECK();
    mult_dev = ws.mult; nacc_dev = ws.nacc;
  }
  const cudaMemcpyKind kd = host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  for (int i = 0; i < num_runs; ++i) {
    const long long S = cfgs[i].num_results;
    if (host) {
      if (bufs[i].samples) ARP_CUDA(cudaMemcpyAsync(bufs[i].samples, slices[i].samples, (size_t)S * C * D * sizeof(real), kd, st));
      if (bufs[i].is_accepted) ARP_CUDA(cudaMemcpyAsync(bufs[i].is_accepted, slices[i].is_accepted, (size_t)S * C, kd, st));
    }
    if (bufs[i].step_mult) ARP_CUDA(cudaMemcpyAsync(bufs[i].step_mult, mult_dev + i * slice_rows, C * sizeof(real), kd, st));
    if (bufs[i].accept_count) ARP_CUDA(cudaMemcpyAsync(bufs[i].accept_count, nacc_dev + i * slice_rows, C * sizeof(int), kd, st));
  }
  ARP_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// ---------------------------------------------------------------- interleaved ---
extern "C" int arp_hmc_interleaved_run(arp_model* m, const arp_ilv_config* cfg, const arp_real* a_a, const arp_real* b_a,
                                       const arp_real* a_b, const arp_real* b_b, int64_t C, const arp_ilv_buffers* buf,
                                       int mem, void* stream) {
  if (!m || !cfg || !a_a || !b_a || !a_b || !b_b || !buf || C <= 0) return fail("arp_hmc_interleaved_run: bad argument");
  if (!buf->x0 || !buf->eps0_a || !buf->eps0_b) return fail("arp_hmc_interleaved_run: x0, eps0_a and eps0_b are required");
  if (cfg->num_leapfrog_steps_a < 1 || cfg->num_leapfrog_steps_b < 1 || cfg->num_results < 1 ||
      cfg->num_burnin_steps < 0 || cfg->num_steps_between_results < 0)
    return fail("arp_hmc_interleaved_run: bad configuration");
  cudaStream_t st = (cudaStream_t)stream;
  const int D = m->dev.D;
  const long long S = cfg->num_results;
  const long long T = 1 + (long long)cfg->num_burnin_steps + (1 + (long long)cfg->num_steps_between_results) * (S - 1);
  const bool host = mem == ARP_MEM_HOST;
  DevBuf da, db, da2, db2, dx0, de1, de2, dmom, dlu, dsamp, dacc1, dacc2;
  if (stage_in(da, a_a, D * sizeof(real), ARP_MEM_HOST, st) || stage_in(db, b_a, D * sizeof(real), ARP_MEM_HOST, st) ||
      stage_in(da2, a_b, D * sizeof(real), ARP_MEM_HOST, st) || stage_in(db2, b_b, D * sizeof(real), ARP_MEM_HOST, st))
    return 1;
  const real *x0 = buf->x0, *e1 = buf->eps0_a, *e2 = buf->eps0_b, *mom = buf->ext_momenta, *lu = buf->ext_log_u;
  real* samples = buf->samples;
  unsigned char *acc1 = buf->is_accepted_a, *acc2 = buf->is_accepted_b;
  if (host) {
    if (stage_in(dx0, x0, (size_t)C * D * sizeof(real), mem, st)) return 1;
    if (stage_in(de1, e1, D * sizeof(real), mem, st) || stage_in(de2, e2, D * sizeof(real), mem, st)) return 1;
    x0 = dx0.as<real>(); e1 = de1.as<real>(); e2 = de2.as<real>();
    if (mom) { if (stage_in(dmom, mom, (size_t)2 * T * C * D * sizeof(real), mem, st)) return 1; mom = dmom.as<real>(); }
    if (lu) { if (stage_in(dlu, lu, (size_t)2 * T * C * sizeof(real), mem, st)) return 1; lu = dlu.as<real>(); }
    if (samples) { ARP_CUDA(dsamp.alloc((size_t)S * C * D * sizeof(real))); samples = dsamp.as<real>(); }
    if (acc1) { ARP_CUDA(dacc1.alloc((size_t)S * C)); acc1 = dacc1.as<unsigned char>(); }
    if (acc2) { ARP_CUDA(dacc2.alloc((size_t)S * C)); acc2 = dacc2.as<unsigned char>(); }
  }
  HmcArgs p{};
  p.C = (int)C; p.D = D; p.L = cfg->num_leapfrog_steps_a; p.T = (int)T; p.t_begin = 0;
  p.num_adapt = cfg->num_adaptation_steps; p.num_burnin = cfg->num_burnin_steps;
  p.stride = 1 + cfg->num_steps_between_results; p.S = (int)S;
  p.seed = cfg->seed; p.chain_offset = (unsigned int)cfg->chain_offset;
  p.target_accept = (real)(cfg->target_accept_prob > 0 ? cfg->target_accept_prob : 0.75);
  p.eps0 = e1; p.a = da.as<real>(); p.b = db.as<real>();
  p.ext_momenta = mom; p.ext_log_u = lu; p.samples = samples; p.samples_orig = nullptr; p.is_accepted = acc1;
  const int lpc = pick_lpc(m, C, cfg->lanes_per_chain);
  const int cpb = ARP_BLOCK / lpc;
  const long long Cpad = round_up(C, cpb), Dpad = round_up(D, 8);
  const size_t vec = (size_t)Cpad * Dpad;
  DevBuf wsbuf, scal, nacc;
  ARP_CUDA(wsbuf.alloc(7 * vec * sizeof(real)));
  ARP_CUDA(cudaMemsetAsync(wsbuf.p, 0, 7 * vec * sizeof(real), st));
  ARP_CUDA(scal.alloc(2 * Cpad * sizeof(real)));
  ARP_CUDA(nacc.alloc(2 * Cpad * sizeof(int)));
  HmcWs ws{};
  real* base = wsbuf.as<real>();
  ws.z = base; ws.g = base + vec; ws.xc = base + 2 * vec; ws.x = base + 3 * vec;
  ws.gx = base + 4 * vec; ws.xcx = base + 5 * vec; ws.v = base + 6 * vec;
  ws.mult = scal.as<real>(); ws.nacc = nacc.as<int>();
  ws.lp = ws.H = ws.lavg = nullptr;
  if (lpc == 1) { ws.sd = (int)Cpad; ws.sc = 1; } else { ws.sd = 1; ws.sc = (int)Dpad; }
  IlvArgs q{};
  q.a2 = da2.as<real>(); q.b2 = db2.as<real>(); q.eps0_2 = e2; q.L2 = cfg->num_leapfrog_steps_b;
  q.rate = (real)(cfg->adaptation_rate > 0 ? cfg->adaptation_rate : 0.05);
  q.is_accepted2 = acc2; q.mult2 = scal.as<real>() + Cpad; q.nacc2 = nacc.as<int>() + Cpad;
  const dim3 grid((unsigned)(Cpad / cpb)), block(ARP_BLOCK);
  const DevModel dm = m->dev;
  const int fp = m->fp;
#define BODY(KIND, LPC, FP) k_hmc_interleaved<KIND, LPC, FP><<<grid, block, 0, st>>>(dm, ws, p, q, x0);
  ARP_DISPATCH(dm.kind, lpc, fp, BODY)
#undef BODY
  ARP_LAUNCH_CHECK();
  const cudaMemcpyKind kd = host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  if (host) {
    if (buf->samples) ARP_CUDA(cudaMemcpyAsync(buf->samples, samples, (size_t)S * C * D * sizeof(real), kd, st));
    if (buf->is_accepted_a) ARP_CUDA(cudaMemcpyAsync(buf->is_accepted_a, acc1, (size_t)S * C, kd, st));
    if (buf->is_accepted_b) ARP_CUDA(cudaMemcpyAsync(buf->is_accepted_b, acc2, (size_t)S * C, kd, st));
  }
  if (buf->step_mult_a) ARP_CUDA(cudaMemcpyAsync(buf->step_mult_a, ws.mult, C * sizeof(real), kd, st));
  if (buf->step_mult_b) ARP_CUDA(cudaMemcpyAsync(buf->step_mult_b, q.mult2, C * sizeof(real), kd, st));
  ARP_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// ------------------------------------------------------------------------ ESS ---
// ARP_ESS_DIRECT=1 forces the direct-summation kernels for every S (A/B measurements, cross-check tests)
static bool ess_force_direct() {
  const char* e = getenv("ARP_ESS_DIRECT");
  return e && e[0] == '1';
}

extern "C" int arp_ess(const arp_real* samples, int64_t S, int64_t C, int64_t D, arp_real* ess, arp_real* mean,
                       arp_real* var, int mem, void* stream) {
  if (!samples || !ess || S < 2 || C <= 0 || D <= 0) return fail("arp_ess: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  DevBuf din, dout;
  const real* in = samples;
  const size_t n = (size_t)C * D;
  real *out = ess, *omean = mean, *ovar = var;
  if (mem == ARP_MEM_HOST) {
    if (stage_in(din, samples, (size_t)S * n * sizeof(real), mem, st)) return 1;
    ARP_CUDA(dout.alloc(3 * n * sizeof(real)));
    in = din.as<real>(); out = dout.as<real>();
    omean = mean ? out + n : nullptr; ovar = var ? out + 2 * n : nullptr;
  }
  DevBuf dxt;
  if (S <= ARP_FFT_MAXS && !ess_force_direct()) {
    // shared-memory FFT (arp_ess_fft.cuh): every lag at once, the samples cross HBM once, no transposed copy
    const size_t smem = sizeof(cplx<real>) * (ARP_FFT_G / 2) * ARP_FFT_BUF + sizeof(double) * 2 * (ARP_FFT_THREADS / 32) * ARP_FFT_G;
    ARP_CUDA(cudaFuncSetAttribute(k_ess_fft<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned fg = (unsigned)((n + ARP_FFT_G - 1) / ARP_FFT_G);
    k_ess_fft<real><<<fg, ARP_FFT_THREADS, smem, st>>>(in, (int)S, (long long)n, out, omean, ovar);
  } else {
    // longer series: direct summation up to the first negative lag on a series-major copy (arp_ess.cuh)
    ARP_CUDA(dxt.alloc((size_t)S * n * sizeof(real)));
    {
      const dim3 tb(32, 8), tg((unsigned)((n + 31) / 32), (unsigned)std::min<long long>((S + 31) / 32, 64));
      k_ess_transpose<<<tg, tb, 0, st>>>(in, (int)S, (long long)n, dxt.as<real>());
      ARP_LAUNCH_CHECK();
    }
    const unsigned eg = (unsigned)((n + ARP_ESS_BLOCK - 1) / ARP_ESS_BLOCK);
    if (!ARP_REAL_IS_DOUBLE && S % 4 == 0)
      k_ess<true><<<eg, ARP_ESS_BLOCK, 0, st>>>(dxt.as<real>(), (int)S, (int)C, (int)D, out, omean, ovar);
    else
      k_ess<false><<<eg, ARP_ESS_BLOCK, 0, st>>>(dxt.as<real>(), (int)S, (int)C, (int)D, out, omean, ovar);
  }
  ARP_LAUNCH_CHECK();
  if (dxt.p) ARP_CUDA(cudaStreamSynchronize(st));  // the transposed copy is freed on return
  if (mem == ARP_MEM_HOST) {
    ARP_CUDA(cudaMemcpyAsync(ess, out, n * sizeof(real), cudaMemcpyDeviceToHost, st));
    if (mean) ARP_CUDA(cudaMemcpyAsync(mean, omean, n * sizeof(real), cudaMemcpyDeviceToHost, st));
    if (var) ARP_CUDA(cudaMemcpyAsync(var, ovar, n * sizeof(real), cudaMemcpyDeviceToHost, st));
    ARP_CUDA(cudaStreamSynchronize(st));
  }
  return 0;
}

// ------------------------------------------------------------------------- VI ---
extern "C" int arp_vi_run(arp_model* m, const arp_vi_config* cfg, const arp_real* a, const arp_real* b,
                          const arp_vi_buffers* buf, int mem, void* stream) {
  if (!m || !cfg || !a || !b || !buf || !buf->loc || !buf->rho || !buf->elbo) return fail("arp_vi_run: bad argument");
  const int P = cfg->num_params;
  if (P < 0 || P > 2 * m->dev.D) return fail("arp_vi_run: num_params out of range");
  if (P > 0 && (!buf->u || !buf->a_index || !buf->b_index)) return fail("arp_vi_run: num_params > 0 needs u, a_index and b_index");
  if (cfg->num_mc_samples < 1 || cfg->num_mc_samples > ARP_VI_MAX_S || cfg->num_optimization_steps < 1)
    return fail("arp_vi_run: num_mc_samples must be in [1, " + std::to_string(ARP_VI_MAX_S) + "]");
  if (cfg->num_runs < 1 || cfg->num_runs > ARP_VI_MAX_RUNS) return fail("arp_vi_run: num_runs out of range");
  cudaStream_t st = (cudaStream_t)stream;
  const int D = m->dev.D, R = cfg->num_runs;
  const int S = cfg->num_mc_samples, steps = cfg->num_optimization_steps;
  const bool host = mem == ARP_MEM_HOST;
  if (P > 0)
    for (int d = 0; d < D; ++d)
      if (buf->a_index[d] < -1 || buf->a_index[d] >= P || buf->b_index[d] < -1 || buf->b_index[d] >= P)
        return fail("arp_vi_run: a_index / b_index entry out of range");
  DevBuf da, db, dloc, drho, du, dia, dib, deps, delbo, dplp, dws;
  if (stage_in(da, a, D * sizeof(real), ARP_MEM_HOST, st)) return 1;
  if (stage_in(db, b, D * sizeof(real), ARP_MEM_HOST, st)) return 1;
  if (P > 0) {
    if (stage_in(dia, buf->a_index, D * sizeof(int), ARP_MEM_HOST, st)) return 1;
    if (stage_in(dib, buf->b_index, D * sizeof(int), ARP_MEM_HOST, st)) return 1;
  }
  real *loc = buf->loc, *rho = buf->rho, *u = buf->u, *elbo = buf->elbo, *plp = buf->prior_logp;
  const real* eps = buf->ext_eps;
  const size_t pbytes = (size_t)R * D * sizeof(real), ubytes = (size_t)R * P * sizeof(real);
  if (host) {
    if (stage_in(dloc, loc, pbytes, mem, st)) return 1;
    if (stage_in(drho, rho, pbytes, mem, st)) return 1;
    loc = dloc.as<real>(); rho = drho.as<real>();
    if (P > 0) { if (stage_in(du, u, ubytes, mem, st)) return 1; u = du.as<real>(); }
    if (eps) { if (stage_in(deps, eps, (size_t)steps * S * D * sizeof(real), mem, st)) return 1; eps = deps.as<real>(); }
    ARP_CUDA(delbo.alloc((size_t)R * steps * sizeof(real)));
    elbo = delbo.as<real>();
    if (plp) { ARP_CUDA(dplp.alloc((size_t)R * steps * sizeof(real))); plp = dplp.as<real>(); }
  }
  ViArgs v{};
  v.D = D; v.S = S; v.steps = steps; v.R = R; v.P = P; v.discrete_prior = cfg->discrete_prior; v.seed = cfg->seed;
  for (int r = 0; r < R; ++r) v.lrs[r] = (real)cfg->learning_rates[r];
  v.loc = loc; v.rho = rho; v.u = u; v.ia = dia.as<int>(); v.ib = dib.as<int>(); v.ext_eps = eps; v.elbo = elbo;
  v.prior_logp = plp;
  v.a_in = da.as<real>(); v.b_in = db.as<real>();
  int rc = vi_launch(m->dev, m->fp, v, st, &dws, &g_launches, &g_last_error);
  if (rc) return rc;
  if (host) {
    ARP_CUDA(cudaMemcpyAsync(buf->loc, loc, pbytes, cudaMemcpyDeviceToHost, st));
    ARP_CUDA(cudaMemcpyAsync(buf->rho, rho, pbytes, cudaMemcpyDeviceToHost, st));
    if (P > 0) ARP_CUDA(cudaMemcpyAsync(buf->u, u, ubytes, cudaMemcpyDeviceToHost, st));
    ARP_CUDA(cudaMemcpyAsync(buf->elbo, elbo, (size_t)R * steps * sizeof(real), cudaMemcpyDeviceToHost, st));
    if (buf->prior_logp) ARP_CUDA(cudaMemcpyAsync(buf->prior_logp, plp, (size_t)R * steps * sizeof(real), cudaMemcpyDeviceToHost, st));
  }
  ARP_CUDA(cudaStreamSynchronize(st));
  return 0;
}
