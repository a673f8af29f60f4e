/* autoreparam_b200 -- C ABI of the B200-native HMC / VI hot path.
 *
 * This is the drop-in boundary for the hot path of mgorinova/autoreparam.  The
 * reference has no FFI of its own (it is pure Python on TF 1.14 / TFP 0.7); each
 * entry point below names the reference interface it replaces (file:line in the
 * reference tree) -- INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *  - plain C: pointers + sizes, no C++/torch types; `stream` is a cudaStream_t
 *    passed as void* (NULL = default stream).
 *  - `mem` says where the caller's buffers live: ARP_MEM_DEVICE (device pointers,
 *    asynchronous on `stream`) or ARP_MEM_HOST (host pointers; the call copies
 *    H2D, runs, copies D2H and synchronises).
 *  - per-chain arrays are row-major [C, D]: chain-major, the D coordinates of a
 *    chain in the reference's trace order (graphs.py:29-44) -- i.e. the
 *    concatenation of the reference's list of [C, *site_shape] state parts.
 *  - `a`, `b` are the per-coordinate parameters of the reparameterisation rule
 *    (program_transformations.py:555-600): CP a=b=1, NCP a=b=0, VIP anything in
 *    [0,1].  They are always HOST arrays of length D (small, copied per call).
 *  - precision: libarp_f32.so computes in float (arp_real = float);
 *    libarp_f64.so is the -DARP_FP64 check build with arp_real = double and the
 *    same symbols.
 *  - every function returns 0 on success, non-zero on error (never throws /
 *    exits); arp_last_error() returns the message of the calling thread's last
 *    failure.
 *  - a handle is used by one host thread at a time; kernels are re-entrant
 *    across streams.
 */
#ifndef AUTOREPARAM_B200_H_
#define AUTOREPARAM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifdef ARP_FP64
typedef double arp_real;
#else
typedef float arp_real;
#endif

#define ARP_MEM_HOST 0
#define ARP_MEM_DEVICE 1

typedef struct arp_model arp_model;

/* Raw model data, as the reference's ModelConfig.model_args / observed_data hold
 * it (models.py:51-54).  Host pointers; copied (and re-organised) at create time.
 * Unused fields are NULL / 0.
 *
 *  8schools   (models.py:131-166)   n=8;  y = treatment_effects, x1 = treatment_stddevs
 *  german_credit_lognormalcentered / german_credit_gammascale (models.py:884-964)
 *             n, f;  X [n, f] row-major design matrix (intercept | numerics | one-hots,
 *             models.py:889-892), y [n] in {0,1}
 *  radon / radon_stddvs (models.py:763-857)
 *             n, j;  idx0 = county [n] 0-based, u [j], x1 = floor x [n], y = log radon [n]
 *  election   (models.py:967-1008)  n, j = n_state;  idx0 = state [n] AS STORED (1-based,
 *             fed to one_hot(depth=j): out-of-range rows select nothing), x1 = female,
 *             x2 = black, y [n]
 *  electric   (models.py:1011-1066) n, j = n_pair, k = n_grade, k2 = n_grade_pair;
 *             idx0 = pair [n], idx1 = grade [n], idx2 = grade_pair [j] (all as stored,
 *             1-based), x1 = treatment [n], y [n]
 *  time_series(models.py:1069-1141) n = T;  x1 = year x [n], y [n]
 */
typedef struct arp_model_data {
  int64_t n, f, j, k, k2;
  const float* X;
  const float* y;
  const float* x1;
  const float* x2;
  const float* u;
  const int32_t* idx0;
  const int32_t* idx1;
  const int32_t* idx2;
} arp_model_data;

/* replaces models.get_model_by_name (models.py:1144-1175) for the in-scope models */
int arp_model_create(const char* model_name, const arp_model_data* data, arp_model** out);
void arp_model_destroy(arp_model* m);
/* number of state coordinates D (sum of the latent site sizes in trace order) */
int arp_model_num_coords(const arp_model* m);

/* replaces target(*params) of graphs.py:37-44 / 84-91 / 139-145 / 197-203 vectorised
 * over chains (inference.vectorize_log_joint_fn, inference.py:172-195) plus the
 * tf.gradients call TFP's HMC makes on it, plus make_to_centered (models.py:59-81).
 *   z [C,D] in;  lp [C], grad [C,D], centered [C,D], abar [C,D] out (each may be NULL).
 *   abar = d log_joint / d a per coordinate (what the cVIP ELBO differentiates). */
int arp_log_joint_grad(arp_model* m, const arp_real* a, const arp_real* b, const arp_real* z, int64_t C,
                       arp_real* lp, arp_real* grad, arp_real* centered, arp_real* abar, int mem, void* stream);

/* Same contract (lp, grad, centered; no abar) through a chosen engine: ARP_ENGINE_AUTO / ARP_ENGINE_SIMT = the call
 * above; ARP_ENGINE_TCGEN05 = the log-joint and gradient exactly as the tensor-core HMC engine computes them
 * (german_credit models, fp32 build only) -- the elementwise parity hook for that engine.  A chain whose coefficients
 * leave the fp16 range of the tensor-core operands is rejected by that engine: lp = -inf, grad = centered = NaN. */
int arp_log_joint_grad_engine(arp_model* m, const arp_real* a, const arp_real* b, const arp_real* z, int64_t C,
                              arp_real* lp, arp_real* grad, arp_real* centered, int engine, int mem, void* stream);

#define ARP_ENGINE_AUTO 0     /* tcgen05 for german_credit with >= 256 chains, SIMT otherwise */
#define ARP_ENGINE_SIMT 1     /* generic FP32 / FP64 SIMT kernels (every model) */
#define ARP_ENGINE_TCGEN05 2  /* tcgen05 tensor-core engine (german_credit, 0/1 outcomes, <= 64 features, any N) */
#define ARP_ENGINE_TCGEN05_STREAM 3  /* alias of 2 (round 1 had a second, shared-memory-resident tcgen05 kernel) */

/* d log_joint / d a and d log_joint / d b per coordinate ([C,D] each, either may be NULL): the adjoints of the site rule's
 * parameters that the cVIP objective differentiates (program_transformations.py:507-523). */
int arp_log_joint_param_grad(arp_model* m, const arp_real* a, const arp_real* b, const arp_real* z, int64_t C,
                             arp_real* abar, arp_real* bbar, int mem, void* stream);

/* HMC configuration: inference.hmc (inference.py:198-242) + main.py flags. */
typedef struct arp_hmc_config {
  int32_t num_leapfrog_steps;   /* --num_leapfrog_steps */
  int32_t num_results;          /* --num_samples : kept samples S */
  int32_t num_burnin_steps;     /* --num_burnin_steps */
  int32_t num_adaptation_steps; /* --num_adaptation_steps (dual averaging, per chain) */
  int32_t num_steps_between_results; /* 1 in the reference (inference.py:234) */
  uint64_t seed;                /* Philox key */
  int64_t chain_offset;         /* global id of chain 0 (multi-GPU sharding: RNG is keyed by global id) */
  double target_accept_prob;    /* 0.75 [TFP default] */
  int32_t lanes_per_chain;      /* 0 = auto; 1,8,32 = force */
  int32_t engine;               /* ARP_ENGINE_* */
  int32_t stream_window;        /* W > 0: streaming statistics of the kept samples with a lag window of W lags (see
                                   arp_hmc_buffers.stream_*); SIMT engine only */
} arp_hmc_config;

/* Buffers of one HMC run.  `mem` applies to every non-NULL pointer here.
 *   z0            [C,D]  in   initial states (util.variational_inits_from_params, util.py:394-410)
 *   eps0          [D]    in   per-coordinate initial step size sigma_q / (L/4)^2 (inference.py:212-216)
 *   ext_momenta   [T,C,D] in  optional injected momenta (T = arp_hmc_num_transitions), else Philox
 *   ext_log_u     [T,C]  in   optional injected log-uniforms for the Metropolis test
 *   samples       [S,C,D] out centred samples (= states_transformed of inference.py:238-239)
 *   samples_orig  [S,C,D] out optional raw (reparameterised-space) samples (= states_orig)
 *   is_accepted   [S,C]  out  uint8, is_accepted of the transition that produced each kept sample
 *   final_z       [C,D]  out  optional final state
 *   step_mult     [C]    out  optional final per-chain step-size multiplier (eps = eps0 * mult)
 *   accept_count  [C]    out  optional number of accepted transitions per chain (all transitions)
 *   stream_*      [C,D]  out  optional streaming statistics, see below
 */
typedef struct arp_hmc_buffers {
  const arp_real* z0;
  const arp_real* eps0;
  const arp_real* ext_momenta;
  const arp_real* ext_log_u;
  arp_real* samples;
  arp_real* samples_orig;
  uint8_t* is_accepted;
  arp_real* final_z;
  arp_real* step_mult;
  int32_t* accept_count;
  /* streaming statistics (cfg.stream_window = W > 0; each pointer optional), for runs whose [S,C,D] traces cannot be
   * stored (BASELINE configs[4]: 65 536 chains x 10 003 coordinates): the kernel keeps, per (chain, coordinate), a ring
   * of the last 2 W kept values, the first W values and W lag-product sums (updated once per block of W kept samples)
   * instead of the trace -- (4 W + 2) floats instead of S. */
  arp_real* stream_mean;      /* [C,D] out  mean of the kept centred samples */
  arp_real* stream_var;       /* [C,D] out  their biased variance */
  arp_real* stream_ess;       /* [C,D] out  ESS (same estimator as arp_ess) from the lags inside the window */
  int32_t* stream_truncated;  /* [C,D] out  1 where no negative autocorrelation appeared inside the window: the ESS then
                                 uses all W lags and is an upper bound */
} arp_hmc_buffers;

/* total transitions of a run: 1 + burnin + (1+between)*(S-1)  [TFP sample_chain] */
int64_t arp_hmc_num_transitions(const arp_hmc_config* cfg);

/* replaces inference.hmc (inference.py:198-242): HamiltonianMonteCarlo +
 * DualAveragingStepSizeAdaptation + sample_chain + transform_mcmc_states, all
 * chains, all transitions, in one persistent kernel launch. */
int arp_hmc_run(arp_model* m, const arp_hmc_config* cfg, const arp_real* a, const arp_real* b, int64_t C,
                const arp_hmc_buffers* buf, int mem, void* stream);

/* Several HMC runs in ONE launch: replaces the loop of `--inference=HMCtuning --num_leapfrog_steps=L` invocations
 * (main.py:316-329, 375-384: one process per L, each appending one entry to `tuning_runs`).  Run i takes
 * num_leapfrog_steps / num_results / num_burnin_steps / num_adaptation_steps from cfgs[i] and eps0 + the output
 * buffers (samples, is_accepted, step_mult, accept_count) from bufs[i]; every run starts from bufs[0].z0 and draws
 * the SAME random streams a separate arp_hmc_run with cfgs[i] would (results are identical to num_runs separate
 * calls).  seed, chain_offset, engine, lanes_per_chain and num_steps_between_results come from cfgs[0] and must
 * agree.  (chains x runs) is the batch axis: 100 chains x 6 values of L fill 600 chain slots of one grid. */
int arp_hmc_run_many(arp_model* m, const arp_hmc_config* cfgs, int32_t num_runs, const arp_real* a, const arp_real* b,
                     int64_t C, const arp_hmc_buffers* bufs, int mem, void* stream);

/* Interleaved CP / NCP sampler (--method=i): replaces inference.hmc_interleaved (inference.py:258-329)
 * + interleaved.Interleaved.one_step (interleaved.py:113-155): per transition one HMC step under rule A
 * (CP), one under rule B (NCP), each preceded by a re-bootstrap of (log-prob, gradient) in that rule's
 * coordinates, each with its own step sizes adapted by SimpleStepSizeAdaptation(adaptation_rate,
 * target_accept_prob).  The chain state lives in the centred space. */
typedef struct arp_ilv_config {
  int32_t num_leapfrog_steps_a;   /* num_leapfrog_steps_cp */
  int32_t num_leapfrog_steps_b;   /* num_leapfrog_steps_ncp */
  int32_t num_results, num_burnin_steps, num_adaptation_steps, num_steps_between_results;
  uint64_t seed;
  int64_t chain_offset;
  double target_accept_prob;      /* 0.75 (inference.py:294,304) */
  double adaptation_rate;         /* 0.05 (inference.py:293,303) */
  int32_t lanes_per_chain;        /* 0 = auto */
} arp_ilv_config;

/*   x0 [C,D] in  initial states in the CENTRED space;  eps0_a / eps0_b [D] base step sizes of the two rules
 *   ext_momenta [2T,C,D], ext_log_u [2T,C] in  optional injected streams, index 2*transition + (0: rule A, 1: rule B)
 *   samples [S,C,D] out centred;  is_accepted_a / _b [S,C] uint8;  step_mult_a / _b [C] final multipliers */
typedef struct arp_ilv_buffers {
  const arp_real* x0;
  const arp_real* eps0_a;
  const arp_real* eps0_b;
  const arp_real* ext_momenta;
  const arp_real* ext_log_u;
  arp_real* samples;
  uint8_t* is_accepted_a;
  uint8_t* is_accepted_b;
  arp_real* step_mult_a;
  arp_real* step_mult_b;
} arp_ilv_buffers;

int arp_hmc_interleaved_run(arp_model* m, const arp_ilv_config* cfg, const arp_real* a_a, const arp_real* b_a,
                            const arp_real* a_b, const arp_real* b_b, int64_t C, const arp_ilv_buffers* buf,
                            int mem, void* stream);

/* replaces tfp.mcmc.effective_sample_size (inference.py:240,327):
 * samples [S,C,D] -> ess [C,D].  Optional extra outputs (NULL to skip): per-series
 * mean [C,D] and biased variance [C,D] -- the per-chain moments R-hat is built
 * from (new capability; the reference has no R-hat). */
int arp_ess(const arp_real* samples, int64_t S, int64_t C, int64_t D, arp_real* ess, arp_real* mean,
            arp_real* var, int mem, void* stream);

/* VI: replaces util.get_mean_field_elbo (util.py:232-268) + the Adam loops of
 * inference.find_best_learning_rate (inference.py:26-154): all `num_runs`
 * learning rates are optimised concurrently (one 8-CTA thread-block cluster each) in one launch.
 *
 * Learnable reparameterisation (cVIP; make_learnable_parametrisation, program_transformations.py:475-533): `num_params`
 * unconstrained parameters u_p with value sigmoid(u_p) (tau = 1).  Coordinate d takes its `a` from slot a_index[d] and its
 * `b` from slot b_index[d]; -1 keeps the value passed in `a` / `b`:
 *   tied as the reference runs it   a_index[d] = d, b_index = -1 with b = 1     
 *   tied, b = a (the paper)         a_index[d] = b_index[d] = d
 *   untied (--tied_pparams=False)   a_index by the shape of the site's loc, b_index by the shape of its scale
 *                                   (a site with a scalar loc / scale shares ONE parameter over its coordinates)
 * discrete_prior = 1 adds the reference's mixture-of-Laplace log-prior of every parameter value (main.py:244-253,
 * inference.py:50-54) to the objective; `elbo` then holds elbo + prior and `prior_logp` the prior term. */
#define ARP_VI_MAX_RUNS 16
typedef struct arp_vi_config {
  int32_t num_mc_samples;         /* --num_mc_samples (256) */
  int32_t num_optimization_steps; /* --num_optimization_steps (3000) */
  int32_t num_runs;               /* number of learning rates R (<= ARP_VI_MAX_RUNS) */
  double learning_rates[ARP_VI_MAX_RUNS]; /* --learning_rates; each is /5 after 1/3 and /20 after 2/3
                                             of the steps (inference.py:69-75) */
  uint64_t seed;
  int32_t num_params;             /* P learnable reparameterisation parameters; 0 = fixed (a, b) */
  int32_t discrete_prior;         /* --discrete_prior */
} arp_vi_config;

typedef struct arp_vi_buffers {
  arp_real* loc;           /* [R,D] in/out  variational means (init 0.01*randn, program_transformations.py:207-210) */
  arp_real* rho;           /* [R,D] in/out  unconstrained scales, scale = softplus(rho) (init -2, :212-215) */
  arp_real* u;             /* [R,P] in/out  unconstrained reparameterisation parameters (init 0: sigmoid = 0.5, :507-510) */
  const int32_t* a_index;  /* [D] HOST  parameter slot of coordinate d's a, or -1 (required if num_params > 0) */
  const int32_t* b_index;  /* [D] HOST  same for b */
  const arp_real* ext_eps; /* [steps,S,D] optional injected standard normals (shared by the R runs) */
  arp_real* elbo;          /* [R,steps] out  objective timeline (value at the pre-update parameters) */
  arp_real* prior_logp;    /* [R,steps] out  optional: the prior term inside `elbo` (0 without discrete_prior) */
} arp_vi_buffers;

int arp_vi_run(arp_model* m, const arp_vi_config* cfg, const arp_real* a, const arp_real* b,
               const arp_vi_buffers* buf, int mem, void* stream);

/* number of kernels this library has launched in this process (bench accounting) */
int64_t arp_kernel_launch_count(void);
const char* arp_last_error(void);
/* scratch buffers are pooled across calls (no cudaMalloc / cudaFree inside a call after warm-up);
 * this returns the pool to the driver */
void arp_release_cached_memory(void);
/* "f32" or "f64" */
const char* arp_precision(void);

#ifdef __cplusplus
}
#endif
#endif /* AUTOREPARAM_B200_H_ */
