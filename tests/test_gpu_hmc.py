"""GPU parity of the persistent HMC kernel against the TFP-order oracle."""
import numpy as np
import pytest

from autoreparam_b200 import engine
from oracle import oracle as O
from tests import common

pytestmark = pytest.mark.gpu


def _setup(model, method, C, seed):
    mc = common.model_config(model)
    raw = common.raw_data(model)
    D = mc.num_coords
    a, b = common.ab_for(method, D)
    z0 = common.random_states(model, D, C, seed=seed, scale=0.3).astype(np.float32).astype(np.float64)
    return mc, raw, D, a, b, z0


CASES = [("8schools", "CP", 0.15), ("8schools", "NCP", 0.2), ("german_credit_lognormalcentered", "VIP_a", 0.01),
         ("german_credit_gammascale", "NCP", 0.002), ("radon", "NCP", 0.02), ("radon_stddvs", "VIP_a", 0.01),
         ("election", "dVIP", 0.01), ("electric", "NCP", 0.01), ("time_series", "NCP", 2e-5),
         ("german_synth", "CP", 0.01)]


@pytest.mark.parametrize("model,method,eps", CASES)
def test_fixed_momenta_trajectory_fp64(model, method, eps):
    """Identical leapfrog trajectories, accept decisions, step-size adaptation and
    thinning given identical momenta / uniforms: fp64 check build vs fp64 oracle."""
    C, L, S, burn, adapt = 5, 3, 4, 3, 5
    mc, raw, D, a, b, z0 = _setup(model, method, C, seed=11)
    omodel = "german_credit_lognormalcentered" if model == "german_synth" else model
    T = O.num_transitions(S, burn)
    rng = np.random.default_rng(5)
    mom = rng.standard_normal((T, C, D))
    lu = np.log(rng.uniform(size=(T, C)))
    eps0 = np.full(D, eps) * rng.uniform(0.5, 1.5, D)
    ref = O.hmc_chain(omodel, raw, z0, eps0, L, S, burn, adapt, a, b, momenta=mom, log_u=lu)
    out = engine.hmc_run(mc, z0, eps0, a, b, num_leapfrog_steps=L, num_results=S, num_burnin_steps=burn,
                         num_adaptation_steps=adapt, ext_momenta=mom, ext_log_u=lu, want_orig=True, precision="f64",
                         engine=engine.ENGINE_SIMT)
    assert out["num_transitions"] == T
    assert (out["is_accepted"].astype(bool) == ref["is_accepted"]).all()
    assert 0 < ref["is_accepted"].mean()  # the case must exercise accepted moves
    assert common.rel_err(out["samples_orig"].reshape(S * C, D), ref["samples_orig"].reshape(S * C, D)).max() < 1e-9
    assert common.rel_err(out["samples"].reshape(S * C, D), ref["samples_centered"].reshape(S * C, D)).max() < 1e-9
    assert common.rel_err(out["final_z"], ref["z"]).max() < 1e-9
    assert common.rel_err(out["step_mult"], ref["step_mult"]).max() < 1e-9


@pytest.mark.parametrize("model,method,eps", CASES)
def test_fixed_momenta_trajectory_fp32(model, method, eps):
    """Same, product (fp32) build: trajectories agree to fp32 round-off over a
    short horizon and every accept decision matches."""
    C, L, S, burn, adapt = 5, 3, 3, 2, 4
    mc, raw, D, a, b, z0 = _setup(model, method, C, seed=12)
    omodel = "german_credit_lognormalcentered" if model == "german_synth" else model
    T = O.num_transitions(S, burn)
    rng = np.random.default_rng(6)
    mom = rng.standard_normal((T, C, D)).astype(np.float32).astype(np.float64)
    lu = np.log(rng.uniform(size=(T, C))).astype(np.float32).astype(np.float64)
    eps0 = (np.full(D, eps) * rng.uniform(0.5, 1.5, D)).astype(np.float32).astype(np.float64)
    ref = O.hmc_chain(omodel, raw, z0, eps0, L, S, burn, adapt, a, b, momenta=mom, log_u=lu)
    for lpc in (1, 8, 32):
        out = engine.hmc_run(mc, z0, eps0, a, b, num_leapfrog_steps=L, num_results=S, num_burnin_steps=burn,
                             num_adaptation_steps=adapt, ext_momenta=mom, ext_log_u=lu, want_orig=True,
                             precision="f32", lanes_per_chain=lpc, engine=engine.ENGINE_SIMT)
        # an accept decision may legitimately flip only if log_alpha is within fp32 noise of log_u
        same = out["is_accepted"].astype(bool) == ref["is_accepted"]
        assert same.all(), (lpc, np.argwhere(~same))
        tol = {"time_series": 5e-4, "german_credit_gammascale": 2e-2}.get(model, 2e-4)  # exp(10 z0 + v) amplifies fp32 round-off
        err = common.rel_err(out["samples"].reshape(S * C, D), ref["samples_centered"].reshape(S * C, D)).max()
        assert err < tol, (lpc, err)


def test_philox_stream_matches_oracle():
    """Internal counter-based RNG: the fp64 build draws the same momenta and
    uniforms as the oracle's numpy Philox, so whole chains coincide."""
    model, method = "8schools", "NCP"
    C, L, S, burn, adapt = 7, 4, 5, 4, 6
    mc, raw, D, a, b, z0 = _setup(model, method, C, seed=13)
    eps0 = np.full(D, 0.2)
    seed, off = 0x1234ABCD5678, 1000
    ref = O.hmc_chain(model, raw, z0, eps0, L, S, burn, adapt, a, b, seed=seed, chain_ids=np.arange(C) + off)
    out = engine.hmc_run(mc, z0, eps0, a, b, num_leapfrog_steps=L, num_results=S, num_burnin_steps=burn,
                         num_adaptation_steps=adapt, seed=seed, chain_offset=off, precision="f64")
    assert (out["is_accepted"].astype(bool) == ref["is_accepted"]).all()
    assert common.rel_err(out["samples"].reshape(S * C, D), ref["samples_centered"].reshape(S * C, D)).max() < 1e-6
    # sharding invariance: chains [3, 7) run alone with chain_offset give the same samples
    sub = engine.hmc_run(mc, z0[3:], eps0, a, b, num_leapfrog_steps=L, num_results=S, num_burnin_steps=burn,
                         num_adaptation_steps=adapt, seed=seed, chain_offset=off + 3, precision="f64")
    assert np.array_equal(sub["samples"], out["samples"][:, 3:])


def test_radon_posterior_matches_closed_form():
    """radon with sigma_y = 1 is linear-Gaussian: the exact posterior mean / sd is
    available in closed form -- an absolute pin for the sampler."""
    import torch
    model = "radon"
    mc = common.model_config(model, "MN")
    raw = common.raw_data(model, "MN")
    D = mc.num_coords
    J = len(raw["u"])
    # posterior precision / mean in the CENTRED space: theta = [mua, b1, b2, m_1..m_J]
    A = np.zeros((D, D)); rhs = np.zeros(D)
    A[0, 0] += 1; A[1, 1] += 1; A[2, 2] += 1                      # N(0,1) priors
    for j in range(J):                                             # m_j ~ N(mua + u_j b1, 1)
        v = np.zeros(D); v[3 + j] = 1; v[0] = -1; v[1] = -raw["u"][j]
        A += np.outer(v, v)
    for n in range(len(raw["y"])):                                 # y_n ~ N(m_c + x_n b2, 1)
        v = np.zeros(D); v[3 + raw["county"][n]] = 1; v[2] = raw["x"][n]
        A += np.outer(v, v); rhs += v * raw["y"][n]
    cov = np.linalg.inv(A); mean = cov @ rhs; sd = np.sqrt(np.diag(cov))
    for method in ("CP", "NCP"):
        a, b = common.ab_for(method, D)
        C, S = 256, 400
        rng = np.random.default_rng(1)
        z0 = rng.standard_normal((C, D)) * 0.1
        eps0 = np.full(D, 0.05) if method == "CP" else np.full(D, 0.03)
        out = engine.hmc_run(mc, z0.astype(np.float32), eps0, a, b, num_leapfrog_steps=8, num_results=S,
                             num_burnin_steps=600, num_adaptation_steps=500, seed=7, precision="f32")
        x = out["samples"].astype(np.float64)            # centred samples [S, C, D]
        acc = out["is_accepted"].mean()
        assert 0.5 < acc < 0.95, acc
        m_hat = x.mean(axis=(0, 1)); s_hat = x.std(axis=(0, 1))
        # Monte-Carlo error: >= C independent chains, thinned samples
        z_score = np.abs(m_hat - mean) / (sd / np.sqrt(C))
        assert z_score.max() < 6.0, (method, z_score.max())
        assert np.abs(s_hat / sd - 1).max() < 0.08, (method, np.abs(s_hat / sd - 1).max())


@pytest.mark.parametrize("model,eng", [("8schools", engine.ENGINE_SIMT), ("radon", engine.ENGINE_SIMT),
                                       ("german_synth", engine.ENGINE_TCGEN05)])
def test_tuning_grid_in_one_launch_equals_separate_runs(model, eng):
    """arp_hmc_run_many: (chains x L) as the batch axis of one launch -- every run must reproduce the separate
    arp_hmc_run with the same L bit for bit (same initial states, same Philox streams, same arithmetic), including
    runs whose sample / burn-in / adaptation counts differ (--count_in_leapfrog_steps)."""
    C = 100 if model != "german_synth" else 128 + 60
    mc = common.model_config(model)
    D = mc.num_coords
    a, b = common.ab_for("NCP", D)
    z0 = common.random_states(model, D, C, seed=5, scale=0.3).astype(np.float32)
    Ls, Ss, Bs, As = [1, 4, 7], [12, 6, 4], [9, 5, 3], [8, 4, 2]
    sig = np.full(D, 0.05)
    eps = [sig / (L / 4.0) ** 2 for L in Ls]
    many = engine.hmc_run_many(mc, z0, eps, a, b, num_leapfrog_steps=Ls, num_results=Ss, num_burnin_steps=Bs,
                               num_adaptation_steps=As, seed=11, chain_offset=7, engine=eng)
    for i, L in enumerate(Ls):
        one = engine.hmc_run(mc, z0, eps[i], a, b, num_leapfrog_steps=L, num_results=Ss[i], num_burnin_steps=Bs[i],
                             num_adaptation_steps=As[i], seed=11, chain_offset=7, engine=eng, want_final=False)
        assert np.array_equal(many[i]["samples"], one["samples"]), L
        assert np.array_equal(many[i]["is_accepted"], one["is_accepted"]), L
        assert np.array_equal(many[i]["step_mult"], one["step_mult"]), L
        assert np.array_equal(many[i]["accept_count"], one["accept_count"]), L
        assert many[i]["num_transitions"] == one["num_transitions"]
    assert many[1]["is_accepted"].mean() > 0.2


@pytest.mark.parametrize("model,lpc", [("8schools", 1), ("radon", 8), ("time_series", 1), ("election", 32)])
@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_streaming_statistics_match_stored_trace(model, lpc, precision):
    """In-kernel streaming mean / variance / windowed ESS (for runs whose traces do not fit, BASELINE configs[4])
    against the same run's stored trace: moments to round-off; ESS equal to arp_ess wherever the first negative
    autocorrelation lies inside the window, flagged as truncated (and an upper bound) elsewhere."""
    C, L, S = 37, 3, 400
    W = 48 if model != "time_series" else 256     # the local-linear-trend chains mix slowly: wider window
    mc = common.model_config(model)
    D = mc.num_coords
    a, b = common.ab_for("NCP" if model != "time_series" else "CP", D)
    z0 = common.random_states(model, D, C, seed=8, scale=0.3)
    eps0 = np.full(D, 0.02 if model != "time_series" else 1e-4)
    out = engine.hmc_run(mc, z0, eps0, a, b, num_leapfrog_steps=L, num_results=S, num_burnin_steps=50,
                         num_adaptation_steps=40, seed=4, engine=engine.ENGINE_SIMT, lanes_per_chain=lpc,
                         precision=precision, stream_window=W)
    x = out["samples"].astype(np.float64)
    tol = 2e-4 if precision == "f32" else 1e-9
    assert np.abs(out["stream_mean"] - x.mean(0)).max() < tol * (1 + np.abs(x).max())
    v = x.var(0)
    okv = v > 1e-12 * (1 + np.abs(x).max() ** 2)
    assert np.abs(out["stream_var"][okv] / v[okv] - 1).max() < 50 * tol
    ref = engine.ess(out["samples"], precision=precision)
    tr = out["stream_truncated"].astype(bool)
    good = ~tr & np.isfinite(ref) & okv
    assert good.mean() > (0.3 if model != "time_series" else 0.02), good.mean()   # the window resolves a share of the series
    etol = 5e-3 if precision == "f32" else 1e-6
    assert np.abs(out["stream_ess"][good] / ref[good] - 1).max() < etol, np.abs(out["stream_ess"][good] / ref[good] - 1).max()
    if tr.any():                                    # truncated window: upper bound
        sel = tr & np.isfinite(ref)
        assert (out["stream_ess"][sel] >= ref[sel] * (1 - 1e-3)).all()
    # and against the oracle's restatement of the windowed estimator on the same trace: values AND truncation flags
    # (the block-wise lag accumulation, the partial last block -- S = 400 is not a multiple of 48 / 256 -- and the
    # head / tail corrections of the mean)
    ess_o, tr_o = O.windowed_ess(x, W)
    flips = (tr_o != tr) & okv                      # rho ~ 0 at the window edge may flip by round-off
    assert flips.mean() < (0.02 if precision == "f32" else 1e-3), flips.mean()
    same = (tr_o == tr) & okv & np.isfinite(ess_o)
    rel = np.abs(out["stream_ess"][same] / ess_o[same] - 1)
    if precision == "f64":
        assert rel.max() < 1e-6, rel.max()
    else:
        assert np.median(rel) < 1e-3 and np.quantile(rel, 0.98) < 0.1, (np.median(rel), np.quantile(rel, 0.98))
