"""GPU parity: effective sample size and the fused VI kernel."""
import numpy as np
import pytest

from autoreparam_b200 import engine
from oracle import oracle as O
from tests import common

pytestmark = pytest.mark.gpu


def _ar1(S, shape, phi, rng):
    x = np.zeros((S,) + shape)
    e = rng.standard_normal((S,) + shape)
    for t in range(1, S):
        x[t] = phi * x[t - 1] + e[t]
    return x


@pytest.mark.parametrize("S", [50, 1000, 1024, 1025, 4097, 50000])
@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_ess_matches_fft_oracle(S, precision):
    """S <= 1024: the shared-memory FFT kernel; longer traces: the direct lag-sum kernels; 50 000 is the reference's
    --num_samples default (main.py:95)."""
    rng = np.random.default_rng(S)
    C, D = (5, 7) if S < 50000 else (3, 4)
    phi = rng.uniform(-0.5, 0.95, (C, D))
    x = _ar1(S, (C, D), phi, rng) + 3.0
    x = x.astype(np.float32).astype(np.float64)
    ref = O.effective_sample_size(x)
    got = engine.ess(x, precision=precision)
    tol = 2e-3 if precision == "f32" else 1e-8
    assert np.abs(got / ref - 1).max() < tol, np.abs(got / ref - 1).max()


@pytest.mark.parametrize("precision,S", [("f32", 200), ("f32", 203), ("f64", 200)])
@pytest.mark.parametrize("C,D", [(70, 3), (33, 300), (1, 1)])
def test_ess_shapes_and_moments(C, D, precision, S):
    """Ragged chain / coordinate counts, the 128-bit-load path (fp32, S % 4 == 0) and the scalar path, and the
    per-chain moments the R-hat reduction consumes."""
    rng = np.random.default_rng(C * 1000 + D)
    phi = rng.uniform(-0.5, 0.9, (C, D))
    x = (_ar1(S, (C, D), phi, rng) + rng.standard_normal((C, D))).astype(np.float32).astype(np.float64)
    ref = O.effective_sample_size(x)
    got, mean, var = engine.ess(x, precision=precision, want_moments=True)
    tol = 2e-3 if precision == "f32" else 1e-8
    assert np.abs(got / ref - 1).max() < tol
    mtol = 1e-5 if precision == "f32" else 1e-10
    assert np.abs(mean - x.mean(0)).max() < mtol * (1 + np.abs(x).max()) and np.abs(var / x.var(0) - 1).max() < 10 * mtol


def test_ess_edge_cases():
    S, C, D = 64, 2, 3
    x = np.zeros((S, C, D))
    x[:, 0, 0] = 1.5                                   # constant series -> NaN (nan_to_num'd by get_min_ess)
    x[:, 0, 1] = np.arange(S) % 2                      # perfectly anti-correlated: rho_1 < 0 -> ESS = S
    x[:, 0, 2] = np.arange(S)                          # trend: long positive autocorrelation
    x[:, 1] = np.random.default_rng(0).standard_normal((S, D))
    ref = O.effective_sample_size(x)
    got = engine.ess(x, precision="f64")
    assert np.isnan(got[0, 0]) and np.isnan(ref[0, 0])
    ok = ~np.isnan(ref)
    assert np.abs(got[ok] / ref[ok] - 1).max() < 1e-8
    assert abs(got[0, 1] - S) < 1e-9


@pytest.mark.parametrize("model,method", [("8schools", "cVIP"), ("8schools", "NCP"), ("radon", "cVIP"),
                                          ("election", "dVIP"), ("electric", "cVIP"),
                                          ("german_credit_lognormalcentered", "cVIP"), ("time_series", "CP")])
def test_vi_steps_match_oracle_fp64(model, method):
    """A few Adam steps on the ELBO with injected normals: ELBO timeline and
    parameters equal the autograd oracle (fp64 build)."""
    mc = common.model_config(model, "MN")
    raw = common.raw_data(model, "MN")
    D = mc.num_coords
    S, steps, lr = 4, 6, 0.05
    rng = np.random.default_rng(21)
    eps = rng.standard_normal((steps, S, D))
    loc0 = 0.01 * rng.standard_normal(D)
    rho0 = np.full(D, -2.0)
    if method == "cVIP":
        a, b, al0 = np.full(D, 0.5), np.ones(D), np.zeros(D)
    else:
        a, b = common.ab_for(method, D)
        al0 = None
    ref = O.vi_run(model, raw, loc0, rho0, eps, lr, steps, a=a, b=b, a_logit0=al0)
    kw = {} if al0 is None else dict(u=al0[None], a_index=np.arange(D), b_index=np.full(D, -1), num_params=D)
    out = engine.vi_run(mc, a, b, loc0[None], rho0[None], [lr], num_mc_samples=S, num_optimization_steps=steps,
                        ext_eps=eps, precision="f64", **kw)
    assert np.abs(out["elbo"][0] / ref["elbo"] - 1).max() < 1e-8, np.abs(out["elbo"][0] / ref["elbo"] - 1).max()
    assert np.abs(out["loc"][0] - ref["loc"]).max() < 1e-7
    assert np.abs(out["rho"][0] - ref["rho"]).max() < 1e-7
    if al0 is not None:
        assert np.abs(out["u"][0] - ref["a_logit"]).max() < 1e-7
        assert np.abs(out["u"][0]).max() > 1e-3   # the parameterisation really moved


@pytest.mark.parametrize("model,S", [("8schools", 20), ("8schools", 70), ("election", 37), ("radon", 300)])
def test_vi_ragged_sample_counts_match_oracle_fp64(model, S):
    """The cluster kernel deals the S Monte-Carlo samples to 8 CTAs x 32 sample slots: sample counts that leave CTAs
    empty (S = 20: 3 per CTA, the last CTA idle), partly filled, or that need a second pass per CTA (S = 300: 38 per
    CTA) give the same ELBO / Adam steps as the oracle."""
    mc = common.model_config(model, "MN")
    raw = common.raw_data(model, "MN")
    D = mc.num_coords
    steps, lr = 3, 0.05
    rng = np.random.default_rng(23)
    eps = rng.standard_normal((steps, S, D))
    loc0 = 0.01 * rng.standard_normal(D)
    rho0 = np.full(D, -2.0)
    a, b, al0 = np.full(D, 0.5), np.ones(D), np.zeros(D)
    ref = O.vi_run(model, raw, loc0, rho0, eps, lr, steps, a=a, b=b, a_logit0=al0)
    out = engine.vi_run(mc, a, b, loc0[None], rho0[None], [lr], num_mc_samples=S, num_optimization_steps=steps,
                        ext_eps=eps, precision="f64", u=al0[None], a_index=np.arange(D), b_index=np.full(D, -1),
                        num_params=D)
    assert np.abs(out["elbo"][0] / ref["elbo"] - 1).max() < 1e-8, np.abs(out["elbo"][0] / ref["elbo"] - 1).max()
    assert np.abs(out["loc"][0] - ref["loc"]).max() < 1e-7
    assert np.abs(out["rho"][0] - ref["rho"]).max() < 1e-7
    assert np.abs(out["u"][0] - ref["a_logit"]).max() < 1e-7


@pytest.mark.parametrize("model", ["8schools", "german_credit_lognormalcentered", "german_credit_gammascale", "election",
                                   "electric", "time_series", "radon_stddvs"])
@pytest.mark.parametrize("mode", ["tied_b_eq_a", "untied", "untied_prior"])
def test_vi_learnable_b_matches_oracle_fp64(model, mode):
    """The paper's tied b = a, the untied mode (independent a / b with the shapes of the site's loc / scale:
    program_transformations.py:512-523) and --discrete_prior (main.py:244-253): a few Adam steps with injected
    normals equal the autograd oracle -- this pins the d/db adjoint of every model, the parameter sharing of scalar
    loc / scale sites and the prior's gradient."""
    from autoreparam_b200 import graphs
    mc = common.model_config(model, "MN")
    raw = common.raw_data(model, "MN")
    D = mc.num_coords
    S, steps, lr = 3, 5, 0.05
    rng = np.random.default_rng(22)
    eps = rng.standard_normal((steps, S, D))
    loc0 = 0.01 * rng.standard_normal(D)
    rho0 = np.full(D, -2.0)
    tgt = graphs.make_cvip_graph(mc, tied_pparams=(mode == "tied_b_eq_a"), tied_b_as_written=False)
    P = tgt.num_params
    assert P > 0 and tgt.a_index.max() < P and tgt.b_index.max() < P
    u0 = 0.3 * rng.standard_normal(P)       # away from the symmetric start so that every adjoint is exercised
    prior = mode == "untied_prior"
    ref = O.vi_run(model, raw, loc0, rho0, eps, lr, steps, a=tgt.a, b=tgt.b, u0=u0, a_index=tgt.a_index,
                   b_index=tgt.b_index, discrete_prior=prior)
    out = engine.vi_run(mc, tgt.a, tgt.b, loc0[None], rho0[None], [lr], num_mc_samples=S, num_optimization_steps=steps,
                        u=u0[None], a_index=tgt.a_index, b_index=tgt.b_index, num_params=P, discrete_prior=prior,
                        ext_eps=eps, precision="f64")
    assert np.abs(out["elbo"][0] / ref["elbo"] - 1).max() < 1e-8, np.abs(out["elbo"][0] / ref["elbo"] - 1).max()
    assert np.abs(out["prior_logp"][0] - ref["prior_logp"]).max() < 1e-9
    assert np.abs(out["loc"][0] - ref["loc"]).max() < 1e-7
    assert np.abs(out["rho"][0] - ref["rho"]).max() < 1e-7
    assert np.abs(out["u"][0] - ref["u"]).max() < 1e-7
    assert np.abs(out["u"][0] - u0).max() > 1e-3
    if prior:
        assert np.abs(ref["prior_logp"]).max() > 0


@pytest.mark.parametrize("model", common.MODELS)
def test_param_adjoints_match_autograd(model):
    """d log_joint / d a and / d b per coordinate (arp_log_joint_param_grad) against autograd, general (a, b)."""
    import torch
    mc = common.model_config(model)
    raw = common.raw_data(model)
    D = mc.num_coords
    a, b = common.ab_for("VIP_ab", D)
    z = common.random_states(model, D, 3, seed=4).astype(np.float32).astype(np.float64)
    ab, bb = engine.log_joint_param_grad(mc, z, a, b, precision="f64")
    for c in range(3):
        at = torch.tensor(a, dtype=torch.float64, requires_grad=True)
        bt = torch.tensor(b, dtype=torch.float64, requires_grad=True)
        tr = O.Tracer(O._split(model, raw, torch.as_tensor(z[c]), torch.float64), O._split(model, raw, at, torch.float64),
                      O._split(model, raw, bt, torch.float64), torch.float64)
        O._BODIES[model](tr, raw)
        ga, gb = torch.autograd.grad(tr.lp, [at, bt], allow_unused=True)
        ga = np.zeros(D) if ga is None else ga.numpy()
        gb = np.zeros(D) if gb is None else gb.numpy()
        assert common.rel_err(ab[c:c + 1], ga[None]).max() < 1e-10, ("abar", c)
        assert common.rel_err(bb[c:c + 1], gb[None]).max() < 1e-10, ("bbar", c, np.abs(bb[c] - gb).argmax())


def test_vi_learning_rates_run_concurrently_and_converge():
    """All learning rates in one launch; 8schools CP ELBO improves and the
    schedule / ELBO bookkeeping of find_best_learning_rate can be applied."""
    mc = common.model_config("8schools")
    D = mc.num_coords
    lrs = [0.02, 0.05, 0.1, 0.2, 0.4]
    rng = np.random.default_rng(2)
    loc0 = 0.01 * rng.standard_normal((len(lrs), D))
    out = engine.vi_run(mc, np.zeros(D), np.zeros(D), loc0, np.full((len(lrs), D), -2.0), lrs, num_mc_samples=256,
                        num_optimization_steps=600, seed=5, precision="f32")
    elbo = out["elbo"]
    assert elbo.shape == (5, 600) and np.isfinite(elbo).all()
    final = elbo[:, -32:].mean(axis=1)
    assert (final > elbo[:, :5].mean(axis=1)).all()
    # 8 schools: ELBO of a good mean-field fit in the non-centred parameterisation is about -31.7
    assert -33.5 < final.max() < -30.5, final
