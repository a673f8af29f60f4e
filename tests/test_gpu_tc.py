"""GPU parity of the tcgen05 German-credit engine (split-fp16 MMA, fp32 accumulate)
against the fp64 oracle, the fp64 SIMT check build and the SIMT fp32 engine.

Engine ids: ENGINE_TCGEN05 (2) and ENGINE_TCGEN05_STREAM (3) name the same kernel (k_german_tcs_hmc);
ENGINE_AUTO picks it for german_credit models with >= 256 chains -- that is what bench.py and main.py run."""
import numpy as np
import pytest

from autoreparam_b200 import engine
from oracle import oracle as O
from tests import common

pytestmark = pytest.mark.gpu
MODEL = "german_credit_lognormalcentered"


def _case(method, C, seed):
    mc = common.model_config("german_synth")
    raw = common.raw_data("german_synth")
    D = mc.num_coords
    a, b = common.ab_for(method, D)
    z0 = common.random_states("german_synth", D, C, seed=seed, scale=0.3).astype(np.float32).astype(np.float64)
    return mc, raw, D, a, b, z0


@pytest.mark.parametrize("method", ["CP", "NCP", "VIP_a", "VIP_ab"])
def test_tc_single_leapfrog_gradient(method):
    """One transition of one leapfrog step with zero momenta and a tiny step:
    the proposal equals z0 + eps^2/2 * grad, which exposes the tensor-core
    gradient itself: 1e-5 relative to the fp64 oracle gradient."""
    C = 9
    mc, raw, D, a, b, z0 = _case(method, C, seed=31)
    _, g_ref = O.log_joint_and_grad(MODEL, raw, z0, a, b)
    eps = 2.0 ** -6
    eps0 = np.full(D, eps)
    mom = np.zeros((1, C, D))
    lu = np.full((1, C), -1e30)   # always accept
    out = engine.hmc_run(mc, z0, eps0, a, b, num_leapfrog_steps=1, num_results=1, num_burnin_steps=0,
                         num_adaptation_steps=0, ext_momenta=mom, ext_log_u=lu, want_orig=True,
                         engine=engine.ENGINE_TCGEN05)
    assert out["is_accepted"].all()
    g_tc = (out["samples_orig"][0].astype(np.float64) - z0) / (0.5 * eps * eps)
    # z0 + eps^2/2 g is rounded to fp32: recover g only to ~ulp(z)/(eps^2/2); compare with that allowance
    err = np.abs(g_tc - g_ref).max(axis=1)
    allow = 1e-5 * np.maximum(np.abs(g_ref).max(axis=1), 1.0) + 2.0 ** -23 * np.abs(z0).max() / (0.5 * eps * eps)
    assert (err < allow).all(), (err, allow)


@pytest.mark.parametrize("method", ["NCP", "VIP_a"])
def test_tc_fixed_momenta_trajectory(method):
    C, L, S, burn, adapt = 6, 3, 3, 2, 4
    mc, raw, D, a, b, z0 = _case(method, C, seed=32)
    T = O.num_transitions(S, burn)
    rng = np.random.default_rng(8)
    mom = rng.standard_normal((T, C, D)).astype(np.float32).astype(np.float64)
    lu = np.log(rng.uniform(size=(T, C))).astype(np.float32).astype(np.float64)
    eps0 = (np.full(D, 0.01) * rng.uniform(0.5, 1.5, D)).astype(np.float32).astype(np.float64)
    ref = O.hmc_chain(MODEL, raw, z0, eps0, L, S, burn, adapt, a, b, momenta=mom, log_u=lu)
    out = engine.hmc_run(mc, z0, eps0, a, b, num_leapfrog_steps=L, num_results=S, num_burnin_steps=burn,
                         num_adaptation_steps=adapt, ext_momenta=mom, ext_log_u=lu, want_orig=True,
                         engine=engine.ENGINE_TCGEN05)
    assert (out["is_accepted"].astype(bool) == ref["is_accepted"]).all()
    assert ref["is_accepted"].mean() > 0
    err = common.rel_err(out["samples"].reshape(S * C, D), ref["samples_centered"].reshape(S * C, D)).max()
    # 8 transitions x 3 leapfrog steps amplify the per-evaluation round-off (which the test above pins at 1e-5)
    assert err < 2e-3, err
    assert common.rel_err(out["step_mult"], ref["step_mult"]).max() < 1e-3


def test_tc_matches_simt_engine_many_chains():
    """Several CTAs incl. a ragged tail, internal Philox momenta: the two engines
    draw identical random numbers, so short runs agree to fp32 round-off."""
    C, L, S, burn, adapt = 128 * 2 + 37, 4, 2, 2, 3
    mc, raw, D, a, b, z0 = _case("NCP", C, seed=33)
    eps0 = np.full(D, 0.02)
    kw = dict(num_leapfrog_steps=L, num_results=S, num_burnin_steps=burn, num_adaptation_steps=adapt, seed=99,
              chain_offset=5)
    o1 = engine.hmc_run(mc, z0, eps0, a, b, engine=engine.ENGINE_SIMT, **kw)
    o2 = engine.hmc_run(mc, z0, eps0, a, b, engine=engine.ENGINE_TCGEN05, **kw)
    same = o1["is_accepted"] == o2["is_accepted"]
    assert same.mean() > 0.995, same.mean()
    ok = same.all(axis=0)
    err = common.rel_err(o2["samples"][:, ok].reshape(-1, D), o1["samples"][:, ok].reshape(-1, D))
    # after the first transition the adapted step is 10 x eps0: a few chains amplify round-off strongly
    assert np.median(err) < 1e-5 and np.quantile(err, 0.99) < 1e-2, (np.median(err), np.quantile(err, 0.99))
    assert (o1["accept_count"][ok] == o2["accept_count"][ok]).all()


# --------------------------------------------------------------------------------------------
# synthetic 1000 x 25 (NF = 32 kernel variant) and the reference's real 1000 x 62 German credit data (NF = 64 variant,
# momentum in the global workspace)
# --------------------------------------------------------------------------------------------
def _case_model(model, method, C, seed):
    mc = common.model_config(model)
    raw = common.raw_data(model)
    D = mc.num_coords
    a, b = common.ab_for(method, D)
    z0 = common.random_states(model, D, C, seed=seed, scale=0.3).astype(np.float32).astype(np.float64)
    return mc, raw, D, a, b, z0


@pytest.mark.parametrize("model", ["german_synth", MODEL, "german_credit_gammascale"])
@pytest.mark.parametrize("method", ["CP", "NCP", "VIP_ab"])
def test_tcs_single_leapfrog_gradient(model, method):
    C = 9
    mc, raw, D, a, b, z0 = _case_model(model, method, C, seed=41)
    _, g_ref = O.log_joint_and_grad(MODEL if model == "german_synth" else model, raw, z0, a, b)
    eps = 2.0 ** -6
    out = engine.hmc_run(mc, z0, np.full(D, eps), a, b, num_leapfrog_steps=1, num_results=1, num_burnin_steps=0,
                         num_adaptation_steps=0, ext_momenta=np.zeros((1, C, D)), ext_log_u=np.full((1, C), -1e30),
                         want_orig=True, engine=engine.ENGINE_TCGEN05_STREAM)
    assert out["is_accepted"].all()
    g_tc = (out["samples_orig"][0].astype(np.float64) - z0) / (0.5 * eps * eps)
    err = np.abs(g_tc - g_ref).max(axis=1)
    allow = 1e-5 * np.maximum(np.abs(g_ref).max(axis=1), 1.0) + 2.0 ** -23 * np.abs(z0).max() / (0.5 * eps * eps)
    assert (err < allow).all(), (err, allow)


@pytest.mark.parametrize("n,f", [(100, 7), (128, 25), (300, 25), (1153, 32), (300, 33), (517, 64), (1000, 3)])
def test_tcs_ragged_shapes(n, f):
    """Chunk counts 1, 3 (odd), 10 and feature counts at the edges of the two kernel variants (F <= 32 / F <= 64,
    balanced ranges 7 = 2+2+2+1, 33 = 9+8+8+8, 3 = 1+1+1+0): gradient to 1e-5 and the log-joint value through a pinned
    Metropolis test, against the fp64 oracle."""
    from autoreparam_b200 import data as arp_data, models as arp_models
    raw = arp_data.synthetic_german_credit(n=n, f=f, seed=n * 100 + f)
    mc = arp_models.from_data(MODEL, raw)
    D, C = mc.num_coords, 130          # two CTAs, the second one with two valid chains
    a, b = common.ab_for("VIP_ab", D)
    z0 = common.random_states(MODEL, D, C, seed=n + f, scale=0.3).astype(np.float32).astype(np.float64)
    fgrad = lambda zz: O.log_joint_and_grad(MODEL, raw, zz, a, b)
    lp0, g_ref = fgrad(z0)
    eps = 2.0 ** -7
    out = engine.hmc_run(mc, z0, np.full(D, eps), a, b, num_leapfrog_steps=1, num_results=1, num_burnin_steps=0,
                         num_adaptation_steps=0, ext_momenta=np.zeros((1, C, D)), ext_log_u=np.full((1, C), -1e30),
                         want_orig=True, engine=engine.ENGINE_TCGEN05_STREAM)
    assert out["is_accepted"].all()
    g_tc = (out["samples_orig"][0].astype(np.float64) - z0) / (0.5 * eps * eps)
    err = np.abs(g_tc - g_ref).max(axis=1)
    allow = 1e-5 * np.maximum(np.abs(g_ref).max(axis=1), 1.0) + 2.0 ** -23 * np.abs(z0).max() / (0.5 * eps * eps)
    assert (err < allow).all(), (err / allow).max()
    # value of the log-joint: two leapfrog steps with random momenta, log u just below / above the oracle's log alpha
    rng = np.random.default_rng(n + 7 * f)
    L, eps = 2, 1e-3
    mom = rng.standard_normal((1, C, D))
    g = g_ref
    v, x = mom[0].copy(), z0.copy()
    for _ in range(L):
        v = v + 0.5 * eps * g
        x = x + eps * v
        lpx, g = fgrad(x)
        v = v + 0.5 * eps * g
    la = lpx - lp0 + 0.5 * (mom[0] ** 2).sum(1) - 0.5 * (v ** 2).sum(1)
    tol = 2e-3 + 5e-6 * np.abs(lpx)
    for sign, want in ((-1.0, 1), (1.0, 0)):
        o2 = engine.hmc_run(mc, z0, np.full(D, eps), a, b, num_leapfrog_steps=L, num_results=1, num_burnin_steps=0,
                            num_adaptation_steps=0, ext_momenta=mom, ext_log_u=(la + sign * tol)[None, :],
                            engine=engine.ENGINE_TCGEN05_STREAM)
        assert (o2["is_accepted"][0] == want).all(), (sign, o2["is_accepted"][0])


@pytest.mark.parametrize("L", [1, 2, 3, 5])
def test_tcs_internal_momenta_any_leapfrog_count(L):
    """The streaming engine draws the Philox momenta of transition t + 1 inside transition t (split over the wait
    windows of its last two leapfrog steps; all in the only step when L = 1): with small steps the run must
    reproduce the SIMT engine, which draws them at the start of each transition, to fp32 round-off."""
    C, S = 128 + 9, 4
    mc, raw, D, a, b, z0 = _case_model("german_synth", "NCP", C, seed=61)
    eps0 = np.full(D, 0.004)
    kw = dict(num_leapfrog_steps=L, num_results=S, num_burnin_steps=1, num_adaptation_steps=0, seed=123, chain_offset=3)
    o1 = engine.hmc_run(mc, z0, eps0, a, b, engine=engine.ENGINE_SIMT, **kw)
    o2 = engine.hmc_run(mc, z0, eps0, a, b, engine=engine.ENGINE_TCGEN05_STREAM, **kw)
    same = (o1["is_accepted"] == o2["is_accepted"]).all(axis=0)
    assert same.mean() > 0.98, same.mean()
    assert o1["is_accepted"].mean() > 0.5
    err = common.rel_err(o2["samples"][:, same].reshape(-1, D), o1["samples"][:, same].reshape(-1, D))
    assert err.max() < 2e-4, err.max()


@pytest.mark.parametrize("s0", [0.2, 0.35, 0.6])
def test_tcs_confident_logits_gradient(s0):
    """Large coefficients (|eta| up to several hundred): the clamp that keeps the product of four sigmoid
    denominators (one shared reciprocal) finite is active, and far tails (sigmoid == 0 or 1 in fp32) must still
    give the fp64 gradient to 1e-5 of its max-norm."""
    C = 37
    mc, raw, D, a, b, z0 = _case_model("german_synth", "NCP", C, seed=47)
    z0[:, 0] = s0          # NCP: overall_log_scale = 10 z  ->  beta ~ exp(10 s0) * 0.3 z
    lp_ref, g_ref = O.log_joint_and_grad(MODEL, raw, z0, a, b)
    eps = 2.0 ** -9
    out = engine.hmc_run(mc, z0, np.full(D, eps), a, b, num_leapfrog_steps=1, num_results=1, num_burnin_steps=0,
                         num_adaptation_steps=0, ext_momenta=np.zeros((1, C, D)), ext_log_u=np.full((1, C), -1e30),
                         want_orig=True, engine=engine.ENGINE_TCGEN05_STREAM)
    assert out["is_accepted"].all()
    g_tc = (out["samples_orig"][0].astype(np.float64) - z0) / (0.5 * eps * eps)
    err = np.abs(g_tc - g_ref).max(axis=1)
    allow = 1e-5 * np.maximum(np.abs(g_ref).max(axis=1), 1.0) + 2.0 ** -23 * np.abs(z0).max() / (0.5 * eps * eps)
    assert (err < allow).all(), (err / allow).max()


@pytest.mark.parametrize("s0", [0.0, 0.35])
def test_tcs_log_likelihood_value_pins_accept(s0):
    """The log-joint value of the proposal (last leapfrog step: ln2 * sum (h + log2 q)) decides the Metropolis test:
    with log u placed just below / above the fp64 oracle's log alpha the chain must accept / reject."""
    C = 64
    mc, raw, D, a, b, z0 = _case_model("german_synth", "NCP", C, seed=48)
    z0[:, 0] = s0
    rng = np.random.default_rng(5)
    # s0 = 0.35: |grad| ~ 1e5, so the step must be tiny for the kinetic energies (fp32) to stay O(1)
    L, eps = 2, (1e-3 if s0 == 0.0 else 2e-6)
    mom = rng.standard_normal((1, C, D))
    f = lambda zz: O.log_joint_and_grad(MODEL, raw, zz, a, b)
    lp0, g = f(z0)
    v, x = mom[0].copy(), z0.copy()
    for _ in range(L):   # TFP op order, fp64
        v = v + 0.5 * eps * g
        x = x + eps * v
        lpx, g = f(x)
        v = v + 0.5 * eps * g
    la = lpx - lp0 + 0.5 * (mom[0] ** 2).sum(1) - 0.5 * (v ** 2).sum(1)
    tol = 2e-3 + 5e-6 * np.abs(lpx)
    e = engine.ENGINE_TCGEN05_STREAM
    for sign, want in ((-1.0, 1), (1.0, 0)):
        out = engine.hmc_run(mc, z0, np.full(D, eps), a, b, num_leapfrog_steps=L, num_results=1, num_burnin_steps=0,
                             num_adaptation_steps=0, ext_momenta=mom, ext_log_u=(la + sign * tol)[None, :], engine=e)
        assert (out["is_accepted"][0] == want).all(), (sign, out["is_accepted"][0], la)


@pytest.mark.parametrize("model", ["german_synth", MODEL, "german_credit_gammascale"])
def test_tcs_fixed_momenta_trajectory(model):
    C, L, S, burn, adapt = 6, 3, 3, 2, 4
    mc, raw, D, a, b, z0 = _case_model(model, "VIP_a", C, seed=42)
    T = O.num_transitions(S, burn)
    rng = np.random.default_rng(9)
    mom = rng.standard_normal((T, C, D)).astype(np.float32).astype(np.float64)
    lu = np.log(rng.uniform(size=(T, C))).astype(np.float32).astype(np.float64)
    eps0 = (np.full(D, 0.002 if "gamma" in model else 0.01) * rng.uniform(0.5, 1.5, D)).astype(np.float32).astype(np.float64)
    ref = O.hmc_chain(MODEL if model == "german_synth" else model, raw, z0, eps0, L, S, burn, adapt, a, b, momenta=mom,
                      log_u=lu)
    out = engine.hmc_run(mc, z0, eps0, a, b, num_leapfrog_steps=L, num_results=S, num_burnin_steps=burn,
                         num_adaptation_steps=adapt, ext_momenta=mom, ext_log_u=lu, want_orig=True,
                         engine=engine.ENGINE_TCGEN05_STREAM)
    assert (out["is_accepted"].astype(bool) == ref["is_accepted"]).all()
    assert ref["is_accepted"].mean() > 0
    err = common.rel_err(out["samples"].reshape(S * C, D), ref["samples_centered"].reshape(S * C, D)).max()
    tol = 2e-2 if "gamma" in model else 2e-3   # exp(10 z0 + v) amplifies fp32 round-off (see test_gpu_hmc)
    assert err < tol, err
    assert common.rel_err(out["final_z"], ref["z"]).max() < tol


@pytest.mark.parametrize("model", ["german_synth", MODEL])
def test_tcs_matches_simt_engine_many_chains(model):
    C, L, S, burn, adapt = 128 * 2 + 37, 4, 2, 2, 3
    mc, raw, D, a, b, z0 = _case_model(model, "NCP", C, seed=43)
    eps0 = np.full(D, 0.01)
    kw = dict(num_leapfrog_steps=L, num_results=S, num_burnin_steps=burn, num_adaptation_steps=adapt, seed=77,
              chain_offset=11)
    o1 = engine.hmc_run(mc, z0, eps0, a, b, engine=engine.ENGINE_SIMT, **kw)
    o2 = engine.hmc_run(mc, z0, eps0, a, b, engine=engine.ENGINE_TCGEN05_STREAM, **kw)
    same = o1["is_accepted"] == o2["is_accepted"]
    assert same.mean() > 0.99, same.mean()
    ok = same.all(axis=0)
    err = common.rel_err(o2["samples"][:, ok].reshape(-1, D), o1["samples"][:, ok].reshape(-1, D))
    assert np.median(err) < 1e-5 and np.quantile(err, 0.99) < 1e-2, (np.median(err), np.quantile(err, 0.99))


# --------------------------------------------------------------------------------------------
# the gradient itself, elementwise (arp_log_joint_grad_engine: the tensor-core kernel's own epilogue / GEMM2 / site
# reverse code evaluated at the input state)
# --------------------------------------------------------------------------------------------
def _elementwise_ok(g, g_ref, rtol=1e-5, floor=1e-6):
    """|g - g_ref| <= rtol |g_ref| elementwise, with an absolute floor of `floor` x the chain's largest gradient entry
    (entries that vanish by cancellation over 1000 observations have no relative accuracy in ANY fp32 evaluation)."""
    g, g_ref = np.asarray(g, np.float64), np.asarray(g_ref, np.float64)
    allow = rtol * np.abs(g_ref) + floor * np.abs(g_ref).max(axis=1, keepdims=True)
    return np.abs(g - g_ref) / allow


@pytest.mark.parametrize("model", ["german_synth", MODEL, "german_credit_gammascale"])
@pytest.mark.parametrize("method", ["CP", "NCP", "VIP_a", "VIP_ab"])
def test_tc_gradient_elementwise(model, method):
    """BASELINE configs[1] shape (synthetic 1000 x 25) and the real 1000 x 62 data: log-joint to 1e-5 relative and the
    gradient ELEMENTWISE to 1e-5 relative against the fp64 oracle, several CTAs incl. a ragged one."""
    C = 128 * 2 + 19
    mc, raw, D, a, b, z0 = _case_model(model, method, C, seed=71)
    name = MODEL if model == "german_synth" else model
    sel = np.r_[0:24, 120:136, C - 19:C]      # oracle (autograd, one chain at a time) on chains of all three CTAs
    lp_ref, g_ref = O.log_joint_and_grad(name, raw, z0[sel], a, b)
    xc_ref = O.to_centered(name, raw, z0[sel], a, b)
    lp, g, xc = engine.log_joint_grad(mc, z0.astype(np.float32), a, b, engine=engine.ENGINE_TCGEN05)
    assert np.isfinite(lp).all() and np.isfinite(g).all()
    assert (np.abs(lp[sel] - lp_ref) <= 1e-5 * np.abs(lp_ref)).all(), np.abs(lp[sel] / lp_ref - 1).max()
    ratio = _elementwise_ok(g[sel], g_ref)
    assert ratio.max() < 1.0, (ratio.max(), np.unravel_index(ratio.argmax(), ratio.shape))
    assert common.rel_err(xc[sel], xc_ref).max() < 1e-5
    # the SIMT fp32 engine meets the same criterion (and the two agree with each other to the same level)
    lp1, g1, _ = engine.log_joint_grad(mc, z0.astype(np.float32), a, b, engine=engine.ENGINE_SIMT)
    assert _elementwise_ok(g1[sel], g_ref).max() < 1.0


def test_tc_gradient_rejects_fp16_overflow():
    """A coefficient outside the fp16 range of the tensor-core A operand is rejected outright by the engine
    (lp = -inf), never silently clamped."""
    C = 5
    mc, raw, D, a, b, z0 = _case_model("german_synth", "CP", C, seed=72)
    z0[2, 1 + 25 + 3] = 1e5      # CP: beta_3 = z itself
    lp, g, _ = engine.log_joint_grad(mc, z0.astype(np.float32), a, b, engine=engine.ENGINE_TCGEN05)
    assert lp[2] == -np.inf and np.isnan(g[2]).all()
    ok = np.arange(C) != 2
    assert np.isfinite(lp[ok]).all() and np.isfinite(g[ok]).all()


# --------------------------------------------------------------------------------------------
# long runs of the engine that ships (ENGINE_AUTO) against the fp64 SIMT check build
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model,eng", [("german_synth", "auto"), ("german_synth", "stream"), (MODEL, "auto")])
def test_default_engine_long_run_statistics(model, eng):
    """>= 2000 transitions x 512 chains (adaptation included): acceptance rate, adapted step sizes and posterior
    moments of the tensor-core engine against the fp64 SIMT build (different seeds, so only statistics can agree).
    This is the test that would have caught the fp16-overflow bug of round 1 (acceptance 0.858 -> 0.878)."""
    C = 512
    mc, raw, D, a, b, z0 = _case_model(model, "NCP", C, seed=34)
    eps0 = np.full(D, 0.1269 if model == "german_synth" else 0.05)
    kw = dict(num_leapfrog_steps=4, num_results=800, num_burnin_steps=500, num_adaptation_steps=400, want_final=False)
    assert engine.hmc_num_transitions(800, 500) >= 2000
    e = engine.ENGINE_AUTO if eng == "auto" else engine.ENGINE_TCGEN05_STREAM
    o_tc = engine.hmc_run(mc, z0 * 0.3, eps0, a, b, engine=e, seed=2, **kw)
    o_64 = engine.hmc_run(mc, z0 * 0.3, eps0, a, b, engine=engine.ENGINE_SIMT, seed=1, precision="f64", **kw)
    T = o_tc["num_transitions"]
    acc_tc, acc_64 = o_tc["accept_count"].sum() / (T * C), o_64["accept_count"].sum() / (T * C)
    # binomial sd of a rate over T x C ~ 1e6 transitions is ~4e-4; chains are autocorrelated: allow 5e-3
    assert abs(acc_tc - acc_64) < 5e-3, (acc_tc, acc_64)
    assert 0.6 < acc_64 < 0.95
    # adapted step-size multipliers: same distribution over chains
    m_tc, m_64 = np.log(o_tc["step_mult"].astype(np.float64)), np.log(o_64["step_mult"])
    assert abs(m_tc.mean() - m_64.mean()) < 6 * np.sqrt(m_tc.var() / C + m_64.var() / C) + 1e-3, (m_tc.mean(), m_64.mean())
    x1, x2 = o_64["samples"].astype(np.float64), o_tc["samples"].astype(np.float64)
    # per-chain means are independent draws: compare the grand means with the between-chain standard error
    c1, c2 = x1.mean(axis=0), x2.mean(axis=0)
    z = np.abs(c1.mean(0) - c2.mean(0)) / np.sqrt(c1.var(0, ddof=1) / C + c2.var(0, ddof=1) / C)
    assert z.max() < 6.0, (z.max(), int(z.argmax()))
    s1, s2 = x1.std(axis=(0, 1)), x2.std(axis=(0, 1))
    assert np.median(np.abs(s1 / s2 - 1)) < 0.03 and np.abs(s1 / s2 - 1).max() < 0.2, np.abs(s1 / s2 - 1).max()


def test_accept_decisions_on_diverging_trajectories():
    """profiles/diag_accept.py as a test: with steps large enough that a few percent of the trajectories diverge,
    the tensor-core engine takes the same accept decisions as the fp64 SIMT build as often as the fp32 SIMT engine
    does (identical Philox streams)."""
    C, L = 128 * 2 + 37, 4
    mc, raw, D, a, b, z0 = _case_model("german_synth", "NCP", C, seed=43)
    for eps in (0.05, 0.2):
        kw = dict(num_leapfrog_steps=L, num_results=1, num_burnin_steps=0, num_adaptation_steps=0, seed=77, chain_offset=11)
        e0 = np.full(D, eps)
        o64 = engine.hmc_run(mc, z0, e0, a, b, engine=engine.ENGINE_SIMT, precision="f64", **kw)
        o32 = engine.hmc_run(mc, z0, e0, a, b, engine=engine.ENGINE_SIMT, **kw)
        otc = engine.hmc_run(mc, z0, e0, a, b, engine=engine.ENGINE_AUTO, **kw)
        same_tc = (otc["is_accepted"] == o64["is_accepted"]).mean()
        same_32 = (o32["is_accepted"] == o64["is_accepted"]).mean()
        assert same_tc >= min(same_32, 0.99) - 0.01, (eps, same_tc, same_32)
        assert abs(otc["is_accepted"].mean() - o64["is_accepted"].mean()) < 0.02
