"""GPU parity of the tcgen05 German-credit engine (split-fp16 MMA, fp32 accumulate)
against the fp64 oracle and against the SIMT fp32 engine."""
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


def test_tc_posterior_agrees_with_simt():
    """Posterior means / sds from the two engines agree within Monte-Carlo error."""
    C = 512
    mc, raw, D, a, b, z0 = _case("NCP", C, seed=34)
    eps0 = np.full(D, 0.05)
    kw = dict(num_leapfrog_steps=4, num_results=150, num_burnin_steps=400, num_adaptation_steps=300)
    o1 = engine.hmc_run(mc, z0 * 0.3, eps0, a, b, engine=engine.ENGINE_SIMT, seed=1, **kw)
    o2 = engine.hmc_run(mc, z0 * 0.3, eps0, a, b, engine=engine.ENGINE_TCGEN05, seed=2, **kw)
    for o in (o1, o2):
        assert 0.5 < o["is_accepted"].mean() < 0.98
    x1, x2 = o1["samples"].astype(np.float64), o2["samples"].astype(np.float64)
    m1, m2 = x1.mean(axis=(0, 1)), x2.mean(axis=(0, 1))
    s1, s2 = x1.std(axis=(0, 1)), x2.std(axis=(0, 1))
    # chain means are independent across chains: sd of the grand mean <= posterior sd / sqrt(C)
    z = np.abs(m1 - m2) / (np.sqrt(s1 ** 2 + s2 ** 2) / np.sqrt(C))
    assert z.max() < 6.0, z.max()
    assert np.median(np.abs(s1 / s2 - 1)) < 0.05 and np.abs(s1 / s2 - 1).max() < 0.25  # slow-mixing log-scales


# --------------------------------------------------------------------------------------------
# streaming variant (X chunk images ring-buffered from L2): synthetic 1000 x 25 (NF = 32) and the
# reference's real 1000 x 62 German credit data (NF = 64, momentum in the global workspace)
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


@pytest.mark.parametrize("eng", ["stream", "dual"])
@pytest.mark.parametrize("s0", [0.2, 0.35, 0.6])
def test_tcs_confident_logits_gradient(eng, s0):
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
                         want_orig=True,
                         engine=engine.ENGINE_TCGEN05_STREAM if eng == "stream" else engine.ENGINE_TCGEN05_DUAL)
    assert out["is_accepted"].all()
    g_tc = (out["samples_orig"][0].astype(np.float64) - z0) / (0.5 * eps * eps)
    err = np.abs(g_tc - g_ref).max(axis=1)
    allow = 1e-5 * np.maximum(np.abs(g_ref).max(axis=1), 1.0) + 2.0 ** -23 * np.abs(z0).max() / (0.5 * eps * eps)
    assert (err < allow).all(), (err / allow).max()


@pytest.mark.parametrize("eng", ["stream", "dual"])
@pytest.mark.parametrize("s0", [0.0, 0.35])
def test_tcs_log_likelihood_value_pins_accept(eng, s0):
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
    e = engine.ENGINE_TCGEN05_STREAM if eng == "stream" else engine.ENGINE_TCGEN05_DUAL
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
# dual-tile variant (two 64-chain M = 64 tiles per CTA, half-warp TMEM accesses): F <= 32
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("method", ["CP", "NCP", "VIP_ab"])
def test_tcd_single_leapfrog_gradient(method):
    C = 128 + 70      # both tiles of CTA 0, tile 0 and part of tile 1 of CTA 1
    mc, raw, D, a, b, z0 = _case_model("german_synth", method, C, seed=51)
    _, g_ref = O.log_joint_and_grad(MODEL, raw, z0, a, b)
    eps = 2.0 ** -6
    out = engine.hmc_run(mc, z0, np.full(D, eps), a, b, num_leapfrog_steps=1, num_results=1, num_burnin_steps=0,
                         num_adaptation_steps=0, ext_momenta=np.zeros((1, C, D)), ext_log_u=np.full((1, C), -1e30),
                         want_orig=True, engine=engine.ENGINE_TCGEN05_DUAL)
    assert out["is_accepted"].all()
    g_tc = (out["samples_orig"][0].astype(np.float64) - z0) / (0.5 * eps * eps)
    err = np.abs(g_tc - g_ref).max(axis=1)
    allow = 1e-5 * np.maximum(np.abs(g_ref).max(axis=1), 1.0) + 2.0 ** -23 * np.abs(z0).max() / (0.5 * eps * eps)
    assert (err < allow).all(), (err, allow)


def test_tcd_fixed_momenta_trajectory():
    C, L, S, burn, adapt = 70, 3, 3, 2, 4
    mc, raw, D, a, b, z0 = _case_model("german_synth", "VIP_a", C, seed=52)
    T = O.num_transitions(S, burn)
    rng = np.random.default_rng(9)
    mom = rng.standard_normal((T, C, D)).astype(np.float32).astype(np.float64)
    lu = np.log(rng.uniform(size=(T, C))).astype(np.float32).astype(np.float64)
    eps0 = (np.full(D, 0.01) * rng.uniform(0.5, 1.5, D)).astype(np.float32).astype(np.float64)
    sel = [0, 15, 16, 63, 64, 69]     # oracle on a few chains of both tiles
    ref = O.hmc_chain(MODEL, raw, z0[sel], eps0, L, S, burn, adapt, a, b, momenta=mom[:, sel], log_u=lu[:, sel])
    out = engine.hmc_run(mc, z0, eps0, a, b, num_leapfrog_steps=L, num_results=S, num_burnin_steps=burn,
                         num_adaptation_steps=adapt, ext_momenta=mom, ext_log_u=lu, want_orig=True,
                         engine=engine.ENGINE_TCGEN05_DUAL)
    assert (out["is_accepted"][:, sel].astype(bool) == ref["is_accepted"]).all()
    err = common.rel_err(out["samples"][:, sel].reshape(-1, D), ref["samples_centered"].reshape(-1, D)).max()
    assert err < 2e-3, err
    assert common.rel_err(out["final_z"][sel], ref["z"]).max() < 2e-3


@pytest.mark.parametrize("model", ["german_synth", "german_credit_gammascale_f24"])
def test_tcd_identical_to_streaming_engine(model):
    """Per chain the dual-tile kernel does the same arithmetic in the same order as the single-tile
    streaming kernel: results must be identical (Philox momenta, several CTAs, ragged tail)."""
    C, L, S, burn, adapt = 128 * 3 + 37, 4, 3, 4, 5
    if model == "german_synth":
        mc, raw, D, a, b, z0 = _case_model(model, "NCP", C, seed=53)
    else:   # gamma-scale prior on the synthetic 1000 x 25 design matrix
        from autoreparam_b200 import models
        base = common.raw_data("german_synth")
        mc = models.from_data("german_credit_gammascale", base)
        D = mc.num_coords
        a, b = common.ab_for("NCP", D)
        z0 = common.random_states("german_credit_gammascale", D, C, seed=53, scale=0.3).astype(np.float32).astype(np.float64)
    eps0 = np.full(D, 0.002 if "gamma" in model else 0.01)
    kw = dict(num_leapfrog_steps=L, num_results=S, num_burnin_steps=burn, num_adaptation_steps=adapt, seed=78,
              chain_offset=5, want_orig=True)
    o1 = engine.hmc_run(mc, z0, eps0, a, b, engine=engine.ENGINE_TCGEN05_STREAM, **kw)
    o2 = engine.hmc_run(mc, z0, eps0, a, b, engine=engine.ENGINE_TCGEN05_DUAL, **kw)
    assert (o1["is_accepted"] == o2["is_accepted"]).all()
    assert np.array_equal(o1["samples"], o2["samples"])
    assert np.array_equal(o1["final_z"], o2["final_z"])
    assert np.array_equal(o1["step_mult"], o2["step_mult"])
    assert o1["is_accepted"].mean() > 0.2
