"""GPU parity of the interleaved CP / NCP sampler (--method=i) against the oracle restatement of
interleaved.py:113-155 / inference.py:258-329."""
import numpy as np
import pytest

from autoreparam_b200 import engine
from oracle import oracle as O
from tests import common

pytestmark = pytest.mark.gpu

CASES = [("8schools", 0.1, 0.15), ("german_credit_lognormalcentered", 0.004, 0.006), ("radon", 0.01, 0.01),
         ("radon_stddvs", 0.006, 0.006), ("election", 0.004, 0.006), ("electric", 0.004, 0.006),
         ("time_series", 1e-5, 1e-5), ("german_credit_gammascale", 0.002, 0.002)]


@pytest.mark.parametrize("model,eps_a,eps_b", CASES)
def test_interleaved_fixed_streams_fp64(model, eps_a, eps_b):
    mc = common.model_config(model, "MN")
    raw = common.raw_data(model, "MN")
    D = mc.num_coords
    C, La, Lb, S, burn, adapt = 4, 2, 3, 3, 2, 4
    T = O.num_transitions(S, burn)
    rng = np.random.default_rng(17)
    x0 = common.random_states(model, D, C, seed=21, scale=0.3)
    mom = rng.standard_normal((2 * T, C, D))
    lu = np.log(rng.uniform(size=(2 * T, C)))
    ea = np.full(D, eps_a) * rng.uniform(0.5, 1.5, D)
    eb = np.full(D, eps_b) * rng.uniform(0.5, 1.5, D)
    rule_a, rule_b = (np.ones(D), np.ones(D)), (np.zeros(D), np.zeros(D))
    ref = O.hmc_interleaved_chain(model, raw, x0, ea, eb, La, Lb, S, burn, adapt, (1.0, 1.0), (0.0, 0.0),
                                  momenta=mom, log_u=lu)
    out = engine.hmc_interleaved_run(mc, x0, ea, eb, rule_a, rule_b, num_leapfrog_steps_a=La,
                                     num_leapfrog_steps_b=Lb, num_results=S, num_burnin_steps=burn,
                                     num_adaptation_steps=adapt, ext_momenta=mom, ext_log_u=lu, precision="f64")
    assert (out["is_accepted_a"].astype(bool) == ref["is_accepted_a"]).all()
    assert (out["is_accepted_b"].astype(bool) == ref["is_accepted_b"]).all()
    assert ref["is_accepted_a"].mean() + ref["is_accepted_b"].mean() > 0
    assert common.rel_err(out["samples"].reshape(S * C, D), ref["samples"].reshape(S * C, D)).max() < 1e-8
    np.testing.assert_allclose(out["step_mult_a"], ref["step_mult_a"], rtol=1e-12)
    np.testing.assert_allclose(out["step_mult_b"], ref["step_mult_b"], rtol=1e-12)


def test_interleaved_philox_and_lanes_fp32():
    """Internal Philox streams (index 2 t + r) and every lanes-per-chain variant agree with the oracle."""
    model = "8schools"
    mc = common.model_config(model)
    raw = common.raw_data(model)
    D = mc.num_coords
    C, La, Lb, S, burn, adapt = 6, 3, 3, 4, 3, 5
    x0 = common.random_states(model, D, C, seed=22, scale=0.5).astype(np.float32).astype(np.float64)
    ea, eb = np.full(D, 0.1), np.full(D, 0.15)
    rule_a, rule_b = (np.ones(D), np.ones(D)), (np.zeros(D), np.zeros(D))
    ref = O.hmc_interleaved_chain(model, raw, x0, ea, eb, La, Lb, S, burn, adapt, seed=5, chain_ids=np.arange(C) + 3)
    for lpc in (1, 8, 32):
        out = engine.hmc_interleaved_run(mc, x0, ea, eb, rule_a, rule_b, num_leapfrog_steps_a=La,
                                         num_leapfrog_steps_b=Lb, num_results=S, num_burnin_steps=burn,
                                         num_adaptation_steps=adapt, seed=5, chain_offset=3, lanes_per_chain=lpc)
        assert (out["is_accepted_a"].astype(bool) == ref["is_accepted_a"]).all(), lpc
        assert (out["is_accepted_b"].astype(bool) == ref["is_accepted_b"]).all(), lpc
        assert common.rel_err(out["samples"].reshape(S * C, D), ref["samples"].reshape(S * C, D)).max() < 5e-4, lpc


def test_interleaved_samples_the_posterior():
    """Interleaving CP and NCP steps targets the same centred posterior: radon (sigma_y = 1) against the
    closed form."""
    model = "radon"
    mc = common.model_config(model, "MN")
    raw = common.raw_data(model, "MN")
    D = mc.num_coords
    J = len(raw["u"])
    A = np.zeros((D, D)); rhs = np.zeros(D)
    A[0, 0] += 1; A[1, 1] += 1; A[2, 2] += 1
    for j in range(J):
        v = np.zeros(D); v[3 + j] = 1; v[0] = -1; v[1] = -raw["u"][j]
        A += np.outer(v, v)
    for n in range(len(raw["y"])):
        v = np.zeros(D); v[3 + raw["county"][n]] = 1; v[2] = raw["x"][n]
        A += np.outer(v, v); rhs += v * raw["y"][n]
    cov = np.linalg.inv(A); mean = cov @ rhs; sd = np.sqrt(np.diag(cov))
    C, S = 256, 300
    x0 = np.random.default_rng(1).standard_normal((C, D)) * 0.1
    out = engine.hmc_interleaved_run(mc, x0.astype(np.float32), np.full(D, 0.05), np.full(D, 0.03),
                                     (np.ones(D), np.ones(D)), (np.zeros(D), np.zeros(D)), num_leapfrog_steps_a=6,
                                     num_leapfrog_steps_b=6, num_results=S, num_burnin_steps=500,
                                     num_adaptation_steps=400, seed=9)
    x = out["samples"].astype(np.float64)
    for k in ("is_accepted_a", "is_accepted_b"):
        assert 0.5 < out[k].mean() < 0.95, (k, out[k].mean())
    z = np.abs(x.mean(axis=(0, 1)) - mean) / (sd / np.sqrt(C))
    assert z.max() < 6.0, z.max()
    assert np.abs(x.std(axis=(0, 1)) / sd - 1).max() < 0.08
