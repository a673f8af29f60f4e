"""GPU: BASELINE configs[4] shapes -- synthetic scaled radon (10^6 observations, 10^4 counties) and the
time-series model at tens of thousands of chains.  Small instances are compared with the oracle; the
full-size instances are checked through size-independent properties (finite differences in the fp64 build,
CP <-> NCP identities, sharding invariance)."""
import numpy as np
import pytest

from autoreparam_b200 import data as arp_data
from autoreparam_b200 import engine, models
from oracle import oracle as O
from tests import common

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("method", ["CP", "NCP", "VIP_a"])
def test_synthetic_radon_small_matches_oracle(method):
    raw = arp_data.synthetic_radon(n=20_000, j=500, seed=5)
    mc = models.from_data("radon", raw)
    D = mc.num_coords
    assert D == 503
    a, b = common.ab_for(method, D)
    z = (0.3 * np.random.default_rng(1).standard_normal((3, D))).astype(np.float32).astype(np.float64)
    lp_ref, g_ref = O.log_joint_and_grad("radon", raw, z, a, b)
    for prec, tol in (("f64", 1e-10), ("f32", 1e-5)):
        lp, g, xc = engine.log_joint_grad(mc, z, a, b, precision=prec)
        assert common.rel_err(lp, lp_ref).max() < tol, (prec, common.rel_err(lp, lp_ref).max())
        assert common.rel_err(g, g_ref).max() < tol, (prec, common.rel_err(g, g_ref).max())


def test_synthetic_radon_full_size_properties():
    raw = arp_data.synthetic_radon(n=1_000_000, j=10_000)
    assert np.all(np.diff(raw["county"]) >= 0) and len(raw["y"]) == 1_000_000
    mc = models.from_data("radon", raw)
    D = mc.num_coords
    assert D == 10_003
    rng = np.random.default_rng(2)
    x = 0.2 * rng.standard_normal((2, D))
    ones, zeros = np.ones(D), np.zeros(D)
    # (1) directional finite difference of the fp64 log joint equals grad . direction
    lp, g, _ = engine.log_joint_grad(mc, x, ones, ones, precision="f64")
    v = rng.standard_normal(D) / np.sqrt(D)
    h = 1e-5
    lp_p, _, _ = engine.log_joint_grad(mc, x + h * v, ones, ones, precision="f64")
    lp_m, _, _ = engine.log_joint_grad(mc, x - h * v, ones, ones, precision="f64")
    fd = (lp_p - lp_m) / (2 * h)
    assert np.abs(fd - g @ v).max() < 1e-6 * np.abs(g @ v).max() + 1e-4
    # (2) every radon scale is 1, so the NCP log joint at to_noncentered(x) equals the CP log joint at x and
    #     its centred output is x again (size-independent identity; to_noncentered restated in numpy here)
    z = x.copy()
    z[:, 3:] = x[:, 3:] - (x[:, [0]] + raw["u"][None, :].astype(np.float64) * x[:, [1]])
    lp_n, g_n, xc_n = engine.log_joint_grad(mc, z, zeros, zeros, precision="f64")
    assert np.abs(lp_n - lp).max() < 1e-9 * np.abs(lp).max()
    assert np.abs(xc_n - x).max() < 1e-12
    # chain rule between the two gradients: d/dm is identical, d/dmua gains the sum over counties
    assert np.abs(g_n[:, 3:] - g[:, 3:]).max() < 1e-8 * np.abs(g).max()
    assert np.abs(g_n[:, 0] - (g[:, 0] + g[:, 3:].sum(1))).max() < 1e-8 * np.abs(g).max()
    # (3) fp32 build against the fp64 build at full size
    lp32, g32, _ = engine.log_joint_grad(mc, x, ones, ones, precision="f32")
    assert common.rel_err(lp32, lp).max() < 1e-5 and common.rel_err(g32, g).max() < 1e-5
    # (4) a short sampler run is finite, accepts, and is invariant to sharding the chains
    z0 = (0.05 * rng.standard_normal((96, D))).astype(np.float32)
    kw = dict(num_leapfrog_steps=2, num_results=3, num_burnin_steps=6, num_adaptation_steps=4, seed=3)
    full = engine.hmc_run(mc, z0, np.full(D, 2e-4), zeros, zeros, **kw)
    part = engine.hmc_run(mc, z0[40:], np.full(D, 2e-4), zeros, zeros, chain_offset=40, **kw)
    assert np.isfinite(full["samples"]).all() and full["accept_count"].sum() > 0
    assert np.array_equal(part["samples"], full["samples"][:, 40:])


def test_time_series_many_chains():
    mc = common.model_config("time_series")
    raw = common.raw_data("time_series")
    D = mc.num_coords
    C = 65_536
    a, b = np.zeros(D), np.zeros(D)
    rng = np.random.default_rng(4)
    z_small = (0.05 * rng.standard_normal((8, D))).astype(np.float32)
    z = np.tile(z_small, (C // 8, 1))
    lp, g, xc = engine.log_joint_grad(mc, z, a, b)
    lp_ref, g_ref = O.log_joint_and_grad("time_series", raw, z_small.astype(np.float64), a, b)
    assert common.rel_err(lp[:8], lp_ref).max() < 1e-5 and common.rel_err(g[:8], g_ref).max() < 1e-5
    assert np.array_equal(lp[:8], lp[-8:]) and np.array_equal(g[:8], g[-8:])      # periodic input, periodic output
    out = engine.hmc_run(mc, z, np.full(D, 2e-4), a, b, num_leapfrog_steps=4, num_results=4, num_burnin_steps=20,
                         num_adaptation_steps=10, seed=6, want_samples=True)
    assert np.isfinite(out["samples"]).all() and 0.3 < out["is_accepted"].mean() <= 1.0
