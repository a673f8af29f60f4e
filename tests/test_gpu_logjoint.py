"""GPU parity: log-joint, gradient, centred values and d/da against the fp64 oracle.

Tolerances (BASELINE.json north_star): 1e-5 relative for the fp32 library,
1e-10 for the -DARP_FP64 check build.  "Relative" is the per-chain max-norm
ratio max_d|x - ref| / max(max_d|ref|, 1)  (see tests/common.rel_err).
"""
import numpy as np
import pytest

from autoreparam_b200 import engine
from oracle import oracle as O
from tests import common

pytestmark = pytest.mark.gpu

METHODS = ["CP", "NCP", "VIP_a", "VIP_ab"]
TOL = {"f32": 1e-5, "f64": 1e-10}
# (Round 1 waived time_series to 2e-4: be * year (~1e3) against an observation scale of 0.12 amplifies fp32 rounding of
# the residual by ~1e4, SURVEY.md appendix B note ii.  Its scans now run in double inside the fp32 build and it meets
# the same 1e-5 as every other model.)
TOL_F32_OVERRIDE = {}


@pytest.mark.parametrize("precision", ["f32", "f64"])
@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("model", common.MODELS)
def test_logjoint_grad_centered(model, method, precision):
    mc = common.model_config(model)
    raw = common.raw_data(model)
    D = mc.num_coords
    assert D == O.num_coords(model, raw)
    a, b = common.ab_for(method, D)
    C = 6
    z = common.random_states(model, D, C, seed=3)
    z = z.astype(np.float32).astype(np.float64)  # identical inputs for both precisions
    lp_ref, g_ref = O.log_joint_and_grad(model, raw, z, a, b)
    xc_ref = O.to_centered(model, raw, z, a, b)
    lp, g, xc, ab = engine.log_joint_grad(mc, z, a, b, precision=precision, want_abar=True)
    tol = TOL[precision]
    if precision == "f32":
        tol = TOL_F32_OVERRIDE.get(model, tol)
    assert common.rel_err(lp, lp_ref).max() < tol, ("lp", common.rel_err(lp, lp_ref).max())
    assert common.rel_err(g, g_ref).max() < tol, ("grad", common.rel_err(g, g_ref).max())
    assert common.rel_err(xc, xc_ref).max() < tol, ("centered", common.rel_err(xc, xc_ref).max())
    # d log_joint / d a  (cVIP ELBO gradient ingredient)
    for c in range(2):
        ab_ref = O.grad_wrt_a(model, raw, z[c], a, b)
        assert common.rel_err(ab[c:c + 1], ab_ref[None]).max() < tol, ("abar", c)


@pytest.mark.parametrize("model", ["german_synth"])
@pytest.mark.parametrize("method", ["CP", "NCP", "VIP_a"])
def test_logjoint_synthetic_german(model, method):
    """BASELINE configs[1] shape: 1000 x 25 synthetic design matrix."""
    mc = common.model_config(model)
    raw = common.raw_data(model)
    D = mc.num_coords
    assert D == 51
    a, b = common.ab_for(method, D)
    z = common.random_states(model, D, 5, seed=5).astype(np.float32).astype(np.float64)
    lp_ref, g_ref = O.log_joint_and_grad("german_credit_lognormalcentered", raw, z, a, b)
    lp, g, xc = engine.log_joint_grad(mc, z, a, b, precision="f32")
    assert common.rel_err(lp, lp_ref).max() < 1e-5
    assert common.rel_err(g, g_ref).max() < 1e-5


@pytest.mark.parametrize("model", ["8schools", "radon", "election"])
def test_many_chains_and_lane_layouts(model):
    """Chain counts that exercise every lanes-per-chain kernel variant (32, 8, 1)
    and ragged tails; results must not depend on the variant."""
    mc = common.model_config(model)
    raw = common.raw_data(model)
    D = mc.num_coords
    a, b = common.ab_for("VIP_a", D)
    z_small = common.random_states(model, D, 4, seed=9).astype(np.float32)
    lp_ref, g_ref = O.log_joint_and_grad(model, raw, z_small.astype(np.float64), a, b)
    for C in (4, 2500, 9500 + 3, 76000):
        z = np.tile(z_small, (C // 4 + 1, 1))[:C]
        lp, g, xc = engine.log_joint_grad(mc, z, a, b, precision="f32")
        assert common.rel_err(lp[:4], lp_ref).max() < 1e-5, C
        assert common.rel_err(g[:4], g_ref).max() < 1e-5, C
        assert common.rel_err(g[-4:], g[(C - 4) % 4: (C - 4) % 4 + 1].repeat(4, 0)).max() >= 0  # finite
        assert np.isfinite(g).all() and np.isfinite(lp).all()
        # periodic input -> periodic output
        k = (C // 4) * 4
        assert np.abs(g[:k].reshape(-1, 4, D) - g[:4][None]).max() < 1e-4 * max(1.0, np.abs(g[:4]).max())
