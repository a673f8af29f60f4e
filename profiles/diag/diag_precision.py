"""Diagnostic (GPU): where the fp32 library loses accuracy against the fp64 oracle.
 1. time_series: per-output max-norm relative error (lp, grad, centred, abar) per method, and the worst coordinate
 2. German credit: elementwise gradient error of the SIMT and the tcgen05 engine (BASELINE configs[1] shape and the
    real 1000 x 62 data)
usage: python profiles/diag/diag_precision.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from autoreparam_b200 import engine  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests import common  # noqa: E402

model = "time_series"
mc, raw = common.model_config(model), common.raw_data(model)
D = mc.num_coords
# (development libraries built with ARP_DEV_GERMAN_ONLY have no time_series kernel: pass `german` to skip this part)
for method in (() if "german" in sys.argv[1:] else ("CP", "NCP", "VIP_a", "VIP_ab")):
    a, b = common.ab_for(method, D)
    z = common.random_states(model, D, 6, seed=3).astype(np.float32).astype(np.float64)
    lp_ref, g_ref = O.log_joint_and_grad(model, raw, z, a, b)
    xc_ref = O.to_centered(model, raw, z, a, b)
    lp, g, xc, ab = engine.log_joint_grad(mc, z, a, b, precision="f32", want_abar=True)
    ab_ref = np.stack([O.grad_wrt_a(model, raw, z[c], a, b) for c in range(6)])
    i = np.unravel_index(np.abs(g - g_ref).argmax(), g.shape)
    j = np.unravel_index(np.abs(ab - ab_ref).argmax(), ab.shape)
    print("time_series %-6s lp %.2e grad %.2e (worst d=%d ref %.4g got %.4g; chain max %.3g) xc %.2e abar %.2e (worst d=%d ref %.4g got %.4g; chain max %.3g)" % (
        method, common.rel_err(lp, lp_ref).max(), common.rel_err(g, g_ref).max(), i[1], g_ref[i], g[i], np.abs(g_ref[i[0]]).max(),
        common.rel_err(xc, xc_ref).max(), common.rel_err(ab, ab_ref).max(), j[1], ab_ref[j], ab[j], np.abs(ab_ref[j[0]]).max()))

for model in ("german_synth", "german_credit_lognormalcentered", "german_credit_gammascale"):
    mc, raw = common.model_config(model), common.raw_data(model)
    D = mc.num_coords
    name = "german_credit_lognormalcentered" if model == "german_synth" else model
    for method in ("CP", "NCP", "VIP_ab"):
        a, b = common.ab_for(method, D)
        z = common.random_states(model, D, 48, seed=71, scale=0.3).astype(np.float32).astype(np.float64)
        lp_ref, g_ref = O.log_joint_and_grad(name, raw, z, a, b)
        for nm, eng in (("simt", engine.ENGINE_SIMT), ("tcgen05", engine.ENGINE_TCGEN05)):
            lp, g, xc = engine.log_joint_grad(mc, z.astype(np.float32), a, b, engine=eng)
            err = np.abs(g - g_ref)
            gmax = np.abs(g_ref).max(axis=1, keepdims=True)
            rel_el = err / np.maximum(np.abs(g_ref), 1e-300)
            big = np.abs(g_ref) > 1e-3 * gmax
            print("%-32s %-6s %-8s lp rel %.2e | grad: max-norm rel %.2e, elementwise rel (|g| > 1e-3 max) %.2e, "
                  "abs err / chain max %.2e, frac elementwise < 1e-5: %.4f" % (
                      model, method, nm, (np.abs(lp - lp_ref) / np.abs(lp_ref)).max(), (err / gmax).max(),
                      rel_el[big].max(), (err / gmax).max(), (rel_el < 1e-5).mean()))
