"""Diagnostic (GPU): gradient accuracy of the SIMT fp32 and tcgen05 engines at TYPICAL-SET states (the states an
HMC run actually visits), against the fp64 check build (== oracle to 1e-10), elementwise and in max-norm.
usage: python profiles/diag/diag_tc_typical.py [features=25|62]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from autoreparam_b200 import data, engine, models  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 25
raw = data.synthetic_german_credit(n=1000, f=F)
mc = models.from_data("german_credit_lognormalcentered", raw)
D = mc.num_coords
C = 2048
rng = np.random.default_rng(0)
for method, (a, b) in (("NCP", (np.zeros(D), np.zeros(D))), ("CP", (np.ones(D), np.ones(D))),
                       ("VIP", (np.full(D, 0.5), np.ones(D)))):
    z0 = (0.1269 * rng.standard_normal((C, D))).astype(np.float32)
    out = engine.hmc_run(mc, z0, np.full(D, 0.1269), a, b, num_leapfrog_steps=4, num_results=2, num_burnin_steps=600,
                         num_adaptation_steps=400, seed=3, engine=engine.ENGINE_SIMT, want_samples=False)
    z = out["final_z"].astype(np.float32)
    lp64, g64, xc64 = engine.log_joint_grad(mc, z.astype(np.float64), a, b, precision="f64")
    eta = xc64[:, 1 + F:] @ raw["X"].T.astype(np.float64)
    print("%s F=%d: %d typical-set states, |eta| median %.2f max %.1f, |g| median %.3g chain-max median %.3g, lp median %.1f" % (
        method, F, C, np.median(np.abs(eta)), np.abs(eta).max(), np.median(np.abs(g64)), np.median(np.abs(g64).max(1)), np.median(lp64)))
    for nm, eng in (("simt", engine.ENGINE_SIMT), ("tcgen05", engine.ENGINE_TCGEN05)):
        lp, g, xc = engine.log_joint_grad(mc, z, a, b, engine=eng)
        err = np.abs(g - g64)
        gmax = np.abs(g64).max(axis=1, keepdims=True)
        el = err / np.maximum(np.abs(g64), 1e-300)
        print("  %-8s lp rel: max %.2e | grad max-norm rel: median %.2e max %.2e | elementwise rel: median %.2e  99%% %.2e  max %.2e | "
              "abs err: median %.2e max %.2e | signed mean err / |g| mean %.2e" % (
                  nm, (np.abs(lp - lp64) / np.abs(lp64)).max(), np.median((err / gmax).max(1)), (err / gmax).max(),
                  np.median(el), np.quantile(el, 0.99), el.max(), np.median(err), err.max(),
                  (g - g64).mean() / np.abs(g64).mean()))
        # per coordinate block: overall scale (1), log-scales (F), coefficients (F)
        for blk, sl in (("s0", slice(0, 1)), ("log-scales", slice(1, 1 + F)), ("beta", slice(1 + F, D))):
            print("     %-10s elementwise rel median %.2e 99%% %.2e; abs err median %.2e" % (
                blk, np.median(el[:, sl]), np.quantile(el[:, sl], 0.99), np.median(err[:, sl])))
