"""Diagnostic: accept decisions of the streaming tcgen05 engine vs the SIMT fp32 and fp64 engines on the
configuration of tests/test_gpu_tc.py::test_tcs_matches_simt_engine_many_chains."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from autoreparam_b200 import engine
from tests import common

C, L, S, burn, adapt = 128 * 2 + 37, 4, 2, 2, 3
model = "german_synth"
mc = common.model_config(model)
D = mc.num_coords
a, b = common.ab_for("NCP", D)
z0 = common.random_states(model, D, C, seed=43, scale=0.3).astype(np.float32).astype(np.float64)
eps0 = np.full(D, 0.01)
for adapt_, tag in ((adapt, "adapt3"), (0, "noadapt")):
    kw = dict(num_leapfrog_steps=L, num_results=S, num_burnin_steps=burn, num_adaptation_steps=adapt_, seed=77, chain_offset=11)
    o64 = engine.hmc_run(mc, z0, eps0, a, b, engine=engine.ENGINE_SIMT, precision="f64", **kw)
    o32 = engine.hmc_run(mc, z0, eps0, a, b, engine=engine.ENGINE_SIMT, **kw)
    otc = engine.hmc_run(mc, z0, eps0, a, b, engine=engine.ENGINE_TCGEN05_STREAM, **kw)
    f = lambda x, y: float((x["is_accepted"] == y["is_accepted"]).mean())
    print(tag, "simt32==f64 %.4f  tc==f64 %.4f  tc==simt32 %.4f" % (f(o32, o64), f(otc, o64), f(otc, o32)),
          "acc f64 %.3f simt %.3f tc %.3f" % (o64["is_accepted"].mean(), o32["is_accepted"].mean(), otc["is_accepted"].mean()),
          "count diff tc-f64", int(np.abs(otc["accept_count"] - o64["accept_count"]).sum()), "simt-f64", int(np.abs(o32["accept_count"] - o64["accept_count"]).sum()))

print("single transition, no adaptation")
for eps in (0.02, 0.05, 0.1, 0.2):
    kw = dict(num_leapfrog_steps=L, num_results=1, num_burnin_steps=0, num_adaptation_steps=0, seed=77, chain_offset=11)
    e0 = np.full(D, eps)
    o64 = engine.hmc_run(mc, z0, e0, a, b, engine=engine.ENGINE_SIMT, precision="f64", **kw)
    otc = engine.hmc_run(mc, z0, e0, a, b, engine=engine.ENGINE_TCGEN05_STREAM, **kw)
    o32 = engine.hmc_run(mc, z0, e0, a, b, engine=engine.ENGINE_SIMT, **kw)
    same = (otc["is_accepted"] == o64["is_accepted"])[0]
    both = (otc["is_accepted"][0] == 1) & (o64["is_accepted"][0] == 1)
    err = common.rel_err(otc["samples"][0][both], o64["samples"][0][both])
    err32 = common.rel_err(o32["samples"][0][both], o64["samples"][0][both])
    print("eps %.2f acc64 %.3f same %.4f simt32-same %.4f  max rel err of accepted samples tc %.2e simt32 %.2e; differing chains %s" %
          (eps, o64["is_accepted"].mean(), same.mean(), (o32["is_accepted"] == o64["is_accepted"]).mean(),
           err.max() if both.any() else -1, err32.max() if both.any() else -1, np.nonzero(~same)[0][:12]))
