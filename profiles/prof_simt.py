"""SIMT engine throughput sweep: gradient evaluations / s per model for several chain counts and lanes-per-chain
layouts (device-resident buffers, CUDA events).  usage: python profiles/prof_simt.py [models,comma] [--ncu MODEL C LPC]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from autoreparam_b200 import engine  # noqa: E402
from tests import common  # noqa: E402


def run(model, C, lpc, S=20, burn=20, L=4, reps=3, method="NCP"):
    mc = common.model_config(model, "PA")
    D = mc.num_coords
    a, b = common.ab_for(method, D)
    rng = np.random.default_rng(0)
    z0 = torch.as_tensor((0.1 * rng.standard_normal((C, D))).astype(np.float32), device="cuda")
    eps0 = np.full(D, 0.01 if model != "time_series" else 1e-4)
    kw = dict(num_leapfrog_steps=L, num_results=S, num_burnin_steps=burn, num_adaptation_steps=burn, seed=1,
              engine=engine.ENGINE_SIMT, lanes_per_chain=lpc, want_final=False, want_samples=False)
    out = engine.hmc_run(mc, z0, eps0, a, b, **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = engine.hmc_run(mc, z0, eps0, a, b, **kw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    T = out["num_transitions"]
    return C * L * T / (min(ts) * 1e-3), min(ts), float(out["accept_count"].float().mean().item()) / T


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--ncu":
        print(run(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), reps=1))
        sys.exit(0)
    models = sys.argv[1].split(",") if len(sys.argv) > 1 else ["8schools", "radon", "radon_stddvs", "election", "electric", "time_series"]
    for model in models:
        for C in (100, 4096, 16384, 131072, 1048576):
            if C == 1048576 and model not in ("8schools",):
                continue
            for lpc in (1, 8, 32):
                if C * lpc > 1048576 * 8:
                    continue
                try:
                    r, ms, acc = run(model, C, lpc, S=10 if C > 100000 else 20, burn=10 if C > 100000 else 20)
                    print("%-14s C %8d lpc %2d: %.3e grad-evals/s  (%.2f ms, accept %.2f)" % (model, C, lpc, r, ms, acc), flush=True)
                except Exception as e:
                    print(model, C, lpc, "failed:", str(e)[:100], flush=True)
