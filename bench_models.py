#!/usr/bin/env python
"""Per-model / per-parameterisation throughput table (north_star: grad-evals/s, ESS/s and
ELBO-iterations/s per model and parameterisation next to a same-run CPU baseline).

    python bench_models.py [--chains 4096] [--out profiles/r01_models.json]

Not the driver's benchmark (that is bench.py); this fills the table in profiles/.
For every in-scope model: VI for CP / NCP / cVIP (all learning rates concurrently, S = 256),
then HMC with the VI step sizes for CP / NCP / cVIP(learned a) / dVIP, L = 4.
CPU baseline = the oracle (fp64 torch autograd, one chain at a time, as tests use it) timed on a
bounded sample of gradient evaluations on this host.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=4096)
    ap.add_argument("--num_samples", type=int, default=500)
    ap.add_argument("--num_burnin_steps", type=int, default=500)
    ap.add_argument("--vi_steps", type=int, default=1000)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_models.json"))
    ap.add_argument("--models", default="")
    args = ap.parse_args()

    import torch
    from autoreparam_b200 import engine, graphs, inference, models
    from oracle import oracle as O
    from tests import common

    assert torch.cuda.is_available()
    names = [m for m in (args.models.split(",") if args.models else common.MODELS + ["german_synth"])]
    L, S, C = 4, args.num_samples, args.chains
    lrs = [0.02, 0.05, 0.1, 0.2, 0.4]
    rows = []
    for name in names:
        mc = common.model_config(name, "PA")
        raw = common.raw_data(name, "PA")
        oname = "german_credit_lognormalcentered" if name == "german_synth" else name
        D = mc.num_coords
        # CPU baseline: oracle gradient evaluations per second (bounded sample)
        zc = common.random_states(name, D, 4, seed=1)
        t0 = time.perf_counter(); n_cpu = 0
        while time.perf_counter() - t0 < 3.0:
            O.log_joint_and_grad(oname, raw, zc, 0.0, 0.0)
            n_cpu += len(zc)
        cpu_rate = n_cpu / (time.perf_counter() - t0)
        learned = None
        for method in ("CP", "NCP", "cVIP", "dVIP"):
            if method == "CP":
                target = graphs.make_cp_graph(mc)
            elif method == "NCP":
                target = graphs.make_ncp_graph(mc)
            elif method == "cVIP":
                target = graphs.make_cvip_graph(mc, "eig", tied_pparams=True)
            else:
                target = graphs.make_dvip_graph(mc, graphs.discretise(learned))
            torch.cuda.synchronize(); t0 = time.perf_counter()
            (elbo, timeline, lr, step0, params, reparam) = inference.find_best_learning_rate(
                target, mc, learning_rates=lrs, num_optimization_steps=args.vi_steps, num_mc_samples=256, seed=1)
            torch.cuda.synchronize(); vi_s = time.perf_counter() - t0
            if method == "cVIP":
                learned = reparam
                hmc_target = graphs.make_dvip_graph(mc, reparam)   # HMC with the learned continuous a
            else:
                hmc_target = target
            rng = np.random.default_rng(2)
            z0 = mc.join([params[n + "_loc"] + params[n + "_scale"] * rng.standard_normal((C,) + tuple(s))
                          for n, s in mc.sites]).astype(np.float32)
            kw = dict(num_leapfrog_steps=L, num_samples=S, num_burnin_steps=args.num_burnin_steps,
                      num_adaptation_steps=int(0.6 * args.num_burnin_steps), seed=3)
            inference.hmc(hmc_target, mc, step0, z0[:64], **dict(kw, num_samples=8, num_burnin_steps=8,
                                                                 num_adaptation_steps=4))   # warm-up
            torch.cuda.synchronize(); t0 = time.perf_counter()
            res = inference.hmc(hmc_target, mc, step0, z0, **kw)
            torch.cuda.synchronize(); hmc_s = time.perf_counter() - t0
            evals = C * L * res.num_transitions
            min_ess = np.nan_to_num(res.ess_flat).min(axis=1)
            row = dict(model=name, method=method, D=D, chains=C, elbo=float(elbo), best_lr=lr,
                       elbo_iters_per_s=len(lrs) * args.vi_steps / vi_s, vi_seconds=vi_s,
                       grad_evals_per_s=evals / hmc_s, hmc_seconds=hmc_s, ess_per_s=float(min_ess.sum() / hmc_s),
                       ess_per_1000_grads=float((1000 * min_ess / (S * L)).mean()),
                       acceptance=float(res.is_accepted.mean()), rhat_max=float(np.nanmax(res.rhat)),
                       cpu_oracle_grad_evals_per_s=cpu_rate)
            rows.append(row)
            print(json.dumps(row), flush=True)
        mc.close()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(dict(note="bench_models.py: end-to-end (host buffers in, ESS out) on one B200; "
                            "elbo_iters_per_s counts all 5 learning rates", rows=rows), f, indent=1)
    hdr = "| model | method | D | ELBO | ELBO it/s | grad-evals/s | ESS/s | ESS/1000 grads | accept | R-hat max | CPU oracle grad-evals/s |"
    print(hdr); print("|" + "---|" * 11)
    for r in rows:
        print("| %s | %s | %d | %.2f | %.3g | %.3g | %.3g | %.3g | %.2f | %.2f | %.3g |" % (
            r["model"], r["method"], r["D"], r["elbo"], r["elbo_iters_per_s"], r["grad_evals_per_s"], r["ess_per_s"],
            r["ess_per_1000_grads"], r["acceptance"], r["rhat_max"], r["cpu_oracle_grad_evals_per_s"]))


if __name__ == "__main__":
    main()
