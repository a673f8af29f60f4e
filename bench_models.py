#!/usr/bin/env python
"""Per-model / per-parameterisation throughput table (north_star: grad-evals/s, ESS/s and ELBO-iterations/s per model
and parameterisation at 1 / 2 / 4 / 8 GPUs, as absolute numbers and as a fraction of roofline, next to a same-run CPU
baseline).

    python bench_models.py [--chains 16384] [--out profiles/r02_models_n1.json]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        bench_models.py --out profiles/r02_models_n8.json

Not the driver's benchmark (that is bench.py, which takes the same models through --model); this fills the tables in
profiles/.  For every in-scope model and method in CP / NCP / cVIP / dVIP:
  VI   find_best_learning_rate exactly as BASELINE configs[3] states it: 5 learning rates x 3000 Adam steps x S = 256,
       one persistent launch (every rank runs the same replica; timed on the device);
  HMC  `--chains` chains PER GPU (weak scaling, chains sharded over the ranks with their global ids), L = 4, with the
       step sizes and initial states of the VI fit; device-timed with CUDA events around the persistent launch,
       barrier + max over ranks; ESS / R-hat through inference.hmc (NCCL all-reduce of the moments).
CPU baseline (rank 0, only when run on one GPU) = the oracle's batched CPU port (fp32 torch, model bodies with dense
one-hot matmuls vectorised over chains with vmap, autograd at every leapfrog step / Adam step), all host threads,
on a bounded sample.
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
    ap.add_argument("--chains", type=int, default=16384, help="chains per GPU")
    ap.add_argument("--num_samples", type=int, default=500)
    ap.add_argument("--num_burnin_steps", type=int, default=500)
    ap.add_argument("--vi_steps", type=int, default=3000)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_models.json"))
    ap.add_argument("--models", default="")
    ap.add_argument("--methods", default="CP,NCP,cVIP,dVIP")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import bench as B
    from autoreparam_b200 import engine, graphs, inference
    from tests import common

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    pk, pk_src = B.peaks()
    f_clk = 1e6 * pk.get("sm_max_mhz", 1965.0)
    peak_fp32 = 148 * 128 * 2 * f_clk / 1e12
    peak_tensor = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    names = [m for m in (args.models.split(",") if args.models else common.MODELS + ["german_synth"])]
    L, S, C = 4, args.num_samples, args.chains
    lrs = [0.02, 0.05, 0.1, 0.2, 0.4]
    rows = []
    for name in names:
        mc = common.model_config(name, "PA")
        raw = common.raw_data(name, "PA")
        oname = "german_credit_lognormalcentered" if name == "german_synth" else name
        D = mc.num_coords
        fake = argparse.Namespace(model=name)
        flop = B.flop_per_grad(fake, raw, D)
        bound = B.MODEL_TABLE[name]["bound"]
        peak = peak_tensor if bound == "tensor" else peak_fp32
        learned = None
        for method in args.methods.split(","):
            if method == "CP":
                target = graphs.make_cp_graph(mc)
            elif method == "NCP":
                target = graphs.make_ncp_graph(mc)
            elif method == "cVIP":
                target = graphs.make_cvip_graph(mc, "eig", tied_pparams=True)
            else:
                if learned is None:
                    continue
                target = graphs.make_dvip_graph(mc, graphs.discretise(learned))
            # ---- VI (every rank: same seed, same result), device-timed
            vi_kw = dict(learning_rates=lrs, num_optimization_steps=args.vi_steps, num_mc_samples=256, seed=1)
            inference.find_best_learning_rate(target, mc, **dict(vi_kw, num_optimization_steps=50))   # warm-up
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            (elbo, timeline, lr, step0, params, reparam) = inference.find_best_learning_rate(target, mc, **vi_kw)
            e1.record()
            torch.cuda.synchronize()
            vi_s = e0.elapsed_time(e1) * 1e-3
            if method == "cVIP":
                learned = reparam
                hmc_target = graphs.make_dvip_graph(mc, reparam)   # HMC with the learned continuous a
            else:
                hmc_target = target
            # ---- HMC: device-resident inputs, CUDA events around the persistent launch, max over ranks
            rng = np.random.default_rng(2 + rank)
            z0 = mc.join([params[n + "_loc"] + params[n + "_scale"] * rng.standard_normal((C,) + tuple(s))
                          for n, s in mc.sites]).astype(np.float32)
            sigma_q = inference._flat_step_sizes(mc, step0)     # inference.py:212-216: sigma_q / (L / 4)^2
            eps0 = sigma_q / (L / 4.0) ** 2
            z_dev = torch.as_tensor(z0, device=dev)
            kw = dict(num_leapfrog_steps=L, num_results=S, num_burnin_steps=args.num_burnin_steps,
                      num_adaptation_steps=int(0.6 * args.num_burnin_steps), chain_offset=rank * C, want_final=False)
            bufs = {"samples": torch.empty((S, C, D), dtype=torch.float32, device=dev),
                    "is_accepted": torch.empty((S, C), dtype=torch.uint8, device=dev)}
            engine.hmc_run(mc, z_dev, eps0, hmc_target.a, hmc_target.b, seed=3, out=bufs, **kw)        # warm-up
            barrier()
            ms = 0.0
            reps = 3
            for i in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = engine.hmc_run(mc, z_dev, eps0, hmc_target.a, hmc_target.b, seed=4 + i, out=bufs, **kw)
                e1.record()
                torch.cuda.synchronize()
                ms += e0.elapsed_time(e1)
            del bufs
            hmc_s = max_over_ranks(ms * 1e-3 / reps)
            T = out["num_transitions"]
            evals_all = world * C * L * T
            # ---- ESS / R-hat / acceptance through the public call (collectives inside)
            ekw = dict(num_leapfrog_steps=L, num_samples=S, num_burnin_steps=args.num_burnin_steps,
                       num_adaptation_steps=int(0.6 * args.num_burnin_steps), chain_offset=rank * C, device=dev,
                       return_is_accepted=False)
            inference.hmc(hmc_target, mc, step0, z0, seed=8, **ekw)    # warm-up: pinned staging buffers, pooled scratch
            barrier()
            t0 = time.perf_counter()
            res = inference.hmc(hmc_target, mc, step0, z0, seed=9, **ekw)
            barrier()
            e2e_s = max_over_ranks(time.perf_counter() - t0)
            min_ess = torch.as_tensor(np.nan_to_num(res.ess_flat).min(axis=1), device=dev)
            ess_sum = torch.tensor([float(min_ess.sum().item())], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ess_sum)
            per_gpu_rate = C * L * T / hmc_s
            row = dict(model=name, method=method, D=D, n_gpus=world, chains_per_gpu=C, elbo=float(elbo), best_lr=lr,
                       elbo_iters_per_s=len(lrs) * args.vi_steps / vi_s, vi_seconds=vi_s,
                       grad_evals_per_s=evals_all / hmc_s, hmc_seconds=hmc_s,
                       grad_evals_per_s_e2e=evals_all / e2e_s,
                       roofline_bound=bound, roofline_achieved_tflops=per_gpu_rate * flop / 1e12,
                       roofline_peak_tflops=peak, roofline_frac=per_gpu_rate * flop / 1e12 / peak,
                       flop_per_grad_eval=flop,
                       ess_per_s=float(ess_sum.item()) / e2e_s,
                       ess_per_1000_grads=float(ess_sum.item()) / (world * C) * 1000.0 / (S * L),
                       acceptance=res.accept_stats[1] / (res.accept_stats[2] * T),
                       rhat_max=float(np.nanmax(res.rhat)))
            if rank == 0 and world == 1 and not args.no_cpu_baseline:
                from oracle import oracle as O
                torch.set_num_threads(os.cpu_count() or 1)
                zc = z0[:256]
                t0 = time.perf_counter()
                n_cpu, _ = O.hmc_cpu_batched(oname, raw, zc, eps0, L, 6, hmc_target.a, hmc_target.b, num_adapt=6)
                row["cpu_grad_evals_per_s"] = n_cpu / (time.perf_counter() - t0)
                vs = 6 if name in ("election",) or name.startswith("german") else 15
                t0 = time.perf_counter()
                O.vi_cpu_batched(oname, raw, 256, vs, 0.05, hmc_target.a, hmc_target.b)
                row["cpu_elbo_iters_per_s"] = vs / (time.perf_counter() - t0)
                row["cpu_cores"] = torch.get_num_threads()
            if rank == 0:
                rows.append(row)
                print(json.dumps(row), flush=True)
        mc.close()
    if rank == 0:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(dict(note="bench_models.py: HMC device-timed (CUDA events, max over ranks), %d chains per GPU, L = 4, "
                                "%d kept samples (thin 2), burn-in %d; VI = 5 learning rates x %d steps x S = 256 in one "
                                "launch; elbo_iters_per_s counts all 5 learning rates; roofline = naive algorithmic flop "
                                "(SURVEY.md 8d) per GPU / FP32 FMA peak (tensor: measured dense bf16 peak)" %
                                (C, S, args.num_burnin_steps, args.vi_steps), peaks_source=pk_src, rows=rows), f, indent=1)
        hdr = ("| model | method | D | GPUs | ELBO | ELBO it/s | grad-evals/s | e2e grad-evals/s | roofline (bound) | ESS/s | "
               "ESS/1000 grads | accept | R-hat max | CPU grad-evals/s | CPU ELBO it/s |")
        print(hdr)
        print("|" + "---|" * 15)
        for r in rows:
            print("| %s | %s | %d | %d | %.2f | %.3g | %.3g | %.3g | %.2f %% (%s) | %.3g | %.3g | %.2f | %.2f | %s | %s |" % (
                r["model"], r["method"], r["D"], r["n_gpus"], r["elbo"], r["elbo_iters_per_s"], r["grad_evals_per_s"],
                r["grad_evals_per_s_e2e"], 100 * r["roofline_frac"], r["roofline_bound"], r["ess_per_s"],
                r["ess_per_1000_grads"], r["acceptance"], r["rhat_max"],
                "%.3g" % r["cpu_grad_evals_per_s"] if "cpu_grad_evals_per_s" in r else "-",
                "%.3g" % r["cpu_elbo_iters_per_s"] if "cpu_elbo_iters_per_s" in r else "-"))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
