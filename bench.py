#!/usr/bin/env python
"""Benchmark of the hot path: batched-chain HMC on German credit (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU reference arm (oracle port)

One "step" = one full pass of the hot path over one batch of chains: a complete
HMC run (1 + burn-in + 2(S-1) transitions of L leapfrog steps, Metropolis accept,
dual-averaging adaptation, thinning, centred-sample store) for C chains per GPU,
in one persistent kernel launch.  metric = leapfrog gradient evaluations / second
(true count: C * L * transitions).

 value  inputs resident in HBM, timed with CUDA events on the launching stream
 e2e    the same run through the public API (autoreparam_b200.inference.hmc) with
        HOST buffers: pinned H2D of the initial states, HMC, ESS kernel, D2H of
        ESS / is_accepted / step sizes -- what main.py's run_hmc consumes
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_GRAD = {25: 1.06e5, 62: 2.55e5}  # 4NF + 12F + 6N
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel at the default workload
# (profiles/r01_bench_kernel_traffic.csv: 0.05 GB read + 3.344 GB written); the algorithmic bytes are the
# thinned sample store 1000 x 16384 x 51 x 4 B + is_accepted = 3.359 GB, i.e. no re-reads
NCU_TRAFFIC_DEFAULT_WORKLOAD = 48132096 + 3344686592


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--chains", type=int, default=16384, help="chains per GPU (weak scaling) / in total (strong scaling)")
    p.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                   help="weak: --chains per GPU; strong: --chains in total, sharded over the GPUs")
    p.add_argument("--method", default="NCP", choices=["CP", "NCP", "cVIP"])
    p.add_argument("--features", type=int, default=25, help="25 = BASELINE synthetic shape")
    p.add_argument("--num_leapfrog_steps", type=int, default=4)
    p.add_argument("--num_samples", type=int, default=1000)
    p.add_argument("--num_burnin_steps", type=int, default=500)
    p.add_argument("--num_adaptation_steps", type=int, default=400)
    p.add_argument("--engine", type=int, default=0, help="0 auto, 1 SIMT fp32, 2 tcgen05")
    p.add_argument("--no_cpu_baseline", action="store_true")
    return p.parse_args()


# ------------------------------------------------------------------ helpers ---
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def method_ab(method, D):
    if method == "CP":
        return np.ones(D), np.ones(D)
    if method == "NCP":
        return np.zeros(D), np.zeros(D)
    # cVIP as the reference runs it: learned a (here a fixed mid-way value), b = 1 (tied as written)
    return np.full(D, 0.5), np.ones(D)


def workload(args):
    from autoreparam_b200 import data
    raw = data.synthetic_german_credit(n=1000, f=args.features)
    D = 1 + 2 * args.features
    a, b = method_ab(args.method, D)
    return raw, D, a, b


def init_states(D, C, rank):
    """Initial states / step sizes of the shape VI hands to HMC (loc ~ 0, sigma_q ~ softplus(-2))."""
    rng = np.random.default_rng(20190603 + 7919 * rank)
    sigma_q = np.full(D, 0.1269)
    z0 = (sigma_q * rng.standard_normal((C, D))).astype(np.float32)
    return z0, sigma_q


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ---------------------------------------------------------- reference arm ---
def cpu_reference_run(args, raw, D, a, b, steps, warmup, chains=1024, transitions=50):
    """The reference's CPU implementation of the path, restated (oracle port:
    PyTorch CPU fp32, [C, D] tensors, autograd gradient at every leapfrog step,
    TFP op order), all host threads, on a bounded sample of the workload."""
    import torch
    from oracle import oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    z0, sigma_q = init_states(D, chains, 0)
    eps0 = sigma_q / (args.num_leapfrog_steps / 4.0) ** 2
    times, evals = [], 0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        n, _ = O.german_hmc_cpu(raw["X"], raw["y"], z0, eps0, args.num_leapfrog_steps, transitions, a, b,
                                seed=i, num_adapt=transitions)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt); evals += n
    total = sum(times)
    return {"value": evals / total, "ms_per_step": 1e3 * total / max(steps, 1), "cores": torch.get_num_threads(),
            "sample": "%d chains x %d transitions x L=%d per step (same model/data/method), %d steps" %
                      (chains, transitions, args.num_leapfrog_steps, steps)}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    raw, D, a, b = workload(args)
    metric = "leapfrog_grad_evals_per_sec"
    config = {"workload": "german_credit_lognormalcentered HMC %s, synthetic 1000x%d (D=%d), %d chains/GPU, L=%d, "
                          "S=%d kept (thin 2), burn-in %d, adapt %d" %
                          (args.method, args.features, D, args.chains, args.num_leapfrog_steps, args.num_samples,
                           args.num_burnin_steps, args.num_adaptation_steps),
              "chains_per_gpu": args.chains, "parallelism": "chains sharded, dp%d" % world,
              "l2": "per-step output (samples) exceeds L2 and a 512 MiB buffer is rewritten between timed steps"}

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(args, raw, D, a, b, args.steps, args.warmup)
        line = {"impl": "reference", "metric": metric, "value": r["value"], "unit": "grad_evals/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": "grad_evals/s", "cores": r["cores"], "kind": "port",
                                 "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "grad_evals/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from autoreparam_b200 import engine, graphs, inference, models, util

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    mc = models.from_data("german_credit_lognormalcentered", raw)
    L, S = args.num_leapfrog_steps, args.num_samples
    if args.scaling == "strong":
        from autoreparam_b200 import distributed
        lo, hi = distributed.shard_range(args.chains, rank, world)
        C, chain_lo = hi - lo, lo
        assert C > 0, "more ranks than chains"
    else:
        C, chain_lo = args.chains, rank * args.chains
    config["scaling"] = args.scaling
    config["chains_total"] = args.chains if args.scaling == "strong" else args.chains * world
    z0, sigma_q = init_states(D, C, rank)
    eps0 = sigma_q / (L / 4.0) ** 2
    T = engine.hmc_num_transitions(S, args.num_burnin_steps)
    evals_per_step = C * L * T            # this rank
    evals_all = config["chains_total"] * L * T   # all ranks
    z_dev = torch.as_tensor(z0, device=dev)
    bufs = {"samples": torch.empty((S, C, D), dtype=torch.float32, device=dev),
            "is_accepted": torch.empty((S, C), dtype=torch.uint8, device=dev)}
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    kw = dict(num_leapfrog_steps=L, num_results=S, num_burnin_steps=args.num_burnin_steps,
              num_adaptation_steps=args.num_adaptation_steps, chain_offset=chain_lo, want_final=False,
              engine=args.engine)

    def step(i):
        return engine.hmc_run(mc, z_dev, eps0, a, b, seed=1000 + i, out=bufs, **kw)

    for i in range(args.warmup):
        step(i)
    # ---- timed region: K steps, CUDA events on the launching stream, max over ranks
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = engine.kernel_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)  # L2 flush, outside the event pair
        ev[i][0].record()
        out = step(args.warmup + i)
        ev[i][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = engine.kernel_launch_count() - launches0
    clocks = sampler.stop()
    ms = sum(s.elapsed_time(e) for s, e in ev)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    value = evals_all * args.steps / (ms * 1e-3)

    # ---- end to end through the public API with host buffers
    target = graphs.TargetGraph(mc, args.method, a, b, False)
    step_sizes = mc.split(sigma_q)
    h2d = z0.nbytes + eps0.astype(np.float32).nbytes + 2 * D * 4
    # what main.py's run_hmc consumes comes back: ESS [C, D], step_mult + accept_count [C], accept counters, R-hat [D] f64.
    # The [S, C, D] centred samples (3.3 GB) and the [S, C] accept flags stay in HBM: the reference's sess.run returns
    # them (main.py:350-360), but the driver only ever reduces them to ESS / an accept count (and saves
    # --num_chains_to_save traces, 0 by default).
    d2h = C * D * 4 + C * 8 + 3 * 8 + D * 8
    inference.hmc(target, mc, step_sizes, z0, num_leapfrog_steps=L, num_samples=S,
                  num_burnin_steps=args.num_burnin_steps, num_adaptation_steps=args.num_adaptation_steps,
                  seed=1, chain_offset=chain_lo, device=dev, engine_kind=args.engine, return_is_accepted=False)
    barrier()
    e0 = time.perf_counter()
    for i in range(args.steps):
        res = inference.hmc(target, mc, step_sizes, z0, num_leapfrog_steps=L, num_samples=S,
                            num_burnin_steps=args.num_burnin_steps, num_adaptation_steps=args.num_adaptation_steps,
                            seed=2000 + i, chain_offset=chain_lo, device=dev, engine_kind=args.engine,
                            return_is_accepted=False)
    barrier()
    e2e_s = time.perf_counter() - e0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e.item())
    e2e_value = evals_all * args.steps / e2e_s
    # ESS / R-hat: per-chain min ESS gathered over ranks (NCCL), as util.get_min_ess consumes it
    min_ess = torch.as_tensor(np.nan_to_num(res.ess_flat).min(axis=1), device=dev)
    if world > 1:
        gathered = [torch.empty_like(min_ess) for _ in range(world)]
        dist.all_gather(gathered, min_ess)
        min_ess = torch.cat(gathered)
    ess_total = float(min_ess.sum().item())
    ess_per_sec = ess_total / (e2e_s / args.steps)
    ess_per_1000 = float((1000.0 * min_ess / (S * L)).mean().item())  # main.py:362-366 normalisation

    if rank == 0:
        pk, pk_src = peaks()
        flop = FLOP_PER_GRAD.get(args.features, 4.0 * 1000 * args.features)
        achieved_tf = evals_per_step * args.steps * flop / (ms * 1e-3) / 1e12   # this rank's kernel
        peak_tf = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        # the pipe that actually binds the dominant kernel: MUFU ops per observation on the XU pipe,
        # 16 results / clk / SM (profiles/micro/pipes.cu), on the SMs the 128-chain tiles occupy
        n_pad = (1000 + 127) // 128 * 128
        mufu_per_obs = (1.0 + 0.25 + 0.25 / L) if args.features <= 32 else (2.0 + 1.0 / L)
        sms_used = min(148, (C + 127) // 128)   # rank 0's tiles
        f_clk = 1e6 * (clocks.get("sm_mhz") or pk.get("sm_max_mhz", 1965.0))
        xu_roof = sms_used * f_clk / (n_pad * mufu_per_obs / 16.0)
        line = {
            "metric": metric, "value": value, "unit": "grad_evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "grad_evals/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h),
                    "returns": "ESS [C,D], step sizes / accept counts [C], R-hat [D] (reduced over all ranks); the [S,C,D] "
                               "samples and [S,C] accept flags stay on the device (main.py reduces them to these)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf,
                         "traffic": NCU_TRAFFIC_DEFAULT_WORKLOAD if (C, S, args.features, args.num_burnin_steps) ==
                         (16384, 1000, 25, 500) else None,
                         "note": "algorithmic fp32 flop (%.3g per grad eval) / measured dense bf16 cuBLAS peak (%s, "
                                 "sustained); per GPU" % (flop, pk_src),
                         "binding_pipe": {"pipe": "xu (MUFU), co-limited by instruction dispatch",
                                          "achieved": evals_per_step * args.steps / (ms * 1e-3), "peak": xu_roof,
                                          "unit": "grad_evals/s per GPU",
                                          "frac": evals_per_step * args.steps / (ms * 1e-3) / xu_roof,
                                          "note": "%.4g MUFU ops per observation x %d padded observations, 16 MUFU "
                                                  "results/clk/SM (measured), %d SMs occupied by the 128-chain tiles, "
                                                  "SM clock sampled under load" % (mufu_per_obs, n_pad, sms_used)}},
            # acceptance rate and R-hat are over the chains of ALL ranks: inference.hmc all-reduces the per-chain
            # moments / accept counters (NCCL) inside the e2e timed region
            "ess": {"ess_per_sec": ess_per_sec, "ess_per_1000_grads_mean": ess_per_1000,
                    "acceptance_rate": res.accept_stats[1] / (res.accept_stats[2] * T),
                    "chains_reduced": int(res.accept_stats[2]),
                    "rhat_max": None if res.rhat is None else float(np.nanmax(res.rhat))},
            "wall_s_timed_region": wall,
        }
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(args, raw, D, a, b, steps=3, warmup=1)   # ~10-15 s of CPU work
            line["cpu_baseline"] = {"value": r["value"], "unit": "grad_evals/s", "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
