#!/usr/bin/env python
"""Benchmark of the hot path: batched-chain HMC on German credit (BASELINE.json configs[1]) by default; the other
BASELINE configs through --model / --inference.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU reference arm (oracle port)
    python bench.py --model 8schools --method CP --chains 1048576                # configs[0] shape, throughput
    python bench.py --model radon --method NCP                                   # configs[2]
    python bench.py --model election --inference VI --method dVIP                # configs[3]: 5 lrs x 3000 steps x S=256
    torchrun ... bench.py --gpus 8 --model radon_synth --chains 8192 --stream_window 64    # configs[4], 65 536 chains
    torchrun ... bench.py --gpus 8 --model time_series --chains 8192                       # configs[4], 65 536 chains

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
    p.add_argument("--model", default="german_synth",
                   help="german_synth (BASELINE configs[1], default) | german_credit_lognormalcentered (real 1000x62) | "
                        "german_credit_gammascale | 8schools | radon | radon_stddvs | election | electric | time_series | "
                        "radon_synth (10^6 observations x 10^4 counties)")
    p.add_argument("--inference", default="HMC", choices=["HMC", "VI"])
    p.add_argument("--method", default="NCP", choices=["CP", "NCP", "cVIP", "dVIP"])
    p.add_argument("--features", type=int, default=25, help="german_synth: 25 = BASELINE synthetic shape")
    p.add_argument("--stream_window", type=int, default=0,
                   help="W > 0: no [S,C,D] trace, ESS / R-hat from in-kernel streaming statistics (radon_synth default 64)")
    p.add_argument("--num_optimization_steps", type=int, default=3000)
    p.add_argument("--num_mc_samples", type=int, default=256)
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
    if method == "dVIP":   # a thresholded pattern (every other coordinate centred), b = 1
        return (np.arange(D) % 2).astype(np.float64), np.ones(D)
    # cVIP as the reference runs it: learned a (here a fixed mid-way value), b = 1 (tied as written)
    return np.full(D, 0.5), np.ones(D)


# algorithmic flop per gradient evaluation of one chain (naive: every observation evaluated), the
# resource that bounds the model's kernel, and the initial q-scale the step sizes are derived from.
#   flop: callable (raw, D) -> flop / grad eval;  bytes: HBM bytes / grad eval when the state is not on chip
MODEL_TABLE = {
    "german_synth": dict(bound="tensor"),
    "german_credit_lognormalcentered": dict(bound="tensor"),
    "german_credit_gammascale": dict(bound="tensor"),
    "8schools": dict(bound="fp32", flop=lambda raw, D: 1.2e2),
    "radon": dict(bound="fp32", flop=lambda raw, D: 6.0 * len(raw["y"]) + 8.0 * len(raw["u"])),
    "radon_stddvs": dict(bound="fp32", flop=lambda raw, D: 10.0 * len(raw["y"]) + 12.0 * len(raw["u"])),
    "election": dict(bound="fp32", flop=lambda raw, D: 10.0 * len(raw["y"])),
    "electric": dict(bound="fp32", flop=lambda raw, D: 14.0 * len(raw["y"]) + 6.0 * D),
    "time_series": dict(bound="fp32", flop=lambda raw, D: 30.0 * len(raw["y"]), sigma_q=1e-3),
    # state not on chip: per leapfrog step read z, v and write z, v ([J] each; the gradient is recomputed): 4 J x 4 B
    "radon_synth": dict(bound="hbm", flop=lambda raw, D: 6.0 * len(raw["y"]) + 8.0 * len(raw["u"]),
                        bytes=lambda raw, D: 16.0 * len(raw["u"]), sigma_q=2e-3),
}


def workload(args):
    """-> (library model name, raw data, D, a, b, description)"""
    from autoreparam_b200 import data
    m = args.model
    if m == "german_synth":
        raw = data.synthetic_german_credit(n=1000, f=args.features)
        name, desc = "german_credit_lognormalcentered", "synthetic 1000x%d" % args.features
    elif m == "radon_synth":
        raw = data.synthetic_radon()
        name, desc = "radon", "synthetic 10^6 observations x 10^4 counties"
    else:
        from tests import common   # the committed data fixtures (tests/golden/data_*.npz)
        raw = common.raw_data(m, "PA")
        name, desc = m, "real data" + (" (PA)" if m.startswith("radon") else "")
    from autoreparam_b200 import models
    mc = models.from_data(name, raw)
    D = mc.num_coords
    a, b = method_ab(args.method, D)
    return name, raw, mc, D, a, b, desc


def flop_per_grad(args, raw, D):
    if args.model.startswith("german"):
        F = raw["X"].shape[1]
        return FLOP_PER_GRAD.get(F, 4.0 * 1000 * F + 12.0 * F + 6.0 * 1000)
    return float(MODEL_TABLE[args.model]["flop"](raw, D))


def init_states(D, C, rank, sigma=0.1269):
    """Initial states / step sizes of the shape VI hands to HMC (loc ~ 0, sigma_q ~ softplus(-2))."""
    rng = np.random.default_rng(20190603 + 7919 * rank)
    sigma_q = np.full(D, sigma)
    z0 = (sigma_q * rng.standard_normal((C, D))).astype(np.float32)
    return z0, sigma_q


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ---------------------------------------------------------- reference arm ---
def cpu_reference_run(args, name, raw, D, a, b, steps, warmup):
    """The reference's CPU implementation of the path, restated (oracle port: PyTorch CPU fp32, [C, D] tensors, dense
    one-hot matmuls, autograd gradient at every leapfrog step / ELBO step, TFP op order), all host threads, on a
    bounded sample of the workload."""
    import torch
    from oracle import oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    sigma = MODEL_TABLE[args.model].get("sigma_q", 0.1269)
    times, units = [], 0
    if args.inference == "VI":
        vi_steps = 8 if args.model in ("election",) else 20
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.vi_cpu_batched(name, raw, args.num_mc_samples, vi_steps, 0.05, a, b, seed=i)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt); units += vi_steps
        sample = "%d Adam steps of one learning rate, S=%d (same model/data/method), %d steps" % (
            vi_steps, args.num_mc_samples, steps)
    else:
        big = args.model == "radon_synth"
        chains = 1024 if args.model.startswith("german") else (8 if big else 256)
        transitions = 50 if args.model.startswith("german") else (2 if big else 20)
        z0, sigma_q = init_states(D, chains, 0, sigma)
        eps0 = sigma_q / (args.num_leapfrog_steps / 4.0) ** 2
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            if args.model.startswith("german_synth"):
                n, _ = O.german_hmc_cpu(raw["X"], raw["y"], z0, eps0, args.num_leapfrog_steps, transitions, a, b,
                                        seed=i, num_adapt=transitions)
            else:
                n, _ = O.hmc_cpu_batched(name, raw, z0, eps0, args.num_leapfrog_steps, transitions, a, b, seed=i,
                                         num_adapt=transitions, gather=big)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt); units += n
        sample = "%d chains x %d transitions x L=%d per step (same model/data/method%s), %d steps" % (
            chains, transitions, args.num_leapfrog_steps, ", county gather instead of the dense one-hot" if big else "",
            steps)
    total = sum(times)
    return {"value": units / total, "ms_per_step": 1e3 * total / max(steps, 1), "cores": torch.get_num_threads(),
            "sample": sample}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.model == "radon_synth" and args.stream_window == 0:
        args.stream_window = 64
    vi = args.inference == "VI"
    name, raw, mc, D, a, b, desc = workload(args)
    metric = "elbo_iterations_per_sec" if vi else "leapfrog_grad_evals_per_sec"
    unit = "elbo_iterations/s" if vi else "grad_evals/s"
    if vi:
        wl = "%s VI %s, %s (D=%d), %d learning rates x %d Adam steps, S=%d" % (
            name, args.method, desc, D, 5, args.num_optimization_steps, args.num_mc_samples)
    else:
        wl = "%s HMC %s, %s (D=%d), %d chains/GPU, L=%d, S=%d kept (thin 2), burn-in %d, adapt %d" % (
            name, args.method, desc, D, args.chains, args.num_leapfrog_steps, args.num_samples, args.num_burnin_steps,
            args.num_adaptation_steps)
        if args.stream_window:
            wl += ", streaming ESS window %d (no trace stored)" % args.stream_window
    config = {"workload": wl, "chains_per_gpu": args.chains, "parallelism": "chains sharded, dp%d" % world,
              "l2": "per-step output (samples) exceeds L2 and a 512 MiB buffer is rewritten between timed steps"}
    if vi:
        config["parallelism"] = "learning rates x Monte-Carlo samples inside one GPU; replicas only across GPUs"
        config.pop("chains_per_gpu")

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(args, name, raw, D, a, b, args.steps, args.warmup)
        line = {"impl": "reference", "metric": metric, "value": r["value"], "unit": unit,
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic" if "synth" in args.model else "real (committed fixture)", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": "port",
                                 "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from autoreparam_b200 import engine, graphs, inference

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

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    pk, pk_src = peaks()
    if vi:
        run_vi(args, rank, world, dev, name, raw, mc, D, a, b, metric, unit, config, barrier, max_over_ranks, flush,
               pk, pk_src)
        if world > 1:
            dist.destroy_process_group()
        return

    L, S, W = args.num_leapfrog_steps, args.num_samples, args.stream_window
    if args.scaling == "strong":
        from autoreparam_b200 import distributed
        lo, hi = distributed.shard_range(args.chains, rank, world)
        C, chain_lo = hi - lo, lo
        assert C > 0, "more ranks than chains"
    else:
        C, chain_lo = args.chains, rank * args.chains
    config["scaling"] = args.scaling
    config["chains_total"] = args.chains if args.scaling == "strong" else args.chains * world
    z0, sigma_q = init_states(D, C, rank, MODEL_TABLE[args.model].get("sigma_q", 0.1269))
    eps0 = sigma_q / (L / 4.0) ** 2
    T = engine.hmc_num_transitions(S, args.num_burnin_steps)
    evals_per_step = C * L * T            # this rank
    evals_all = config["chains_total"] * L * T   # all ranks
    z_dev = torch.as_tensor(z0, device=dev)
    bufs = {"is_accepted": torch.empty((S, C), dtype=torch.uint8, device=dev)}
    if not W:
        bufs["samples"] = torch.empty((S, C, D), dtype=torch.float32, device=dev)
    kw = dict(num_leapfrog_steps=L, num_results=S, num_burnin_steps=args.num_burnin_steps,
              num_adaptation_steps=args.num_adaptation_steps, chain_offset=chain_lo, want_final=False)
    if W:
        kw.update(engine=engine.ENGINE_SIMT, want_samples=False, want_is_accepted=True, stream_window=W)
    else:
        kw.update(engine=args.engine)

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
    ms = max_over_ranks(sum(s.elapsed_time(e) for s, e in ev))
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
    ekw = dict(num_leapfrog_steps=L, num_samples=S, num_burnin_steps=args.num_burnin_steps,
               num_adaptation_steps=args.num_adaptation_steps, chain_offset=chain_lo, device=dev,
               engine_kind=args.engine, return_is_accepted=False, stream_window=W)
    inference.hmc(target, mc, step_sizes, z0, seed=1, **ekw)
    barrier()
    e0 = time.perf_counter()
    for i in range(args.steps):
        res = inference.hmc(target, mc, step_sizes, z0, seed=2000 + i, **ekw)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - e0)
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
        flop = flop_per_grad(args, raw, D)
        rate = evals_per_step * args.steps / (ms * 1e-3)          # this rank's kernel, grad evals / s
        f_clk = 1e6 * (clocks.get("sm_mhz") or pk.get("sm_max_mhz", 1965.0))
        bound = MODEL_TABLE[args.model]["bound"]
        if bound == "tensor":
            F = raw["X"].shape[1]
            achieved_tf = rate * flop / 1e12
            peak_tf = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
            # the pipe that actually binds the dominant kernel: MUFU ops per observation on the XU
            # pipe, 16 results / clk / SM (profiles/micro/pipes.cu), on the SMs the 128-chain tiles occupy
            N = raw["X"].shape[0]
            n_pad = (N + 127) // 128 * 128
            mufu_per_obs = (1.0 + 0.25 + 0.25 / L) if F <= 32 else (2.0 + 1.0 / L)
            sms_used = min(148, (C + 127) // 128)   # rank 0's tiles
            xu_roof = sms_used * f_clk / (n_pad * mufu_per_obs / 16.0)
            roof = {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": achieved_tf / peak_tf,
                    "traffic": NCU_TRAFFIC_DEFAULT_WORKLOAD if (args.model, C, S, F, args.num_burnin_steps) ==
                    ("german_synth", 16384, 1000, 25, 500) else None,
                    "traffic_source": "profiles/r01_bench_kernel_traffic.csv (ncu capture of this command, not re-measured by this run)",
                    "note": "algorithmic fp32 flop (%.3g per grad eval) / measured dense bf16 cuBLAS peak (%s, "
                            "sustained); per GPU" % (flop, pk_src),
                    "binding_pipe": {"pipe": "xu (MUFU), co-limited by instruction dispatch",
                                     "achieved": rate, "peak": xu_roof, "unit": "grad_evals/s per GPU",
                                     "frac": rate / xu_roof,
                                     "note": "%.4g MUFU ops per observation x %d padded observations, 16 MUFU "
                                             "results/clk/SM (measured), %d SMs occupied by the 128-chain tiles, "
                                             "SM clock sampled under load" % (mufu_per_obs, n_pad, sms_used)}}
        elif bound == "hbm":
            by = float(MODEL_TABLE[args.model]["bytes"](raw, D))
            if W:   # streaming statistics, per kept sample and coordinate: pivot read + ring write, and once per block of W
                    # kept samples 2 W ring reads + W lag sums read-modify-written + the running sum
                by += (S / float(L * T)) * D * 4.0 * (2.0 + (4.0 * W + 2.0) / W)
            achieved = rate * by / 1e9
            peak = pk.get("hbm_gbs_sustained", pk["hbm_gbs"])
            roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None,
                    "note": "algorithmic bytes %.4g per grad eval (state vectors not on chip: read z, v + write z, v per "
                            "leapfrog step%s) / %s HBM copy bandwidth; naive-flop rate %.3g TFLOP/s (the kernel works from "
                            "per-county sufficient statistics)" % (by, " + the streaming-ESS planes" if W else "", pk_src,
                                                                   rate * flop / 1e12)}
        else:
            # FP32 issue: 148 SMs x 128 lanes x 2 flop / clk.  The contract's bounds are hbm | tensor; these small models
            # run from on-chip state with no tensor-core work, so the honest bound is the FP32 pipe and it is named so.
            peak_tf = 148 * 128 * 2 * f_clk / 1e12
            achieved_tf = rate * flop / 1e12
            roof = {"bound": "fp32", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": achieved_tf / peak_tf, "traffic": None,
                    "note": "naive algorithmic flop (%.3g per grad eval, SURVEY.md 8d) / FP32 FMA peak at the sampled SM "
                            "clock; radon / election evaluate merged cells or sufficient statistics, so the fraction can "
                            "exceed what the executed instruction count would give" % flop}
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic" if "synth" in args.model else "real (committed fixture)",
            "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h),
                    "returns": "ESS [C,D], step sizes / accept counts [C], R-hat [D] (reduced over all ranks); the [S,C,D] "
                               "samples and [S,C] accept flags stay on the device (main.py reduces them to these; the "
                               "reference's sess.run returns them, main.py:350-360)"},
            "gpu_launches": int(launches),
            "roofline": roof,
            # acceptance rate and R-hat are over the chains of ALL ranks: inference.hmc all-reduces the per-chain
            # moments / accept counters (NCCL) inside the e2e timed region
            "ess": {"ess_per_sec": ess_per_sec, "ess_per_1000_grads_mean": ess_per_1000,
                    "acceptance_rate": res.accept_stats[1] / (res.accept_stats[2] * T),
                    "chains_reduced": int(res.accept_stats[2]),
                    "rhat_max": None if res.rhat is None else float(np.nanmax(res.rhat))},
            "wall_s_timed_region": wall,
        }
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(args, name, raw, D, a, b, steps=3, warmup=1)   # ~10-15 s of CPU work
            line["cpu_baseline"] = {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_vi(args, rank, world, dev, name, raw, mc, D, a, b, metric, unit, config, barrier, max_over_ranks, flush, pk,
           pk_src):
    """BASELINE configs[3]: find_best_learning_rate (inference.py:26-154) -- 5 learning rates x num_optimization_steps
    Adam steps x S Monte-Carlo samples in one persistent launch.  A step of the bench = one such call; every rank runs
    an independent replica (VI does not shard: SURVEY.md 8e)."""
    import torch
    from autoreparam_b200 import engine, graphs, inference
    lrs = [0.02, 0.05, 0.1, 0.2, 0.4]
    steps_vi, S = args.num_optimization_steps, args.num_mc_samples
    rng = np.random.default_rng(7 + rank)
    loc0 = 1e-2 * rng.standard_normal((len(lrs), D))
    rho0 = np.full((len(lrs), D), -2.0)
    iters = len(lrs) * steps_vi

    def step(i):
        return engine.vi_run(mc, a, b, loc0, rho0, lrs, num_mc_samples=S, num_optimization_steps=steps_vi, seed=100 + i)

    for i in range(args.warmup):
        step(i)
    sampler = ClockSampler(dev.index)
    barrier()
    sampler.start()
    launches0 = engine.kernel_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        ev[i][0].record()
        out = step(args.warmup + i)
        ev[i][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = engine.kernel_launch_count() - launches0
    clocks = sampler.stop()
    ms = max_over_ranks(sum(s.elapsed_time(e) for s, e in ev))
    value = world * iters * args.steps / (ms * 1e-3)
    # end to end: the public call with its host-side selection of the best learning rate
    target = graphs.TargetGraph(mc, args.method, a, b, False)
    barrier()
    e0 = time.perf_counter()
    for i in range(args.steps):
        best = inference.find_best_learning_rate(target, mc, learning_rates=lrs, num_optimization_steps=steps_vi,
                                                 num_mc_samples=S, seed=200 + i)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - e0)
    if rank == 0:
        flop = flop_per_grad(args, raw, D)
        f_clk = 1e6 * (clocks.get("sm_mhz") or pk.get("sm_max_mhz", 1965.0))
        peak_tf = 148 * 128 * 2 * f_clk / 1e12
        achieved_tf = (iters * args.steps / (ms * 1e-3)) * S * flop / 1e12
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic" if "synth" in args.model else "real (committed fixture)",
            "config": config, "clocks": clocks,
            "e2e": {"value": world * iters * args.steps / e2e_s, "unit": unit,
                    "h2d_bytes_per_step": int(2 * len(lrs) * D * 4 + 2 * D * 4),
                    "d2h_bytes_per_step": int(2 * len(lrs) * D * 4 + 2 * len(lrs) * steps_vi * 4)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp32", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf, "traffic": None,
                         "note": "S x naive flop per gradient evaluation (%.3g) per ELBO iteration / FP32 FMA peak; the "
                                 "optimisation is %d dependent steps, so latency per step, not throughput, is what "
                                 "bounds it (SURVEY.md 8d)" % (flop, steps_vi)},
            "vi": {"best_elbo": float(best[0]), "best_lr": float(best[2]),
                   "elbo_last32_per_lr": [float(np.mean(out["elbo"][r][-32:])) for r in range(len(lrs))],
                   "seconds_per_call": ms * 1e-3 / args.steps},
            "wall_s_timed_region": wall,
        }
        if world == 1 and not args.no_cpu_baseline:
            from types import SimpleNamespace  # noqa: F401
            r = cpu_reference_run(args, name, raw, D, a, b, steps=2, warmup=1)
            line["cpu_baseline"] = {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"]}
        print(json.dumps(line))


if __name__ == "__main__":
    main()
