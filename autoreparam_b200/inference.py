"""Inference engines: host-side mirror of the reference's ``inference.py``.

``hmc`` replaces ``inference.hmc`` (``inference.py:198-242``) and
``find_best_learning_rate`` replaces ``inference.find_best_learning_rate``
(``inference.py:26-154``).  Both are thin: every transition / optimisation step
runs inside the CUDA library; this module stages buffers, calls the C ABI and
shapes the results the way ``main.py`` consumes them.  The reference reads its
sizes from global absl FLAGS; here they are keyword arguments with the flag
names.
"""
from __future__ import annotations

import collections

import numpy as np

from . import distributed, engine, util


class HmcResult(collections.namedtuple(
        "HmcResult", "ess is_accepted samples rhat step_mult accept_count num_transitions ess_flat accept_stats")):
    """ess: list of [C, *site] (un-normalised, as tfp.mcmc.effective_sample_size);
    is_accepted [S, C] bool; samples: list of [S, n_save, *site] centred traces
    (None unless chains were requested); rhat [D] per coordinate over the chains of ALL ranks;
    accept_stats = (accepted kept transitions, accepted transitions, chains) summed over ALL ranks."""


_PINNED = {}


def _pinned_like(t, tag=0):
    """Reusable pinned host staging buffer for a device tensor (allocated once per tag / shape / dtype)."""
    import torch
    key = (tag, tuple(t.shape), t.dtype)
    if key not in _PINNED:
        _PINNED[key] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    return _PINNED[key]


def _flat_step_sizes(model_config, step_size_init):
    """initial_step_size is a list per site in trace order (main.py:283-284)."""
    if isinstance(step_size_init, np.ndarray) and step_size_init.shape == (model_config.num_coords,):
        return step_size_init.astype(np.float64)
    parts = []
    for (name, shape), s in zip(model_config.sites, step_size_init):
        parts.append(np.broadcast_to(np.asarray(s, dtype=np.float64), shape).reshape(-1))
    return np.concatenate(parts)


def hmc(target, model_config, step_size_init, initial_states, reparam=None, *, num_leapfrog_steps, num_samples,
        num_burnin_steps, num_adaptation_steps, num_chains_to_save=0, seed=0, chain_offset=0, device="cuda",
        engine_kind=engine.ENGINE_AUTO, precision="f32", keep_on_device=False, return_is_accepted=True,
        stream_window=0):
    """HMC + dual-averaging adaptation + thinning + to-centred + ESS (``inference.py:198-242``).

    target            TargetGraph (graphs.py) -- carries the (a, b) rule, so ``reparam`` is accepted
                      for signature parity only.
    step_size_init    list per site (VI's sigma_q); eps0 = sigma_q / (L / 4)**2  (inference.py:212-216)
    initial_states    list of [C, *site] arrays (util.variational_inits_from_params) or flat [C, D]
    return_is_accepted  False: the [S, C] accept flags stay on the device (``is_accepted`` is None); their sum is in
                      ``accept_stats`` -- all the drop-in driver reads (main.py:372-373)
    stream_window     W > 0: no [S, C, D] trace is stored; mean / variance / ESS come from in-kernel streaming
                      statistics with a W-lag window (for runs whose traces do not fit: BASELINE configs[4]).
                      ``num_chains_to_save`` must then be 0.
    """
    import torch

    mc = model_config
    if isinstance(initial_states, np.ndarray) and initial_states.ndim == 2:
        z0 = initial_states
    else:
        z0 = mc.join(list(initial_states))
    C, D = z0.shape
    eps0 = _flat_step_sizes(mc, step_size_init) / (float(num_leapfrog_steps) / 4.0) ** 2
    tdt = torch.float32 if precision == "f32" else torch.float64
    dev = torch.device(device)
    # host -> device from pinned memory
    z_np = np.ascontiguousarray(z0, dtype=np.float32 if precision == "f32" else np.float64)
    z_pin = _pinned_like(torch.from_numpy(z_np), "z0")
    z_pin.copy_(torch.from_numpy(z_np))
    z_dev = z_pin.to(dev, non_blocking=True)
    if stream_window > 0:
        assert num_chains_to_save == 0, "streaming statistics keep no traces"
        out = engine.hmc_run(mc, z_dev, eps0, target.a, target.b, num_leapfrog_steps=num_leapfrog_steps,
                             num_results=num_samples, num_burnin_steps=num_burnin_steps,
                             num_adaptation_steps=num_adaptation_steps, seed=seed, chain_offset=chain_offset,
                             want_final=False, engine=engine.ENGINE_SIMT, precision=precision, want_samples=False,
                             want_is_accepted=True, stream_window=stream_window)
        ess_dev, mean_dev, var_dev = out["stream_ess"], out["stream_mean"], out["stream_var"]
    else:
        out = engine.hmc_run(mc, z_dev, eps0, target.a, target.b, num_leapfrog_steps=num_leapfrog_steps,
                             num_results=num_samples, num_burnin_steps=num_burnin_steps,
                             num_adaptation_steps=num_adaptation_steps, seed=seed, chain_offset=chain_offset,
                             want_final=False, engine=engine_kind, precision=precision)
        ess_dev, mean_dev, var_dev = engine.ess(out["samples"], precision=precision, want_moments=True)
    # R-hat from the per-chain moments: [C, D] -> packed sums [3 D + 1] on the device, summed over the ranks (one small
    # NCCL all-reduce when the chains are sharded over several GPUs), -> [D].  Accept counters ride in a second
    # 3-element all-reduce.  These are the only collectives of an HMC run.
    stats = distributed.rhat_stats(mean_dev, var_dev)
    acc = torch.stack([out["is_accepted"].sum(dtype=torch.float64), out["accept_count"].sum(dtype=torch.float64),
                       torch.tensor(float(C), dtype=torch.float64, device=dev)])
    distributed.allreduce_sum_(stats)
    distributed.allreduce_sum_(acc)
    rhat_dev = distributed.rhat_from_stats(stats, num_samples) if (C > 1 or distributed.is_multi()) else None
    # device -> host through pinned buffers, one synchronisation: only what the caller consumes
    dev_out = [ess_dev, out["step_mult"], out["accept_count"], acc] + \
        ([rhat_dev] if rhat_dev is not None else [None]) + ([out["is_accepted"]] if return_is_accepted else [None])
    host = [None if t is None else _pinned_like(t, tag) for tag, t in enumerate(dev_out)]
    for h, t in zip(host, dev_out):
        if t is not None:
            h.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    ess_flat, step_mult, accept_count, acc_host, rhat, is_acc_u8 = [None if h is None else h.numpy().copy() for h in host]
    is_acc = None if is_acc_u8 is None else is_acc_u8.view(np.bool_)
    samples = None
    if num_chains_to_save > 0:
        samples = mc.split(out["samples"][:, :num_chains_to_save].cpu().numpy())
    res = HmcResult(ess=mc.split(ess_flat), is_accepted=is_acc, samples=samples,
                    rhat=rhat,
                    step_mult=step_mult, accept_count=accept_count,
                    num_transitions=out["num_transitions"], ess_flat=ess_flat,
                    accept_stats=tuple(float(v) for v in acc_host))
    if keep_on_device:
        return res, out
    return res


def hmc_tuning_grid(target, model_config, step_size_init, initial_states, *, num_leapfrog_steps, num_samples,
                    num_burnin_steps, num_adaptation_steps, seed=0, chain_offset=0, device="cuda",
                    engine_kind=engine.ENGINE_AUTO, precision="f32"):
    """The whole ``num_leapfrog_steps`` tuning grid in ONE launch (``arp_hmc_run_many``): what the reference does
    with one ``--inference=HMCtuning --num_leapfrog_steps=L`` process per L (``main.py:316-329, 375-384``).
    All arguments that the reference rescales per L (``--count_in_leapfrog_steps``, ``main.py:318-324``) are lists,
    one entry per L.  Each run equals ``hmc(...)`` with that L (same initial states, same random streams).
    Returns a list of ``HmcResult`` (``samples`` / ``is_accepted`` None: the tuning driver only reads ESS)."""
    import torch

    mc = model_config
    Ls = [int(l) for l in num_leapfrog_steps]
    n = len(Ls)
    bc = lambda v: [int(x) for x in v] if isinstance(v, (list, tuple, np.ndarray)) else [int(v)] * n
    Ss, Bs, As = bc(num_samples), bc(num_burnin_steps), bc(num_adaptation_steps)
    z0 = initial_states if (isinstance(initial_states, np.ndarray) and initial_states.ndim == 2) \
        else mc.join(list(initial_states))
    C, D = z0.shape
    sig = _flat_step_sizes(mc, step_size_init)
    eps = [sig / (float(l) / 4.0) ** 2 for l in Ls]                      # inference.py:212-216
    dev = torch.device(device)
    tdt = torch.float32 if precision == "f32" else torch.float64
    z_dev = torch.as_tensor(np.ascontiguousarray(z0), dtype=tdt).to(dev)
    outs = engine.hmc_run_many(mc, z_dev, eps, target.a, target.b, num_leapfrog_steps=Ls, num_results=Ss,
                               num_burnin_steps=Bs, num_adaptation_steps=As, seed=seed, chain_offset=chain_offset,
                               engine=engine_kind, precision=precision)
    results = []
    for o, S in zip(outs, Ss):
        ess_dev, mean_dev, var_dev = engine.ess(o["samples"], precision=precision, want_moments=True)
        stats = distributed.allreduce_sum_(distributed.rhat_stats(mean_dev, var_dev))
        acc = torch.stack([o["is_accepted"].sum(dtype=torch.float64), o["accept_count"].sum(dtype=torch.float64),
                           torch.tensor(float(C), dtype=torch.float64, device=dev)])
        distributed.allreduce_sum_(acc)
        rhat = distributed.rhat_from_stats(stats, S).cpu().numpy() if (C > 1 or distributed.is_multi()) else None
        ess_flat = ess_dev.cpu().numpy()
        results.append(HmcResult(ess=mc.split(ess_flat), is_accepted=None, samples=None, rhat=rhat,
                                 step_mult=o["step_mult"].cpu().numpy(), accept_count=o["accept_count"].cpu().numpy(),
                                 num_transitions=o["num_transitions"], ess_flat=ess_flat,
                                 accept_stats=tuple(float(v) for v in acc.cpu().numpy())))
    return results


def find_best_learning_rate(target, model_config, *, learning_rates, num_optimization_steps, num_mc_samples,
                            seed=0, precision="f32", init_rng=None, log_fn=None, discrete_prior=False):
    """Adam on the mean-field ELBO for every learning rate (``inference.py:26-154``).

    All learning rates run concurrently in one kernel launch.  Returns the
    reference's tuple ``(best_elbo, best_timeline, best_lr, step_size_init,
    learned_variational_params, learned_reparam)``; ``learned_reparam`` is None
    unless ``target.learnable`` (cVIP).  ``discrete_prior`` (``main.py:244-253``): the mixture-of-Laplace
    prior on the learnable parameters enters the objective; the returned ELBO is the 'pure' one
    (``inference.py:150``: objective minus the prior term, both averaged over the last 32 steps)."""
    from . import graphs
    mc = model_config
    D = mc.num_coords
    R = len(learning_rates)
    rng = np.random.default_rng(seed) if init_rng is None else init_rng
    # program_transformations.py:207-215: loc = 1e-2 * randn, scale = softplus(-2); re-initialised per run
    loc0 = 1e-2 * rng.standard_normal((R, D))
    rho0 = np.full((R, D), -2.0)
    P = target.num_params if target.learnable else 0
    out = engine.vi_run(mc, target.a, target.b, loc0, rho0, [float(l) for l in learning_rates],
                        num_mc_samples=num_mc_samples, num_optimization_steps=num_optimization_steps,
                        u=np.zeros((R, P)) if P else None,        # sigmoid(0) = 0.5, :507-510
                        a_index=target.a_index, b_index=target.b_index, num_params=P,
                        discrete_prior=bool(discrete_prior and P), seed=seed, precision=precision)
    best = None
    for r, lr in enumerate(learning_rates):
        timeline = out["elbo"][r]
        this_elbo = float(np.mean(timeline[-32:]))                     # inference.py:121
        this_plp = float(np.mean(out["prior_logp"][r][-32:]))          # inference.py:122
        if log_fn is not None:
            for step in range(0, num_optimization_steps, 100):         # inference.py:107-108
                log_fn("step {} elbo {}".format(step, timeline[step]))
            log_fn("     finished optimization with elbo {} vs best ELBO {}".format(
                this_elbo, None if best is None else best[0]))
        if not np.isfinite(this_elbo):                                 # inference.py:127
            continue
        if best is None or best[0] < this_elbo:
            best = (this_elbo, r, float(lr), this_plp)
    if best is None:
        raise FloatingPointError("no learning rate produced a finite ELBO")
    best_elbo_with_prior, r, best_lr, best_plp = best
    best_elbo = best_elbo_with_prior - best_plp                        # inference.py:150
    scale = np.logaddexp(out["rho"][r].astype(np.float64), 0.0)          # softplus
    params = collections.OrderedDict()
    for (name, shape), lo, sc in zip(mc.sites, mc.split(out["loc"][r]), mc.split(scale)):
        params[name + "_loc"] = np.asarray(lo, dtype=np.float32)
        params[name + "_scale"] = np.asarray(sc, dtype=np.float32)
    step_size_init = util.get_approximate_step_size(params, num_leapfrog_steps=1)  # inference.py:42-43
    learned_reparam = None
    if target.learnable:
        vals = 1.0 / (1.0 + np.exp(-out["u"][r].astype(np.float64)))
        learned_reparam = graphs.learned_reparam_from_params(target, vals)
    return (best_elbo, list(out["elbo"][r]), best_lr, step_size_init, params, learned_reparam)


class InterleavedResult(collections.namedtuple(
        "InterleavedResult", "ess is_accepted_cp is_accepted_ncp samples rhat ess_flat num_transitions")):
    pass


def hmc_interleaved(model_config, target_cp, target_ncp, num_leapfrog_steps_cp, num_leapfrog_steps_ncp,
                    step_size_cp, step_size_ncp, initial_states_cp, *, num_samples, num_burnin_steps,
                    num_adaptation_steps, num_chains_to_save=0, seed=0, chain_offset=0, device="cuda",
                    precision="f32"):
    """Interleaved CP / NCP HMC (``inference.py:258-329`` + ``interleaved.py``): per transition one CP step and
    one NCP step, each with ``SimpleStepSizeAdaptation(adaptation_rate=0.05, target_accept_prob=0.75)`` and base
    step sizes ``sigma_q / (L / 4)**2`` of its own VI fit.  States / samples are in the centred space."""
    import torch

    mc = model_config
    x0 = initial_states_cp if (isinstance(initial_states_cp, np.ndarray) and initial_states_cp.ndim == 2) \
        else mc.join(list(initial_states_cp))
    C, D = x0.shape
    eps_cp = _flat_step_sizes(mc, step_size_cp) / (float(num_leapfrog_steps_cp) / 4.0) ** 2
    eps_ncp = _flat_step_sizes(mc, step_size_ncp) / (float(num_leapfrog_steps_ncp) / 4.0) ** 2
    dev = torch.device(device)
    x_np = np.ascontiguousarray(x0, dtype=np.float32 if precision == "f32" else np.float64)
    x_pin = _pinned_like(torch.from_numpy(x_np), "x0i")
    x_pin.copy_(torch.from_numpy(x_np))
    out = engine.hmc_interleaved_run(mc, x_pin.to(dev, non_blocking=True), eps_cp, eps_ncp,
                                     (target_cp.a, target_cp.b), (target_ncp.a, target_ncp.b),
                                     num_leapfrog_steps_a=num_leapfrog_steps_cp,
                                     num_leapfrog_steps_b=num_leapfrog_steps_ncp, num_results=num_samples,
                                     num_burnin_steps=num_burnin_steps, num_adaptation_steps=num_adaptation_steps,
                                     seed=seed, chain_offset=chain_offset, precision=precision)
    ess_dev, mean_dev, var_dev = engine.ess(out["samples"], precision=precision, want_moments=True)
    ess_flat = ess_dev.cpu().numpy()
    stats = distributed.allreduce_sum_(distributed.rhat_stats(mean_dev, var_dev))   # R-hat over the chains of all ranks
    rhat = distributed.rhat_from_stats(stats, num_samples).cpu().numpy() if (C > 1 or distributed.is_multi()) else None
    samples = None
    if num_chains_to_save > 0:
        samples = mc.split(out["samples"][:, :num_chains_to_save].cpu().numpy())
    return InterleavedResult(
        ess=mc.split(ess_flat), is_accepted_cp=out["is_accepted_a"].cpu().numpy().astype(bool),
        is_accepted_ncp=out["is_accepted_b"].cpu().numpy().astype(bool), samples=samples,
        rhat=rhat,
        ess_flat=ess_flat, num_transitions=out["num_transitions"])
