"""Drop-in for the reference's ``main.py``: same flags, same results-dir files.

    python -m autoreparam_b200.main --model=8schools --inference=VI  --method=CP
    python -m autoreparam_b200.main --model=8schools --inference=HMC --method=CP --num_leapfrog_steps=4

Flags mirror ``main.py:37-113``; the file-name scheme ``main.py:208-219``; the JSON
keys ``main.py:277-290`` (VI) and ``:375-391,531-550`` (HMC / HMCtuning, appended
to lists); ``<base>_ess.npz / _ess.txt / _traces.npz`` ``main.py:552-585``; the
log lines ``util.print`` emits (``main.py:195,328,370``, ``inference.py:108,123``).
Extra flags (not in the reference): ``--seed``, ``--data_dir``, ``--precision``,
``--tied_b_as_written``.  Under ``torchrun`` the chains of an HMC run are sharded
over the ranks (one GPU each) and rank 0 writes the files.
"""
from __future__ import annotations

import argparse
import collections
import io
import json
import os
import time

import numpy as np

from . import distributed, graphs, inference, models, util


def _bool(v):
    if isinstance(v, bool):
        return v
    return str(v).lower() in ("1", "true", "t", "yes", "y")


def build_parser():
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    a = p.add_argument
    a("--model", default="8schools", help="Model to be used.")
    a("--dataset", default="", help="Dataset to be used.")
    a("--inference", default="VI", help="Inference method to be used: VI, HMCtuning, or HMC.")
    a("--method", default="CP", help="Method to be used: CP, NCP, i (only if inference = HMC), cVIP, dVIP.")
    a("--learnable_parameterisation_type", default="eig",
      help='Type of learnable parameterisation (only affects file names for the in-scope models).')
    a("--reparameterise_variational", type=_bool, nargs="?", const=True, default=False)
    a("--discrete_prior", type=_bool, nargs="?", const=True, default=False)
    a("--tied_pparams", type=_bool, nargs="?", const=True, default=True)
    a("--results_dir", default="", help="File to write results.")
    a("--learning_rates", default="0.02,0.05,0.1,0.2,0.4", help="Learning rates (list)")
    a("--num_optimization_steps", type=int, default=3000)
    a("--num_mc_samples", type=int, default=256)
    a("--num_leapfrog_steps", default=None,
      help="int; with --inference=HMCtuning also a comma-separated list: the whole grid then runs in ONE launch")
    a("--count_in_leapfrog_steps", type=_bool, nargs="?", const=True, default=False)
    a("--num_samples", type=int, default=50000)
    a("--num_chains", type=int, default=100)
    a("--num_burnin_steps", type=int, default=10000)
    a("--num_adaptation_steps", type=int, default=6000)
    a("--num_chains_to_save", type=int, default=0)
    # extensions
    a("--seed", type=int, default=0, help="Philox / initialisation seed (the reference is unseeded).")
    a("--data_dir", default=None, help="Directory holding the reference's data files (default ./data/).")
    a("--precision", default="f32", choices=["f32", "f64"])
    a("--stream_window", type=int, default=0,
      help="W > 0: keep no [S, C, D] trace; ESS / R-hat from in-kernel streaming statistics with a W-lag window "
           "(chain counts whose traces do not fit in memory).")
    a("--tied_b_as_written", type=_bool, nargs="?", const=True, default=True,
      help="True: tied VIP uses b = 1 as the reference does as written; False: the paper's b = a.")
    return p


def results_filename(FLAGS):
    """main.py:208-219."""
    return "{}{}{}{}{}.json".format(
        FLAGS.method,
        ("_" + FLAGS.learnable_parameterisation_type if "VIP" in FLAGS.method else ""),
        ("_tied" if FLAGS.tied_pparams else ""),
        ("_reparam_variational" if "VIP" in FLAGS.method and FLAGS.reparameterise_variational else ""),
        ("_discrete_prior" if "VIP" in FLAGS.method and FLAGS.discrete_prior else ""))


def cvip_path(FLAGS, results_dir):
    """main.py:119-124."""
    return os.path.join(results_dir, "cVIP_{}{}{}{}.json".format(
        FLAGS.learnable_parameterisation_type,
        "_tied" if FLAGS.tied_pparams else "",
        "_reparam_variational" if FLAGS.reparameterise_variational else "",
        "_discrete_prior" if FLAGS.discrete_prior else ""))


def create_target_graph(FLAGS, model_config, results_dir):
    """main.py:117-187: method -> target (+ the reparameterisation actually used)."""
    actual_reparam = None
    if FLAGS.method == "CP":
        target, actual_reparam = graphs.make_cp_graph(model_config), "CP"
    elif FLAGS.method == "NCP":
        target, actual_reparam = graphs.make_ncp_graph(model_config), "NCP"
    elif FLAGS.method == "cVIP":
        if FLAGS.inference == "VI":
            target = graphs.make_cvip_graph(model_config, FLAGS.learnable_parameterisation_type,
                                            tied_pparams=FLAGS.tied_pparams,
                                            tied_b_as_written=FLAGS.tied_b_as_written)
        else:
            with open(cvip_path(FLAGS, results_dir)) as f:
                actual_reparam = json.load(f)["learned_reparam"]
            target = graphs.make_dvip_graph(model_config, actual_reparam, FLAGS.learnable_parameterisation_type)
    elif FLAGS.method == "dVIP":
        path = cvip_path(FLAGS, results_dir)
        if not os.path.exists(path):
            raise Exception("Run cVIP first to find reparameterisation")
        with open(path) as f:
            reparam = json.load(f)["learned_reparam"]
        discrete = graphs.discretise(reparam)
        print("discrete parameterisation is", discrete)
        target = graphs.make_dvip_graph(model_config, discrete, FLAGS.learnable_parameterisation_type)
        actual_reparam = discrete
    elif FLAGS.method == "i":
        if FLAGS.inference == "VI":
            raise Exception("Cannot run interleaved VI. Use `i` method with HMC only.")   # main.py:136-137
        return (graphs.make_cp_graph(model_config), graphs.make_ncp_graph(model_config)), None
    else:
        raise Exception("unknown method {}".format(FLAGS.method))
    return target, actual_reparam


def _clean(d):
    if d is None:
        return None
    return collections.OrderedDict((k, np.asarray(v).item() if np.ndim(v) == 0 else np.asarray(v).tolist())
                                   for k, v in d.items())


def run_vi(FLAGS, model_config, results_dir, file_path):
    """main.py:234-290."""
    if os.path.exists(file_path):
        util.print("Already ran experiment {}-{} on model {} with dataset {}. Skipping".format(
            FLAGS.inference, FLAGS.method, FLAGS.model, FLAGS.dataset))
        return
    target, actual_reparam = create_target_graph(FLAGS, model_config, results_dir)
    lrs = [float(x) for x in (FLAGS.learning_rates.split(",") if isinstance(FLAGS.learning_rates, str)
                              else FLAGS.learning_rates)]
    start_time = time.time()
    (elbo_final, elbo_timeline, learning_rate, initial_step_size, learned_variational_params,
     learned_reparam) = inference.find_best_learning_rate(
        target, model_config, learning_rates=lrs, num_optimization_steps=FLAGS.num_optimization_steps,
        num_mc_samples=FLAGS.num_mc_samples, seed=FLAGS.seed, precision=FLAGS.precision, log_fn=util.print,
        discrete_prior=FLAGS.discrete_prior)
    end_time = time.time()
    if learned_reparam is None and isinstance(actual_reparam, dict):
        learned_reparam = actual_reparam  # main.py:266-267: save actual parameters used for dVIP
    results = {
        "elbo": float(elbo_final),
        "variational_fit_time_secs": end_time - start_time,
        "actual_num_variational_steps": len(elbo_timeline),
        "estimated_elbo_std": float(np.std(elbo_timeline[-32:])),
        "learning_rate": learning_rate,
        "initial_step_size": [np.asarray(i).item() if np.ndim(i) == 0 else np.asarray(i).tolist()
                              for i in initial_step_size],
        "learned_reparam": _clean(learned_reparam),
        "learned_variational_params": _clean(learned_variational_params),
    }
    with open(file_path, "w") as outfile:
        json.dump(results, outfile)


def get_best_num_leapfrog_steps_from_tuning_runs(tuning_runs):
    """main.py:292-294."""
    best_run = max(tuning_runs, key=lambda d: d["ess_min"])
    return best_run["num_leapfrog_steps"]


def _parse_leapfrog_steps(FLAGS):
    v = FLAGS.num_leapfrog_steps
    if v is None or v == "":
        return []
    if isinstance(v, (int, np.integer)):
        return [int(v)]
    return [int(x) for x in str(v).split(",") if x.strip()]


def run_hmc_tuning_grid(FLAGS, model_config, results_dir, file_path, grid):
    """--inference=HMCtuning --num_leapfrog_steps=L1,L2,...: the reference runs one process per L (main.py:316-329,
    375-384); here (chains x L) is the batch axis of ONE launch and one `tuning_runs` entry per L is appended, in the
    order given, exactly as the separate invocations would have written them."""
    if not os.path.exists(file_path):
        raise Exception("Run VI first to find initial step sizes")
    with open(file_path) as f:
        prev_results = json.load(f)
    done = set(r["num_leapfrog_steps"] for r in prev_results.get("tuning_runs", []))
    todo = []
    for L in grid:
        if L in done:
            print("A tuning run already exists for HMC with {} leapfrog steps skipping.".format(L))
        else:
            todo.append(L)
    if not todo:
        return
    rank, world = distributed.rank_world()
    _check_sharding(FLAGS, world)
    rng = np.random.default_rng(FLAGS.seed + 1)
    initial_states = list(util.variational_inits_from_params(
        prev_results["learned_variational_params"], param_names=model_config.param_names,
        num_inits=FLAGS.num_chains, rng=rng).values())
    target, _ = create_target_graph(FLAGS, model_config, results_dir)
    lo, hi = distributed.shard_range(FLAGS.num_chains, rank, world)
    z0 = model_config.join(initial_states)[lo:hi]
    scale = (lambda v, L: int(v / float(L))) if FLAGS.count_in_leapfrog_steps else (lambda v, L: v)   # main.py:318-324
    Ss = [scale(FLAGS.num_samples, L) for L in todo]
    Bs = [scale(FLAGS.num_burnin_steps, L) for L in todo]
    As = [scale(FLAGS.num_adaptation_steps, L) for L in todo]
    util.print("\nNumber of leaprog steps: grid {} in one launch.\n".format(todo))
    start_time = time.time()
    results = inference.hmc_tuning_grid(target, model_config, prev_results["initial_step_size"], z0,
                                        num_leapfrog_steps=todo, num_samples=Ss, num_burnin_steps=Bs,
                                        num_adaptation_steps=As, seed=FLAGS.seed, chain_offset=lo,
                                        device=distributed.local_device(), precision=FLAGS.precision)
    mcmc_time = time.time() - start_time
    for L, S, B, res in zip(todo, Ss, Bs, results):
        ess_flat = distributed.gather_chains(res.ess_flat, distributed.local_device())
        if rank != 0:
            continue
        normalized = [1000 * e / (S * L) for e in model_config.split(ess_flat)]
        ess_min, sem_min = util.get_min_ess(normalized, FLAGS.num_chains)
        util.print("L = {}: ESS per 1000 gradients: {} +/- {}".format(L, ess_min, sem_min))
        acceptance_rate = res.accept_stats[0] * 100. / float(S * FLAGS.num_chains)
        # mcmc_time: the grid shares one launch; every entry carries the time of the whole grid divided evenly
        save_hmc_results(file_path=file_path, tuning_runs={
            "num_leapfrog_steps": L, "ess_min": float(ess_min), "sem_min": float(sem_min),
            "acceptance_rate": float(acceptance_rate), "mcmc_time": mcmc_time / len(todo), "num_samples": S,
            "num_burnin_steps": B})


def run_hmc(FLAGS, model_config, results_dir, file_path, tuning=False):
    """main.py:296-398."""
    grid = _parse_leapfrog_steps(FLAGS)
    if tuning and len(grid) > 1:
        return run_hmc_tuning_grid(FLAGS, model_config, results_dir, file_path, grid)
    FLAGS.num_leapfrog_steps = grid[0] if grid else None
    if os.path.exists(file_path):
        with open(file_path) as f:
            prev_results = json.load(f)
    else:
        raise Exception("Run VI first to find initial step sizes")
    param_names = model_config.param_names
    rank, world = distributed.rank_world()
    initial_step_size = prev_results["initial_step_size"]
    rng = np.random.default_rng(FLAGS.seed + 1)
    initial_states = list(util.variational_inits_from_params(
        prev_results["learned_variational_params"], param_names=param_names, num_inits=FLAGS.num_chains,
        rng=rng).values())
    if tuning:
        if not FLAGS.num_leapfrog_steps:
            raise ValueError("You must specify the number of leapfrog steps for a tuning run.")
        for existing_run in prev_results.get("tuning_runs", []):
            if existing_run["num_leapfrog_steps"] == FLAGS.num_leapfrog_steps:
                print("A tuning run already exists for HMC with {} leapfrog steps skipping. ({})".format(
                    FLAGS.num_leapfrog_steps, existing_run))
                return
    if not FLAGS.num_leapfrog_steps:
        FLAGS.num_leapfrog_steps = get_best_num_leapfrog_steps_from_tuning_runs(prev_results["tuning_runs"])
    util.print("\nNumber of leaprog steps is set to {}.\n".format(FLAGS.num_leapfrog_steps))
    if FLAGS.count_in_leapfrog_steps:
        FLAGS.num_samples = int(FLAGS.num_samples / float(FLAGS.num_leapfrog_steps))
        FLAGS.num_burnin_steps = int(FLAGS.num_burnin_steps / float(FLAGS.num_leapfrog_steps))
        FLAGS.num_adaptation_steps = int(FLAGS.num_adaptation_steps / float(FLAGS.num_leapfrog_steps))

    target, actual_reparam = create_target_graph(FLAGS, model_config, results_dir)
    _check_sharding(FLAGS, world)
    lo, hi = distributed.shard_range(FLAGS.num_chains, rank, world)
    z0 = model_config.join(initial_states)[lo:hi]
    device = distributed.local_device()
    start_time = time.time()
    # traces are only kept for the first --num_chains_to_save chains (main.py:578-585): they live on rank 0
    n_save = min(FLAGS.num_chains_to_save, hi - lo) if rank == 0 else 0
    if rank == 0 and n_save < FLAGS.num_chains_to_save:
        util.print("note: --num_chains_to_save={} exceeds rank 0's shard of {} chains; saving {}".format(
            FLAGS.num_chains_to_save, hi - lo, n_save))
    res = inference.hmc(target, model_config, initial_step_size, z0, reparam=actual_reparam,
                        num_leapfrog_steps=FLAGS.num_leapfrog_steps, num_samples=FLAGS.num_samples,
                        num_burnin_steps=FLAGS.num_burnin_steps, num_adaptation_steps=FLAGS.num_adaptation_steps,
                        num_chains_to_save=n_save, seed=FLAGS.seed, chain_offset=lo, device=device,
                        precision=FLAGS.precision, return_is_accepted=False, stream_window=FLAGS.stream_window)
    ess_flat = distributed.gather_chains(res.ess_flat, device)            # [C, D] over all ranks
    n_accepted = res.accept_stats[0]     # accepted kept transitions of ALL ranks (all-reduced inside inference.hmc)
    samples = res.samples
    mcmc_time = time.time() - start_time
    if rank != 0:
        return
    ess_final = model_config.split(ess_flat)
    # report effective samples per 1000 gradient evals (main.py:362-366)
    normalized_ess_final = [1000 * e / (FLAGS.num_samples * FLAGS.num_leapfrog_steps) for e in ess_final]
    ess_min, sem_min = util.get_min_ess(normalized_ess_final, FLAGS.num_chains)
    util.print("ESS per 1000 gradients: {} +/- {}".format(ess_min, sem_min))
    acceptance_rate = n_accepted * 100. / float(FLAGS.num_samples * FLAGS.num_chains)
    if tuning:
        save_hmc_results(file_path=file_path, tuning_runs={
            "num_leapfrog_steps": FLAGS.num_leapfrog_steps, "ess_min": float(ess_min), "sem_min": float(sem_min),
            "acceptance_rate": float(acceptance_rate), "mcmc_time": mcmc_time, "num_samples": FLAGS.num_samples,
            "num_burnin_steps": FLAGS.num_burnin_steps})
    else:
        # superset of main.py:375-391: analyze.py:44-50 reads num_leapfrog_steps; ESS per second (un-normalised minimum
        # ESS summed over chains / wall time) and the largest R-hat over the chains of ALL ranks are new
        min_ess_raw = np.nan_to_num(ess_flat).min(axis=1)
        save_hmc_results(file_path=file_path, ess_min=float(ess_min), sem_min=float(sem_min),
                         acceptance_rate=float(acceptance_rate), mcmc_time_sec=mcmc_time,
                         num_leapfrog_steps=FLAGS.num_leapfrog_steps,
                         ess_per_sec=float(min_ess_raw.sum() / mcmc_time),
                         rhat_max=None if res.rhat is None else float(np.nanmax(res.rhat)))
        save_ess(file_path_base=file_path[:-5], samples=samples, param_names=param_names,
                 normalized_ess_final=normalized_ess_final, num_chains_to_save=FLAGS.num_chains_to_save)


def _check_sharding(FLAGS, world):
    """Every rank must own at least one chain: an empty shard would fail its kernel launch and leave the other
    ranks waiting in the collectives.  Raised on ALL ranks, before any work."""
    if FLAGS.num_chains < world:
        raise ValueError("--num_chains={} is smaller than the number of ranks ({}): run with fewer GPUs".format(
            FLAGS.num_chains, world))


def _first_existing(results_dir, names):
    for n in names:
        p = os.path.join(results_dir, n)
        if os.path.exists(p):
            return p
    return None


def run_interleaved_hmc(FLAGS, model_config, results_dir, file_path):
    """main.py:452-528.  The reference reads literally ``CP.json`` / ``NCP.json`` (the names VI writes with
    ``--tied_pparams=False``); the default-flag names ``CP_tied.json`` / ``NCP_tied.json`` are accepted too.
    (The published driver then dies on a NameError, ``main.py:515``; this one saves what it evidently meant.)"""
    file_path_cp = _first_existing(results_dir, ["CP.json", "CP_tied.json"])
    file_path_ncp = _first_existing(results_dir, ["NCP.json", "NCP_tied.json"])
    if file_path_cp is None or file_path_ncp is None:
        raise Exception("Run VI first to find initial step sizes, and HMCfirst to find num_leapfrog_steps.")
    param_names = model_config.param_names
    with open(file_path_cp) as f:
        prev = json.load(f)
    initial_step_size_cp = prev["initial_step_size"]
    num_leapfrog_steps_cp = get_best_num_leapfrog_steps_from_tuning_runs(prev["tuning_runs"])
    learned_variational_params_cp = prev["learned_variational_params"]
    with open(file_path_ncp) as f:
        prev = json.load(f)
    initial_step_size_ncp = prev["initial_step_size"]
    num_leapfrog_steps_ncp = get_best_num_leapfrog_steps_from_tuning_runs(prev["tuning_runs"])
    rng = np.random.default_rng(FLAGS.seed + 1)
    initial_states_cp = list(util.variational_inits_from_params(
        learned_variational_params_cp, param_names=param_names, num_inits=FLAGS.num_chains, rng=rng).values())
    (target_cp, target_ncp), _ = create_target_graph(FLAGS, model_config, results_dir)
    rank, world = distributed.rank_world()
    _check_sharding(FLAGS, world)
    lo, hi = distributed.shard_range(FLAGS.num_chains, rank, world)
    device = distributed.local_device()
    x0 = model_config.join(initial_states_cp)[lo:hi]
    best_ess_min, results = 0, None
    for num_ls in sorted(set([num_leapfrog_steps_ncp, num_leapfrog_steps_cp])):
        FLAGS.num_leapfrog_steps = num_ls + num_ls
        util.print("\nNumber of leaprog steps is set to {}.\n".format(FLAGS.num_leapfrog_steps))
        start_time = time.time()
        res = inference.hmc_interleaved(
            model_config, target_cp, target_ncp, num_ls, num_ls, initial_step_size_cp, initial_step_size_ncp, x0,
            num_samples=FLAGS.num_samples, num_burnin_steps=FLAGS.num_burnin_steps,
            num_adaptation_steps=FLAGS.num_adaptation_steps,
            num_chains_to_save=min(FLAGS.num_chains_to_save, hi - lo) if rank == 0 else 0, seed=FLAGS.seed,
            chain_offset=lo, device=device, precision=FLAGS.precision)
        ess_flat = distributed.gather_chains(res.ess_flat, device)
        n_cp = distributed.sum_scalar(float(res.is_accepted_cp.sum()), device)
        n_ncp = distributed.sum_scalar(float(res.is_accepted_ncp.sum()), device)
        mcmc_time = time.time() - start_time
        normalized = [1000 * e / (FLAGS.num_samples * FLAGS.num_leapfrog_steps) for e in model_config.split(ess_flat)]
        ess_min, sem_min = util.get_min_ess(normalized, FLAGS.num_chains)
        util.print("ESS: {} +/- {}".format(ess_min, sem_min))
        denom = float(FLAGS.num_samples * FLAGS.num_chains)
        if results is None or float(ess_min) > best_ess_min:
            best_ess_min = float(ess_min)
            results = (num_ls, ess_min, sem_min, n_cp * 100. / denom, n_ncp * 100. / denom, mcmc_time, res.samples,
                       normalized)
    if rank != 0:
        return
    (best_num_ls, ess_min, sem_min, acc_cp, acc_ncp, mcmc_time, samples, normalized) = results
    FLAGS.num_leapfrog_steps = best_num_ls + best_num_ls
    save_hmc_results(file_path=file_path, initial_step_size_ncp=initial_step_size_ncp,
                     initial_step_size_cp=initial_step_size_cp, num_leapfrog_steps=best_num_ls,
                     ess_min=float(ess_min), sem_min=float(sem_min), acceptance_rate_cp=float(acc_cp),
                     acceptance_rate_ncp=float(acc_ncp), mcmc_time_sec=mcmc_time)
    save_ess(file_path_base=file_path[:-5], samples=samples, param_names=param_names,
             normalized_ess_final=normalized, num_chains_to_save=FLAGS.num_chains_to_save)


def save_hmc_results(file_path, **kwargs):
    """main.py:531-550: read-modify-append."""
    try:
        with open(file_path) as f:
            results = json.load(f)
    except IOError:
        results = {}
    for k in kwargs:
        if k not in results:
            results[k] = []
    for k, v in kwargs.items():
        results.get(k).append(v)
    with open(file_path, "w") as outfile:
        json.dump(results, outfile)


def save_ess(file_path_base, samples, normalized_ess_final, param_names, num_chains_to_save=0):
    """main.py:552-585."""
    dict_ess = dict((param_names[i], np.array(normalized_ess_final[i])) for i in range(len(param_names)))
    with open(file_path_base + "_ess.npz", "wb") as out_f:
        io_buffer = io.BytesIO()
        np.savez(io_buffer, **dict_ess)
        out_f.write(io_buffer.getvalue())
    with open(file_path_base + "_ess.txt", "w") as out_f:
        for k, v in dict_ess.items():
            out_f.write("{}: {}\n\n".format(k, v))
        out_f.write("\n\n")
        for k, v in dict_ess.items():
            out_f.write("{} mean: {}\n".format(k, np.mean(v, axis=0)))
            out_f.write("{} stddev: {}\n\n".format(k, np.std(v, axis=0)))
    if num_chains_to_save > 0 and samples is not None:
        dict_res = dict((param_names[i], samples[i][:, :num_chains_to_save]) for i in range(len(param_names)))
        with open(file_path_base + "_traces.npz", "wb") as out_f:
            io_buffer = io.BytesIO()
            np.savez(io_buffer, **dict_res)
            out_f.write(io_buffer.getvalue())


def check_supported(FLAGS):
    """Flag combinations of the reference driver that are outside the accelerated hot path are rejected up front,
    before any work (README: parity table)."""
    if FLAGS.model in models.OUT_OF_SCOPE_MODELS:
        raise NotImplementedError("model {} is outside the accelerated hot path (GP / MVN / funnel models: a per-step "
                                  "Cholesky / eigh; see DESIGN.md)".format(FLAGS.model))
    if FLAGS.reparameterise_variational:
        raise NotImplementedError("--reparameterise_variational (make_variational_model_special, util.py:235-239) is "
                                  "outside the accelerated hot path")


def main(argv=None):
    """main.py:190-231."""
    FLAGS = build_parser().parse_args(argv)
    check_supported(FLAGS)
    util.print("Loading model {} with dataset {}.".format(FLAGS.model, FLAGS.dataset))
    model_config = models.get_model_by_name(FLAGS.model, dataset=FLAGS.dataset, data_dir=FLAGS.data_dir)
    results_dir = FLAGS.results_dir if FLAGS.results_dir != "" else FLAGS.model + "_" + FLAGS.dataset
    distributed.init_if_needed()
    rank, _ = distributed.rank_world()
    if rank == 0:
        os.makedirs(results_dir, exist_ok=True)
    distributed.barrier()
    file_path = os.path.join(results_dir, results_filename(FLAGS))
    if FLAGS.inference == "VI":
        if rank == 0:
            run_vi(FLAGS, model_config, results_dir, file_path)
    elif FLAGS.inference == "HMC":
        if FLAGS.method == "i":
            run_interleaved_hmc(FLAGS, model_config, results_dir, file_path)
        else:
            run_hmc(FLAGS, model_config, results_dir, file_path, tuning=False)
    elif FLAGS.inference == "HMCtuning":
        run_hmc(FLAGS, model_config, results_dir, file_path, tuning=True)
    else:
        raise Exception("unknown inference {}".format(FLAGS.inference))
    distributed.barrier()
    distributed.shutdown()


if __name__ == "__main__":
    main()
