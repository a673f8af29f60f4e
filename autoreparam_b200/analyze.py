"""Reader for the results directories ``autoreparam_b200.main`` (and the reference's ``main.py``) write:
the host-side mirror of the reference's ``analyze.py``.

    python -m autoreparam_b200.analyze --results_dir=RESULTS --elbos
    python -m autoreparam_b200.analyze --results_dir=RESULTS --ess [--normalize_times] [--model=radon_PA]
    python -m autoreparam_b200.analyze --results_dir=RESULTS --reparams
    python -m autoreparam_b200.analyze --results_dir=RESULTS --validate

Same flags and the same printed lines as ``analyze.py:10-15, 102-182``.  Two things differ, on purpose:

* **File names.**  The reference reader opens ``<results_dir>/<model>/<m>.json`` for ``m`` in ``CP, NCP, cVIP_exp,
  cVIP_exp_tied, i`` (``analyze.py:19-26, 38-40``) but the reference writer names files
  ``<method>[_<lpt>][_tied][...].json`` (``main.py:208-219``): with default flags ``CP_tied.json``,
  ``cVIP_eig_tied.json`` ... -- the published reader cannot find what the published writer wrote.  Here every logical
  method is resolved through a list of candidates (``candidates()``): the reader's literal name first, then the
  writer's names for the default and non-default flag combinations.
* ``--validate`` (new): checks every JSON of a model directory for the keys the reader needs and for equal list
  lengths of the appended HMC fields; ``--ess`` also prints ESS per second and R-hat when the file carries them.

Nothing here touches the GPU.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import sys

import numpy as np

METHODS = ["CP", "NCP", "cVIP_exp", "cVIP_exp_tied"]      # analyze.py:19-24
CVIP_METHODS = ["cVIP_exp_tied"]                           # analyze.py:26
MODEL_NAMES = [                                             # analyze.py:95-102
    "8schools_data", "german_credit_lognormalcentered_data", "radon_MA", "radon_AZ", "radon_IN", "radon_MO",
    "radon_MN", "radon_PA", "radon_ND", "radon_stddvs_MA", "radon_stddvs_AZ", "radon_stddvs_IN", "radon_stddvs_MO",
    "radon_stddvs_MN", "radon_stddvs_PA", "radon_stddvs_ND", "election_data", "electric_data", "time_series_data"]

VI_KEYS = ["elbo", "estimated_elbo_std", "variational_fit_time_secs", "learning_rate", "initial_step_size",
           "learned_variational_params"]                    # main.py:277-290
HMC_LIST_KEYS = ["ess_min", "sem_min", "acceptance_rate", "mcmc_time_sec"]   # main.py:375-391


def candidates(method):
    """File stems tried, in order, for one of the reader's logical methods."""
    if method in ("CP", "NCP", "i"):
        return [method, method + "_tied"]
    if method.startswith("cVIP"):
        tied = method.endswith("_tied")
        lpt = method[len("cVIP_"):-len("_tied")] if tied else method[len("cVIP_"):]
        stems = [method]
        if tied:
            stems += ["cVIP_%s_tied" % lpt, "cVIP_*_tied"]
        else:
            stems += ["cVIP_%s" % lpt, "cVIP_*[!d]"]          # any lpt, not ending in "_tied"
        return stems
    return [method]


def find_result_file(results_dir, model_name, method):
    base = os.path.join(results_dir, model_name)
    for stem in candidates(method):
        if "*" in stem:
            hits = sorted(h for h in glob.glob(os.path.join(base, stem + ".json"))
                          if not os.path.basename(h).startswith("dVIP"))
            if hits:
                return hits[0]
        else:
            path = os.path.join(base, stem + ".json")
            if os.path.exists(path):
                return path
    raise FileNotFoundError("no results file for method %s under %s (tried %s)" %
                            (method, base, ", ".join(s + ".json" for s in candidates(method))))


def load(results_dir, model_name, method):
    with open(find_result_file(results_dir, model_name, method)) as f:
        return json.load(f)


def get_ess(results_dir, model_name, log=print):
    """analyze.py:30-57: per method the appended lists ess_min / sem_min, the leapfrog step count, VI and MCMC times."""
    ess, sem, leapfrog_steps, vi_times, mcmc_times, extra = {}, {}, {}, {}, {}, {}
    for m in METHODS + ["i"]:
        try:
            r = load(results_dir, model_name, m)
            e, s = r["ess_min"], r["sem_min"]
            lf = r["num_leapfrog_steps"] if "num_leapfrog_steps" in r else r["num_leapfrog_steps_cp"]
            if isinstance(lf, list):       # autoreparam_b200.main appends it like every other HMC field
                lf = lf[0]
            mt = r["mcmc_time_sec"]
        except Exception as exc:   # the reference prints and carries on
            log(exc)
            continue
        ess[m], sem[m], leapfrog_steps[m], mcmc_times[m] = e, s, lf, mt
        vi_times[m] = r.get("variational_fit_time_secs", None)
        extra[m] = {k: r[k] for k in ("rhat_max", "ess_per_sec", "acceptance_rate") if k in r}
    return ess, sem, leapfrog_steps, vi_times, mcmc_times, extra


def get_elbos(results_dir, model_name, log=print):
    """analyze.py:60-74, except that a missing method is reported and skipped (the reference drops the whole model)."""
    elbos, stds = {}, {}
    for m in METHODS:
        try:
            r = load(results_dir, model_name, m)
            elbos[m], stds[m] = r["elbo"], r["estimated_elbo_std"]
        except Exception as exc:
            log(exc)
    return elbos, stds


def get_reparam(results_dir, model_name):
    out = {}
    for m in CVIP_METHODS:
        reparam = load(results_dir, model_name, m)["learned_reparam"]
        out[m] = {k: np.array(v).astype(np.float32) for k, v in reparam.items()}
    return out


def validate(results_dir, model_name):
    """Problems found in one model directory (empty list = the reader and ``main.py``'s HMC stage will work)."""
    problems = []
    base = os.path.join(results_dir, model_name)
    files = sorted(glob.glob(os.path.join(base, "*.json")))
    if not files:
        return ["%s: no results files" % base]
    for path in files:
        name = os.path.basename(path)
        try:
            with open(path) as f:
                r = json.load(f)
        except Exception as exc:
            problems.append("%s: unreadable (%s)" % (name, exc))
            continue
        if not name.startswith("i"):
            missing = [k for k in VI_KEYS if k not in r]
            if missing:
                problems.append("%s: VI keys missing: %s" % (name, ", ".join(missing)))
        if "VIP" in name and "learned_reparam" not in r:
            problems.append("%s: learned_reparam missing" % name)
        need = [k for k in HMC_LIST_KEYS if not (name.startswith("i") and k == "acceptance_rate")]   # i: _cp / _ncp
        present = [k for k in need if k in r]
        if present:
            if len(present) != len(need):
                problems.append("%s: HMC keys missing: %s" % (name, ", ".join(sorted(set(need) - set(present)))))
            lengths = {k: len(r[k]) if isinstance(r[k], list) else -1 for k in present}
            if len(set(lengths.values())) != 1 or -1 in lengths.values():
                problems.append("%s: appended HMC lists differ in length: %s" % (name, lengths))
            if "num_leapfrog_steps" not in r and "num_leapfrog_steps_cp" not in r:
                problems.append("%s: num_leapfrog_steps missing (the reference writer never stores it for "
                                "non-interleaved runs; autoreparam_b200.main does)" % name)
    return problems


def parse(argv=None):
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    boolean = lambda v: str(v).lower() in ("1", "true", "yes", "")
    for flag in ("elbos", "ess", "reparams", "normalize_times", "validate"):   # absl-style booleans
        p.add_argument("--" + flag, nargs="?", const=True, default=False, type=boolean)
    p.add_argument("--model", default="all")
    p.add_argument("--results_dir", default="")
    return p.parse_args(argv)


def main(argv=None, log=print):
    F = parse(argv)
    names = MODEL_NAMES if F.model == "all" else [F.model]
    rc = 0
    if F.elbos:                                                     # analyze.py:107-121
        for name in names:
            log(" ******  {}  ****** ".format(name))
            try:
                e, s = get_elbos(F.results_dir, name, log)
            except Exception as exc:
                log(exc)
                continue
            for key in e.keys():
                log("{0:.4f} +/- {1:.2f}   : {2}".format(e[key], s[key], key))
            log("\n\n")
    if F.reparams:                                                  # analyze.py:123-138
        for name in names:
            log(" ******  {}  ****** ".format(name))
            try:
                reparam = get_reparam(F.results_dir, name)
            except Exception as exc:
                log(exc)
                continue
            for m in CVIP_METHODS:
                log("   {}".format(m))
                for k, v in reparam[m].items():
                    log("{:>10}: {}".format(k, v))
                log("\n")
    if F.ess:                                                       # analyze.py:140-182
        for name in names:
            log(" ******  {}  ****** ".format(name))
            try:
                ess, sem, leapfrog_steps, vi_times, mcmc_times, extra = get_ess(F.results_dir, name, log)
            except Exception as exc:
                log(exc)
                continue
            for key in ess.keys():
                if F.normalize_times:
                    mcmc_time, my_ess, my_sem = mcmc_times[key][0], ess[key][0], sem[key][0]
                    if key == "i":
                        leapfrog_steps_per_sample = leapfrog_steps[key] * 2
                        vi_time = vi_times["CP"] + vi_times["NCP"]
                    else:
                        leapfrog_steps_per_sample = leapfrog_steps[key]
                        vi_time = vi_times[key]
                    total_time = vi_time + mcmc_time
                    total_grad_evals = leapfrog_steps_per_sample * 10000.
                    total_effective_samples = my_ess * total_grad_evals / 1000.
                    sampling_stderr = my_sem * total_grad_evals / 1000.
                    time_per_variational_step = vi_time / 3000.
                    time_per_variational_step_cp = vi_times["CP"] / 3000.
                    time_per_step = mcmc_time / (10000 * (2 if key == "i" else 1) * leapfrog_steps[key])
                    time_per_step_cp = mcmc_times["CP"][0] / (10000 * leapfrog_steps["CP"])
                    log("{} +/- {} in {}s ({}s VI + {}s MCMC): {} ({} leapfrog steps, {:.2f}x/{:.2f}x CP time per "
                        "VI/MCMC step)".format(total_effective_samples, sampling_stderr, total_time, vi_time, mcmc_time,
                                               key, leapfrog_steps[key],
                                               time_per_variational_step / time_per_variational_step_cp,
                                               time_per_step / time_per_step_cp))
                else:
                    log("{} +/- {} : {} ({} leapfrog steps)".format(ess[key], sem[key], key, leapfrog_steps[key]))
                if extra.get(key):
                    log("      " + ", ".join("{} {}".format(k, v) for k, v in sorted(extra[key].items())))
            log("\n\n")
    if F.validate:
        for name in names:
            base = os.path.join(F.results_dir, name)
            if F.model == "all" and not os.path.isdir(base):
                continue
            problems = validate(F.results_dir, name)
            log(" ******  {}  ****** {}".format(name, "ok" if not problems else ""))
            for pr in problems:
                log("  " + pr)
                rc = 1
    return rc


if __name__ == "__main__":
    sys.exit(main())
