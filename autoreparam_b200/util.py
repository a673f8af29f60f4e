"""Host-side helpers mirroring the reference's ``util.py`` (the parts main.py uses)."""
from __future__ import annotations

import builtins
import collections
import logging

import numpy as np


def print(*args):  # pylint: disable=redefined-builtin
    """util.print (util.py:413-415): stdout + log."""
    builtins.print(*args)
    logging.info(" ".join(str(a) for a in args))


def get_approximate_step_size(variational_parameters, num_leapfrog_steps):
    """util.py:271-276: every ``*_scale`` entry divided by num_leapfrog_steps**2."""
    return [np.asarray(variational_parameters[k]) / num_leapfrog_steps ** 2
            for k in variational_parameters.keys() if k.endswith("_scale")]


def variational_inits_from_params(learned_variational_params, param_names, num_inits, rng=None):
    """util.py:394-410: z0 = loc + scale * randn, shape (num_inits,) + site shape.
    The reference uses the unseeded global numpy RNG; pass ``rng`` for reproducibility."""
    rng = np.random if rng is None else rng
    locs, stddevs, samples = collections.OrderedDict(), collections.OrderedDict(), collections.OrderedDict()
    for k, v in learned_variational_params.items():
        if k.endswith("_loc"):
            locs[k[:-4]] = v
        elif k.endswith("_scale"):
            stddevs[k[:-6]] = v
    for k in param_names:
        shape = (num_inits,) + np.asarray(locs[k]).shape
        noise = rng.standard_normal(shape) if hasattr(rng, "standard_normal") else rng.randn(*shape)
        samples[k] = (noise * stddevs[k] + locs[k]).astype(np.float32)
    return samples


def get_min_ess(ess, num_chains):
    """util.py:445-460: nan->0, per-chain minimum over every coordinate of every
    site, then mean and std/sqrt(n) over chains.  ``ess``: list of [C, *site]."""
    ess = [np.nan_to_num(e) for e in ess]
    min_ess = []
    for c in range(num_chains):
        min_ess.append(min(np.array(e[c]).min() for e in ess))
    mean_ess = np.mean(min_ess)
    sem_ess = np.std(min_ess) / np.sqrt(len(min_ess))
    return mean_ess, sem_ess


def rhat_from_moments(chain_mean, chain_var, num_samples):
    """Potential scale reduction (new capability -- the reference has none).
    chain_mean / chain_var: [C, D] per-chain mean and biased variance over S samples."""
    chain_mean = np.asarray(chain_mean, dtype=np.float64)
    chain_var = np.asarray(chain_var, dtype=np.float64)
    s = float(num_samples)
    w = (chain_var * s / (s - 1.0)).mean(axis=0)
    b_over_n = chain_mean.var(axis=0, ddof=1)
    return np.sqrt(((s - 1.0) / s * w + b_over_n) / w)
