"""Shared helpers for the tests: fixtures -> raw data -> ModelConfig, parameterisations."""
import os

import numpy as np

from autoreparam_b200 import data as arp_data
from autoreparam_b200 import models as arp_models

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

MODELS = ["8schools", "german_credit_lognormalcentered", "german_credit_gammascale", "radon", "radon_stddvs",
          "election", "electric", "time_series"]


def _npz(name):
    with np.load(os.path.join(GOLDEN, "data_%s.npz" % name)) as f:
        return {k: f[k] for k in f.files}


def raw_data(model, dataset="PA"):
    """Raw arrays in the loaders' format, from the committed fixtures."""
    if model == "8schools":
        return arp_data.eight_schools()
    if model == "time_series":
        return arp_data.time_series()
    if model.startswith("german_credit"):
        d = _npz("german_credit")
        return {"X": d["X"].astype(np.float32), "y": d["y"].astype(np.float32)}
    if model in ("radon", "radon_stddvs"):
        return _npz("radon_" + dataset)
    if model == "election":
        d = _npz("election")
        return {"n_state": int(d["n_state"]), "state": d["state"].astype(np.int32),
                "female": d["female"].astype(np.float32), "black": d["black"].astype(np.float32),
                "y": d["y"].astype(np.float32)}
    if model == "electric":
        d = _npz("electric")
        out = {k: d[k] for k in d}
        for k in ("n_pair", "n_grade", "n_grade_pair"):
            out[k] = int(out[k])
        return out
    if model == "german_synth":
        return arp_data.synthetic_german_credit()
    raise KeyError(model)


def model_config(model, dataset="PA"):
    name = "german_credit_lognormalcentered" if model == "german_synth" else model
    return arp_models.from_data(name, raw_data(model, dataset))


def ab_for(method, D, seed=0):
    """(a, b) [D] for a named parameterisation used across the tests."""
    rng = np.random.default_rng(1000 + seed)
    if method == "CP":
        return np.ones(D), np.ones(D)
    if method == "NCP":
        return np.zeros(D), np.zeros(D)
    if method == "VIP_a":      # as-written tied VIP: learned a, b = 1
        return rng.uniform(0.05, 0.95, D), np.ones(D)
    if method == "VIP_ab":     # general (a, b)
        return rng.uniform(0.05, 0.95, D), rng.uniform(0.05, 0.95, D)
    if method == "dVIP":       # thresholded
        return (rng.uniform(0, 1, D) >= 0.5).astype(float), np.ones(D)
    raise KeyError(method)


def random_states(model, D, C, seed=0, scale=0.5):
    """Moderate random states (all in-scope models are well behaved near 0)."""
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((C, D)) * scale
    if model == "time_series":
        # keep the random walk near the data so residual / 0.12 stays O(1e3), not O(1e4)
        z *= 0.2
    return z


def rel_err(x, ref):
    """max-norm relative error per row (chain): max_d |x - ref| / max(max_d |ref|, 1)."""
    x, ref = np.asarray(x, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    if x.ndim == 1:
        return np.abs(x - ref) / np.maximum(np.abs(ref), 1.0)
    num = np.abs(x - ref).reshape(x.shape[0], -1).max(axis=1)
    den = np.maximum(np.abs(ref).reshape(ref.shape[0], -1).max(axis=1), 1.0)
    return num / den
