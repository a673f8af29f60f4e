"""Method -> target: host-side mirror of the reference's ``graphs.py``.

The reference builds, per ``--method``, a TF closure ``target(*params)`` plus a
mean-field ELBO (``graphs.py:14-213``).  Here a target is just the model handle
plus the per-coordinate parameters ``(a, b)`` of the site rule
(``program_transformations.py:555-600``) that the CUDA kernels take as kernel
parameters:

    CP    a = b = 1                    (make_cp_graph,   graphs.py:14-54)
    NCP   a = b = 0                    (make_ncp_graph,  graphs.py:57-101; ncp :262-279)
    cVIP  a = sigmoid(u) learnable     (make_cvip_graph, graphs.py:104-161)
    dVIP  a = 1[a_cVIP >= 0.5]         (make_dvip_graph, graphs.py:164-213; main.py:170-172)

``b``: with ``--tied_pparams`` (the default) the reference returns the tied
``b = a`` only from the call that *creates* the variable; every later trace of
the model -- i.e. every graph that is actually optimised or sampled -- gets
``learnable_parameters.get(name + '_b', 1.)`` = 1 (``program_transformations.py:
495-500,512-514``).  ``tied_b_as_written=True``
reproduces that; ``False`` gives the paper's intent ``b = a``.
"""
from __future__ import annotations

import collections

import numpy as np


class TargetGraph:
    def __init__(self, model_config, method, a, b, learnable):
        self.model_config = model_config
        self.method = method
        self.a = np.asarray(a, dtype=np.float64)
        self.b = np.asarray(b, dtype=np.float64)
        self.learnable = learnable  # True: a is optimised by VI (cVIP)


def make_cp_graph(model_config):
    d = model_config.num_coords
    return TargetGraph(model_config, "CP", np.ones(d), np.ones(d), False)


def make_ncp_graph(model_config):
    d = model_config.num_coords
    return TargetGraph(model_config, "NCP", np.zeros(d), np.zeros(d), False)


def make_cvip_graph(model_config, parameterisation_type="exp", tied_pparams=False, tied_b_as_written=True):
    """Learnable a, initialised at sigmoid(0) = 0.5 (program_transformations.py:507-510)."""
    d = model_config.num_coords
    a = np.full(d, 0.5)
    if tied_pparams:
        b = np.ones(d) if tied_b_as_written else a.copy()
    else:
        # untied: an independent learnable b = sigmoid(0) (:517-523); the VI kernel only learns `a`
        raise NotImplementedError("untied VIP (independent learnable b) is not implemented; see DESIGN.md")
    g = TargetGraph(model_config, "cVIP", a, b, True)
    g.tie_b = tied_pparams and not tied_b_as_written
    return g


def reparam_to_ab(model_config, reparam):
    """``learned_reparam`` dict ({site}_a [, {site}_b]) -> flat (a, b).

    Missing ``_b`` -> 1.0, exactly ``learnable_parameters.get(scale_name, 1.)``
    (program_transformations.py:496)."""
    per_a = {n: reparam[n + "_a"] for n, _ in model_config.sites if (n + "_a") in reparam}
    per_b = {n: reparam[n + "_b"] for n, _ in model_config.sites if (n + "_b") in reparam}
    # sites without a learned parameter (non-Normal sites) are never reparameterised
    a = model_config.broadcast_site_params(per_a, 1.0)
    b = model_config.broadcast_site_params(per_b, 1.0)
    return a, b


def make_dvip_graph(model_config, reparam, parameterisation_type="exp"):
    a, b = reparam_to_ab(model_config, reparam)
    return TargetGraph(model_config, "dVIP", a, b, False)


def discretise(reparam):
    """main.py:170-172: threshold every entry of the cVIP result at 0.5."""
    return collections.OrderedDict(
        (k, (np.array(v) >= 0.5).astype(np.float32)) for k, v in reparam.items()
        if not (k.endswith("_prior_mean") or k.endswith("_prior_scale")))
