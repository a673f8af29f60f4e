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
    def __init__(self, model_config, method, a, b, learnable, a_index=None, b_index=None, params=()):
        self.model_config = model_config
        self.method = method
        self.a = np.asarray(a, dtype=np.float64)
        self.b = np.asarray(b, dtype=np.float64)
        self.learnable = learnable  # True: the reparameterisation is optimised by VI (cVIP)
        # learnable parameters: slot of every coordinate's a / b (-1 = fixed at self.a / self.b) and, per slot range,
        # the key under which the reference stores it in learned_reparam: [(key, first_slot, shape), ...]
        self.a_index = None if a_index is None else np.asarray(a_index, dtype=np.int32)
        self.b_index = None if b_index is None else np.asarray(b_index, dtype=np.int32)
        self.params = list(params)
        self.num_params = sum(int(np.prod(shape)) if len(shape) else 1 for _, _, shape in self.params)


def make_cp_graph(model_config):
    d = model_config.num_coords
    return TargetGraph(model_config, "CP", np.ones(d), np.ones(d), False)


def make_ncp_graph(model_config):
    d = model_config.num_coords
    return TargetGraph(model_config, "NCP", np.zeros(d), np.zeros(d), False)


def make_cvip_graph(model_config, parameterisation_type="exp", tied_pparams=False, tied_b_as_written=True):
    """Learnable reparameterisation, every parameter initialised at sigmoid(0) = 0.5
    (program_transformations.py:486-533).

    tied_pparams=True, tied_b_as_written=True   one `a` per coordinate, b = 1: what the reference's tied mode does
                                                as written
    tied_pparams=True, tied_b_as_written=False  one `a` per coordinate and b = a (the paper's intent)
    tied_pparams=False                          `a` with the shape of the site's loc, an independent `b` with the
                                                shape of its scale (:517-523: recenter passes 'scalar', so `_b` is
                                                always created); a scalar loc / scale shares one parameter
    Non-Normal sites (german_credit_gammascale's beta_log_scales) are never reparameterised."""
    mc = model_config
    d = mc.num_coords
    a, b = np.ones(d), np.ones(d)
    ia, ib = np.full(d, -1, dtype=np.int32), np.full(d, -1, dtype=np.int32)
    params, slot = [], 0
    for name, shape in mc.sites:
        if name in mc.non_normal:
            continue
        o, size = mc.offsets[name]
        if tied_pparams:
            ia[o:o + size] = slot + np.arange(size)
            if not tied_b_as_written:
                ib[o:o + size] = slot + np.arange(size)
            params.append((name + "_a", slot, shape))
            slot += size
        else:
            if name in mc.scalar_loc:
                ia[o:o + size] = slot
                params.append((name + "_a", slot, ()))
                slot += 1
            else:
                ia[o:o + size] = slot + np.arange(size)
                params.append((name + "_a", slot, shape))
                slot += size
            if name in mc.scalar_scale:
                ib[o:o + size] = slot
                params.append((name + "_b", slot, ()))
                slot += 1
            else:
                ib[o:o + size] = slot + np.arange(size)
                params.append((name + "_b", slot, shape))
                slot += size
        a[o:o + size] = 0.5
        if ib[o] >= 0:
            b[o:o + size] = 0.5
    g = TargetGraph(mc, "cVIP", a, b, True, ia, ib, params)
    g.tie_b = bool(tied_pparams and not tied_b_as_written)
    return g


def learned_reparam_from_params(target, values):
    """Parameter values [P] of a cVIP fit -> the reference's ``learned_reparam`` dict ({site}_a [, {site}_b]).
    With the paper's tie (b = a) `_b` is stored too, so that a later HMC / dVIP run rebuilds the same (a, b)."""
    out = collections.OrderedDict()
    values = np.asarray(values, dtype=np.float64)
    for key, slot, shape in target.params:
        size = int(np.prod(shape)) if len(shape) else 1
        out[key] = np.asarray(values[slot:slot + size].reshape(shape), dtype=np.float32)
        if getattr(target, "tie_b", False) and key.endswith("_a"):
            out[key[:-2] + "_b"] = out[key].copy()
    return out


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
