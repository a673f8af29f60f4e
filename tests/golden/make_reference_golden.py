"""Golden vectors from the reference's OWN model / interceptor code.

Executes, unmodified and in place, /root/reference/{models.py, program_transformations.py}
on top of the torch-float64 TF/TFP/Edward2 stand-in in tests/golden/_tfshim, and freezes
into tests/golden/reference_logjoint.npz, per (model, rule):

    z        [C, D]  state coordinates (trace order)
    a, b     [D]     rule parameters per coordinate
    lp       [C]     target(*z) as graphs.py:37-44 / 84-91 / 197-203 build it
                     (make_log_joint_fn over the model wrapped in ncp / recenter)
    centered [C, D]  make_to_centered(**reparam)(z)  (models.py:59-81)
    noncentered [C, D] to_noncentered(centered)      (models.py:84-102), NCP only

Run here only (the GPU box has no /root/reference):
    python tests/golden/make_reference_golden.py
The TF/TFP arithmetic itself (Normal / Bernoulli log-probs, one_hot, the interceptor
stack) is restated by the stand-in; the model bodies, site transformations, argument
wiring and every index quirk come from the reference's code as it runs.
"""
import collections
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("ARP_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "_tfshim"))
os.chdir(REF)  # models.py reads ./data/

import pandas as pd  # noqa: E402
import torch  # noqa: E402

# pandas 3 compatibility for models.py:863,874 (delim_whitespace / object dtype)
pd.set_option("future.infer_string", False)
_read_csv = pd.read_csv


def read_csv(f, *args, **kw):
    if kw.pop("delim_whitespace", False):
        kw["sep"] = r"\s+"
    return _read_csv(f, *args, **kw)


pd.read_csv = read_csv
_replace = pd.Series.replace


def replace(self, to_replace=None, *a, **k):   # models.py:736 str -> int replace via dict
    if isinstance(to_replace, dict) and not a and not k:
        return self.map(lambda v: to_replace.get(v, v))
    return _replace(self, to_replace, *a, **k)


pd.Series.replace = replace

import tensorflow.compat.v1 as tf  # noqa: E402  (the stand-in)
tf.app.flags.FLAGS.learnable_parameterisation_type = "eig"   # main.py:49-52 default
tf.app.flags.FLAGS.num_chains = 4

import models  # noqa: E402  (reference)
import program_transformations as ed_transforms  # noqa: E402  (reference)
from tensorflow_probability import edward2 as ed  # noqa: E402

MODELS = [("8schools", None), ("german_credit_lognormalcentered", None), ("german_credit_gammascale", None),
          ("radon", "PA"), ("radon_stddvs", "PA"), ("election", None), ("electric", None), ("time_series", None)]
C = 3


def trace_names(model_config):
    with ed.tape() as model_tape:
        model_config.model(*model_config.model_args)
    return [(k, tuple(v.shape)) for k, v in model_tape.items() if k not in model_config.observed_data]


def make_target(model_config, interceptor):
    """graphs.py:57-91 / 164-203 without the ELBO: log joint of the (wrapped) model, latents positional."""
    def wrapped(*params):
        if interceptor is None:
            return model_config.model(*params)
        with ed.interception(interceptor):
            return model_config.model(*params)

    log_joint = ed_transforms.make_log_joint_fn(wrapped)
    with ed.tape() as model_tape:
        wrapped(*model_config.model_args)
    names = list(model_tape.keys())

    def target(*param_args):
        kwargs, i = {}, 0
        for name in names:
            if name in model_config.observed_data:
                kwargs[name] = model_config.observed_data[name]
            else:
                kwargs[name] = param_args[i]
                i += 1
        return log_joint(*model_config.model_args, **kwargs)
    return target, [n for n in names if n not in model_config.observed_data]


def split(z, sites):
    parts, o = [], 0
    for _, shape in sites:
        n = int(np.prod(shape)) if len(shape) else 1
        parts.append(torch.as_tensor(z[o:o + n].reshape(shape), dtype=torch.float64))
        o += n
    return parts


def flat(parts):
    return np.concatenate([np.asarray(p.detach() if torch.is_tensor(p) else p, dtype=np.float64).reshape(-1)
                           for p in parts])


out = collections.OrderedDict()
notes = []
for mname, dataset in MODELS:
    mc = models.get_model_by_name(mname, dataset=dataset)
    sites = trace_names(mc)
    D = sum(int(np.prod(s)) if len(s) else 1 for _, s in sites)
    rng = np.random.default_rng(abs(hash(mname)) % (2 ** 31))
    rng = np.random.default_rng(sum(map(ord, mname)))
    scale = 0.1 if mname == "time_series" else 0.5
    Z = (scale * rng.standard_normal((C, D))).astype(np.float32).astype(np.float64)   # float32-representable inputs
    rules = collections.OrderedDict()
    rules["CP"] = (None, np.ones(D), np.ones(D), None)
    # reparam dicts as main.py hands them to make_dvip_graph / make_to_centered
    ra = collections.OrderedDict((n + "_a", rng.uniform(0.05, 0.95, s).astype(np.float32).astype(np.float64)) for n, s in sites)
    rab = collections.OrderedDict(ra)
    rab.update((n + "_b", rng.uniform(0.05, 0.95, s).astype(np.float32).astype(np.float64)) for n, s in sites)
    rd = collections.OrderedDict((n + "_a", (rng.uniform(0, 1, s) >= 0.5).astype(np.float64)) for n, s in sites)
    rules["NCP"] = (ed_transforms.ncp, np.zeros(D), np.zeros(D), None)
    for key, rp in (("VIP_a", ra), ("VIP_ab", rab), ("dVIP", rd)):
        _, interceptor, _ = ed_transforms.make_learnable_parametrisation(learnable_parameters=dict(rp))
        a = flat([np.broadcast_to(rp[n + "_a"], s) for n, s in sites])
        b = flat([np.broadcast_to(rp.get(n + "_b", 1.0), s) for n, s in sites])
        rules[key] = (interceptor, a, b, rp)
    for rule, (interceptor, a, b, rp) in rules.items():
        try:
            target, latent_names = make_target(mc, interceptor)
            assert latent_names == [n for n, _ in sites] or rule == "NCP", (latent_names, sites)
            lp = np.array([float(target(*split(Z[c], sites))) for c in range(C)])
            if rule == "CP":
                cen = Z.copy()
            elif rule == "NCP":
                cen = np.array([flat(mc.to_centered(split(Z[c], sites))) for c in range(C)])
            else:
                to_c = mc.make_to_centered(**dict(rp))
                cen = np.array([flat(to_c(split(Z[c], sites))) for c in range(C)])
            key = "%s/%s" % (mname, rule)
            out[key + "/z"], out[key + "/a"], out[key + "/b"] = Z, a, b
            out[key + "/lp"], out[key + "/centered"] = lp, cen
            if rule == "NCP":
                out[key + "/noncentered"] = np.array([flat(mc.to_noncentered(split(cen[c], sites))) for c in range(C)])
            print("%-34s %-7s D=%3d lp[0]=% .10f" % (mname, rule, D, lp[0]))
        except Exception as e:  # recorded: the reference itself cannot run this combination
            notes.append("%s/%s: reference raises %s: %s" % (mname, rule, type(e).__name__, e))
            print("%-34s %-7s REFERENCE FAILS: %s: %s" % (mname, rule, type(e).__name__, e))

# tied-parameter quirk: create the cVIP variables with tied_pparams=True,
# then evaluate the target the way graphs.make_cvip_graph does (a second trace).
mc = models.get_model_by_name("8schools")
sites = trace_names(mc)
learnable, interceptor, _ = ed_transforms.make_learnable_parametrisation(tau=1., parameterisation_type="eig",
                                                                         tied_pparams=True)
target, _ = make_target(mc, interceptor)   # first trace creates a = sigmoid(0) = 0.5 (tied b = a only here)
z = np.concatenate([[0.7, -0.3], np.linspace(-1, 1, 8)])
out["8schools/cVIP_tied_as_written/z"] = z[None]
out["8schools/cVIP_tied_as_written/lp"] = np.array([float(target(*split(z, sites)))])
out["8schools/cVIP_tied_as_written/keys"] = np.array(sorted(learnable.keys()))
print("tied cVIP second-trace lp", out["8schools/cVIP_tied_as_written/lp"], sorted(learnable.keys()))
out["notes"] = np.array(notes)
np.savez_compressed(os.path.join(HERE, "reference_logjoint.npz"), **out)
print("wrote", os.path.join(HERE, "reference_logjoint.npz"), len(out), "arrays;", len(notes), "notes")
