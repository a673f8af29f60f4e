"""Model zoo: host-side mirror of the reference's ``models.py`` for the in-scope models.

``get_model_by_name(model, dataset)`` has the reference's signature and accepted
names (``models.py:1144-1175``) and returns a ``ModelConfig`` that plays the role
of the reference's namedtuple (``models.py:51-54``): it knows the latent sites in
trace order (= HMC state parts, ``graphs.py:29-44``), owns the raw data, and hands
the data to the CUDA library (``arp_model_create``).  The model *bodies* live in
CUDA (``csrc/arp_models.cuh``); nothing here evaluates a density.
"""
from __future__ import annotations

import collections
import ctypes as C

import numpy as np

from . import _lib, data as _data

IN_SCOPE_MODELS = ("8schools", "radon", "radon_stddvs", "german_credit_lognormalcentered",
                   "german_credit_gammascale", "election", "electric", "time_series")
# reference models.py:1144-1175 dispatches these too; they are a different hot path
OUT_OF_SCOPE_MODELS = ("gp_rhizoc", "gp_classification", "gp_poisson", "gp_poisson_fixed", "police",
                       "neals_funnel")


def _ptr(arr):
    return None if arr is None else arr.ctypes.data_as(C.c_void_p)


class ModelConfig:
    """One model + data set.

    Attributes
      name            model name as accepted by ``--model``
      sites           ``[(site_name, shape), ...]`` latent sites in trace order
      observed_data   ``{'y': ...}`` as in the reference's ModelConfig
      num_coords      D = total number of state coordinates
    """

    def __init__(self, name, raw, sites, observed, c_fields, scalar_loc=(), scalar_scale=(), non_normal=()):
        # learnable-parameter shapes of the site rule (program_transformations.py:563-566): `a` has the shape of the
        # site's loc, `b` of its scale; a site listed in scalar_loc / scalar_scale passes a scalar there although the
        # site itself is a vector.  non_normal sites are never reparameterised (:315-316,782-783).
        self.scalar_loc, self.scalar_scale, self.non_normal = set(scalar_loc), set(scalar_scale), set(non_normal)
        self.name = name
        self.raw = raw
        self.sites = [(n, tuple(s)) for n, s in sites]
        self.observed_data = observed
        self._c_fields = c_fields  # kwargs for arp_model_data, numpy arrays kept alive here
        self._handles = {}
        self.offsets = collections.OrderedDict()
        o = 0
        for n, s in self.sites:
            size = int(np.prod(s)) if len(s) else 1
            self.offsets[n] = (o, size)
            o += size
        self.num_coords = o

    # -- layout helpers: list of [C, *site] parts  <->  flat [C, D] ----------
    @property
    def param_names(self):
        return [n for n, _ in self.sites]

    def join(self, parts):
        """List of per-site arrays ``[C, *site_shape]`` -> ``[C, D]``."""
        parts = [np.asarray(p) for p in parts]
        c = parts[0].shape[0]
        return np.concatenate([p.reshape(c, -1) for p in parts], axis=1)

    def split(self, flat):
        """``[..., D]`` -> list of per-site arrays ``[..., *site_shape]``."""
        flat = np.asarray(flat)
        out = []
        for n, s in self.sites:
            o, size = self.offsets[n]
            out.append(flat[..., o:o + size].reshape(flat.shape[:-1] + s))
        return out

    def broadcast_site_params(self, per_site, default):
        """``{site: scalar or array}`` -> flat [D] (site value broadcast over the site)."""
        out = np.empty(self.num_coords, dtype=np.float64)
        for n, s in self.sites:
            o, size = self.offsets[n]
            v = per_site.get(n, default) if per_site is not None else default
            out[o:o + size] = np.broadcast_to(np.asarray(v, dtype=np.float64), s if len(s) else ()).reshape(-1)
        return out

    # -- CUDA handle ---------------------------------------------------------
    def handle(self, precision="f32"):
        if precision not in self._handles:
            lib = _lib.load(precision)
            md = _lib.ModelData()
            for k, v in self._c_fields.items():
                setattr(md, k, _ptr(v) if isinstance(v, np.ndarray) else int(v))
            h = C.c_void_p()
            _lib.check(lib, lib.arp_model_create(self.name.encode(), C.byref(md), C.byref(h)), "arp_model_create")
            assert lib.arp_model_num_coords(h) == self.num_coords, (lib.arp_model_num_coords(h), self.num_coords)
            self._handles[precision] = h
        return self._handles[precision]

    def close(self):
        for prec, h in self._handles.items():
            _lib.load(prec).arp_model_destroy(h)
        self._handles = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _f32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


def _i32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.int32))


def from_data(model, raw):
    """Build a ModelConfig from already-loaded raw arrays (``autoreparam_b200.data``)."""
    if model == "8schools":
        y, s = _f32(raw["y"]), _f32(raw["sigma"])
        sites = [("mu", ()), ("log_tau", ()), ("theta", (8,))]
        return ModelConfig(model, raw, sites, {"y": y}, dict(n=8, y=y, x1=s))
    if model in ("german_credit_lognormalcentered", "german_credit_gammascale"):
        X, y = _f32(raw["X"]), _f32(raw["y"])
        n, f = X.shape
        sites = [("overall_log_scale", ()), ("beta_log_scales", (f,)), ("beta", (f,))]
        gamma = model == "german_credit_gammascale"
        # models.py:894-901: beta_log_scales ~ N(loc=overall_log_scale [scalar], scale=ones(F)); beta ~ N(zeros(F), exp(.))
        return ModelConfig(model, raw, sites, {"y": y[None, :]}, dict(n=n, f=f, X=X, y=y),
                           scalar_loc=() if gamma else ("beta_log_scales",),
                           non_normal=("beta_log_scales",) if gamma else ())
    if model in ("radon", "radon_stddvs"):
        c, u, x, y = _i32(raw["county"]), _f32(raw["u"]), _f32(raw["x"]), _f32(raw["y"])
        j = len(u)
        sites = [("mua", ()), ("b1", ()), ("b2", ()), ("m", (j,))]
        if model == "radon_stddvs":
            sites.append(("log_m_stddv", (j,)))
        return ModelConfig(model, raw, sites, {"y": y.reshape(-1, 1)}, dict(n=len(y), j=j, idx0=c, u=u, x1=x, y=y))
    if model == "election":
        k = int(raw["n_state"])
        st, fe, bl, y = _i32(raw["state"]), _f32(raw["female"]), _f32(raw["black"]), _f32(raw["y"])
        sites = [("mua", ()), ("log_sigma_a", ()), ("a", (k,)), ("b1", ()), ("b2", ())]
        # models.py:974: a ~ N(loc=mua [scalar], scale=ones(n_state) * exp(log_sigma_a))
        return ModelConfig(model, raw, sites, {"y": y.reshape(-1, 1)},
                           dict(n=len(y), j=k, idx0=st, x1=fe, x2=bl, y=y), scalar_loc=("a",))
    if model == "electric":
        npair, ng, ngp = int(raw["n_pair"]), int(raw["n_grade"]), int(raw["n_grade_pair"])
        pair, grade, gp = _i32(raw["pair"]), _i32(raw["grade"]), _i32(raw["grade_pair"])
        tr, y = _f32(raw["treatment"]), _f32(raw["y"])
        sites = [("mua", (ngp,)), ("sigma_y", (ng,)), ("a", (npair, 1)), ("b", (ng,))]
        # models.py:1021-1028: mua, sigma_y, b have loc=0. (scalar) and vector scales; a has scale=1. (scalar)
        return ModelConfig(model, raw, sites, {"y": y},
                           dict(n=len(y), j=npair, k=ng, k2=ngp, idx0=pair, idx1=grade, idx2=gp, x1=tr, y=y),
                           scalar_loc=("mua", "sigma_y", "b"), scalar_scale=("a",))
    if model == "time_series":
        x, y = _f32(raw["x"]), _f32(raw["y"])
        t = len(x)
        sites = [("sigma_alpha", ()), ("sigma_mu", ()), ("alpha0", ()), ("mu0", ())]
        for i in range(1, t):
            sites += [("alpha%d" % i, ()), ("mu%d" % i, ())]
        sites.append(("beta", ()))
        return ModelConfig(model, raw, sites, {"y": y}, dict(n=t, x1=x, y=y))
    if model in OUT_OF_SCOPE_MODELS:
        raise NotImplementedError(
            "model {} is outside the accelerated hot path (GP / MVN models need a per-step "
            "Cholesky; see DESIGN.md)".format(model))
    raise Exception("unknown model {}".format(model))  # reference models.py:1173-1174


def load_raw(model, dataset=None, data_dir=None):
    if model == "8schools":
        return _data.eight_schools()
    if model in ("german_credit_lognormalcentered", "german_credit_gammascale"):
        return _data.load_german_credit(data_dir)
    if model in ("radon", "radon_stddvs"):
        return _data.load_radon(dataset if dataset else "MN", data_dir)  # default state_code='MN', models.py:763,809
    if model == "election":
        return _data.load_election(data_dir)
    if model == "electric":
        return _data.load_electric(data_dir)
    if model == "time_series":
        return _data.time_series()
    return None


def get_model_by_name(model, dataset=None, data_dir=None):
    """Reference ``models.get_model_by_name`` (``models.py:1144-1175``)."""
    return from_data(model, load_raw(model, dataset, data_dir))
