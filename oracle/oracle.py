"""CPU ORACLE -- test infrastructure only (parity UNPINNED, see below).

This file restates, on the CPU (PyTorch float64/float32 + numpy), the algorithm
of the reference's hot path.  It is the checker the CUDA path is compared
against.  It must never be imported by the product package
``autoreparam_b200``; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may use it.

Parity status: **unpinned by the reference's own tests** -- the reference ships
no golden vectors (its single test file ``models_test.py`` does not parse) and
its arithmetic lives in TensorFlow 1.14 / TensorFlow-Probability 0.7.0
(``README.md:40``), neither of which is installable here.  What pins this
oracle instead:

* ``tests/golden/reference_logjoint.npz`` -- log-joint / centred values
  produced by executing the reference's OWN ``models.py`` /
  ``program_transformations.py`` / ``graphs.py`` model bodies and interceptors
  through a minimal torch-backed stand-in for the TF/TFP/Edward2 API
  (``tests/golden/make_reference_golden.py``);
* the survey's scipy known-answer table;
* closed-form posterior of the linear-Gaussian radon model.

Everything tagged [TFP] restates TensorFlow-Probability 0.7 behaviour from its
published algorithm (hmc.py, metropolis_hastings.py,
dual_averaging_step_size_adaptation.py, sample.py, diagnostic.py,
stats/sample_stats.py); call sites in the reference are cited per function.

Every model body below is written the way the reference writes it (dense
one-hot matmuls, same op order), NOT the way the CUDA kernels compute it
(gather / segmented sums), so the two are independent derivations.
"""
from __future__ import annotations

import math

import numpy as np
import torch

LOG_2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------- #
# Distributions (TFP log_prob formulas)
# --------------------------------------------------------------------------- #
def normal_lp(x, loc, scale):
    """tfd.Normal.log_prob: -0.5 ((x-loc)/scale)^2 - log scale - 0.5 log 2pi."""
    u = (x - loc) / scale
    return -0.5 * u * u - torch.log(scale) - 0.5 * LOG_2PI


def bernoulli_lp(y, logits):
    """tfd.Bernoulli(logits).log_prob(y) = -sigmoid_cross_entropy(labels=y, logits)
    = y*eta - max(eta,0) - log1p(exp(-|eta|))."""
    return y * logits - torch.clamp(logits, min=0.0) - torch.log1p(torch.exp(-torch.abs(logits)))


def one_hot(idx, depth, dtype):
    """tf.one_hot: out-of-range indices give an all-zero row."""
    idx = torch.as_tensor(np.asarray(idx), dtype=torch.int64)
    out = torch.zeros((idx.shape[0], depth), dtype=dtype)
    ok = (idx >= 0) & (idx < depth)
    rows = torch.arange(idx.shape[0])[ok]
    out[rows, idx[ok]] = 1.0
    return out


# --------------------------------------------------------------------------- #
# Site rule  (program_transformations.py:555-600 ; NCP special case :262-279)
# --------------------------------------------------------------------------- #
class Tracer:
    """Plays the role of the interceptor stack for one evaluation of a model.

    ``site(name, loc, scale)`` does what ``recenter`` + the log-joint
    interceptor do for one ``ed.Normal`` site: looks up the state coordinate
    block ``z`` for the site, adds ``N(z; a*loc, scale**b)`` to the log joint
    (``program_transformations.py:569-572,125-127``) and returns the centred
    value ``loc + scale/scale**b * (z - a*loc)`` (``:574-576,600``).
    a=b=1 is CP (identity), a=b=0 is ``ncp`` (``:262-279``).
    """

    def __init__(self, z_parts, a_parts, b_parts, dtype):
        self.z = z_parts
        self.a = a_parts
        self.b = b_parts
        self.dtype = dtype
        self.lp = torch.zeros((), dtype=dtype)
        self.centered = {}
        self.site_lp = {}

    def _t(self, v):
        return torch.as_tensor(v, dtype=self.dtype)

    def site(self, name, loc, scale):
        z = self.z[name]
        a = self._t(self.a[name])
        b = self._t(self.b[name])
        loc = self._t(loc) * torch.ones_like(z)
        scale = self._t(scale) * torch.ones_like(z)
        std_loc = loc * a
        std_scale = torch.pow(scale, b)
        lp = normal_lp(z, std_loc, std_scale).sum()
        self.lp = self.lp + lp
        self.site_lp[name] = lp
        aff_scale = scale / std_scale
        aff_shift = loc - aff_scale * std_loc
        x = aff_shift + aff_scale * z
        self.centered[name] = x
        return x

    def raw(self, name):
        """A latent that is not a Normal site (never reparameterised,
        ``program_transformations.py:315-316,782-783``)."""
        x = self.z[name]
        self.centered[name] = x
        return x

    def add(self, lp):
        self.lp = self.lp + lp


# --------------------------------------------------------------------------- #
# Model bodies (one evaluation, one chain).  `d` = data dict from the loaders.
# --------------------------------------------------------------------------- #
def _m_8schools(tr, d):
    """models.py:139-147."""
    dt = tr.dtype
    sig = torch.as_tensor(d["sigma"], dtype=dt)
    y = torch.as_tensor(d["y"], dtype=dt)
    mu = tr.site("mu", 0.0, 5.0)
    log_tau = tr.site("log_tau", 0.0, 5.0)
    theta = tr.site("theta", mu * torch.ones(8, dtype=dt), torch.exp(log_tau) * torch.ones(8, dtype=dt))
    tr.add(normal_lp(y, theta, sig).sum())


def _m_german_lognormal(tr, d):
    """models.py:888-904."""
    dt = tr.dtype
    X = torch.as_tensor(d["X"], dtype=dt)
    y = torch.as_tensor(d["y"], dtype=dt)
    F = X.shape[1]
    s0 = tr.site("overall_log_scale", 0.0, 10.0)
    s = tr.site("beta_log_scales", s0 * torch.ones(F, dtype=dt), torch.ones(F, dtype=dt))
    beta = tr.site("beta", torch.zeros(F, dtype=dt), torch.exp(s))
    logits = torch.einsum("nd,md->mn", X, beta[None, :])
    tr.add(bernoulli_lp(y[None, :], logits).sum())


def _m_german_gamma(tr, d):
    """models.py:930-945.  beta_log_scales ~ log Gamma(1/2, 1/2): density of
    v = log g, g ~ Gamma(alpha, rate): alpha v - rate e^v + alpha log rate - lgamma(alpha)."""
    dt = tr.dtype
    X = torch.as_tensor(d["X"], dtype=dt)
    y = torch.as_tensor(d["y"], dtype=dt)
    F = X.shape[1]
    s0 = tr.site("overall_log_scale", 0.0, 10.0)
    v = tr.raw("beta_log_scales")
    tr.add((0.5 * v - 0.5 * torch.exp(v) + 0.5 * math.log(0.5) - math.lgamma(0.5)).sum())
    beta = tr.site("beta", torch.zeros(F, dtype=dt), torch.exp(s0 + v))
    logits = torch.einsum("nd,md->mn", X, beta[None, :])
    tr.add(bernoulli_lp(y[None, :], logits).sum())


def _m_radon(tr, d):
    """models.py:826-837, sigma_y = 1 (:839)."""
    dt = tr.dtype
    u = torch.as_tensor(d["u"], dtype=dt)
    x = torch.as_tensor(d["x"], dtype=dt)
    y = torch.as_tensor(d["y"], dtype=dt).reshape(-1, 1)
    J = u.shape[0]
    mua = tr.site("mua", 0.0, 1.0)
    b1 = tr.site("b1", 0.0, 1.0)
    b2 = tr.site("b2", 0.0, 1.0)
    m = tr.site("m", mua + u * b1, torch.ones(J, dtype=dt))
    C = one_hot(d["county"], J, dt)
    y_mu = C @ m[:, None] + x[:, None] * b2
    tr.add(normal_lp(y, y_mu, torch.ones((), dtype=dt)).sum())


def _m_radon_stddvs(tr, d):
    """models.py:772-788."""
    dt = tr.dtype
    u = torch.as_tensor(d["u"], dtype=dt)
    x = torch.as_tensor(d["x"], dtype=dt)
    y = torch.as_tensor(d["y"], dtype=dt).reshape(-1, 1)
    J = u.shape[0]
    mua = tr.site("mua", 0.0, 1.0)
    b1 = tr.site("b1", 0.0, 1.0)
    b2 = tr.site("b2", 0.0, 1.0)
    m = tr.site("m", mua + u * b1, torch.ones(J, dtype=dt))
    C = one_hot(d["county"], J, dt)
    ls = tr.site("log_m_stddv", torch.zeros(J, dtype=dt), torch.ones(J, dtype=dt))
    y_mu = C @ m[:, None] + x[:, None] * b2
    y_sd = C @ torch.exp(ls)[:, None]
    tr.add(normal_lp(y, y_mu, y_sd).sum())


def _m_election(tr, d):
    """models.py:969-982.  `state` is 1-based and fed to one_hot(depth=51)."""
    dt = tr.dtype
    n_state = int(d["n_state"])
    black = torch.as_tensor(d["black"], dtype=dt)
    female = torch.as_tensor(d["female"], dtype=dt)
    y = torch.as_tensor(d["y"], dtype=dt).reshape(-1, 1)
    mua = tr.site("mua", 0.0, 100.0)
    lsa = tr.site("log_sigma_a", 0.0, 10.0)
    a = tr.site("a", mua * torch.ones(n_state, dtype=dt), torch.ones(n_state, dtype=dt) * torch.exp(lsa))
    b1 = tr.site("b1", 0.0, 100.0)
    b2 = tr.site("b2", 0.0, 100.0)
    C = one_hot(d["state"], n_state, dt)
    y_hat = C @ a[:, None] + female[:, None] * b2 + black[:, None] * b1
    tr.add(bernoulli_lp(y, y_hat).sum())


def _m_electric(tr, d):
    """models.py:1013-1035.  All three index vectors are 1-based into
    one_hot(depth=K); `a` has shape (96, 1)."""
    dt = tr.dtype
    n_pair, n_grade, n_gp = int(d["n_pair"]), int(d["n_grade"]), int(d["n_grade_pair"])
    N = len(d["y"])
    treatment = torch.as_tensor(d["treatment"], dtype=dt)
    y = torch.as_tensor(d["y"], dtype=dt)
    C_pair = one_hot(d["pair"], n_pair, dt)
    C_grade = one_hot(d["grade"], n_grade, dt)
    C_gp = one_hot(d["grade_pair"], n_gp, dt)
    mua = tr.site("mua", 0.0, torch.ones(n_gp, dtype=dt))
    mua_hat = 100.0 * (C_gp @ mua[:, None])
    sigma_y = tr.site("sigma_y", 0.0, torch.ones(n_grade, dtype=dt))
    sigma_y_hat = C_grade @ sigma_y[:, None]
    a = tr.site("a", mua_hat, 1.0)  # (96, 1)
    b = tr.site("b", 0.0, 100.0 * torch.ones(n_grade, dtype=dt))
    y_hat_a = (C_pair @ a).reshape(N)
    y_hat_b = (C_grade @ b[:, None]).reshape(N)
    y_hat_sigma = torch.exp(sigma_y_hat.reshape(N))
    y_hat = y_hat_a + y_hat_b * treatment
    tr.add(normal_lp(y, y_hat, y_hat_sigma).sum())


def _m_time_series(tr, d):
    """models.py:1071-1096 (interleaved alpha_t, mu_t trace order)."""
    dt = tr.dtype
    x = torch.as_tensor(d["x"], dtype=dt)
    y = torch.as_tensor(d["y"], dtype=dt)
    T = x.shape[0]
    sp = torch.nn.functional.softplus
    sa = tr.site("sigma_alpha", 0.0, 1.0)
    sm = tr.site("sigma_mu", 0.0, 1.0)
    alpha = [tr.site("alpha0", 0.0, sp(sa))]
    mu = [tr.site("mu0", 0.0, sp(sm))]
    for t in range(1, T):
        alpha.append(tr.site("alpha%d" % t, alpha[t - 1] + mu[t - 1], sp(sa)))
        mu.append(tr.site("mu%d" % t, mu[t - 1], sp(sm)))
    beta = tr.site("beta", 0.0, 1.0)
    # scale=0.12 enters the TF graph as a float32 constant
    tr.add(normal_lp(y, torch.stack(alpha) + beta * x, torch.as_tensor(float(np.float32(0.12)), dtype=dt)).sum())


def site_table(model, d):
    """(name, shape) of the latent sites in trace order (= state-part order,
    ``graphs.py:29-44``)."""
    if model == "8schools":
        return [("mu", ()), ("log_tau", ()), ("theta", (8,))]
    if model in ("german_credit_lognormalcentered", "german_credit_gammascale"):
        F = d["X"].shape[1]
        return [("overall_log_scale", ()), ("beta_log_scales", (F,)), ("beta", (F,))]
    if model == "radon":
        J = len(d["u"])
        return [("mua", ()), ("b1", ()), ("b2", ()), ("m", (J,))]
    if model == "radon_stddvs":
        J = len(d["u"])
        return [("mua", ()), ("b1", ()), ("b2", ()), ("m", (J,)), ("log_m_stddv", (J,))]
    if model == "election":
        return [("mua", ()), ("log_sigma_a", ()), ("a", (int(d["n_state"]),)), ("b1", ()), ("b2", ())]
    if model == "electric":
        return [("mua", (int(d["n_grade_pair"]),)), ("sigma_y", (int(d["n_grade"]),)),
                ("a", (int(d["n_pair"]), 1)), ("b", (int(d["n_grade"]),))]
    if model == "time_series":
        T = len(d["x"])
        out = [("sigma_alpha", ()), ("sigma_mu", ()), ("alpha0", ()), ("mu0", ())]
        for t in range(1, T):
            out += [("alpha%d" % t, ()), ("mu%d" % t, ())]
        return out + [("beta", ())]
    raise KeyError(model)


_BODIES = {
    "8schools": _m_8schools,
    "german_credit_lognormalcentered": _m_german_lognormal,
    "german_credit_gammascale": _m_german_gamma,
    "radon": _m_radon,
    "radon_stddvs": _m_radon_stddvs,
    "election": _m_election,
    "electric": _m_electric,
    "time_series": _m_time_series,
}


def num_coords(model, d):
    return int(sum(int(np.prod(s)) for _, s in site_table(model, d)))


def _split(model, d, flat, dtype):
    """Flat [D] vector -> dict site -> tensor of the site's shape."""
    out, o = {}, 0
    for name, shape in site_table(model, d):
        n = int(np.prod(shape))
        v = flat[o:o + n]
        out[name] = v.reshape(shape) if not isinstance(v, float) else v
        o += n
    return out


def _as_flat(model, d, v, dtype, default):
    D = num_coords(model, d)
    if v is None:
        return torch.full((D,), float(default), dtype=dtype)
    if isinstance(v, (int, float)):
        return torch.full((D,), float(v), dtype=dtype)
    t = torch.as_tensor(np.asarray(v), dtype=dtype).reshape(-1)
    assert t.shape[0] == D
    return t


def trace(model, d, z, a=None, b=None, dtype=torch.float64):
    """Evaluate one chain: returns the Tracer (lp, centred values)."""
    z = z if torch.is_tensor(z) else torch.as_tensor(np.asarray(z), dtype=dtype)
    z = z.to(dtype)
    a = _as_flat(model, d, a, dtype, 1.0)
    b = _as_flat(model, d, b, dtype, 1.0)
    tr = Tracer(_split(model, d, z, dtype), _split(model, d, a, dtype), _split(model, d, b, dtype), dtype)
    _BODIES[model](tr, d)
    return tr


def log_joint(model, d, z, a=None, b=None, dtype=torch.float64):
    """target(*z) of graphs.py:37-44 for one chain; z flat [D]."""
    return trace(model, d, z, a, b, dtype).lp


def log_joint_and_grad(model, d, Z, a=None, b=None, dtype=torch.float64):
    """Batched over chains: Z [C, D] -> (lp [C], grad [C, D]) via autograd.
    (vectorize_log_joint_fn, inference.py:172-195 + tf.gradients in TFP HMC.)"""
    Z = np.asarray(Z)
    lps, grads = [], []
    for c in range(Z.shape[0]):
        z = torch.tensor(Z[c], dtype=dtype, requires_grad=True)
        lp = log_joint(model, d, z, a, b, dtype)
        (g,) = torch.autograd.grad(lp, z)
        lps.append(lp.detach().numpy())
        grads.append(g.numpy())
    return np.array(lps), np.array(grads)


def grad_wrt_a(model, d, z, a, b, dtype=torch.float64):
    """d log_joint / d a (per coordinate) -- what the cVIP ELBO needs."""
    z = torch.tensor(np.asarray(z), dtype=dtype)
    a = torch.tensor(np.asarray(a), dtype=dtype, requires_grad=True)
    D = num_coords(model, d)
    bb = _as_flat(model, d, b, dtype, 1.0)
    tr = Tracer(_split(model, d, z, dtype), _split(model, d, a, dtype), _split(model, d, bb, dtype), dtype)
    _BODIES[model](tr, d)
    (g,) = torch.autograd.grad(tr.lp, a)
    return g.numpy().reshape(D)


def to_centered(model, d, Z, a=None, b=None, dtype=torch.float64):
    """make_to_centered (models.py:59-81): state -> centred values, per chain."""
    Z = np.asarray(Z)
    out = []
    for c in range(Z.shape[0]):
        tr = trace(model, d, torch.tensor(Z[c], dtype=dtype), a, b, dtype)
        out.append(torch.cat([tr.centered[n].reshape(-1) for n, _ in site_table(model, d)]).numpy())
    return np.array(out)


def to_noncentered(model, d, X, a=None, b=None, dtype=torch.float64):
    """make_to_noncentered / make_to_partially_noncentered (models.py:84-128):
    centred values -> state coordinates under rule (a, b) (default NCP a=b=0).
    Each site inverts its affine map given the centred values of its parents."""
    X = np.asarray(X)
    a = 0.0 if a is None else a
    b = 0.0 if b is None else b
    out = []
    for c in range(X.shape[0]):
        xs = _split(model, d, torch.tensor(X[c], dtype=dtype), dtype)

        class Inv(Tracer):
            def site(self, name, loc, scale):
                x = xs[name]
                aa = self._t(self.a[name])
                bb = self._t(self.b[name])
                loc = self._t(loc) * torch.ones_like(x)
                scale = self._t(scale) * torch.ones_like(x)
                std_loc = loc * aa
                aff_scale = scale / torch.pow(scale, bb)
                aff_shift = loc - aff_scale * std_loc
                self.centered[name] = (x - aff_shift) / aff_scale  # the state coordinate
                return x

            def raw(self, name):
                self.centered[name] = xs[name]
                return xs[name]

        D = num_coords(model, d)
        af = _as_flat(model, d, a, dtype, 0.0)
        bf = _as_flat(model, d, b, dtype, 0.0)
        tr = Inv(xs, _split(model, d, af, dtype), _split(model, d, bf, dtype), dtype)
        _BODIES[model](tr, d)
        out.append(torch.cat([tr.centered[n].reshape(-1) for n, _ in site_table(model, d)]).numpy())
    return np.array(out)


# --------------------------------------------------------------------------- #
# Philox4x32-10 (our own stream convention; the reference is unseeded)
# --------------------------------------------------------------------------- #
PHILOX_M0, PHILOX_M1 = 0xD2511F53, 0xCD9E8D57
PHILOX_W0, PHILOX_W1 = 0x9E3779B9, 0xBB67AE85
STREAM_MOMENTUM, STREAM_ACCEPT, STREAM_VI = 0, 1, 2


def philox4x32(ctr, key):
    """Vectorised Philox4x32-10.  ctr: uint32 [..., 4], key: (k0, k1)."""
    c = [np.asarray(ctr[..., i], dtype=np.uint64) for i in range(4)]
    k0, k1 = np.uint64(key[0] & 0xFFFFFFFF), np.uint64(key[1] & 0xFFFFFFFF)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(PHILOX_M0) * c[0]
        p1 = np.uint64(PHILOX_M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = [(hi1 ^ c[1] ^ k0) & mask, lo1, (hi0 ^ c[3] ^ k1) & mask, lo0]
        k0 = (k0 + np.uint64(PHILOX_W0)) & mask
        k1 = (k1 + np.uint64(PHILOX_W1)) & mask
    return np.stack(c, axis=-1).astype(np.uint32)


def _u01(bits):
    """uint32 -> float32 uniform in (0, 1): (top 24 bits + 0.5) * 2^-24."""
    return ((bits >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)


def philox_normals(seed, chain_ids, step, stream, D):
    """Standard normals [len(chain_ids), D] for one transition / VI step.
    Counter = (chain, step, block j = d // 4, stream); Box-Muller on pairs
    (u0,u1)->(n0,n1), (u2,u3)->(n2,n3)."""
    chain_ids = np.asarray(chain_ids, dtype=np.uint32)
    nb = (D + 3) // 4
    ctr = np.zeros((len(chain_ids), nb, 4), dtype=np.uint32)
    ctr[..., 0] = chain_ids[:, None]
    ctr[..., 1] = np.uint32(step)
    ctr[..., 2] = np.arange(nb, dtype=np.uint32)[None, :]
    ctr[..., 3] = np.uint32(stream)
    r = philox4x32(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    u = _u01(r).astype(np.float64)
    rad0 = np.sqrt(-2.0 * np.log(u[..., 0]))
    rad1 = np.sqrt(-2.0 * np.log(u[..., 2]))
    n = np.stack([rad0 * np.cos(2 * np.pi * u[..., 1]), rad0 * np.sin(2 * np.pi * u[..., 1]),
                  rad1 * np.cos(2 * np.pi * u[..., 3]), rad1 * np.sin(2 * np.pi * u[..., 3])], axis=-1)
    return n.reshape(len(chain_ids), nb * 4)[:, :D]


def philox_log_uniform(seed, chain_ids, step):
    """log U for the Metropolis test of one transition, [len(chain_ids)]."""
    chain_ids = np.asarray(chain_ids, dtype=np.uint32)
    ctr = np.zeros((len(chain_ids), 4), dtype=np.uint32)
    ctr[:, 0] = chain_ids
    ctr[:, 1] = np.uint32(step)
    ctr[:, 3] = np.uint32(STREAM_ACCEPT)
    r = philox4x32(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    return np.log(_u01(r[:, 0]).astype(np.float64))


# --------------------------------------------------------------------------- #
# HMC  [TFP]  (call sites inference.py:218-234)
# --------------------------------------------------------------------------- #
def num_transitions(num_results, num_burnin, num_steps_between_results=1):
    """[TFP sample_chain] first result after 1+burnin steps, then one every
    1+num_steps_between_results (inference.py:229-234)."""
    return 1 + num_burnin + (1 + num_steps_between_results) * (num_results - 1)


def hmc_chain(model, d, z0, eps0, L, num_results, num_burnin, num_adapt,
              a=None, b=None, momenta=None, log_u=None, seed=0, chain_ids=None,
              dtype=torch.float64, thin=1, target_accept=0.75, gamma=0.05, t0=10.0, kappa=0.75,
              return_all=False):
    """Batched-chain HMC in TFP 0.7 op order.

    z0 [C, D]; eps0 [D] (per-coordinate base step, inference.py:212-216) ->
    returns dict(samples_centered [S, C, D], samples_orig, is_accepted [S, C],
    step_mult [C], lp [C]).  `momenta` [T, C, D] / `log_u` [T, C] inject the
    random streams; otherwise the Philox convention above is used.
    Per-chain dual averaging [TFP DualAveragingStepSizeAdaptation,
    inference.py:224-226]: step size parts have the state's rank so no
    cross-chain reduction happens; the per-(c,d) log-step recursion factorises
    into a per-chain scalar multiplier on eps0 which is what is tracked here.
    """
    np_dt = np.float64 if dtype == torch.float64 else np.float32
    z = np.array(z0, dtype=np_dt)
    C, D = z.shape
    eps0 = np.asarray(eps0, dtype=np_dt).reshape(1, D)
    if chain_ids is None:
        chain_ids = np.arange(C)
    T = num_transitions(num_results, num_burnin, thin)
    f = lambda zz: tuple(np.asarray(v, dtype=np_dt) for v in log_joint_and_grad(model, d, zz, a, b, dtype))
    lp, g = f(z)
    H = np.zeros(C, dtype=np_dt)
    log_avg = np.zeros(C, dtype=np_dt)       # log of the averaged multiplier (relative to eps0)
    mult = np.ones(C, dtype=np_dt)           # current multiplier on eps0
    samples, samples_orig, accepted = [], [], []
    all_states = []
    next_keep = num_burnin  # 0-based transition index after which a sample is kept
    for t in range(T):
        eps = eps0 * mult[:, None]
        v0 = (np.asarray(momenta[t], dtype=np_dt) if momenta is not None
              else philox_normals(seed, chain_ids, t, STREAM_MOMENTUM, D).astype(np_dt))
        v, x, gx = v0.copy(), z.copy(), g.copy()
        for _ in range(L):
            v = v + np_dt(0.5) * eps * gx
            x = x + eps * v
            lpx, gx = f(x)
            v = v + np_dt(0.5) * eps * gx
        log_alpha = lpx - lp + np_dt(0.5) * (v0 * v0).sum(1) - np_dt(0.5) * (v * v).sum(1)
        log_alpha = np.where(np.isfinite(log_alpha) | (log_alpha == np.inf), log_alpha, -np.inf)
        lu = (np.asarray(log_u[t], dtype=np_dt) if log_u is not None
              else philox_log_uniform(seed, chain_ids, t).astype(np_dt))
        acc = lu < log_alpha
        z = np.where(acc[:, None], x, z)
        g = np.where(acc[:, None], gx, g)
        lp = np.where(acc, lpx, lp)
        # dual averaging, per chain
        t1 = t + 1
        if t1 <= num_adapt:
            H = H + np_dt(target_accept) - np.exp(np.minimum(log_alpha, 0.0))
            log_step = np_dt(math.log(10.0)) - H * np_dt(math.sqrt(t1)) / np_dt((t1 + t0) * gamma)
            eta = np_dt(t1 ** (-kappa))
            log_avg = eta * log_step + (1 - eta) * log_avg
            mult = np.exp(log_step) if t1 < num_adapt else np.exp(log_avg)
        if return_all:
            all_states.append(z.copy())
        if t == next_keep:
            samples_orig.append(z.copy())
            samples.append(to_centered(model, d, z, a, b, dtype).astype(np_dt))
            accepted.append(acc.copy())
            next_keep += 1 + thin
    out = dict(samples_centered=np.array(samples), samples_orig=np.array(samples_orig),
               is_accepted=np.array(accepted), step_mult=mult, lp=lp, z=z)
    if return_all:
        out["all_states"] = np.array(all_states)
    return out


def hmc_interleaved_chain(model, d, x0, eps0_a, eps0_b, L_a, L_b, num_results, num_burnin, num_adapt,
                          rule_a=(1.0, 1.0), rule_b=(0.0, 0.0), momenta=None, log_u=None, seed=0, chain_ids=None,
                          thin=1, target_accept=0.75, rate=0.05):
    """Interleaved CP / NCP sampler: interleaved.py:113-155 + inference.py:258-329 [TFP
    SimpleStepSizeAdaptation].  The state x lives in the centred space.  Per transition, for rule r in
    (A = CP, B = NCP): z = to_rule_r(x) (to_noncentered); (lp, g) re-bootstrapped at z; one HMC step of L_r
    leapfrog steps with step eps0_r * mult_r; x = to_centered(z'); while transition < num_adapt,
    mult_r *= (1 + rate) if min(1, exp(log_alpha)) > target else 1 / (1 + rate).
    Streams are indexed 2 * transition + r."""
    x = np.array(x0, dtype=np.float64)
    C, D = x.shape
    if chain_ids is None:
        chain_ids = np.arange(C)
    T = num_transitions(num_results, num_burnin, thin)
    rules = (rule_a, rule_b)
    eps0 = (np.asarray(eps0_a, dtype=np.float64).reshape(1, D), np.asarray(eps0_b, dtype=np.float64).reshape(1, D))
    Ls = (L_a, L_b)
    mult = [np.ones(C), np.ones(C)]
    samples, acc_a, acc_b = [], [], []
    next_keep = num_burnin
    for t in range(T):
        accs = []
        for r in (0, 1):
            a, b = rules[r]
            f = lambda zz: log_joint_and_grad(model, d, zz, a, b)
            z = to_noncentered(model, d, x, a, b)
            lp, g = f(z)
            sid = 2 * t + r
            v0 = (np.asarray(momenta[sid], dtype=np.float64) if momenta is not None
                  else philox_normals(seed, chain_ids, sid, STREAM_MOMENTUM, D))
            eps = eps0[r] * mult[r][:, None]
            v, xx, gx = v0.copy(), z.copy(), g.copy()
            for _ in range(Ls[r]):
                v = v + 0.5 * eps * gx
                xx = xx + eps * v
                lpx, gx = f(xx)
                v = v + 0.5 * eps * gx
            la = lpx - lp + 0.5 * (v0 * v0).sum(1) - 0.5 * (v * v).sum(1)
            la = np.where(np.isfinite(la) | (la == np.inf), la, -np.inf)
            lu = (np.asarray(log_u[sid], dtype=np.float64) if log_u is not None
                  else philox_log_uniform(seed, chain_ids, sid))
            acc = lu < la
            z = np.where(acc[:, None], xx, z)
            x = to_centered(model, d, z, a, b)
            if t < num_adapt:
                pacc = np.exp(np.minimum(la, 0.0))
                mult[r] = np.where(pacc > target_accept, mult[r] * (1 + rate), mult[r] / (1 + rate))
            accs.append(acc)
        if t == next_keep:
            samples.append(x.copy()); acc_a.append(accs[0]); acc_b.append(accs[1])
            next_keep += 1 + thin
    return dict(samples=np.array(samples), is_accepted_a=np.array(acc_a), is_accepted_b=np.array(acc_b),
                step_mult_a=mult[0], step_mult_b=mult[1], x=x)


# --------------------------------------------------------------------------- #
# ESS  [TFP effective_sample_size, inference.py:240] + util.get_min_ess
# --------------------------------------------------------------------------- #
def effective_sample_size(states):
    """states [S, ...] -> ESS [...].  FFT autocorrelation (centre, zero-pad to a
    power of two >= 2S, divide lag k by (S-k), normalise by lag 0), zero from the
    first lag with rho < 0 on (filter_threshold=0), ESS = S / (-1 + 2 sum_k
    (S-k)/S rho_k)."""
    x = np.asarray(states, dtype=np.float64)
    S = x.shape[0]
    x = x - x.mean(axis=0, keepdims=True)
    n_fft = 1 << int(math.ceil(math.log2(2 * S)))
    fx = np.fft.rfft(x, n=n_fft, axis=0)
    acov = np.fft.irfft(fx * np.conj(fx), n=n_fft, axis=0)[:S]
    k = np.arange(S).reshape((S,) + (1,) * (x.ndim - 1))
    acov = acov / (S - k)
    with np.errstate(invalid="ignore", divide="ignore"):
        rho = acov / acov[:1]
        mask = np.maximum(1.0 - np.cumsum(rho < 0.0, axis=0), 0.0)
        rho = rho * mask
        return S / (-1.0 + 2.0 * np.sum((S - k) / S * rho, axis=0))


def windowed_ess(states, W):
    """Bounded-memory ESS as the streaming statistics of arp_hmc_run compute it (stream_window = W): the same estimator as
    effective_sample_size, but only the lags k < W are available -- the sum stops at the first lag with rho < 0 or, if
    none lies inside the window, uses all W lags (then `truncated` is True and the value is an upper bound).  Restated
    from the centred lag products directly (no FFT), in the streaming form: y = x - pivot (pivot = first sample),
    A_k = sum_t y_t y_{t-k},  c_k = A_k - m [(sum - head_k) + (sum - tail_k)] + (S - k) m^2 with m = mean(y).
    states [S, ...] -> (ess [...], truncated [...] bool)."""
    x = np.asarray(states, dtype=np.float64)
    S = x.shape[0]
    y = x - x[:1]
    tot = y.sum(axis=0)
    m = tot / S
    K = min(W, S)
    ess = np.full(x.shape[1:], np.nan)
    trunc = np.zeros(x.shape[1:], dtype=bool)
    acc = np.zeros(x.shape[1:])
    alive = np.ones(x.shape[1:], dtype=bool)
    acov0 = None
    for k in range(K):
        A = (y[k:] * y[:S - k]).sum(axis=0)
        head = y[:k].sum(axis=0)           # sum of the first k values
        tail = y[S - k:].sum(axis=0) if k > 0 else 0.0
        ck = A - m * ((tot - head) + (tot - tail)) + (S - k) * m * m
        if k == 0:
            acov0 = ck / S
            alive &= acov0 > 0
        with np.errstate(invalid="ignore", divide="ignore"):
            rho = (ck / (S - k)) / acov0
        neg = alive & (rho < 0)
        alive &= ~neg
        acc = np.where(alive, acc + (S - k) / S * rho, acc)
    ok = acov0 > 0
    with np.errstate(invalid="ignore", divide="ignore"):
        ess = np.where(ok, S / (-1.0 + 2.0 * acc), np.nan)
    trunc = ok & alive & (K < S)
    return ess, trunc


def get_min_ess(ess_parts, num_chains):
    """util.py:445-460: nan->0, per-chain min over all coordinates, mean and
    std/sqrt(n) over chains.  ess_parts: list of [C, *site]."""
    ess_parts = [np.nan_to_num(e) for e in ess_parts]
    mins = [min(np.array(e[c]).min() for e in ess_parts) for c in range(num_chains)]
    return float(np.mean(mins)), float(np.std(mins) / np.sqrt(len(mins)))


# --------------------------------------------------------------------------- #
# VI  (util.py:232-268, program_transformations.py:192-241, inference.py:26-154)
# --------------------------------------------------------------------------- #
def discrete_prior_logp(p):
    """main.py:244-253: Mixture(Categorical(logits=[0, 5, 0]), [Laplace(0, 0.1), Uniform(0, 1), Laplace(1, 0.1)])
    .log_prob(p) for p in (0, 1) -- TFP Mixture.log_prob = logsumexp(log cat_probs + component log_probs)
    [TFP-from-memory]."""
    logw = torch.log_softmax(torch.tensor([0.0, 5.0, 0.0], dtype=p.dtype), dim=0)
    lap0 = -torch.abs(p) / 0.1 - math.log(0.2)
    lap1 = -torch.abs(p - 1.0) / 0.1 - math.log(0.2)
    uni = torch.zeros_like(p)
    return torch.logsumexp(torch.stack([logw[0] + lap0, logw[1] + uni, logw[2] + lap1]), dim=0)


def elbo_and_grads(model, d, loc, rho, eps, a=None, b=None, a_logit=None, dtype=torch.float64, u=None, a_index=None,
                   b_index=None, discrete_prior=False):
    """One ELBO evaluation with S = eps.shape[0] injected standard normals.

    q = prod N(loc, softplus(rho)); z_s = loc + scale*eps_s;
    ELBO = mean_s[log_joint(z_s) - log q(z_s)] with -log q differentiated
    through z AND directly (total-gradient estimator, util.py:253-266).
    If `a_logit` is given the rule is cVIP: a = sigmoid(a_logit)
    (program_transformations.py:507-510), and d/d a_logit is returned too.
    General learnable reparameterisation (tied b = a, untied a / b; program_transformations.py:512-523): `u` [P]
    unconstrained parameters, coordinate d takes a = sigmoid(u[a_index[d]]) (b likewise) where the index is >= 0.
    discrete_prior adds sum_p log prior(sigmoid(u_p)) to the objective (inference.py:50-54).
    Returns objective, dict of gradients of **-objective** (what Adam minimises) [+ "prior_logp"].
    """
    loc = torch.tensor(np.asarray(loc), dtype=dtype, requires_grad=True)
    rho = torch.tensor(np.asarray(rho), dtype=dtype, requires_grad=True)
    eps = torch.as_tensor(np.asarray(eps), dtype=dtype)
    params = [loc, rho]
    b_t = _as_flat(model, d, b, dtype, 1.0)
    plp = torch.zeros((), dtype=dtype)
    if u is not None:
        ut = torch.tensor(np.asarray(u), dtype=dtype, requires_grad=True)
        pv = torch.sigmoid(ut)
        a_fix = _as_flat(model, d, a, dtype, 1.0)
        ia = torch.as_tensor(np.asarray(a_index), dtype=torch.long)
        ib = torch.as_tensor(np.asarray(b_index), dtype=torch.long)
        a_t = torch.where(ia >= 0, pv[ia.clamp(min=0)], a_fix)
        b_t = torch.where(ib >= 0, pv[ib.clamp(min=0)], b_t)
        params.append(ut)
        if discrete_prior:
            plp = discrete_prior_logp(pv).sum()
    elif a_logit is not None:
        ut = torch.tensor(np.asarray(a_logit), dtype=dtype, requires_grad=True)
        a_t = torch.sigmoid(ut)
        params.append(ut)
    else:
        a_t = _as_flat(model, d, a, dtype, 1.0)
    scale = torch.nn.functional.softplus(rho)
    total = torch.zeros((), dtype=dtype)
    for s in range(eps.shape[0]):
        z = loc + scale * eps[s]
        tr = Tracer(_split(model, d, z, dtype), _split(model, d, a_t, dtype), _split(model, d, b_t, dtype), dtype)
        _BODIES[model](tr, d)
        entropy = -normal_lp(z, loc, scale).sum()
        total = total + tr.lp + entropy
    elbo = total / eps.shape[0] + plp
    grads = torch.autograd.grad(-elbo, params)
    out = {"loc": grads[0].numpy(), "rho": grads[1].numpy(), "prior_logp": float(plp.detach())}
    if u is not None:
        out["u"] = grads[2].numpy()
    elif a_logit is not None:
        out["a_logit"] = grads[2].numpy()
    return float(elbo.detach()), out


def adam_step(theta, g, m, v, t, lr, beta1=0.9, beta2=0.999, eps_hat=1e-8):
    """TF1 AdamOptimizer update (inference.py:47): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    m,v moving averages; theta -= lr_t * m / (sqrt(v) + eps_hat).  NaN grads -> 0
    (inference.py:62)."""
    g = np.where(np.isnan(g), 0.0, g)
    lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    m = beta1 * m + (1.0 - beta1) * g
    v = beta2 * v + (1.0 - beta2) * g * g
    theta = theta - lr_t * m / (np.sqrt(v) + eps_hat)
    return theta, m, v


def lr_schedule(step, base_lr, num_steps):
    """inference.py:69-75."""
    if step > 2 * num_steps / 3:
        return base_lr / 20
    if step > num_steps / 3:
        return base_lr / 5
    return base_lr


def vi_run(model, d, loc0, rho0, eps_all, lr, num_steps, a=None, b=None, a_logit0=None,
           dtype=torch.float64, u0=None, a_index=None, b_index=None, discrete_prior=False):
    """num_steps of Adam on -objective with injected eps_all [num_steps, S, D]."""
    loc, rho = np.array(loc0, dtype=np.float64), np.array(rho0, dtype=np.float64)
    key = "u" if u0 is not None else "a_logit"
    ul = None
    if u0 is not None:
        ul = np.array(u0, dtype=np.float64)
    elif a_logit0 is not None:
        ul = np.array(a_logit0, dtype=np.float64)
    st = {k: (np.zeros_like(loc), np.zeros_like(loc)) for k in ("loc", "rho")}
    if ul is not None:
        st[key] = (np.zeros_like(ul), np.zeros_like(ul))
    timeline, plps = [], []
    for step in range(num_steps):
        if u0 is not None:
            e, g = elbo_and_grads(model, d, loc, rho, eps_all[step], a, b, None, dtype, u=ul, a_index=a_index,
                                  b_index=b_index, discrete_prior=discrete_prior)
        else:
            e, g = elbo_and_grads(model, d, loc, rho, eps_all[step], a, b, ul, dtype)
        timeline.append(e)
        plps.append(g["prior_logp"])
        cur = lr_schedule(step, lr, num_steps)
        loc, m, v = adam_step(loc, g["loc"], *st["loc"], step + 1, cur); st["loc"] = (m, v)
        rho, m, v = adam_step(rho, g["rho"], *st["rho"], step + 1, cur); st["rho"] = (m, v)
        if ul is not None:
            ul, m, v = adam_step(ul, g[key], *st[key], step + 1, cur); st[key] = (m, v)
    return dict(loc=loc, rho=rho, a_logit=ul, u=ul, elbo=np.array(timeline), prior_logp=np.array(plps))


# --------------------------------------------------------------------------- #
# Fast batched CPU implementation (float32, all cores) -- the CPU BASELINE leg.
# Same op order as the reference's TF graph: [C, D] tensors, dense one-hot
# matmuls, value+gradient by autograd at every leapfrog step.
# --------------------------------------------------------------------------- #
def german_batched_target(X, y):
    """Batched log-joint of german_credit_lognormalcentered under rule (a, b)
    for Z [C, D] -> [C]; torch ops only so autograd gives the gradient in one
    backward pass, as tf.gradients does for the pfor-vectorised target."""
    F = X.shape[1]

    def site(z, loc, scale, a, b):
        std_loc = loc * a
        std_scale = torch.pow(scale, b)
        u = (z - std_loc) / std_scale
        lp = (-0.5 * u * u - torch.log(std_scale) - 0.5 * LOG_2PI)
        aff = scale / std_scale
        return lp, loc - aff * std_loc + aff * z

    def target(Z, a, b):
        z0, zs, zb = Z[:, :1], Z[:, 1:1 + F], Z[:, 1 + F:]
        lp0, s0 = site(z0, torch.zeros_like(z0), torch.full_like(z0, 10.0), a[:1], b[:1])
        lp1, s = site(zs, s0.expand(-1, F), torch.ones_like(zs), a[1:1 + F], b[1:1 + F])
        lp2, beta = site(zb, torch.zeros_like(zb), torch.exp(s), a[1 + F:], b[1 + F:])
        logits = torch.einsum("nd,md->mn", X, beta)
        ll = y[None, :] * logits - torch.clamp(logits, min=0.0) - torch.log1p(torch.exp(-torch.abs(logits)))
        return lp0.sum(1) + lp1.sum(1) + lp2.sum(1) + ll.sum(1)

    return target


def german_hmc_cpu(X, y, z0, eps0, L, T, a, b, seed=0, num_adapt=0):
    """float32 batched HMC for the CPU baseline: returns (#grad evals, final z)."""
    torch.manual_seed(seed)
    X = torch.as_tensor(X, dtype=torch.float32)
    y = torch.as_tensor(y, dtype=torch.float32)
    a = torch.as_tensor(np.asarray(a), dtype=torch.float32)
    b = torch.as_tensor(np.asarray(b), dtype=torch.float32)
    target = german_batched_target(X, y)

    def vg(Z):
        Z = Z.detach().requires_grad_(True)
        lp = target(Z, a, b)
        (g,) = torch.autograd.grad(lp.sum(), Z)
        return lp.detach(), g

    z = torch.as_tensor(np.asarray(z0), dtype=torch.float32)
    eps = torch.as_tensor(np.asarray(eps0), dtype=torch.float32)[None, :]
    C = z.shape[0]
    lp, g = vg(z)
    H = torch.zeros(C)
    log_avg = torch.zeros(C)
    mult = torch.ones(C)
    n_evals = C
    for t in range(T):
        e = eps * mult[:, None]
        v0 = torch.randn_like(z)
        v, x, gx = v0.clone(), z.clone(), g.clone()
        for _ in range(L):
            v = v + 0.5 * e * gx
            x = x + e * v
            lpx, gx = vg(x)
            v = v + 0.5 * e * gx
        n_evals += C * L
        la = lpx - lp + 0.5 * (v0 * v0).sum(1) - 0.5 * (v * v).sum(1)
        la = torch.where(torch.isnan(la), torch.full_like(la, -float("inf")), la)
        acc = torch.log(torch.rand(C)) < la
        z = torch.where(acc[:, None], x, z)
        g = torch.where(acc[:, None], gx, g)
        lp = torch.where(acc, lpx, lp)
        t1 = t + 1
        if t1 <= num_adapt:
            H = H + 0.75 - torch.exp(torch.clamp(la, max=0.0))
            ls = math.log(10.0) - H * math.sqrt(t1) / ((t1 + 10.0) * 0.05)
            eta = t1 ** -0.75
            log_avg = eta * ls + (1 - eta) * log_avg
            mult = torch.exp(ls) if t1 < num_adapt else torch.exp(log_avg)
    return n_evals, z.numpy()


# --------------------------------------------------------------------------- #
# Batched CPU baseline for ANY model (float32, all cores): the model bodies above
# (dense one-hot matmuls, as the reference writes them) vectorised over the chain
# axis with torch.func.vmap -- the role pfor plays in inference.py:172-195 -- and
# differentiated with autograd at every leapfrog step, as tf.gradients does.
# --------------------------------------------------------------------------- #
def _m_radon_gather(tr, d):
    """models.py:826-837 with the county look-up as a gather (the dense one-hot of a 10^6 x 10^4 problem would be
    4 x 10^10 floats; SURVEY.md 8d allows the gather variant for the CPU baseline)."""
    dt = tr.dtype
    u = torch.as_tensor(d["u"], dtype=dt)
    x = torch.as_tensor(d["x"], dtype=dt)
    y = torch.as_tensor(d["y"], dtype=dt)
    county = torch.as_tensor(np.asarray(d["county"]), dtype=torch.int64)
    J = u.shape[0]
    mua = tr.site("mua", 0.0, 1.0)
    b1 = tr.site("b1", 0.0, 1.0)
    b2 = tr.site("b2", 0.0, 1.0)
    m = tr.site("m", mua + u * b1, torch.ones(J, dtype=dt))
    tr.add(normal_lp(y, m[county] + x * b2, torch.ones((), dtype=dt)).sum())


def batched_value_and_grad(model, d, a, b, dtype=torch.float32, gather=False):
    """Z [C, D] tensor -> (lp [C], grad [C, D]) for all chains at once."""
    a_t = _as_flat(model, d, a, dtype, 1.0)
    b_t = _as_flat(model, d, b, dtype, 1.0)
    body = _m_radon_gather if (gather and model == "radon") else _BODIES[model]

    def f(z):
        tr = Tracer(_split(model, d, z, dtype), _split(model, d, a_t, dtype), _split(model, d, b_t, dtype), dtype)
        body(tr, d)
        return tr.lp

    vg = torch.func.vmap(torch.func.grad_and_value(f))

    def call(Z):
        g, lp = vg(Z)
        return lp, g
    return call


def hmc_cpu_batched(model, d, z0, eps0, L, T, a, b, seed=0, num_adapt=0, gather=False):
    """float32 batched HMC of any model for the CPU baseline: returns (#grad evals, final z).  Same transition as
    hmc_chain (TFP op order, per-chain dual averaging) with torch's own RNG."""
    torch.manual_seed(seed)
    vg = batched_value_and_grad(model, d, a, b, torch.float32, gather)
    z = torch.as_tensor(np.asarray(z0), dtype=torch.float32)
    eps = torch.as_tensor(np.asarray(eps0), dtype=torch.float32)[None, :]
    C = z.shape[0]
    lp, g = vg(z)
    H, log_avg, mult = torch.zeros(C), torch.zeros(C), torch.ones(C)
    n_evals = C
    for t in range(T):
        e = eps * mult[:, None]
        v0 = torch.randn_like(z)
        v, x, gx = v0.clone(), z.clone(), g.clone()
        for _ in range(L):
            v = v + 0.5 * e * gx
            x = x + e * v
            lpx, gx = vg(x)
            v = v + 0.5 * e * gx
        n_evals += C * L
        la = lpx - lp + 0.5 * (v0 * v0).sum(1) - 0.5 * (v * v).sum(1)
        la = torch.where(torch.isnan(la), torch.full_like(la, -float("inf")), la)
        acc = torch.log(torch.rand(C)) < la
        z = torch.where(acc[:, None], x, z)
        g = torch.where(acc[:, None], gx, g)
        lp = torch.where(acc, lpx, lp)
        t1 = t + 1
        if t1 <= num_adapt:
            H = H + 0.75 - torch.exp(torch.clamp(la, max=0.0))
            ls = math.log(10.0) - H * math.sqrt(t1) / ((t1 + 10.0) * 0.05)
            eta = t1 ** -0.75
            log_avg = eta * ls + (1 - eta) * log_avg
            mult = torch.exp(ls) if t1 < num_adapt else torch.exp(log_avg)
    return n_evals, z.numpy()


def vi_cpu_batched(model, d, S, num_steps, lr, a, b, seed=0):
    """float32 mean-field VI for the CPU baseline (fixed (a, b): CP / NCP / dVIP): the S Monte-Carlo samples of a
    step evaluated as one batch, total-gradient estimator and TF1 Adam as in vi_run.  Returns the ELBO timeline."""
    torch.manual_seed(seed)
    dtype = torch.float32
    D = num_coords(model, d)
    a_t = _as_flat(model, d, a, dtype, 1.0)
    b_t = _as_flat(model, d, b, dtype, 1.0)

    def f(z):
        tr = Tracer(_split(model, d, z, dtype), _split(model, d, a_t, dtype), _split(model, d, b_t, dtype), dtype)
        _BODIES[model](tr, d)
        return tr.lp

    fb = torch.func.vmap(f)
    loc = (0.01 * torch.randn(D)).requires_grad_(True)
    rho = torch.full((D,), -2.0, requires_grad=True)
    st = [(torch.zeros(D), torch.zeros(D)), (torch.zeros(D), torch.zeros(D))]
    timeline = []
    for step in range(num_steps):
        eps = torch.randn(S, D)
        scale = torch.nn.functional.softplus(rho)
        z = loc + scale * eps
        elbo = (fb(z) - normal_lp(z, loc, scale).sum(1)).mean()
        grads = torch.autograd.grad(-elbo, [loc, rho])
        timeline.append(float(elbo.detach()))
        cur = lr_schedule(step, lr, num_steps)
        with torch.no_grad():
            for i, (p, g) in enumerate(zip((loc, rho), grads)):
                g = torch.nan_to_num(g, nan=0.0)
                m, v = st[i]
                m = 0.9 * m + 0.1 * g
                v = 0.999 * v + 0.001 * g * g
                st[i] = (m, v)
                t = step + 1
                lr_t = cur * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
                p -= lr_t * m / (torch.sqrt(v) + 1e-8)
    return np.array(timeline)
