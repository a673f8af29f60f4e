"""Thin Python layer over the C ABI: numpy (host buffers) or torch.cuda tensors
(device buffers, current torch stream).  No arithmetic happens here."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05, ENGINE_TCGEN05_STREAM = 0, 1, 2, 3   # include/autoreparam_b200.h: ARP_ENGINE_*


def _is_torch(x):
    return x is not None and type(x).__module__.startswith("torch")


def _np(x, dtype):
    return None if x is None else np.ascontiguousarray(np.asarray(x, dtype=dtype))


def _p(x):
    if x is None:
        return None
    if _is_torch(x):
        return C.c_void_p(x.data_ptr())
    return x.ctypes.data_as(C.c_void_p)


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _torch_dtype(precision):
    import torch
    return torch.float32 if precision == "f32" else torch.float64


def log_joint_grad(model, z, a, b, precision="f32", want_abar=False, engine=ENGINE_AUTO):
    """``z`` [C, D] -> (lp [C], grad [C, D], centered [C, D][, abar [C, D]]).
    ``engine=ENGINE_TCGEN05``: the values as the tensor-core HMC engine computes them (``arp_log_joint_grad_engine``)."""
    lib = _lib.load(precision)
    dt = _lib.np_dtype(precision)
    a_h, b_h = _np(a, dt), _np(b, dt)
    D = model.num_coords
    assert a_h.shape == (D,) and b_h.shape == (D,)
    if _is_torch(z):
        import torch
        z = z.contiguous()
        assert z.dtype == _torch_dtype(precision) and z.is_cuda
        Cn = z.shape[0]
        lp = torch.empty(Cn, dtype=z.dtype, device=z.device)
        g, xc = torch.empty_like(z), torch.empty_like(z)
        ab = torch.empty_like(z) if want_abar else None
        mem, st = _lib.ARP_MEM_DEVICE, _stream()
    else:
        z = _np(z, dt)
        Cn = z.shape[0]
        lp = np.empty(Cn, dtype=dt)
        g, xc = np.empty_like(z), np.empty_like(z)
        ab = np.empty_like(z) if want_abar else None
        mem, st = _lib.ARP_MEM_HOST, None
    assert z.shape == (Cn, D)
    if engine in (ENGINE_TCGEN05, ENGINE_TCGEN05_STREAM):
        assert not want_abar, "the tcgen05 engine does not produce d/da"
        rc = lib.arp_log_joint_grad_engine(model.handle(precision), _p(a_h), _p(b_h), _p(z), Cn, _p(lp), _p(g), _p(xc),
                                           engine, mem, st)
        _lib.check(lib, rc, "arp_log_joint_grad_engine")
        return lp, g, xc
    rc = lib.arp_log_joint_grad(model.handle(precision), _p(a_h), _p(b_h), _p(z), Cn, _p(lp), _p(g), _p(xc), _p(ab),
                                mem, st)
    _lib.check(lib, rc, "arp_log_joint_grad")
    return (lp, g, xc, ab) if want_abar else (lp, g, xc)


def hmc_num_transitions(num_results, num_burnin_steps, num_steps_between_results=1):
    return 1 + num_burnin_steps + (1 + num_steps_between_results) * (num_results - 1)


def hmc_run(model, z0, eps0, a, b, *, num_leapfrog_steps, num_results, num_burnin_steps, num_adaptation_steps,
            num_steps_between_results=1, seed=0, chain_offset=0, target_accept_prob=0.75, ext_momenta=None,
            ext_log_u=None, want_samples=True, want_orig=False, want_final=True, engine=ENGINE_AUTO,
            lanes_per_chain=0, precision="f32", out=None, stream_window=0, want_is_accepted=None):
    """Run every transition of every chain in one launch (``arp_hmc_run``).

    Host mode: numpy in, numpy out.  Device mode: ``z0`` is a torch.cuda tensor;
    outputs are torch.cuda tensors (``out`` may carry preallocated ``samples`` /
    ``is_accepted`` tensors to reuse across calls).
    Returns a dict: samples [S,C,D] (centred), samples_orig, is_accepted [S,C] uint8,
    final_z [C,D], step_mult [C], accept_count [C].
    ``stream_window`` = W > 0 (SIMT engine): additionally stream_mean / stream_var / stream_ess [C,D] and
    stream_truncated [C,D] int32 from in-kernel streaming statistics with a W-lag window -- combine with
    ``want_samples=False`` for runs whose traces do not fit in memory.
    """
    lib = _lib.load(precision)
    dt = _lib.np_dtype(precision)
    D = model.num_coords
    cfg = _lib.HmcConfig(num_leapfrog_steps, num_results, num_burnin_steps, num_adaptation_steps,
                         num_steps_between_results, seed, chain_offset, target_accept_prob, lanes_per_chain, engine,
                         int(stream_window))
    T = lib.arp_hmc_num_transitions(C.byref(cfg))
    a_h, b_h = _np(a, dt), _np(b, dt)
    out = dict(out or {})
    S = num_results
    if _is_torch(z0):
        import torch
        tdt = _torch_dtype(precision)
        z0 = z0.contiguous()
        Cn = z0.shape[0]
        dev = z0.device
        eps0 = torch.as_tensor(eps0, dtype=tdt, device=dev).contiguous()
        mom = None if ext_momenta is None else torch.as_tensor(ext_momenta, dtype=tdt, device=dev).contiguous()
        lu = None if ext_log_u is None else torch.as_tensor(ext_log_u, dtype=tdt, device=dev).contiguous()
        mk = lambda shape, dtype: torch.empty(shape, dtype=dtype, device=dev)
        u8, i32 = torch.uint8, torch.int32
        mem, st = _lib.ARP_MEM_DEVICE, _stream()
    else:
        tdt = dt
        z0 = _np(z0, dt)
        Cn = z0.shape[0]
        eps0 = _np(eps0, dt)
        mom, lu = _np(ext_momenta, dt), _np(ext_log_u, dt)
        mk = lambda shape, dtype: np.empty(shape, dtype=dtype)
        u8, i32 = np.uint8, np.int32
        mem, st = _lib.ARP_MEM_HOST, None
    assert tuple(z0.shape) == (Cn, D) and tuple(eps0.shape) == (D,)
    if mom is not None:
        assert tuple(mom.shape) == (T, Cn, D)
    if lu is not None:
        assert tuple(lu.shape) == (T, Cn)
    if want_samples and "samples" not in out:
        out["samples"] = mk((S, Cn, D), tdt)
    if (want_samples if want_is_accepted is None else want_is_accepted) and "is_accepted" not in out:
        out["is_accepted"] = mk((S, Cn), u8)
    if want_orig and "samples_orig" not in out:
        out["samples_orig"] = mk((S, Cn, D), tdt)
    if want_final:
        out["final_z"] = mk((Cn, D), tdt)
    out["step_mult"] = mk((Cn,), tdt)
    out["accept_count"] = mk((Cn,), i32)
    if stream_window > 0:
        for k in ("stream_mean", "stream_var", "stream_ess"):
            out[k] = mk((Cn, D), tdt)
        out["stream_truncated"] = mk((Cn, D), i32)
    buf = _lib.HmcBuffers(_p(z0), _p(eps0), _p(mom), _p(lu), _p(out.get("samples")), _p(out.get("samples_orig")),
                          _p(out.get("is_accepted")), _p(out.get("final_z")), _p(out["step_mult"]),
                          _p(out["accept_count"]), _p(out.get("stream_mean")), _p(out.get("stream_var")),
                          _p(out.get("stream_ess")), _p(out.get("stream_truncated")))
    rc = lib.arp_hmc_run(model.handle(precision), C.byref(cfg), _p(a_h), _p(b_h), Cn, C.byref(buf), mem, st)
    _lib.check(lib, rc, "arp_hmc_run")
    out["num_transitions"] = int(T)
    return out


def hmc_run_many(model, z0, eps0_list, a, b, *, num_leapfrog_steps, num_results, num_burnin_steps,
                 num_adaptation_steps, num_steps_between_results=1, seed=0, chain_offset=0, target_accept_prob=0.75,
                 engine=ENGINE_AUTO, lanes_per_chain=0, precision="f32", want_samples=True):
    """Several HMC runs in ONE launch (``arp_hmc_run_many``): the leapfrog-step tuning grid.

    ``num_leapfrog_steps`` / ``num_results`` / ``num_burnin_steps`` / ``num_adaptation_steps`` are lists (one entry
    per run; a scalar is broadcast), ``eps0_list`` one [D] array per run; every run starts from ``z0`` [C, D]
    (numpy: host buffers; torch.cuda: device buffers).  Returns a list of dicts like ``hmc_run``."""
    lib = _lib.load(precision)
    dt = _lib.np_dtype(precision)
    D = model.num_coords
    n = len(eps0_list)
    bc = lambda v: list(v) if isinstance(v, (list, tuple, np.ndarray)) else [v] * n
    Ls, Ss, Bs, As = bc(num_leapfrog_steps), bc(num_results), bc(num_burnin_steps), bc(num_adaptation_steps)
    assert len(Ls) == len(Ss) == len(Bs) == len(As) == n
    cfgs = (_lib.HmcConfig * n)()
    for i in range(n):
        cfgs[i] = _lib.HmcConfig(int(Ls[i]), int(Ss[i]), int(Bs[i]), int(As[i]), num_steps_between_results, seed,
                                 chain_offset, target_accept_prob, lanes_per_chain, engine, 0)
    a_h, b_h = _np(a, dt), _np(b, dt)
    if _is_torch(z0):
        import torch
        tdt = _torch_dtype(precision)
        z0 = z0.contiguous()
        dev = z0.device
        conv = lambda v: torch.as_tensor(v, dtype=tdt, device=dev).contiguous()
        mk = lambda shape, dtype: torch.empty(shape, dtype=dtype, device=dev)
        u8, i32 = torch.uint8, torch.int32
        mem, st = _lib.ARP_MEM_DEVICE, _stream()
    else:
        tdt = dt
        z0 = _np(z0, dt)
        conv = lambda v: _np(v, dt)
        mk = lambda shape, dtype: np.empty(shape, dtype=dtype)
        u8, i32 = np.uint8, np.int32
        mem, st = _lib.ARP_MEM_HOST, None
    Cn = z0.shape[0]
    assert tuple(z0.shape) == (Cn, D)
    bufs = (_lib.HmcBuffers * n)()
    outs, keep = [], []
    for i in range(n):
        eps = conv(eps0_list[i])
        assert tuple(eps.shape) == (D,)
        keep.append(eps)
        o = dict(step_mult=mk((Cn,), tdt), accept_count=mk((Cn,), i32))
        if want_samples:
            o["samples"] = mk((int(Ss[i]), Cn, D), tdt)
            o["is_accepted"] = mk((int(Ss[i]), Cn), u8)
        bufs[i] = _lib.HmcBuffers(_p(z0), _p(eps), None, None, _p(o.get("samples")), None, _p(o.get("is_accepted")),
                                  None, _p(o["step_mult"]), _p(o["accept_count"]), None, None, None, None)
        o["num_transitions"] = hmc_num_transitions(int(Ss[i]), int(Bs[i]), num_steps_between_results)
        outs.append(o)
    rc = lib.arp_hmc_run_many(model.handle(precision), cfgs, n, _p(a_h), _p(b_h), Cn, bufs, mem, st)
    _lib.check(lib, rc, "arp_hmc_run_many")
    return outs


def hmc_interleaved_run(model, x0, eps0_a, eps0_b, rule_a, rule_b, *, num_leapfrog_steps_a, num_leapfrog_steps_b,
                        num_results, num_burnin_steps, num_adaptation_steps, num_steps_between_results=1, seed=0,
                        chain_offset=0, target_accept_prob=0.75, adaptation_rate=0.05, ext_momenta=None,
                        ext_log_u=None, lanes_per_chain=0, precision="f32"):
    """Interleaved sampler (``arp_hmc_interleaved_run``).  ``x0`` [C, D] centred initial states (numpy or
    torch.cuda); ``rule_a`` / ``rule_b`` = (a, b) arrays.  Returns dict(samples [S,C,D] centred,
    is_accepted_a, is_accepted_b [S,C], step_mult_a, step_mult_b [C])."""
    lib = _lib.load(precision)
    dt = _lib.np_dtype(precision)
    D = model.num_coords
    cfg = _lib.IlvConfig(num_leapfrog_steps_a, num_leapfrog_steps_b, num_results, num_burnin_steps,
                         num_adaptation_steps, num_steps_between_results, seed, chain_offset, target_accept_prob,
                         adaptation_rate, lanes_per_chain)
    T = hmc_num_transitions(num_results, num_burnin_steps, num_steps_between_results)
    S = num_results
    aa, ba, ab, bb = [_np(v, dt) for v in (rule_a[0], rule_a[1], rule_b[0], rule_b[1])]
    if _is_torch(x0):
        import torch
        tdt = _torch_dtype(precision)
        x0 = x0.contiguous()
        dev = x0.device
        conv = lambda v: None if v is None else torch.as_tensor(v, dtype=tdt, device=dev).contiguous()
        mk = lambda shape, dtype: torch.empty(shape, dtype=dtype, device=dev)
        u8 = torch.uint8
        mem, st = _lib.ARP_MEM_DEVICE, _stream()
    else:
        tdt = dt
        x0 = _np(x0, dt)
        conv = lambda v: _np(v, dt)
        mk = lambda shape, dtype: np.empty(shape, dtype=dtype)
        u8 = np.uint8
        mem, st = _lib.ARP_MEM_HOST, None
    Cn = x0.shape[0]
    e1, e2, mom, lu = conv(eps0_a), conv(eps0_b), conv(ext_momenta), conv(ext_log_u)
    if mom is not None:
        assert tuple(mom.shape) == (2 * T, Cn, D)
    if lu is not None:
        assert tuple(lu.shape) == (2 * T, Cn)
    out = dict(samples=mk((S, Cn, D), tdt), is_accepted_a=mk((S, Cn), u8), is_accepted_b=mk((S, Cn), u8),
               step_mult_a=mk((Cn,), tdt), step_mult_b=mk((Cn,), tdt))
    buf = _lib.IlvBuffers(_p(x0), _p(e1), _p(e2), _p(mom), _p(lu), _p(out["samples"]), _p(out["is_accepted_a"]),
                          _p(out["is_accepted_b"]), _p(out["step_mult_a"]), _p(out["step_mult_b"]))
    rc = lib.arp_hmc_interleaved_run(model.handle(precision), C.byref(cfg), _p(aa), _p(ba), _p(ab), _p(bb), Cn,
                                     C.byref(buf), mem, st)
    _lib.check(lib, rc, "arp_hmc_interleaved_run")
    out["num_transitions"] = int(T)
    return out


def ess(samples, precision="f32", want_moments=False):
    """samples [S, C, D] -> ESS [C, D] (``arp_ess``; TFP effective_sample_size semantics).
    With ``want_moments`` also returns the per-chain mean and (biased) variance [C, D]."""
    lib = _lib.load(precision)
    dt = _lib.np_dtype(precision)
    if _is_torch(samples):
        import torch
        samples = samples.contiguous()
        S, Cn, D = samples.shape
        mk = lambda: torch.empty((Cn, D), dtype=samples.dtype, device=samples.device)
        mem, st = _lib.ARP_MEM_DEVICE, _stream()
    else:
        samples = _np(samples, dt)
        S, Cn, D = samples.shape
        mk = lambda: np.empty((Cn, D), dtype=dt)
        mem, st = _lib.ARP_MEM_HOST, None
    out = mk()
    mean = mk() if want_moments else None
    var = mk() if want_moments else None
    _lib.check(lib, lib.arp_ess(_p(samples), S, Cn, D, _p(out), _p(mean), _p(var), mem, st), "arp_ess")
    return (out, mean, var) if want_moments else out


def log_joint_param_grad(model, z, a, b, precision="f32"):
    """``z`` [C, D] (numpy) -> (abar, bbar) [C, D]: d log_joint / d a and / d b per coordinate
    (``arp_log_joint_param_grad``), the adjoints the cVIP objective differentiates."""
    lib = _lib.load(precision)
    dt = _lib.np_dtype(precision)
    a_h, b_h, z = _np(a, dt), _np(b, dt), _np(z, dt)
    ab, bb = np.empty_like(z), np.empty_like(z)
    rc = lib.arp_log_joint_param_grad(model.handle(precision), _p(a_h), _p(b_h), _p(z), z.shape[0], _p(ab), _p(bb),
                                      _lib.ARP_MEM_HOST, None)
    _lib.check(lib, rc, "arp_log_joint_param_grad")
    return ab, bb


def vi_run(model, a, b, loc, rho, learning_rates, *, num_mc_samples, num_optimization_steps, u=None, a_index=None,
           b_index=None, num_params=0, discrete_prior=False, seed=0, ext_eps=None, precision="f32"):
    """All learning rates at once (``arp_vi_run``), host buffers.

    loc / rho: [R, D] initial values, u: [R, P] unconstrained reparameterisation parameters (copied, not modified);
    a_index / b_index [D]: slot of every coordinate's a / b, -1 = fixed (see include/autoreparam_b200.h).
    Returns dict(loc, rho, u, elbo [R, steps], prior_logp [R, steps])."""
    lib = _lib.load(precision)
    dt = _lib.np_dtype(precision)
    D = model.num_coords
    R = len(learning_rates)
    P = int(num_params)
    cfg = _lib.ViConfig()
    cfg.num_mc_samples, cfg.num_optimization_steps, cfg.num_runs = num_mc_samples, num_optimization_steps, R
    for i, lr in enumerate(learning_rates):
        cfg.learning_rates[i] = float(lr)
    cfg.seed = seed
    cfg.num_params = P
    cfg.discrete_prior = 1 if discrete_prior else 0
    loc = np.array(np.broadcast_to(np.asarray(loc, dtype=dt), (R, D)))
    rho = np.array(np.broadcast_to(np.asarray(rho, dtype=dt), (R, D)))
    uu = ia = ib = None
    if P > 0:
        uu = np.array(np.broadcast_to(np.asarray(np.zeros(P) if u is None else u, dtype=dt), (R, P)))
        ia = np.ascontiguousarray(np.asarray(a_index, dtype=np.int32))
        ib = np.ascontiguousarray(np.asarray(b_index, dtype=np.int32))
        assert ia.shape == (D,) and ib.shape == (D,)
    eps = _np(ext_eps, dt)
    if eps is not None:
        assert eps.shape == (num_optimization_steps, num_mc_samples, D)
    elbo = np.empty((R, num_optimization_steps), dtype=dt)
    plp = np.zeros((R, num_optimization_steps), dtype=dt)
    buf = _lib.ViBuffers(_p(loc), _p(rho), _p(uu), _p(ia), _p(ib), _p(eps), _p(elbo), _p(plp))
    a_h, b_h = _np(a, dt), _np(b, dt)
    rc = lib.arp_vi_run(model.handle(precision), C.byref(cfg), _p(a_h), _p(b_h), C.byref(buf), _lib.ARP_MEM_HOST, None)
    _lib.check(lib, rc, "arp_vi_run")
    return dict(loc=loc, rho=rho, u=uu, elbo=elbo, prior_logp=plp)


def kernel_launch_count(precision="f32"):
    return int(_lib.load(precision).arp_kernel_launch_count())
