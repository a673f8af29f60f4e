"""ctypes binding of the C ABI declared in ``include/autoreparam_b200.h``.

The CUDA library is the product: there is NO CPU fallback.  If the shared
library has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)
importing a symbol from here raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ARP_MEM_HOST, ARP_MEM_DEVICE = 0, 1
ARP_VI_MAX_RUNS = 16


class ModelData(C.Structure):
    _fields_ = [("n", C.c_int64), ("f", C.c_int64), ("j", C.c_int64), ("k", C.c_int64), ("k2", C.c_int64),
                ("X", C.c_void_p), ("y", C.c_void_p), ("x1", C.c_void_p), ("x2", C.c_void_p), ("u", C.c_void_p),
                ("idx0", C.c_void_p), ("idx1", C.c_void_p), ("idx2", C.c_void_p)]


class HmcConfig(C.Structure):
    _fields_ = [("num_leapfrog_steps", C.c_int32), ("num_results", C.c_int32), ("num_burnin_steps", C.c_int32),
                ("num_adaptation_steps", C.c_int32), ("num_steps_between_results", C.c_int32),
                ("seed", C.c_uint64), ("chain_offset", C.c_int64), ("target_accept_prob", C.c_double),
                ("lanes_per_chain", C.c_int32), ("engine", C.c_int32), ("stream_window", C.c_int32)]


class HmcBuffers(C.Structure):
    _fields_ = [("z0", C.c_void_p), ("eps0", C.c_void_p), ("ext_momenta", C.c_void_p), ("ext_log_u", C.c_void_p),
                ("samples", C.c_void_p), ("samples_orig", C.c_void_p), ("is_accepted", C.c_void_p),
                ("final_z", C.c_void_p), ("step_mult", C.c_void_p), ("accept_count", C.c_void_p),
                ("stream_mean", C.c_void_p), ("stream_var", C.c_void_p), ("stream_ess", C.c_void_p),
                ("stream_truncated", C.c_void_p)]


class IlvConfig(C.Structure):
    _fields_ = [("num_leapfrog_steps_a", C.c_int32), ("num_leapfrog_steps_b", C.c_int32), ("num_results", C.c_int32),
                ("num_burnin_steps", C.c_int32), ("num_adaptation_steps", C.c_int32),
                ("num_steps_between_results", C.c_int32), ("seed", C.c_uint64), ("chain_offset", C.c_int64),
                ("target_accept_prob", C.c_double), ("adaptation_rate", C.c_double), ("lanes_per_chain", C.c_int32)]


class IlvBuffers(C.Structure):
    _fields_ = [("x0", C.c_void_p), ("eps0_a", C.c_void_p), ("eps0_b", C.c_void_p), ("ext_momenta", C.c_void_p),
                ("ext_log_u", C.c_void_p), ("samples", C.c_void_p), ("is_accepted_a", C.c_void_p),
                ("is_accepted_b", C.c_void_p), ("step_mult_a", C.c_void_p), ("step_mult_b", C.c_void_p)]


class ViConfig(C.Structure):
    _fields_ = [("num_mc_samples", C.c_int32), ("num_optimization_steps", C.c_int32), ("num_runs", C.c_int32),
                ("learning_rates", C.c_double * ARP_VI_MAX_RUNS), ("seed", C.c_uint64), ("num_params", C.c_int32),
                ("discrete_prior", C.c_int32)]


class ViBuffers(C.Structure):
    _fields_ = [("loc", C.c_void_p), ("rho", C.c_void_p), ("u", C.c_void_p), ("a_index", C.c_void_p),
                ("b_index", C.c_void_p), ("ext_eps", C.c_void_p), ("elbo", C.c_void_p), ("prior_logp", C.c_void_p)]


# every symbol include/autoreparam_b200.h declares
EXPORTS = ["arp_model_create", "arp_model_destroy", "arp_model_num_coords", "arp_log_joint_grad", "arp_log_joint_grad_engine", "arp_log_joint_param_grad",
           "arp_hmc_num_transitions", "arp_hmc_run", "arp_hmc_run_many", "arp_hmc_interleaved_run", "arp_ess", "arp_vi_run", "arp_kernel_launch_count",
           "arp_last_error", "arp_precision", "arp_release_cached_memory"]

_libs = {}


def lib_path(precision="f32"):
    # ARP_LIB_F32 / ARP_LIB_F64 override the path (kernel A/B experiments); default is the in-tree build
    return os.environ.get("ARP_LIB_%s" % precision.upper(), os.path.join(_HERE, "libarp_%s.so" % precision))


def _try_build():
    """A fresh checkout has no .so (they are git-ignored): compile them once with nvcc if it is available.
    This only builds the CUDA library; nothing here can compute without it."""
    root = os.path.dirname(_HERE)
    entry = os.path.join(root, "__graft_entry__.py")
    if not os.path.exists(entry):
        return
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("__graft_entry__", entry)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.compile_libraries()
    except Exception as e:  # reported by the caller as "library missing"
        import sys
        sys.stderr.write("autoreparam_b200: building the CUDA library failed: %s\n" % e)


def load(precision="f32"):
    """Load (once) and return the CUDA library for ``precision`` in {"f32","f64"}."""
    if precision in _libs:
        return _libs[precision]
    path = lib_path(precision)
    if not os.path.exists(path):
        _try_build()
    if not os.path.exists(path):
        raise RuntimeError(
            "autoreparam_b200: CUDA library %s is missing -- build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'`; there is no CPU fallback" % path)
    lib = C.CDLL(path)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    lib.arp_model_create.argtypes = [C.c_char_p, C.POINTER(ModelData), C.POINTER(vp)]
    lib.arp_model_create.restype = i32
    lib.arp_model_destroy.argtypes = [vp]
    lib.arp_model_destroy.restype = None
    lib.arp_model_num_coords.argtypes = [vp]
    lib.arp_model_num_coords.restype = i32
    lib.arp_log_joint_grad.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp, vp, i32, vp]
    lib.arp_log_joint_grad.restype = i32
    if os.environ.get("ARP_LIB_%s" % precision.upper()):
        # kernel A/B experiments may load a library built from an older tree: tolerate exports it does not have yet
        # (calling one of them then fails with AttributeError; the in-tree product library must export everything,
        # tests/test_cpu_host.py checks that)
        for name in EXPORTS:
            if not hasattr(lib, name):
                setattr(lib, name, None)
    if lib.arp_log_joint_grad_engine is not None:
        lib.arp_log_joint_grad_engine.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp, i32, i32, vp]
        lib.arp_log_joint_grad_engine.restype = i32
    if lib.arp_log_joint_param_grad is not None:
        lib.arp_log_joint_param_grad.argtypes = [vp, vp, vp, vp, i64, vp, vp, i32, vp]
        lib.arp_log_joint_param_grad.restype = i32
    lib.arp_hmc_num_transitions.argtypes = [C.POINTER(HmcConfig)]
    lib.arp_hmc_num_transitions.restype = i64
    lib.arp_hmc_run.argtypes = [vp, C.POINTER(HmcConfig), vp, vp, i64, C.POINTER(HmcBuffers), i32, vp]
    lib.arp_hmc_run.restype = i32
    if lib.arp_hmc_run_many is not None:
        lib.arp_hmc_run_many.argtypes = [vp, C.POINTER(HmcConfig), i32, vp, vp, i64, C.POINTER(HmcBuffers), i32, vp]
        lib.arp_hmc_run_many.restype = i32
    lib.arp_hmc_interleaved_run.argtypes = [vp, C.POINTER(IlvConfig), vp, vp, vp, vp, i64, C.POINTER(IlvBuffers), i32, vp]
    lib.arp_hmc_interleaved_run.restype = i32
    lib.arp_ess.argtypes = [vp, i64, i64, i64, vp, vp, vp, i32, vp]
    lib.arp_ess.restype = i32
    lib.arp_vi_run.argtypes = [vp, C.POINTER(ViConfig), vp, vp, C.POINTER(ViBuffers), i32, vp]
    lib.arp_vi_run.restype = i32
    lib.arp_kernel_launch_count.argtypes = []
    lib.arp_kernel_launch_count.restype = i64
    lib.arp_last_error.argtypes = []
    lib.arp_last_error.restype = C.c_char_p
    lib.arp_release_cached_memory.argtypes = []
    lib.arp_release_cached_memory.restype = None
    lib.arp_precision.argtypes = []
    lib.arp_precision.restype = C.c_char_p
    assert lib.arp_precision().decode() == precision
    _libs[precision] = lib
    return lib


def np_dtype(precision):
    return np.float32 if precision == "f32" else np.float64


def check(lib, rc, what):
    if rc != 0:
        raise RuntimeError("%s failed: %s" % (what, lib.arp_last_error().decode()))
