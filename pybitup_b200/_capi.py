"""ctypes binding of the C ABI declared in ``include/pybitup_b200.h``.

The shared library is built in-tree (``pybitup_b200/_lib/libpybitup_b200.so``,
see ``__graft_entry__.build``).  There is NO CPU fallback: if the library is
missing, importing this module raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PBU_LIB: an experimental build of the same library (tools/build_variant.sh), for A/B measurements
LIB_PATH = os.environ.get("PBU_LIB") or os.path.join(_HERE, "_lib", "libpybitup_b200.so")

PBU_MAX_PARAM = 16
PBU_MAX_DIM = 16

PBU_OK = 0
PBU_ERR_INVALID = -1
PBU_ERR_UNSUPPORTED = -2
PBU_ERR_CUDA = -3
PBU_ERR_NOT_PD = -4
PBU_ERR_STREAM_EXHAUSTED = -5

MODEL_IDS = {"linear": 0, "poly": 1, "arrhenius": 2, "heat": 3, "soize": 4}
ALGO_IDS = {"RWMH": 0, "AMH": 1, "DR": 2, "DRAM": 3, "ISDE": 4, "HMC": 5}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


class pbu_problem(C.Structure):
    _fields_ = [("model", C.c_int32), ("n_param", C.c_int32), ("n_unc", C.c_int32), ("n_points", C.c_int32),
                ("n_runs", C.c_int32), ("n_design_cols", C.c_int32),
                ("theta_nom", _dp), ("unc_idx", _ip), ("design", _dp), ("y", _dp), ("std_y", _dp),
                ("lb", _dp), ("ub", _dp), ("gamma", C.c_double), ("consts", C.c_double * 8),
                ("log_const", C.c_double), ("has_log_const", C.c_int32), ("n_models", C.c_int32),
                ("model_first", _ip), ("model_count", _ip), ("model_theta_nom", _dp), ("model_unc_idx", _ip),
                ("model_consts", _dp), ("par_kind", _ip), ("par_a", _dp), ("par_b", _dp), ("par_c", _dp), ("par_ldj_s", _dp),
                ("par_ldj_c0", C.c_double)]


class pbu_sampler_cfg(C.Structure):
    _fields_ = [("algo", C.c_int32), ("updating_it", C.c_int32), ("eps_v", C.c_double), ("gamma_dr", C.c_double),
                ("isde_h", C.c_double), ("isde_f0", C.c_double), ("isde_gradient", C.c_int32),
                ("adapt_start", C.c_int32), ("adapt_update", C.c_int32), ("adapt_end", C.c_int32),
                ("am_welford", C.c_int32), ("nan_rejects", C.c_int32), ("lanes_per_chain", C.c_int32),
                ("block_threads", C.c_int32), ("chains_per_thread", C.c_int32), ("tile_points", C.c_int32),
                ("seed", C.c_uint64), ("chain_offset", C.c_uint64), ("hmc_step_size", C.c_double),
                ("hmc_num_steps", C.c_int32), ("cooperative", C.c_int32), ("track_map", C.c_int32), ("reserved0", C.c_int32)]


class pbu_run_outputs(C.Structure):
    _fields_ = [("samples", C.c_void_p), ("lp_kept", C.c_void_p), ("trace_lp1", C.c_void_p),
                ("trace_lp2", C.c_void_p), ("trace_acc", C.c_void_p), ("aux", C.c_void_p),
                ("n_sample_chains", C.c_int64)]


class EngineError(RuntimeError):
    """A pbu_* call failed (status, message)."""

    def __init__(self, status, message):
        super().__init__("pybitup_b200 C ABI error %d: %s" % (status, message))
        self.status = status


class UnsupportedProblem(EngineError, ValueError):
    """No device kernel exists for the requested (model, P, d, ...) -- there is no CPU fallback."""


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "pybitup_b200: CUDA library %s is missing. Build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or `make -C pybitup_b200/csrc`). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    E = C.c_void_p
    sig = {
        "pbu_abi_version": (C.c_int, []),
        "pbu_last_error": (C.c_char_p, []),
        "pbu_list_kernels": (C.c_int, [C.c_char_p, C.c_int]),
        "pbu_engine_create": (C.c_int, [C.POINTER(pbu_problem), C.POINTER(pbu_sampler_cfg), C.c_int64, C.c_int,
                                        C.POINTER(E)]),
        "pbu_engine_destroy": (C.c_int, [E]),
        "pbu_engine_set_stream": (C.c_int, [E, C.c_void_p]),
        "pbu_engine_init_chains": (C.c_int, [E, _dp, _dp]),
        "pbu_engine_set_streams": (C.c_int, [E, _dp, C.c_int64, _dp, C.c_int64]),
        "pbu_engine_run": (C.c_int, [E, C.c_int64, C.c_int64, C.POINTER(pbu_run_outputs)]),
        "pbu_engine_n_keep": (C.c_int64, [E, C.c_int64, C.c_int64]),
        "pbu_engine_synchronize": (C.c_int, [E]),
        "pbu_engine_get_state": (C.c_int, [E, _dp, _dp]),
        "pbu_engine_get_counters": (C.c_int, [E, _lp, _dp, _lp, _ip]),
        "pbu_engine_read_state": (C.c_int, [E, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
        "pbu_engine_get_proposal": (C.c_int, [E, _dp, _dp, _dp]),
        "pbu_engine_iteration": (C.c_int64, [E]),
        "pbu_engine_get_isde_state": (C.c_int, [E, _dp, _dp]),
        "pbu_engine_get_launch": (C.c_int, [E, _ip, _ip, _ip, _ip, _lp]),
        "pbu_engine_get_variant": (C.c_int, [E, _ip, _ip, _ip, _ip]),
        "pbu_eval_logpost": (C.c_int, [E, _dp, C.c_int64, _dp]),
        "pbu_eval_model": (C.c_int, [E, _dp, C.c_int64, _dp]),
        "pbu_engine_chain_sums_len": (C.c_int, [E]),
        "pbu_engine_chain_sums": (C.c_int, [E, _dp, _dp]),
        "pbu_engine_reset_moments": (C.c_int, [E]),
        "pbu_engine_pooled_stats_len": (C.c_int, [E]),
        "pbu_engine_chain_sums_dev": (C.c_int, [E, C.c_void_p, C.c_void_p, C.c_int]),
        "pbu_engine_pooled_mean_dev": (C.c_int, [E, C.c_void_p, C.c_void_p, C.c_void_p]),
        "pbu_engine_pooled_finalize_dev": (C.c_int, [E, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p]),
        "pbu_engine_set_proposal": (C.c_int, [E, _dp]),
        "pbu_engine_get_map": (C.c_int, [E, _dp, _dp]),
        "pbu_propagation_create": (C.c_int, [E, C.c_int64, C.POINTER(E)]),
        "pbu_propagation_destroy": (C.c_int, [E]),
        "pbu_propagation_eval_host": (C.c_int, [E, _dp]),
        "pbu_propagation_eval_gaussian": (C.c_int, [E, _dp, _dp, C.c_uint64, C.c_uint64]),
        "pbu_propagation_buffers": (C.c_int, [E, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
        "pbu_propagation_moments": (C.c_int, [E, _dp, _dp, _dp]),
        "pbu_propagation_select_begin": (C.c_int, [E, _lp, C.c_int, _dp, _dp, C.POINTER(C.c_int)]),
        "pbu_propagation_select_pass": (C.c_int, [E, C.c_int, C.POINTER(C.c_uint64)]),
        "pbu_propagation_select_update": (C.c_int, [E, C.c_int, C.POINTER(C.c_uint64)]),
        "pbu_propagation_select_pass_dev": (C.c_int, [E, C.c_int, C.c_void_p]),
        "pbu_propagation_select_update_dev": (C.c_int, [E, C.c_int, C.c_void_p]),
        "pbu_propagation_select_result": (C.c_int, [E, _dp]),
        "pbu_propagation_bracket_sample": (C.c_int, [E, _dp, C.c_int, C.c_int, _dp]),
        "pbu_propagation_bracket_pass": (C.c_int, [E, _dp, C.c_int, C.c_int64, _dp, _dp, _dp, _lp, _lp, C.POINTER(C.c_int)]),
        "pbu_propagation_bracket_select_begin": (C.c_int, [E, _lp, _dp, C.POINTER(C.c_int)]),
        "pbu_propagation_get_evals": (C.c_int, [E, C.c_int64, C.c_int64, _dp]),
        "pbu_propagation_get_samples": (C.c_int, [E, C.c_int64, C.c_int64, _dp]),
        "pbu_measure_fp64_peak": (C.c_int, [C.c_int, C.c_int, C.c_int, _dp, _dp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)        # AttributeError if the library does not export the symbol
        fn.restype = res
        fn.argtypes = args
    return lib, sig


lib, SIGNATURES = _load()
EXPORTED = tuple(SIGNATURES)


def check(status):
    if status == PBU_OK:
        return
    msg = (lib.pbu_last_error() or b"").decode("utf-8", "replace")
    if status == PBU_ERR_NOT_PD:
        raise np.linalg.LinAlgError(msg)          # what scipy.linalg.cholesky raises in the reference
    if status == PBU_ERR_UNSUPPORTED:
        raise UnsupportedProblem(status, msg)
    raise EngineError(status, msg)


def dptr(a):
    """float64 C-contiguous numpy array -> double* (the array must outlive the call)."""
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def list_kernels():
    buf = C.create_string_buffer(1 << 16)
    n = lib.pbu_list_kernels(buf, len(buf))
    return n, buf.value.decode()


def measure_fp64_peak(device=0, iters=4096, reps=5):
    """Achieved FP64 TFLOP/s of a register-resident DFMA loop (FMA = 2 flops); reps < 0: sustained mean of -reps
    back-to-back launches."""
    tf, ms = C.c_double(0.0), C.c_double(0.0)
    check(lib.pbu_measure_fp64_peak(device, iters, reps, C.byref(tf), C.byref(ms)))
    return tf.value, ms.value
