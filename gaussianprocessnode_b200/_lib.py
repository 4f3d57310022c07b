"""ctypes binding of libsgp.so -- the same C ABI a Julia host reaches with ``ccall`` (include/sgp.h).

There is no fallback: if the shared library is missing or no CUDA device is present, calls raise."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SGP_LIB_PATH", os.path.join(_HERE, "libsgp.so"))   # override: instrumented builds for profiling

c_double_p = ctypes.POINTER(ctypes.c_double)
c_float_p = ctypes.POINTER(ctypes.c_float)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_void_pp = ctypes.POINTER(ctypes.c_void_p)

# name -> (restype, argtypes); must list every symbol include/sgp.h declares (tests/test_abi.py checks it)
SIGNATURES = {
    "sgp_create": (ctypes.c_int, [c_void_pp, ctypes.c_int]),
    "sgp_destroy": (None, [ctypes.c_void_p]),
    "sgp_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "sgp_version": (ctypes.c_char_p, []),
    "sgp_pinned_alloc": (ctypes.c_int, [ctypes.c_size_t, c_void_pp]),
    "sgp_pinned_free": (None, [ctypes.c_void_p]),
    "sgp_set_kernel": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, c_double_p]),
    "sgp_set_inducing": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_double_p]),
    "sgp_set_data": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, c_double_p, c_double_p, c_double_p, c_double_p]),
    "sgp_set_targets": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p]),
    "sgp_set_data_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "sgp_sweep_psi": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "sgp_sweep_psi_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                          c_double_p, c_double_p]),
    "sgp_sweep_psi_host_packed": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                                 c_double_p, c_double_p, c_double_p]),
    "sgp_fetch_psi2_packed": (ctypes.c_int, [ctypes.c_void_p, c_double_p]),
    "sgp_sweep_psi_uncertain": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, c_double_p, c_double_p,
                                               ctypes.c_int, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "sgp_kuu_factor": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, c_double_p]),
    "sgp_kuu_solve": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_double_p]),
    "sgp_posterior_v": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p, ctypes.c_double, c_double_p, c_double_p, c_double_p]),
    "sgp_prior_set": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p]),
    "sgp_prior_set_isotropic": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double]),
    "sgp_posterior_v_stream": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_int, c_double_p, c_double_p, c_double_p]),
    "sgp_w_terms": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "sgp_predict_mean": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, c_double_p, c_double_p, c_double_p]),
    "sgp_predict_probit": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, c_double_p, c_double_p, ctypes.c_double, c_double_p, c_double_p, c_double_p]),
    "sgp_dense_timed": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int, c_float_p]),
    "sgp_fp64_peak": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_double_p]),
    "sgp_theta_objective": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p, ctypes.c_double, ctypes.c_double, c_double_p, c_double_p, c_double_p]),
    "sgp_comm_unique_id": (ctypes.c_int, [ctypes.c_char_p]),
    "sgp_comm_init": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p]),
    "sgp_uncertain_node_terms": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "sgp_in_logmessage": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, c_double_p, ctypes.c_int, c_double_p, c_double_p, c_double_p,
                          ctypes.c_double, c_double_p, c_double_p, c_double_p]),
    "sgp_sweep_timed": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_float_p, c_float_p]),
    "sgp_sweep_timed_flushed": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, c_float_p, c_float_p]),
    "sgp_last_sweep_info": (ctypes.c_int, [ctypes.c_void_p, c_int_p, c_int_p, c_int_p, c_int_p]),
    "sgp_debug_p2_plan": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_int_p, ctypes.c_int,
                                        c_int_p]),
    "sgp_sweep_debug_clocks": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.c_int, c_int_p]),
    "sgp_stats_dev": (ctypes.c_int, [ctypes.c_void_p, c_void_pp, c_void_pp, c_void_pp]),
}

_lib = None


def load():
    """Loads libsgp.so (building nothing: see build.py / __graft_entry__.build).  Raises if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libsgp.so not found at %s -- run `python -m gaussianprocessnode_b200.build` "
                           "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
