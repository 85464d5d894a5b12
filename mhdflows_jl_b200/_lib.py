"""ctypes binding of include/mhdflows_b200.h.  There is no CPU fallback: if the CUDA library is
missing or no device is present every entry point fails loudly."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MHDF_LIB") or os.path.join(HERE, "libmhdflows_b200.so")   # MHDF_LIB: A/B builds when tuning

OK, ERR_INVALID, ERR_CUDA, ERR_NCCL, ERR_NONFINITE, ERR_STATE = 0, -1, -2, -3, -4, -5
F32, F64 = 0, 1
HD, MHD, EMHD = 0, 1, 2
RK4, LSRK54, HM89 = 0, 1, 2
FRESH, STALE, STAGE = 0, 1, 2
FORCING_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_double)      # mhdf_forcing_fn

# every symbol include/mhdflows_b200.h declares
SYMBOLS = [
    "mhdf_create", "mhdf_destroy", "mhdf_last_error", "mhdf_nccl_unique_id", "mhdf_set_real", "mhdf_get_real",
    "mhdf_set_spectral", "mhdf_get_spectral", "mhdf_step", "mhdf_calcN", "mhdf_set_dt", "mhdf_set_clock",
    "mhdf_get_clock", "mhdf_cfl_dt", "mhdf_energy", "mhdf_helicity", "mhdf_spectrum", "mhdf_stale_stats",
    "mhdf_step_timed", "mhdf_stepper_stats", "mhdf_profile", "mhdf_profile_get", "mhdf_launch_count", "mhdf_info",
    "mhdf_ipc_blob_size", "mhdf_ipc_export", "mhdf_ipc_import", "mhdf_set_forcing",
    "mhdf_set_forcing_spectral", "mhdf_set_forcing_callback",
    "mhdf_set_forcing_a99", "mhdf_forcing_a99_calls", "mhdf_div_correction", "mhdf_set_vp_field",
    "mhdf_set_random_phase", "mhdf_scale_decomposition", "mhdf_vector_potential", "mhdf_correlation", "mhdf_set_forcing_nd",
]
A99_HOST, A99_GPU = 1, 2


class Config(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("Lx", C.c_double), ("Ly", C.c_double), ("Lz", C.c_double),
                ("nu", C.c_double), ("eta", C.c_double), ("n_nu", C.c_int), ("dt", C.c_double),
                ("physics", C.c_int), ("stepper", C.c_int), ("dtype", C.c_int), ("device", C.c_int),
                ("rank", C.c_int), ("nranks", C.c_int), ("nccl_id", C.c_void_p), ("vp", C.c_int), ("nd", C.c_int)]


class A99(C.Structure):   # mhdf_a99
    _fields_ = [("variant", C.c_int), ("amp", C.c_double), ("kf", C.c_double), ("sigma2", C.c_double), ("b", C.c_double),
                ("seed", C.c_ulonglong), ("call", C.c_ulonglong)]


class MHDFlowsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mhdflows_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MHDFlowsError(ERR_STATE, f"{LIB_PATH} is missing: build it with `python -m mhdflows_jl_b200.build` "
                                       "(nvcc, sm_100a).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i, d, ll = C.c_void_p, C.c_int, C.c_double, C.c_longlong
    pd, pll, pi = C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_int)
    sig = {
        "mhdf_create": (i, [C.POINTER(Config), C.POINTER(vp)]),
        "mhdf_destroy": (i, [vp]),
        "mhdf_last_error": (C.c_char_p, [vp]),
        "mhdf_nccl_unique_id": (i, [vp]),
        "mhdf_set_real": (i, [vp, i, vp]),
        "mhdf_get_real": (i, [vp, i, i, vp]),
        "mhdf_set_spectral": (i, [vp, i, vp]),
        "mhdf_get_spectral": (i, [vp, i, i, vp]),
        "mhdf_step": (i, [vp, i]),
        "mhdf_stepper_stats": (i, [vp, pll, pd]),
        "mhdf_calcN": (i, [vp, vp]),
        "mhdf_set_dt": (i, [vp, d]),
        "mhdf_set_clock": (i, [vp, d, ll]),
        "mhdf_get_clock": (i, [vp, pd, pd, pll]),
        "mhdf_cfl_dt": (i, [vp, d, d, pd]),
        "mhdf_energy": (i, [vp, i, pd, pd]),
        "mhdf_helicity": (i, [vp, pd, pd, pd]),
        "mhdf_spectrum": (i, [vp, i, pd, i]),
        "mhdf_stale_stats": (i, [vp, pd, pd]),
        "mhdf_step_timed": (i, [vp, i, pd]),
        "mhdf_profile": (i, [vp, i]),
        "mhdf_profile_get": (i, [vp, pd, pll, i]),
        "mhdf_launch_count": (ll, [vp]),
        "mhdf_info": (i, [vp, pi, pi, pi, pi, pi, pll]),
        "mhdf_ipc_blob_size": (i, [vp]),
        "mhdf_ipc_export": (i, [vp, vp]),
        "mhdf_ipc_import": (i, [vp, vp]),
        "mhdf_set_forcing": (i, [vp, i, vp]),
        "mhdf_set_forcing_spectral": (i, [vp, i, vp]),
        "mhdf_set_forcing_callback": (i, [vp, FORCING_FN, vp]),
        "mhdf_set_forcing_a99": (i, [vp, C.POINTER(A99)]),
        "mhdf_forcing_a99_calls": (i, [vp, C.POINTER(C.c_ulonglong)]),
        "mhdf_div_correction": (i, [vp, i]),
        "mhdf_set_vp_field": (i, [vp, i, vp]),
        "mhdf_set_random_phase": (i, [vp, i, C.c_ulonglong, d, d, d]),
        "mhdf_scale_decomposition": (i, [vp, i, i, d, d, vp]),
        "mhdf_vector_potential": (i, [vp, i, vp]),
        "mhdf_correlation": (i, [vp, i, i, vp]),
        "mhdf_set_forcing_nd": (i, [vp, d, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        if not hasattr(L, name) and os.environ.get("MHDF_LIB"):
            continue                     # A/B against an older tuning build
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(handle, code):
    if code != OK:
        msg = lib().mhdf_last_error(handle)
        raise MHDFlowsError(code, (msg or b"").decode("utf-8", "replace"))
