"""ctypes binding of libnlo_b200.so (include/nlo_b200.h).  No CPU fallback: if the library is
missing or no CUDA device is usable, every compute call raises."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libnlo_b200.so"

NLO_MAX_CIRCLES = 8
PREC_FP32_SIMT, PREC_TC_3XF16, PREC_AUTO = 0, 1, 2


class NloError(RuntimeError):
    pass


class SdfDesc(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("hidden", C.c_uint32), ("n_hidden_mats", C.c_uint32),
                ("act0", C.c_uint32), ("act", C.c_uint32), ("p0", C.c_float), ("p", C.c_float)]


class NlpDesc(C.Structure):
    _fields_ = [("dynamics", C.c_uint32), ("shape", C.c_uint32), ("N", C.c_uint32), ("use_slack", C.c_uint32),
                ("use_smooth", C.c_uint32), ("enforce_heading", C.c_uint32), ("sdf_mode", C.c_uint32),
                ("n_circles", C.c_uint32), ("dt", C.c_float), ("slack_penalty", C.c_float),
                ("smooth_weight", C.c_float), ("length", C.c_float), ("width", C.c_float), ("wheelbase", C.c_float),
                ("circles", (C.c_float * 4) * NLO_MAX_CIRCLES), ("obstacle_kind", C.c_uint32 * NLO_MAX_CIRCLES)]


class CompactCounts(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in ("n_g_var", "n_g_copy", "n_jac_var", "n_jac_const", "n_grad_var", "n_grad_lin")]


class RrtObstacle(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("first_vertex", C.c_uint32), ("n_vertices", C.c_uint32), ("pad", C.c_uint32),
                ("cx", C.c_double), ("cy", C.c_double), ("size", C.c_double), ("margin", C.c_double)]


class IpOptions(C.Structure):
    _fields_ = [("tol", C.c_double), ("max_iter", C.c_int), ("mu0", C.c_double), ("ls_multipliers", C.c_int), ("compact", C.c_int),
                ("verbose", C.c_int)]


class IpStats(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("iterations", "evaluations", "hessians", "trials", "compactions")] + [("trial_problems", C.c_longlong), ("phase_ms", C.c_double * 9), ("kkt_problems", C.c_longlong),
                                                                                                               ("kkt_retries", C.c_longlong), ("kkt_retry_hist", C.c_longlong * 16)]


_P = C.c_void_p
_F = C.c_void_p       # float* (device or host address passed as integer)
_SZ = C.c_size_t
_LL = C.c_longlong
_CASADI = [C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.POINTER(C.c_double)), C.POINTER(_LL), C.POINTER(C.c_double), C.c_int]

# name -> (restype, argtypes); every symbol include/nlo_b200.h declares
SIGNATURES = {
    "nlo_version": (C.c_int, []),
    "nlo_last_error": (C.c_char_p, []),
    "nlo_device_count": (C.c_int, []),
    "nlo_device_sm_count": (C.c_int, [C.c_int]),
    "nlo_sdf_weight_count": (_SZ, [C.POINTER(SdfDesc)]),
    "nlo_sdf_create": (C.c_int, [C.POINTER(SdfDesc), _F, _SZ, C.c_int, C.POINTER(_P)]),
    "nlo_sdf_load": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(_P)]),
    "nlo_sdf_save": (C.c_int, [C.c_char_p, C.POINTER(SdfDesc), _F, _SZ]),
    "nlo_sdf_destroy": (None, [_P]),
    "nlo_sdf_set_precision": (C.c_int, [_P, C.c_int]),
    "nlo_sdf_get_precision": (C.c_int, [_P]),
    "nlo_sdf_describe": (C.c_int, [_P, C.POINTER(SdfDesc)]),
    "nlo_sdf_eval": (C.c_int, [_P, _F, _F, _F, _SZ, _F, _F, _F, _P]),
    "nlo_sdf_hess": (C.c_int, [_P, _F, _F, _F, _SZ, _F, _F, _F, _P]),
    "nlo_sdf_eval_host": (C.c_int, [_P, _F, _F, _F, _SZ, _F, _F, _F]),
    "nlo_sdf_hess_host": (C.c_int, [_P, _F, _F, _F, _SZ, _F, _F, _F]),
    "nlo_launch_count": (C.c_ulonglong, []),
    "nlo_casadi_bind": (C.c_int, [_P]),
    "nlo_casadi_set_batch": (C.c_int, [_LL]),
    "nlo_nlp_create": (C.c_int, [C.POINTER(NlpDesc), _P, C.c_int, C.POINTER(_P)]),
    "nlo_nlp_destroy": (None, [_P]),
    "nlo_nlp_n_w": (_LL, [_P]),
    "nlo_nlp_n_g": (_LL, [_P]),
    "nlo_nlp_nnz_jac": (_LL, [_P]),
    "nlo_nlp_n_sdf_points": (_LL, [_P]),
    "nlo_nlp_jac_sparsity": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "nlo_nlp_reserve": (C.c_int, [_P, _SZ]),
    "nlo_nlp_eval": (C.c_int, [_P, _F, _SZ, _SZ, _F, _F, _F, _F, _P]),
    "nlo_nlp_eval_dynamics": (C.c_int, [_P, _F, _SZ, _SZ, _F, _F, _P]),
    "nlo_nlp_nnz_hess": (_LL, [_P]),
    "nlo_nlp_hess_sparsity": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "nlo_nlp_hess": (C.c_int, [_P, _F, _F, _F, _SZ, _SZ, _F, _P]),
    "nlo_nlp_jac_tvec": (C.c_int, [_P, _F, _F, _F, _SZ, _SZ, _F, _P]),
    "nlo_nlp_violation": (C.c_int, [_P, _F, _F, _F, _SZ, _SZ, _F, _P]),
    "nlo_nlp_eval_host": (C.c_int, [_P, _F, _SZ, _F, _F, _F, _F]),
    "nlo_nlp_compact_counts": (C.c_int, [_P, C.POINTER(CompactCounts)]),
    "nlo_nlp_compact_layout": (C.c_int, [_P] + [_F] * 9),
    "nlo_nlp_eval_host_compact": (C.c_int, [_P, _F, _SZ, _F, _F, _F, _F]),
    "nlo_ip_create": (C.c_int, [_P, _F, _F, _SZ, C.POINTER(_P)]),
    "nlo_ip_destroy": (None, [_P]),
    "nlo_ip_capacity": (_SZ, [_P]),
    "nlo_ip_solve": (C.c_int, [_P, _F, _SZ, C.POINTER(IpOptions), _F, _F, _F, _F, _F, _F, _F, C.POINTER(IpStats)]),
    "nlo_ip_kkt_step": (C.c_int, [_P, _F, _F, _F, _F, _F, _SZ, _SZ, _F, _F, _P]),
    "nlo_rrt_paths": (C.c_int, [C.POINTER(RrtObstacle), C.c_int, _F, C.c_int, _F, _F, _F, _F, _F, _SZ, C.c_double, C.c_int, C.c_double, C.c_double,
                                C.c_int, C.c_int, C.c_int, _F, _F]),
    "nlo_transpose_to_soa": (C.c_int, [_F, _F, _SZ, _SZ, _SZ, _P]),
    "nlo_transpose_to_aos": (C.c_int, [_F, _F, _SZ, _SZ, _SZ, _P]),
}
for _base in ("nn_sdf", "jac_nn_sdf", "adj1_nn_sdf", "jac_adj1_nn_sdf"):
    for _sfx in ("", "_batch"):
        _n = _base + _sfx
        SIGNATURES[_n] = (C.c_int, _CASADI)
        SIGNATURES[_n + "_n_in"] = (_LL, [])
        SIGNATURES[_n + "_n_out"] = (_LL, [])
        if _sfx or _base == "nn_sdf":
            SIGNATURES[_n + "_sparsity_in"] = (C.POINTER(_LL), [_LL])
            SIGNATURES[_n + "_sparsity_out"] = (C.POINTER(_LL), [_LL])
SIGNATURES["nn_sdf_incref"] = (None, [])
SIGNATURES["nn_sdf_decref"] = (None, [])

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise NloError(f"{LIB_PATH} is missing: build it with `python -m nlotrajectories_b200.build` "
                       "(nlotrajectories_b200 has no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise NloError(load().nlo_last_error().decode() or f"libnlo_b200 call failed (rc={rc})")


def require_gpu() -> int:
    n = load().nlo_device_count()
    if n <= 0:
        raise NloError("no CUDA device visible: nlotrajectories_b200 evaluates on B200 only (no CPU fallback)")
    return n


def ptr(t) -> int:
    """Address of a torch tensor / numpy array / None."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data
