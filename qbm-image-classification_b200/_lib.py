"""ctypes binding of libqbm_b200.so (the C ABI declared in include/qbm_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# QBM_B200_LIB: an instrumented build of the same library (tools/probe_tile_prof.py); there is still no fallback
LIB_PATH = os.environ.get("QBM_B200_LIB") or os.path.join(HERE, "libqbm_b200.so")

QBM_OK = 0
QBM_EINVAL = -1
QBM_EUNSUPPORTED = -2
QBM_ECUDA = -3
QBM_EWORKSPACE = -4
QBM_SA_MAX_N = 2048

_c_i = ctypes.c_int
_c_ll = ctypes.c_longlong
_c_u64 = ctypes.c_uint64
_c_sz = ctypes.c_size_t
_c_p = ctypes.c_void_p
_c_u = ctypes.c_uint

# name -> (restype, argtypes); must list every symbol include/qbm_b200.h declares
SIGNATURES = {
    "qbm_version": (_c_i, []),
    "qbm_last_error": (ctypes.c_char_p, []),
    "qbm_device_info": (_c_i, [_c_p, _c_p, _c_p]),
    "qbm_qubo_to_ising": (_c_i, [_c_p, _c_i, _c_ll, _c_p, _c_p, _c_p, _c_p, _c_p]),
    "qbm_beta_schedule": (_c_i, [_c_p, _c_ll, _c_i, _c_p, _c_p]),
    "qbm_sa_workspace_bytes": (_c_sz, [_c_i, _c_ll]),
    "qbm_sa_workspace_bytes_two_phase": (_c_sz, [_c_i, _c_ll, _c_ll]),
    "qbm_sa_sample": (_c_i, [_c_p, _c_p, _c_i, _c_i, _c_ll, _c_p, _c_ll, _c_i, _c_i, _c_ll, _c_u64, _c_u64,
                             _c_p, _c_p, _c_p, _c_p, _c_sz, _c_u, _c_p]),
    "qbm_qubo_energy": (_c_i, [_c_p, _c_i, _c_ll, _c_p, _c_ll, _c_p, _c_p]),
    "qbm_phase_stats_workspace_bytes": (_c_sz, [_c_ll, _c_ll, _c_i]),
    "qbm_phase_stats": (_c_i, [_c_p, _c_ll, _c_ll, _c_i, _c_p, _c_p, _c_p, _c_sz, _c_p]),
    "qbm_gemm_tf32": (_c_i, [_c_p, _c_ll, _c_p, _c_ll, _c_i, _c_i, _c_i, ctypes.c_float, ctypes.c_float, _c_p, _c_ll, _c_p,
                             _c_i, _c_p, _c_ll, _c_p, _c_ll, _c_p]),
    "qbm_rbm_workspace_bytes": (_c_sz, [_c_i, _c_i, _c_i, _c_i]),
    "qbm_rbm_sample_hidden": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_p, _c_p]),
    "qbm_rbm_sample_visible": (_c_i, [_c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_p, _c_p]),
    "qbm_rbm_sample_class": (_c_i, [_c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_p, _c_p]),
    "qbm_rbm_class_given_x": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_p, _c_p, _c_sz, _c_p]),
    "qbm_rbm_disc_step": (_c_i, [_c_p] * 8 + [_c_i] * 4 + [ctypes.c_float] * 3 + [_c_p, _c_p, _c_p, _c_p, _c_sz, _c_p]),
    "qbm_rbm_cd1_step": (_c_i, [_c_p] * 8 + [_c_i] * 4 + [ctypes.c_float] * 2 + [_c_u64, _c_u, _c_p, _c_sz, _c_p]),
    "qbm_rbm_cd1_step_dev": (_c_i, [_c_p] * 8 + [_c_i] * 4 + [ctypes.c_float] * 2 + [_c_u64, _c_u, _c_p, _c_p, _c_sz, _c_p]),
    "qbm_rbm_grad_count": (_c_sz, [_c_i, _c_i, _c_i]),
    "qbm_rbm_disc_grad": (_c_i, [_c_p] * 6 + [_c_i] * 4 + [_c_p, _c_p, _c_p, _c_p, _c_sz, _c_p]),
    "qbm_rbm_cd1_grad": (_c_i, [_c_p] * 8 + [_c_i] * 4 + [_c_u64, _c_u, _c_p, _c_p, _c_sz, _c_p]),
    "qbm_rbm_cd1_grad_dev": (_c_i, [_c_p] * 8 + [_c_i] * 4 + [_c_u64, _c_u, _c_p, _c_p, _c_p, _c_sz, _c_p]),
    "qbm_rbm_peer_bytes": (_c_sz, [_c_i, _c_i, _c_i]),
    "qbm_peer_alloc": (_c_i, [_c_sz, _c_p]),
    "qbm_peer_free": (_c_i, [_c_p]),
    "qbm_peer_export": (_c_i, [_c_p, _c_p]),
    "qbm_peer_import": (_c_i, [_c_p, _c_p]),
    "qbm_peer_close": (_c_i, [_c_p]),
    "qbm_rbm_peer_error": (_c_i, [_c_p, _c_i, _c_i, _c_i, _c_p]),
    "qbm_rbm_apply_grad_peer": (_c_i, [_c_p] * 7 + [_c_i] * 6 + [ctypes.c_float, ctypes.c_float, _c_p, ctypes.c_float, _c_u, _c_p, _c_p]),
    "qbm_rbm_apply_grad": (_c_i, [_c_p] * 7 + [_c_i] * 3 + [ctypes.c_float, ctypes.c_float, _c_p, ctypes.c_float, _c_p]),
    "qbm_disc_param_count": (_c_ll, [_c_i] * 4),
    "qbm_disc_build_qubo": (_c_i, [_c_p, _c_i, _c_i, _c_i, _c_i, _c_p, _c_p, _c_ll, ctypes.c_double, _c_p, _c_p]),
    "qbm_disc_errors": (_c_i, [_c_i] * 5 + [_c_p, _c_p, _c_ll, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p]),
    "qbm_sgd_apply": (_c_i, [_c_p, _c_p, _c_ll, ctypes.c_double, ctypes.c_double, _c_p]),
    "qbm_convdeep_num_pooled": (_c_i, [_c_i] * 5),
    "qbm_convdeep_context": (_c_i, [_c_p, _c_p, _c_ll, _c_i, _c_i, _c_i, _c_i, _c_i, _c_p, _c_p, _c_p, _c_p]),
    "qbm_convdeep_param_count": (_c_ll, [_c_i, _c_i, _c_p, _c_i, _c_i, _c_i, _c_i]),
    "qbm_convdeep_build_qubo": (_c_i, [_c_p, _c_i, _c_i, _c_p, _c_i, _c_i, _c_i, _c_i, _c_p, _c_i, _c_p, _c_p, _c_ll,
                                       ctypes.c_double, _c_p, _c_p]),
    "qbm_convdeep_errors": (_c_i, [_c_i, _c_i, _c_p, _c_i, _c_i, _c_i, _c_i, _c_i, _c_i, _c_p, _c_p, _c_p, _c_ll, _c_p, _c_p,
                                   _c_p, _c_p, _c_p, _c_p]),
    "qbm_test_philox": (_c_i, [_c_p, _c_p, _c_p, _c_ll, _c_p]),
    "qbm_test_neg_log": (_c_i, [_c_p, _c_p, _c_ll, _c_p]),
    "qbm_rbm_workspace_layout": (_c_i, [_c_i, _c_i, _c_i, _c_i, _c_p]),
    "qbm_probe_onchip_peaks": (_c_i, [_c_p, _c_p, _c_sz, _c_p]),
}

_lib = None


class QbmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libqbm_b200 error {code}: {msg}")
        self.code = code


def load() -> ctypes.CDLL:
    """Load the shared library (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python __graft_entry__.py` (or "
                "`python qbm-image-classification_b200/build.py`) to compile the CUDA extension; "
                "there is no CPU fallback")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != QBM_OK:
        msg = load().qbm_last_error().decode("utf-8", "replace")
        if rc in (QBM_EINVAL, QBM_EUNSUPPORTED, QBM_EWORKSPACE):
            raise ValueError(f"libqbm_b200 error {rc}: {msg}")
        raise QbmError(rc, msg)
