"""ctypes binding of ``libb200pinn.so`` (C ABI declared in ``include/b200pinn.h``).

There is no CPU fallback: if the shared library is missing this module raises at
import of the first symbol, and every entry point raises ``RuntimeError`` on a
non-zero return code.  The library is built in-tree by ``build.py`` /
``__graft_entry__.build()`` with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200pinn.so")

ABI_VERSION = 5
N_IN = 8
MAX_HIDDEN = 8
N_LAMBDA = 17

# families / flags / slots: keep in sync with include/b200pinn.h (tests check the header)
FAM_V, FAM_TS, FAM_T, FAM_H, FAM_O, FAM_DATA = 1, 2, 4, 8, 16, 32
RES_ACCURATE_MATH, RES_NO_MODE_A, RES_NO_MODE_B, RES_NO_CLUSTER = 1, 2, 4, 8
# pinn_net_t.flags: per-call path selection and model options (the library keeps no process-global switches)
NET_NO_TC_FWD, NET_NO_TC_BWD, NET_NO_WIDE_TC, NET_PDL_NEVER, NET_PDL_ALWAYS, NET_NO_LOGVAR, NET_NO_FUSED_BWD = 1, 2, 4, 8, 16, 32, 64
NET_NO_WIDE_RESIDENT, NET_NO_TMA_INPUT, NET_NO_TC3 = 128, 256, 512
SUM_NAMES = ["N", "FV2", "EA2", "DATA2", "GA1", "GA2", "GA3", "GB1", "GB2", "GB3",
             "FT2", "FTABS", "GT1", "GT3", "GT5", "FTE2",
             "FH2", "GH1", "GH2", "GH3", "HACT", "HTGT",
             "FO2", "GO1", "GO2", "GO3", "OACT", "OTGT"]
S = {n: i for i, n in enumerate(SUM_NAMES)}
S_COUNT = len(SUM_NAMES)
COL_NAMES = ["FV", "VACT", "VOHM", "VCONC", "ENERNST", "VEST5", "I", "VOUT5",
             "FTS", "TS_PRED", "T_REAL", "FT", "T_PRED",
             "FH", "H_ACT", "H_TGT", "I_TOTAL",
             "FO", "O_ACT", "O_TGT", "O_Q", "O2"]
COL = {n: i for i, n in enumerate(COL_NAMES)}
C_COUNT = len(COL_NAMES)


class PinnNet(C.Structure):
    _fields_ = [("n_in", C.c_int32), ("width", C.c_int32), ("n_hidden", C.c_int32), ("flags", C.c_int32),
                ("W", C.c_void_p * MAX_HIDDEN), ("b", C.c_void_p * MAX_HIDDEN),
                ("Wp", C.c_void_p), ("bp", C.c_void_p), ("Wv0", C.c_void_p), ("bv0", C.c_void_p),
                ("Wv1", C.c_void_p), ("bv1", C.c_void_p), ("Wv2", C.c_void_p), ("bv2", C.c_void_p)]


class PinnDropout(C.Structure):
    _fields_ = [("p", C.c_float), ("reserved", C.c_int32), ("seed", C.c_uint64),
                ("sample_offset", C.c_int64), ("pass_offset", C.c_int64),
                ("mask_sample_stride_n", C.c_int64), ("masks", C.c_void_p)]


class PinnScalers(C.Structure):
    _fields_ = [("x_inv_scale", C.c_float * N_IN), ("x_off", C.c_float * N_IN),
                ("y_inv_scale", C.c_float), ("y_off", C.c_float),
                ("scale_y", C.c_float), ("min_y", C.c_float),
                ("p_h2o", C.c_float), ("reserved", C.c_float)]


class PinnExportScalers(C.Structure):
    _fields_ = [("x_min", C.c_double * N_IN), ("x_scale", C.c_double * N_IN), ("y_min", C.c_double),
                ("y_scale", C.c_double), ("min_y", C.c_double), ("scale_y", C.c_double)]


class PinnRfParams(C.Structure):
    _fields_ = [("z_safe", C.c_double), ("lambda_decay", C.c_double), ("k_logistic", C.c_double),
                ("c0_logistic", C.c_double), ("c_max", C.c_double), ("alpha_smooth", C.c_double),
                ("warn_threshold", C.c_double)]


_vp, _i64, _i32, _u32, _sz, _dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32, C.c_size_t, C.c_double
_SIGNATURES = {
    "pinn_abi_version": (C.c_int, []),
    "pinn_device_sm_count": (C.c_int, []),
    "pinn_adam_step_p2p": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _u32, _vp, _vp, _i64, _vp, _dbl, _dbl, _i64, _vp]),
    "pinn_error_string": (C.c_char_p, [C.c_int]),
    "pinn_param_count": (_i64, [_i32, _i32]),
    "pinn_mlp_fwd_workspace_bytes": (_sz, [_i32, _i32, _i64]),
    "pinn_mlp_bwd_workspace_bytes": (_sz, [_i32, _i32, _i64]),
    "pinn_mlp_bwd_workspace_bytes_flags": (_sz, [_i32, _i32, _i64, _i32]),
    "pinn_mlp_fwd_workspace_bytes_flags": (_sz, [_i32, _i32, _i64, _i32]),
    "pinn_mc_workspace_bytes_flags": (_sz, [_i32, _i32, _i64, _i32]),
    "pinn_mc_workspace_bytes": (_sz, [_i32, _i32, _i64]),
    "pinn_residuals_workspace_bytes": (_sz, [_i64]),
    "pinn_mlp_fwd": (C.c_int, [C.POINTER(PinnNet), _vp, _i64, C.POINTER(PinnDropout), _vp, _vp, _vp, _sz, _vp]),
    "pinn_mlp_bwd": (C.c_int, [C.POINTER(PinnNet), _vp, _i64, C.POINTER(PinnDropout), _vp, _vp, _vp, _i64,
                               _vp, _vp, _vp, _sz, _vp]),
    "pinn_train_dnn_step": (C.c_int, [C.POINTER(PinnNet), _vp, _i64, C.POINTER(PinnDropout), _vp, _i64, _vp, _vp, _vp, _vp,
                                      _dbl, _dbl, _i64, _vp, _vp, _vp, _sz, _vp]),
    "pinn_train_dnn_steps": (C.c_int, [C.POINTER(PinnNet), _vp, _i64, C.POINTER(PinnDropout), _vp, _i64, _vp, _vp, _vp, _vp,
                                       _dbl, _dbl, _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "pinn_dp_bucket_words": (_i64, [_i32, _i32, _i32]),
    "pinn_train_dnn_steps_dp": (C.c_int, [C.POINTER(PinnNet), _vp, _i64, C.POINTER(PinnDropout), _vp, _i64, _vp, _vp, _vp, _vp,
                                          _dbl, _dbl, _i64, _i64, _vp, _i32, _i32, _u32, _vp, _vp, _vp, _sz, _vp]),
    "pinn_residuals": (C.c_int, [_vp, _vp, _vp, _i64, C.POINTER(PinnScalers), _vp, _u32, _u32, _vp, _vp,
                                 _vp, _vp, _vp, _sz, _vp]),
    "pinn_mc_dropout": (C.c_int, [C.POINTER(PinnNet), _vp, _i64, _i32, C.POINTER(PinnDropout),
                                  _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pinn_export_rows": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, C.POINTER(PinnExportScalers),
                                   _i64, _vp, _vp, _vp]),
    "pinn_rf_workspace_bytes": (_sz, [_i64, _i32]),
    "pinn_rf_stats": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "pinn_rf_series": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, C.POINTER(PinnRfParams), _vp, _vp, _vp, _vp, _vp, _vp, _sz,
                                 _vp]),
    "pinn_gmm_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "pinn_gmm_pass": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pinn_adam_step": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _dbl, _dbl, _i64, _dbl, _vp, _vp, _vp,
                                 _i32, _vp]),
    "pinn_scalar_phase_workspace_bytes": (_sz, []),
    "pinn_scalar_phase": (C.c_int, [_vp, _vp, _vp, _i64, C.POINTER(PinnScalers), _vp, _u32, _u32, _i32, _i32,
                                    C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_float), _vp, _vp, _vp,
                                    _dbl, _dbl, _i64, _i64, _vp, _vp, _sz, _vp]),
    "pinn_adam_step_from_sums": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _dbl, _dbl, _i64, _vp, _vp,
                                           _vp]),
}
EXPORTS = tuple(_SIGNATURES)

_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"b200pinn: {LIB_PATH} not found -- build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a).  There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if handle.pinn_abi_version() != ABI_VERSION:
            raise RuntimeError("b200pinn: ABI version mismatch between _abi.py and libb200pinn.so")
        _lib = handle
    return _lib


def check(code: int, what: str):
    if code != 0:
        msg = lib().pinn_error_string(code)
        raise RuntimeError(f"b200pinn: {what} failed ({code}): {msg.decode() if msg else '?'}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())
