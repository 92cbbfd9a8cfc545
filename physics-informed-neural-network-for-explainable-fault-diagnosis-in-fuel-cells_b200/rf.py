"""RF(t) risk function -- drop-in for ``estimate_mu_sigma_normal``, ``compute_rf_time_series``
and ``find_first_alarm_index`` of ``04_risk_function_early_warning_index.py`` (04:181-300).

Inputs are ``comprehensive_results`` matrices ``[N, 22]`` (one stack) or ``[S, N, 22]`` (a
fleet of independent stacks, BASELINE config 5); the two first-order recurrences are scans on
the device in float64 like the reference.  Constants default to the script's (04:84-101,163).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi, kernels as K
from ._abi import PinnRfParams, check, ptr

RF_Z_SAFE, RF_LAMBDA_DECAY = 2.0, 0.9971                     # 04:97-98
RF_K_LOGISTIC, RF_C0_LOGISTIC, RF_C_MAX = 0.0005, 500.0, 1000.0   # 04:99-101
RF_ALPHA_SMOOTH, RF_WARN_THRESHOLD = 0.2, 0.3                # 04:112,163
RF_RES_KEYS = ("res", "pV", "pT", "pH", "pO")                # 04:80  (columns 12..16)
NORMAL_LABELS = (0,)                                         # 04:79
RF_LAYER_CONFIG = {"voltage": ["res", "pV"], "gas": ["pH", "pO"], "temp": ["pT"]}     # 04:84-88
RF_LAYER_WEIGHTS = {"voltage": 1.0, "gas": 1.0, "temp": 1.0}                          # 04:92-96
RF_FEATURE_WEIGHTS = np.array([1.0, 1.0, 1.0, 1.0, 1.0])                              # 04:90
RF_P_LAYER = 2.0                                                                      # 04:97


def _only_default(name, value, default):
    """The kernels implement the script's configuration; other layer layouts are not built."""
    same = np.array_equal(np.asarray(value, dtype=object), np.asarray(default, dtype=object)) \
        if not isinstance(default, dict) else value == default
    if value is not None and not same:
        raise NotImplementedError(f"b200pinn.rf: only the reference's default `{name}` is implemented on the device")


def _as_device(results, device=None):
    t = results if torch.is_tensor(results) else torch.as_tensor(np.ascontiguousarray(results, np.float64))
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("b200pinn.rf: needs a CUDA device; there is no CPU path")
        t = t.to(device or torch.device("cuda", torch.cuda.current_device()))
    t = t.to(torch.float64).contiguous()
    return (t.unsqueeze(0), True) if t.dim() == 2 else (t, False)


def rf_device(results: torch.Tensor, z_safe=RF_Z_SAFE, lambda_decay=RF_LAMBDA_DECAY, k_logistic=RF_K_LOGISTIC,
              C0_logistic=RF_C0_LOGISTIC, C_max=RF_C_MAX, alpha_smooth=RF_ALPHA_SMOOTH,
              warn_threshold=RF_WARN_THRESHOLD, mu_sigma=None, want_extra=False):
    """``results``: CUDA float64 ``[S, N, 22]`` (``comprehensive_results`` rows) or the compact ``[S, N, 6]`` form
    (columns 12..17 only, as ``export_rows_device(..., want_rf_cols=True)`` emits).  Returns dict of CUDA tensors:
    ``mu_sigma [S,10]``, ``rf_inst, rf_smooth [S,N]``, ``first_alarm [S]`` (-1 = never), optionally ``C, S_tot``."""
    S_, n = results.shape[0], results.shape[1]
    row_cols = results.shape[2]
    if row_cols not in (22, 6):
        raise ValueError("b200pinn.rf: rows must have 22 columns (comprehensive_results) or 6 (compact RF columns)")
    col0 = 12 if row_cols == 22 else 0
    dev = results.device
    L = _abi.lib()
    nb = L.pinn_rf_workspace_bytes(n, S_)
    ws = K._workspace("rf", nb, dev)
    prm = PinnRfParams(z_safe, lambda_decay, k_logistic, C0_logistic, C_max, alpha_smooth, warn_threshold)
    out = {}
    with torch.cuda.device(dev):
        if mu_sigma is None:
            mu_sigma = torch.empty(S_, 10, device=dev, dtype=torch.float64)
            check(L.pinn_rf_stats(ptr(results), n, S_, row_cols, col0, ptr(mu_sigma), ptr(ws), nb, K._stream()), "pinn_rf_stats")
            K.LAUNCHES += 2
        out["mu_sigma"] = mu_sigma
        out["rf_inst"] = torch.empty(S_, n, device=dev, dtype=torch.float64)
        out["rf_smooth"] = torch.empty(S_, n, device=dev, dtype=torch.float64)
        out["first_alarm"] = torch.empty(S_, device=dev, dtype=torch.int64)
        if want_extra:
            out["C"] = torch.empty(S_, n, device=dev, dtype=torch.float64)
            out["S_tot"] = torch.empty(S_, n, device=dev, dtype=torch.float64)
        check(L.pinn_rf_series(ptr(results), n, S_, row_cols, col0, ptr(mu_sigma), C.byref(prm), ptr(out["rf_inst"]), ptr(out["rf_smooth"]),
                               ptr(out.get("C")), ptr(out.get("S_tot")), ptr(out["first_alarm"]), ptr(ws), nb, K._stream()),
              "pinn_rf_series")
        K.LAUNCHES += 6
    return out


def estimate_mu_sigma_normal(results, res_keys=RF_RES_KEYS, normal_labels=NORMAL_LABELS):
    """04:181-197 -> ``(mu[5], sigma[5])`` numpy (order res, pV, pT, pH, pO)."""
    _only_default("res_keys", tuple(res_keys), RF_RES_KEYS)
    _only_default("normal_labels", tuple(normal_labels), NORMAL_LABELS)
    r, _ = _as_device(results)
    dev = r.device
    L = _abi.lib()
    nb = L.pinn_rf_workspace_bytes(r.shape[1], r.shape[0])
    ws = K._workspace("rf", nb, dev)
    ms = torch.empty(r.shape[0], 10, device=dev, dtype=torch.float64)
    with torch.cuda.device(dev):
        check(L.pinn_rf_stats(ptr(r), r.shape[1], r.shape[0], 22, 12, ptr(ms), ptr(ws), nb, K._stream()), "pinn_rf_stats")
    m = ms[0].cpu().numpy()
    return m[:5].copy(), m[5:].copy()


def compute_rf_time_series(results, mu, sigma, res_keys=RF_RES_KEYS, feature_weights=RF_FEATURE_WEIGHTS,
                           layer_config=RF_LAYER_CONFIG, layer_weights=RF_LAYER_WEIGHTS, p_layer=RF_P_LAYER,
                           z_safe=RF_Z_SAFE, lambda_decay=RF_LAMBDA_DECAY, k_logistic=RF_K_LOGISTIC,
                           C0_logistic=RF_C0_LOGISTIC, C_max=RF_C_MAX, alpha_smooth=RF_ALPHA_SMOOTH):
    """04:201-285 -> ``(RF_inst, RF_smooth, extra)`` numpy, ``extra`` holding ``S_tot`` and ``C``."""
    _only_default("res_keys", tuple(res_keys), RF_RES_KEYS)
    _only_default("feature_weights", feature_weights, RF_FEATURE_WEIGHTS)
    _only_default("layer_config", layer_config, RF_LAYER_CONFIG)
    _only_default("layer_weights", layer_weights, RF_LAYER_WEIGHTS)
    _only_default("p_layer", p_layer, RF_P_LAYER)
    r, _ = _as_device(results)
    ms = torch.tensor(np.concatenate([np.asarray(mu, np.float64), np.asarray(sigma, np.float64)])[None, :],
                      device=r.device)
    o = rf_device(r, z_safe, lambda_decay, k_logistic, C0_logistic, C_max, alpha_smooth, mu_sigma=ms, want_extra=True)
    # per-layer strengths (04:243-254) for API parity: plain elementwise torch on the five residual columns (not a hot path)
    mu_t, sg_t = ms[0, :5], ms[0, 5:]
    a_tr = torch.clamp(((r[0][:, 12:17] - mu_t) / sg_t).abs() - z_safe, min=0.0)
    a_tr = torch.where(torch.isnan(r[0][:, 12:17]), torch.full_like(a_tr, float("nan")), a_tr)
    idx = {k: i for i, k in enumerate(RF_RES_KEYS)}
    s_layers = {name: torch.sqrt((a_tr[:, [idx[k] for k in keys]] ** 2).sum(dim=1)).cpu().numpy()
                for name, keys in RF_LAYER_CONFIG.items()}
    return (o["rf_inst"][0].cpu().numpy(), o["rf_smooth"][0].cpu().numpy(),
            {"S_layers": s_layers, "S_tot": o["S_tot"][0].cpu().numpy(), "C": o["C"][0].cpu().numpy()})


def find_first_alarm_index(series, threshold, mode="above"):
    """04:289-300 (host helper for already-downloaded series)."""
    s = np.asarray(series)
    idx = np.where(s >= threshold)[0] if mode == "above" else np.where(s <= threshold)[0]
    return None if len(idx) == 0 else int(idx[0])
