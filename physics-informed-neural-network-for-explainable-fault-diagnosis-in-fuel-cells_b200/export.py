"""Export row writer -- drop-in for ``create_comprehensive_results_array_v2``,
``smooth_by_segments`` and ``create_fault_labels`` (01:1830-2047).

The reference un-scales on the host, calls ``get_MC_samples`` and the four ``net_f_*``
separately (each with its own host round trips), smooths with pandas and fills a numpy
array.  Here: one MC sweep (K4), one residual launch in export form (K3, all families + 22
column rows), one row-writer launch (K5) that un-scales, smooths per segment and assembles
the float64 ``[N, 22]`` matrix on the device; the host only receives the finished array.
Column meaning: 01:2162-2183.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi, kernels as K
from ._abi import PinnExportScalers, check, ptr
from .mc import mc_dropout_device

SMOOTH_WINDOW = 200          # 01:1972


def _export_scalers(scaler_X, scaler_Y) -> PinnExportScalers:
    s = PinnExportScalers()
    for j in range(_abi.N_IN):
        s.x_min[j] = float(scaler_X.min_[j])
        s.x_scale[j] = float(scaler_X.scale_[j])
    s.y_min = float(np.asarray(scaler_Y.min_).reshape(-1)[0])
    s.y_scale = float(np.asarray(scaler_Y.scale_).reshape(-1)[0])
    lo, hi = float(scaler_Y.feature_range[0]), float(scaler_Y.feature_range[1])
    dmin = float(np.asarray(scaler_Y.data_min_, np.float64).reshape(-1)[0])
    dmax = float(np.asarray(scaler_Y.data_max_, np.float64).reshape(-1)[0])
    s.scale_y = (hi - lo) / (dmax - dmin + 1e-12)            # 01:1924
    s.min_y = lo - dmin * s.scale_y                          # 01:1925
    return s


def export_rows_device(model, x, y, boundaries, n_labeled, mc_times, dropout, scaler_X, scaler_Y, masks=None,
                       window=SMOOTH_WINDOW, seed=None, sample_offset=0, pass_offset=0, want_rf_cols=False):
    """Device-level export of one stack: ``x [n,8]``, ``y [n]`` CUDA tensors (normalised);
    returns a CUDA float64 tensor ``[n, 22]``.  ``pass_offset``: first pass index of the sweep in the
    network's dropout stream (``get_MC_samples`` at the same ``_drop_calls`` draws the same masks).
    ``want_rf_cols``: also return the dense ``[n, 6]`` copy of columns 12..17 that ``rf.rf_device(..., compact=True)`` consumes
    (the fleet pipeline of config 5: the risk series then reads 48 dense bytes per row instead of sparse 176-byte rows)."""
    dnn = model.dnn
    n, dev = x.shape[0], x.device
    mc = mc_dropout_device(dnn, x, mc_times, float(dropout), seed=seed, sample_offset=sample_offset, pass_offset=pass_offset,
                           masks=masks)
    net = K.net_from_module(dnn)
    u, _ = K.mlp_forward(net, x)                              # eval-mode prediction feeding net_f_V (01:1944-1948)
    fam = _abi.FAM_V | _abi.FAM_TS | _abi.FAM_H | _abi.FAM_O
    _, cols = K.residuals(x, u, None, model._scalers(scaler_X), model._lambdas(), fam, want_cols=True)
    out = torch.empty(n, 22, device=dev, dtype=torch.float64)
    rf_cols = torch.empty(n, 6, device=dev, dtype=torch.float64) if want_rf_cols else None
    seg = torch.tensor(list(boundaries), device=dev, dtype=torch.int64) if boundaries else None
    sc = _export_scalers(scaler_X, scaler_Y)
    with torch.cuda.device(dev):
        check(_abi.lib().pinn_export_rows(ptr(x), ptr(y), ptr(mc["pred_mean"]), ptr(mc["a_u"]), ptr(mc["e_u"]), ptr(cols),
                                          ptr(seg), 0 if seg is None else seg.numel(), int(n_labeled), int(window),
                                          C.byref(sc), n, ptr(out), ptr(rf_cols), K._stream()), "pinn_export_rows")
    K.LAUNCHES += 1
    return (out, rf_cols) if want_rf_cols else out


def create_comprehensive_results_array_v2(model, dataset, mc_times=2000, dropout=0.2):
    """01:1877-2010: returns the host float64 ``[N, 22]`` array (``comprehensive_results``)."""
    if len(dataset) == 9:
        x_train, y_train, x_val, y_val, x_test, y_test, scaler_X, scaler_Y, data_info = dataset
    else:
        x_train, y_train, x_test, y_test, scaler_X, scaler_Y, data_info = dataset
    dev = model.device
    x = x_test.detach().to(dev, torch.float32).contiguous()
    y = y_test.detach().to(dev, torch.float32).reshape(-1).contiguous()
    n = x.shape[0]
    boundaries, n_labeled = None, 0
    if data_info and "boundary_lines" in data_info and len(data_info["boundary_lines"]) > 0:
        boundaries = list(data_info["boundary_lines"])
        if boundaries[-1] != n:                               # 01:1977-1978
            boundaries = boundaries + [n]
        n_labeled = len(data_info.get("fault_data_list", []))
    model.dnn.eval()
    masks = getattr(model.dnn, "_injected_mc", None)
    calls = getattr(model.dnn, "_drop_calls", 0)
    out = export_rows_device(model, x, y, boundaries, n_labeled, mc_times, dropout, scaler_X, scaler_Y, masks=masks,
                             pass_offset=calls)
    if hasattr(model.dnn, "_drop_calls"):
        model.dnn._drop_calls = calls + int(mc_times)
    return out.cpu().numpy()


def create_fault_labels(n_samples, data_info):
    """01:2013-2047 (host helper kept for API parity; the row writer labels on the device)."""
    labels = np.zeros(n_samples)
    if data_info and "boundary_lines" in data_info and "fault_data_list" in data_info:
        for i in range(len(data_info["fault_data_list"])):
            labels[data_info["boundary_lines"][i]:data_info["boundary_lines"][i + 1]] = i + 1
    return labels
