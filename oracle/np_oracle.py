"""CPU oracle: numpy restatement of the reference's PINN hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may.  It restates, in plain numpy, the algorithms of
``/root/reference/01_train_pinn_multiphysics_model.py`` ("01:" below) that the
CUDA path replaces.  Every function runs in ``dtype=np.float32`` (mirrors the
reference's fp32 op order) or ``np.float64`` (tie-breaker when the reference's
own fp32 noise exceeds the tolerance, SURVEY.md section 8c / hazard H2).

Pinning: the reference ships no tests or golden vectors,
so this oracle is pinned against outputs of the reference itself, generated in
the build container by ``tests/golden/make_golden.py``, ``make_golden_export.py``
and ``make_golden_gmm.py`` (they import the unmodified reference scripts 01, 04
and 03 by path) and committed as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` holds the comparison.  The export / RF(t) / GMM
restatements at the end of the file follow 01:1830-2047, 04:181-300 and
03:360-426 (the latter on top of sklearn's published GaussianMixture arithmetic).

Parameter dictionaries use the reference's ``state_dict`` keys
(``layers.layer_{i}.weight`` ... see 01:399-419); weights are ``[out, in]``.
"""
from __future__ import annotations

import numpy as np

A_CELL = 270.0
FARADAY = 96485.0
R_GAS = 8.314
N_CELLS = 5.0
ALPHA = 0.5
GF_LIQ = -220170.0


# --------------------------------------------------------------------------- net
def n_hidden_layers(params) -> int:
    return sum(1 for k in params if k.startswith("layers.layer_") and k.endswith(".weight"))


def dropout_scale(p: float, dtype=np.float32):
    """Scaled-mask value torch uses: ``bernoulli_(1-p).div_(1-p)``."""
    keep = dtype(1.0 - p)
    return dtype(1.0) / keep


def softplus_log(v, dtype):
    """``log(softplus(v) + 1e-6)`` with torch's threshold 20 (01:432-434)."""
    v = v.astype(dtype)
    sp = np.where(v > 20, v, np.log1p(np.exp(np.minimum(v, dtype(20)))).astype(dtype))
    return np.log(sp + dtype(1e-6)).astype(dtype), sp


def dnn_forward(params, x, masks=None, dtype=np.float32, return_cache=False, logvar=True):
    """``DNN.forward`` (01:421-438).

    ``masks``: ``None`` (eval mode) or a list of ``L+1`` *scaled* masks (values in
    ``{0, 1/(1-p)}``): one per trunk layer ``[N,H]`` and one for the variance
    head ``[N,H/2]`` -- the order in which ``nn.Dropout`` modules fire.
    ``logvar=False``: the constructor flag of 01:428-436 -- the log-variance output is zeros.
    """
    logvar_flag = logvar
    P = {k: np.asarray(v, dtype=dtype) for k, v in params.items()}
    L = n_hidden_layers(P)
    h = np.asarray(x, dtype=dtype)
    acts, hs = [], [h]
    for i in range(L):
        a = np.tanh(h @ P[f"layers.layer_{i}.weight"].T + P[f"layers.layer_{i}.bias"])
        acts.append(a)
        h = a * masks[i].astype(dtype) if masks is not None else a
        hs.append(h)
    out = h @ P["predict.weight"].T + P["predict.bias"]
    a_v0 = np.tanh(h @ P["var_layers.0.weight"].T + P["var_layers.0.bias"])
    v0 = a_v0 * masks[L].astype(dtype) if masks is not None else a_v0
    v1 = np.tanh(v0 @ P["var_layers.3.weight"].T + P["var_layers.3.bias"])
    v = v1 @ P["var_layers.5.weight"].T + P["var_layers.5.bias"]
    logvar, sp = softplus_log(v, dtype)
    if not logvar_flag:
        logvar = np.zeros_like(out)                                     # 01:436
    if return_cache:
        return out, logvar, dict(acts=acts, hs=hs, a_v0=a_v0, v0=v0, v1=v1, v=v, sp=sp)
    return out, logvar


def aleatoric_loss(y, u, s, dtype=np.float32):
    """``aleatoric_loss`` (01:916-927)."""
    y, u, s = (np.asarray(t, dtype=dtype) for t in (y, u, s))
    nll = np.mean(dtype(0.5) * np.exp(-s) * (y - u) ** 2 + dtype(0.5) * s, dtype=dtype)
    return nll + dtype(0.01) * np.mean(np.abs(s), dtype=dtype)


def aleatoric_loss_grads(y, u, s, dtype=np.float64):
    """dL/du, dL/ds of :func:`aleatoric_loss`."""
    y, u, s = (np.asarray(t, dtype=dtype) for t in (y, u, s))
    n = dtype(y.shape[0])
    e = np.exp(-s)
    du = -e * (y - u) / n
    ds = (-0.5 * e * (y - u) ** 2 + 0.5 + 0.01 * np.sign(s)) / n
    return du, ds


def dnn_backward(params, x, masks, du, ds, dtype=np.float64, logvar=True):
    """Gradients of ``sum(du*out) + sum(ds*logvar)`` w.r.t. every DNN parameter
    (what ``loss.backward()`` at 01:953 produces through autograd).  ``logvar=False``: the log-variance is a
    constant (01:436), the variance head gets no gradient (keys absent, like ``grad is None``)."""
    P = {k: np.asarray(v, dtype=dtype) for k, v in params.items()}
    L = n_hidden_layers(P)
    _, _, c = dnn_forward(params, x, masks, dtype, return_cache=True)
    du = np.asarray(du, dtype=dtype)
    ds = np.asarray(ds, dtype=dtype) * (1.0 if logvar else 0.0)
    g = {}
    v, sp = c["v"], c["sp"]
    dsp_dv = np.where(v > 20, 1.0, 1.0 / (1.0 + np.exp(-v)))
    dv = ds * dsp_dv / (sp + dtype(1e-6))
    g["var_layers.5.weight"] = dv.T @ c["v1"]
    g["var_layers.5.bias"] = dv.sum(0)
    dz1 = (dv @ P["var_layers.5.weight"]) * (1 - c["v1"] ** 2)
    g["var_layers.3.weight"] = dz1.T @ c["v0"]
    g["var_layers.3.bias"] = dz1.sum(0)
    dv0 = dz1 @ P["var_layers.3.weight"]
    if masks is not None:
        dv0 = dv0 * masks[L].astype(dtype)
    dz0 = dv0 * (1 - c["a_v0"] ** 2)
    hL = c["hs"][L]
    g["var_layers.0.weight"] = dz0.T @ hL
    g["var_layers.0.bias"] = dz0.sum(0)
    g["predict.weight"] = du.T @ hL
    g["predict.bias"] = du.sum(0)
    dh = dz0 @ P["var_layers.0.weight"] + du @ P["predict.weight"]
    for i in reversed(range(L)):
        if masks is not None:
            dh = dh * masks[i].astype(dtype)
        dz = dh * (1 - c["acts"][i] ** 2)
        g[f"layers.layer_{i}.weight"] = dz.T @ c["hs"][i]
        g[f"layers.layer_{i}.bias"] = dz.sum(0)
        dh = dz @ P[f"layers.layer_{i}.weight"]
    if not logvar:
        g = {k: v for k, v in g.items() if not k.startswith("var_layers")}
    return g


# ------------------------------------------------------------------------ scalers
def inverse_transform(scaler, x, dtype=np.float32):
    """sklearn ``MinMaxScaler.inverse_transform`` on an fp32 array: each in-place
    op is computed in fp64 and rounded back to fp32."""
    x = np.asarray(x)
    if dtype == np.float64:
        return (x.astype(np.float64) - scaler.min_) / scaler.scale_
    t = (x.astype(np.float64) - scaler.min_).astype(np.float32)
    return (t.astype(np.float64) / scaler.scale_).astype(np.float32)


def y_affine(u_scal, dtype=np.float32):
    """``scale_y, min_y`` as rebuilt every step in ``train_lambda`` (01:1017-1022)."""
    lo, hi = float(u_scal.feature_range[0]), float(u_scal.feature_range[1])
    dmin = np.asarray(u_scal.data_min_, dtype=dtype)
    dmax = np.asarray(u_scal.data_max_, dtype=dtype)
    scale_y = dtype(hi - lo) / (dmax - dmin + dtype(1e-12))
    min_y = dtype(lo) - dmin * scale_y
    return scale_y.astype(dtype), min_y.astype(dtype)


# ---------------------------------------------------------------------- residuals
def p_h2o(dtype=np.float32):
    """Water-vapour pressure at the constant Tc=55 (01:745,752-753)."""
    Tc = dtype(55)
    x = dtype(-2.1794) + dtype(0.02953) * Tc - dtype(9.1837e-5) * (Tc ** 2) + dtype(1.4454e-7) * (Tc ** 3)
    return (dtype(10) ** x).astype(dtype) if hasattr(x, "astype") else dtype(dtype(10) ** x)


def net_f_V(x_norm, u_norm, x_scal, u_scal, lam, dtype=np.float32):
    """``net_f_V`` (01:724-765).  ``u_norm`` is the DNN prediction (normalised);
    ``lam = (lambda_1, lambda_2, lambda_3)``.  Returns the reference's 9-tuple."""
    f = dtype
    r = inverse_transform(x_scal, x_norm, f)
    i = r[:, 0:1] / f(A_CELL) + f(1e-5)
    T_out = r[:, 5:6]
    V_out = inverse_transform(u_scal, np.asarray(u_norm).reshape(-1, 1), f) / f(N_CELLS)
    lam1, lam2, lam3 = (f(v) for v in lam)
    P_H2 = r[:, 3:4] / f(101) + f(1)
    P_air = r[:, 4:5] / f(101) + f(1)
    Tk = T_out + f(273.15)
    PH2O = p_h2o(f)
    tkp = Tk ** f(1.334)
    pp_H2 = f(0.5) * (P_H2 / np.exp(f(1.653) * i / tkp) - PH2O)
    pp_O2 = P_air / np.exp(f(4.192) * i / tkp) - PH2O
    b = f(R_GAS) * Tk / (f(2.0) * f(ALPHA) * f(FARADAY))
    V_act = -b * np.log(i / lam2)
    V_ohm = -(i * lam1)
    V_conc = f(ALPHA) * b * np.log(f(1) - i / lam3)
    E = -f(GF_LIQ) / (f(2) * f(FARADAY)) - (f(R_GAS) * Tk) * np.log(PH2O / (pp_H2 * pp_O2 ** f(0.5))) / (f(2) * f(FARADAY))
    V_est = E + V_act + V_ohm + V_conc
    fV = V_est - V_out
    return tuple(np.asarray(t, dtype=f) for t in
                 (fV, V_act, V_ohm, V_conc, E, V_est * f(5), i, np.array([lam3]), V_out * f(5)))


def net_f_T_simple(x_norm, x_scal, lamT, dtype=np.float32):
    """``net_f_T_simple`` (01:869-914); ``lamT = (T1..T5)``; T2, T4 unused."""
    f = dtype
    r = inverse_transform(x_scal, x_norm, f)
    i = r[:, 0:1] / f(A_CELL) + f(1e-6)
    m = r[:, 1:2] + f(1e-6)
    T_in, T_real = r[:, 2:3], r[:, 5:6]
    I_t = i * f(A_CELL)
    T_pred = f(lamT[0]) * I_t + f(lamT[2]) * m + f(0.5) * T_in + f(lamT[4])
    return (T_real - T_pred).astype(f), T_pred.astype(f), T_real.astype(f)


def net_f_T(x_norm, u_norm, x_scal, u_scal, lamT, dtype=np.float32):
    """``net_f_T`` (01:767-867): Euler energy balance, rows t-1 -> t."""
    f = dtype
    n = x_norm.shape[0]
    if n < 2:
        z = np.zeros((n, 1), f)
        return z, z.copy(), z.copy()
    r = inverse_transform(x_scal, x_norm, f)
    i = r[:, 0:1] / f(A_CELL) + f(1e-5)
    m = r[:, 1:2] + f(1e-6)
    T_in, T_out = r[:, 2:3], r[:, 5:6]
    ip, mp, Tinp, Toutp = i[:-1], m[:-1], T_in[:-1], T_out[:-1]
    I_t = ip * f(A_CELL)
    V_rev = f(1.229) - f(0.0009) * ((Toutp + f(273.15)) - f(298.15))
    V_cell = inverse_transform(u_scal, np.asarray(u_norm).reshape(-1, 1)[:-1], f) / f(N_CELLS)
    Q_e = (I_t * V_rev - I_t * V_cell) * f(lamT[3])
    Q_c = mp * f(4180.0) * (Toutp - Tinp) * f(lamT[0])
    Q_r = f(20.0) * f(0.2) * (Toutp - f(25.0)) * f(lamT[2])
    dT = (Q_e - Q_c - Q_r) / f(lamT[1])
    T_next = Toutp + dT * f(0.1)
    T_pred = np.concatenate([T_out[0:1], T_next], axis=0)
    return (T_out - T_pred).astype(f), T_pred.astype(f), T_out.astype(f)


def net_f_H(x_norm, x_scal, lamH, dtype=np.float32):
    """``net_f_H`` (01:621-722); ``lamH = (H1..H4)``; H4 unused."""
    f = dtype
    r = inverse_transform(x_scal, x_norm, f)
    i = r[:, 0:1] / f(A_CELL) + f(1e-5)
    h2 = r[:, 6:7] + f(1e-6)
    I_t = i * f(A_CELL)
    Q = I_t / (f(2) * f(FARADAY)) * f(N_CELLS) * f(22.4) * f(60)
    Q = np.maximum(Q, f(1e-8))
    H1, H2, H3 = f(lamH[0]), f(lamH[1]), f(lamH[2])
    target = np.where(I_t <= H3, H1 + H2 * (I_t / f(100.0)), H1 + H2 * (H3 / f(100.0)))
    actual = h2 / Q
    return ((actual - target).astype(f), actual.astype(f), target.astype(f), I_t.astype(f),
            np.array([H3], f))


def net_f_O(x_norm, x_scal, lamO, dtype=np.float32):
    """``net_f_O`` (01:535-619); ``lamO = (O1..O4)``; O4 unused."""
    f = dtype
    r = inverse_transform(x_scal, x_norm, f)
    i = r[:, 0:1] / f(A_CELL) + f(1e-5)
    air = r[:, 7:8] + f(1e-6)
    I_t = i * f(A_CELL)
    Q = (I_t * f(N_CELLS)) / (f(4) * f(FARADAY)) * f(22.4) * f(60)
    Q = np.maximum(Q, f(1e-8))
    O1, O2, th = f(lamO[0]), f(lamO[1]), abs(f(lamO[2]))
    target = np.where(I_t <= th, O1 + O2 * (I_t / f(100.0)), O1 + O2 * (th / f(100.0)))
    target = np.clip(target, f(1.05), f(15.0))
    o2 = air * f(0.21)
    actual = o2 / Q
    fO = actual - target + np.maximum(f(1.0) - actual, f(0.0)) * f(10.0)
    return fO.astype(f), actual.astype(f), target.astype(f), Q.astype(f), o2.astype(f)


# ------------------------------------------------------- losses + lambda gradients
def lambda_losses(x_norm, y_norm, u_norm, x_scal, u_scal, lam, dnn_para, dtype=np.float64):
    """``train_lambda`` loss terms (01:1009-1034): returns
    ``(total, physics, data, grads[3])`` with analytic d total/d(lambda_1..3)
    (the residual is detached from the DNN, 01:734-737)."""
    f = dtype
    fV, _, _, _, _, V5, i, _, _ = net_f_V(x_norm, u_norm, x_scal, u_scal, lam, f)
    y = np.asarray(y_norm, f).reshape(-1, 1)
    u = np.asarray(u_norm, f).reshape(-1, 1)
    n = f(y.shape[0])
    scale_y, min_y = y_affine(u_scal, f)
    if dnn_para:
        physics = np.mean(fV ** 2, dtype=f)
        dV = 2 * fV / n                                  # d physics / d V_est (per cell)
    else:
        e = y - (V5 * scale_y + min_y)
        physics = np.mean(e ** 2, dtype=f)
        dV = -2 * f(5) * scale_y * e / n
    data = np.mean((y - u) ** 2, dtype=f)
    r = inverse_transform(x_scal, x_norm, f)
    Tk = r[:, 5:6] + f(273.15)
    b = f(R_GAS) * Tk / (f(2.0) * f(ALPHA) * f(FARADAY))
    l2, l3 = f(lam[1]), f(lam[2])
    g1 = np.sum(dV * (-i))
    g2 = np.sum(dV * (b / l2))
    g3 = np.sum(dV * (f(ALPHA) * b * i / (l3 * (l3 - i))))
    return physics + data, physics, data, np.array([g1, g2, g3], f)


def thermal_loss(x_norm, x_scal, lamT, dtype=np.float64):
    """``train_thermal`` loss (01:1109-1112) and d/d(T1,T3,T5); also mean|f|."""
    f = dtype
    fT, _, _ = net_f_T_simple(x_norm, x_scal, lamT, f)
    r = inverse_transform(x_scal, x_norm, f)
    I_t = (r[:, 0:1] / f(A_CELL) + f(1e-6)) * f(A_CELL)
    m = r[:, 1:2] + f(1e-6)
    g = np.array([np.mean(-2 * fT * I_t), np.mean(-2 * fT * m), np.mean(-2 * fT)], f)
    return np.mean(fT ** 2, dtype=f), g, np.mean(np.abs(fT), dtype=f)


def hydrogen_loss(x_norm, x_scal, lamH, dtype=np.float64):
    """``train_hydrogen`` loss (01:1357-1360) and d/d(H1,H2,H3)."""
    f = dtype
    fH, _, _, I_t, _ = net_f_H(x_norm, x_scal, lamH, f)
    H2, H3 = f(lamH[1]), f(lamH[2])
    lin = I_t <= H3
    g = np.array([np.mean(-2 * fH),
                  np.mean(-2 * fH * np.where(lin, I_t, H3) / f(100.0)),
                  np.mean(-2 * fH * np.where(lin, f(0), H2 / f(100.0)))], f)
    return np.mean(fH ** 2, dtype=f), g


def oxygen_loss(x_norm, x_scal, lamO, dtype=np.float64):
    """``train_oxygen`` loss (01:1207-1222) and d/d(O1,O2,O3)."""
    f = dtype
    fO, _, tgt, _, _ = net_f_O(x_norm, x_scal, lamO, f)
    r = inverse_transform(x_scal, x_norm, f)
    I_t = (r[:, 0:1] / f(A_CELL) + f(1e-5)) * f(A_CELL)
    O1, O2, O3 = f(lamO[0]), f(lamO[1]), f(lamO[2])
    th = abs(O3)
    lin = I_t <= th
    raw = np.where(lin, O1 + O2 * (I_t / f(100.0)), O1 + O2 * (th / f(100.0)))
    gate = ((raw >= f(1.05)) & (raw <= f(15.0))).astype(f)   # torch.clamp passes grad on the bounds
    sgn = np.sign(O3)
    g = np.array([np.mean(-2 * fO * gate),
                  np.mean(-2 * fO * gate * np.where(lin, I_t, th) / f(100.0)),
                  np.mean(-2 * fO * gate * np.where(lin, f(0), O2 * sgn / f(100.0)))], f)
    return np.mean(fO ** 2, dtype=f), g


# ----------------------------------------------------------------- optimiser steps
class Adam:
    """``torch.optim.Adam`` defaults (betas .9/.999, eps 1e-8) + ``StepLR``; used to
    restate the phase trainers' update (01:939-940, 999-1002, 1098-1102, ...)."""

    def __init__(self, n, lr, step_size=1000, gamma=0.8, dtype=np.float64):
        self.m = np.zeros(n, dtype)
        self.v = np.zeros(n, dtype)
        self.t = 0
        self.lr0, self.step_size, self.gamma = lr, step_size, gamma
        self.dtype = dtype

    def lr(self):
        return self.lr0 * self.gamma ** (self.t // self.step_size)

    def step(self, p, g, mask=None):
        """In torch a parameter whose ``.grad`` is ``None`` is skipped entirely;
        ``mask`` marks the entries that do receive a gradient."""
        f = self.dtype
        lr = self.lr()
        self.t += 1
        b1, b2 = 0.9, 0.999
        g = np.asarray(g, f)
        sel = np.ones(p.shape, bool) if mask is None else np.asarray(mask, bool)
        self.m[sel] = b1 * self.m[sel] + (1 - b1) * g[sel]
        self.v[sel] = b2 * self.v[sel] + (1 - b2) * g[sel] ** 2
        bc1, bc2 = 1 - b1 ** self.t, 1 - b2 ** self.t
        denom = np.sqrt(self.v[sel]) / np.sqrt(bc2) + 1e-8
        p = np.array(p, f)
        p[sel] = p[sel] - (lr / bc1) * self.m[sel] / denom
        return p


# ------------------------------------------------------------------ MC statistics
def mc_statistics(pred_eval_t, pred_drop_t, logvar_drop_t, dtype=np.float32):
    """Tail of ``get_MC_samples`` (01:1475-1491) on stacked ``(T, N, 1)`` arrays."""
    pe = np.asarray(pred_eval_t, dtype)
    pd = np.asarray(pred_drop_t, dtype)
    lv = np.asarray(logvar_drop_t, dtype)
    pred_mean = np.mean(pe, axis=0)
    a_u = np.sqrt(np.exp(np.mean(lv, axis=0)))
    e_u = np.sqrt(np.var(pd, axis=0))
    return pred_mean.squeeze(), a_u.squeeze(), e_u.squeeze()


def mc_dropout(params, x, masks_t, dtype=np.float32, logvar=True):
    """``get_MC_samples`` with injected masks.  ``masks_t[t]`` is the list of
    ``L+1`` scaled masks of pass ``t`` (the first forward of each ``predict``;
    the second, inside ``net_f_V``, is discarded: 01:1407)."""
    T = len(masks_t)
    pe, _ = dnn_forward(params, x, None, dtype, logvar=logvar)
    us, ss = [], []
    for t in range(T):
        u, s = dnn_forward(params, x, masks_t[t], dtype, logvar=logvar)
        us.append(u)
        ss.append(s)
    return mc_statistics(np.stack([pe] * T), np.stack(us), np.stack(ss), dtype)


# ------------------------------------------------------------- export rows (f1) / RF(t) (f2)
def moving_average_centered(arr, window):
    """``_moving_average_centered`` (01:1830-1846): pandas ``rolling(window, center=True,
    min_periods=1).mean()``; for an even window the span of element i is [i-w/2, i+w/2-1]
    (SURVEY section 5 [probe])."""
    a = np.asarray(arr, np.float64)
    n, half = a.shape[0], window // 2
    out = np.empty(n, np.float64)
    c = np.concatenate([[0.0], np.cumsum(a)])
    for i in range(n):
        lo, hi = max(0, i - half), min(n, i + (window - half))
        out[i] = (c[hi] - c[lo]) / (hi - lo)
    return out


def smooth_by_segments(values, boundary_lines, window):
    """``smooth_by_segments`` (01:1848-1872): smooth each [start, end) segment on its own."""
    v = np.asarray(values, np.float64)
    n = v.shape[0]
    if not boundary_lines or boundary_lines[-1] < n:
        return moving_average_centered(v, window)
    bl = [b for b in boundary_lines if 0 < b <= n] if boundary_lines[-1] != n else list(boundary_lines)
    out = np.empty(n, np.float64)
    for s, e in zip([0] + bl[:-1], bl):
        out[s:e] = moving_average_centered(v[s:e], window)
    return out


def fault_labels(n, boundary_lines, n_faults):
    """``create_fault_labels`` (01:2013-2047): 0 for the normal block, i+1 for fault segment i."""
    lab = np.zeros(n)
    for i in range(n_faults):
        lab[boundary_lines[i]:boundary_lines[i + 1]] = i + 1
    return lab


def export_rows(x_norm, y_norm, pred_mean, a_u, e_u, fV, fT, fH, fO, V5, T_pred, actH, actO, labels, x_scal, u_scal,
                boundaries, window=200):
    """22-column ``comprehensive_results`` of ``create_comprehensive_results_array_v2``
    (01:1907-2010); inputs are the fp32 arrays the reference has at that point."""
    n = x_norm.shape[0]
    out = np.zeros((n, 22))
    out[:, 0:8] = inverse_transform(x_scal, x_norm, np.float32)
    yr = inverse_transform(u_scal, np.asarray(y_norm).reshape(-1, 1), np.float32).flatten()
    lo, hi = float(u_scal.feature_range[0]), float(u_scal.feature_range[1])
    dmin, dmax = u_scal.data_min_.astype(np.float64), u_scal.data_max_.astype(np.float64)
    scale_y = (hi - lo) / (dmax - dmin + 1e-12)
    min_y = lo - dmin * scale_y
    pm = ((np.asarray(pred_mean) - min_y) / (scale_y + 1e-12)).reshape(-1)
    ale = (np.asarray(a_u) / (scale_y + 1e-12)).reshape(-1)
    epi = (np.asarray(e_u) / (scale_y + 1e-12)).reshape(-1)
    out[:, 8], out[:, 9] = yr, pm
    out[:, 10] = smooth_by_segments(ale, boundaries, window)
    out[:, 11] = smooth_by_segments(epi, boundaries, window)
    out[:, 12] = yr - pm
    for c, v in ((13, fV), (14, fT), (15, fH), (16, fO), (17, labels), (18, V5), (19, T_pred), (20, actH), (21, actO)):
        out[:, c] = np.asarray(v).reshape(-1)
    return out


RF_COLS = (12, 13, 14, 15, 16)      # res, pV, pT, pH, pO   (04:58-62, 80)


def rf_mu_sigma(results, normal_labels=(0,)):
    """``estimate_mu_sigma_normal`` (04:181-197)."""
    lab = results[:, 17].astype(int)
    R = results[np.isin(lab, normal_labels)][:, RF_COLS].astype(float)
    mu, sigma = np.nanmean(R, axis=0), np.nanstd(R, axis=0, ddof=1)
    sigma[sigma == 0] = 1e-6
    return mu, sigma


def rf_series(results, mu, sigma, z_safe=2.0, lam=0.9971, k=0.0005, C0=500.0, C_max=1000.0, alpha=0.2):
    """``compute_rf_time_series`` (04:201-285) with the script's constants (04:84-101): layers
    {res,pV}, {pH,pO}, {pT}, 2-norm per layer, unit weights."""
    R = results[:, RF_COLS].astype(float)
    a = np.maximum(0.0, np.abs((R - mu) / sigma) - z_safe)
    S = np.sqrt(a[:, 0] ** 2 + a[:, 1] ** 2) + np.sqrt(a[:, 3] ** 2 + a[:, 4] ** 2) + np.sqrt(a[:, 2] ** 2)
    n = R.shape[0]
    C = np.zeros(n)
    for t in range(1, n):
        C[t] = lam * C[t - 1] + S[t]
    L0 = 1.0 / (1.0 + np.exp(-k * (0.0 - C0)))
    Lm = 1.0 / (1.0 + np.exp(-k * (C_max - C0)))
    rf = np.clip((1.0 / (1.0 + np.exp(-k * (np.clip(C, 0.0, C_max) - C0))) - L0) / (Lm - L0), 0.0, 1.0)
    sm = np.zeros(n)
    sm[0] = rf[0]
    for t in range(1, n):
        sm[t] = alpha * rf[t] + (1 - alpha) * sm[t - 1]
    return rf, sm, S, C


def first_alarm(series, threshold):
    """``find_first_alarm_index`` (04:289-300), mode 'above'; -1 when never reached."""
    idx = np.where(np.asarray(series) >= threshold)[0]
    return int(idx[0]) if len(idx) else -1


# ------------------------------------------------------------------ GMM diagnosis (03:360-426)
# sklearn.mixture.GaussianMixture (covariance_type="full") is a third-party dependency of the
# reference (version unpinned by it; the container's sklearn 1.9.0 is what the goldens were made
# with).  The functions below restate its published E-step / M-step arithmetic
# (_estimate_log_gaussian_prob, _estimate_log_prob_resp, _estimate_gaussian_parameters) and the
# reference's own label calibration and class-probability mapping.
def gmm_resp(X, weights, means, prec_chol):
    """``(log_prob_norm[n], resp[n, C])`` of sklearn's ``_estimate_log_prob_resp`` (float64)."""
    X = np.asarray(X, np.float64)
    n, d = X.shape
    C = means.shape[0]
    log_prob = np.empty((n, C))
    for c in range(C):
        y = (X - means[c]) @ prec_chol[c]
        log_det = np.sum(np.log(np.diag(prec_chol[c])))
        log_prob[:, c] = -0.5 * (d * np.log(2.0 * np.pi) + np.sum(y * y, axis=1)) + log_det
    wlp = log_prob + np.log(weights)
    mx = wlp.max(axis=1, keepdims=True)
    lpn = mx[:, 0] + np.log(np.exp(wlp - mx).sum(axis=1))
    return lpn, np.exp(wlp - lpn[:, None])


def gmm_m_step(X, resp, reg_covar=1e-6):
    """``(weights, means, covariances)`` of ``_estimate_gaussian_parameters`` + the weight normalisation of ``_m_step``."""
    X = np.asarray(X, np.float64)
    nk = resp.sum(axis=0) + 10 * np.finfo(np.float64).eps
    means = resp.T @ X / nk[:, None]
    C, d = means.shape
    cov = np.empty((C, d, d))
    for c in range(C):
        diff = X - means[c]
        cov[c] = (resp[:, c] * diff.T) @ diff / nk[c]
        cov[c].flat[:: d + 1] += reg_covar
    return nk / nk.sum(), means, cov


def gmm_calibrate(resp_tr, y_tr, n_classes):
    """P(fault | component) from responsibilities and labels, 03:394-412."""
    C = resp_tr.shape[1]
    P = np.zeros((C, n_classes))
    for c in range(C):
        w = resp_tr[:, c]
        if w.sum() <= 0:
            P[c] = 1.0 / n_classes
            continue
        for k in range(n_classes):
            P[c, k] = w[y_tr == k].sum()
        s = P[c].sum()
        P[c] = P[c] / s if s > 0 else 1.0 / n_classes
    return P


def gmm_class_prob(resp_te, comp_fault_prob):
    """``(y_prob, y_pred)`` of 03:415-423."""
    y = np.clip(resp_te @ comp_fault_prob, 1e-12, 1.0)
    y /= y.sum(axis=1, keepdims=True)
    return y, y.argmax(axis=1)
