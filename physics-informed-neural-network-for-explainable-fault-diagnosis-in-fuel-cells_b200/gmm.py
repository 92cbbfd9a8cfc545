"""GMM fault diagnosis on the device -- drop-in for ``fit_gmm_and_get_probabilities``
(03_unsupervised_gmm_fault_diagnosis 03:360-426).

The reference fits ``sklearn.mixture.GaussianMixture(covariance_type="full")`` on a few per-sample
features (pV, pT, pH, pO), calibrates every component against the training labels and maps test
responsibilities to fault-class probabilities.  Here every pass over the samples -- the E-step, the
M-step's sufficient statistics, the calibration sums and the class-probability mapping -- is one launch
of ``pinn_gmm_pass``; what stays on the host is O(components x d^2) arithmetic per EM iteration
(means / covariances from the statistics, the Cholesky factors: sklearn's own helper) and sklearn's
k-means initialisation, so the fitted object that is returned IS a ``GaussianMixture``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _abi
from ._abi import check, ptr
from . import kernels as K


def _dev64(a, device):
    t = a if torch.is_tensor(a) else torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64))
    return t.to(device=device, dtype=torch.float64).contiguous()


def _cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("b200pinn.gmm needs a CUDA device (B200, sm_100a); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def gmm_pass(X, weights, means, prec_chol, labels=None, n_classes=0, comp_class_prob=None, want_resp=False,
             want_stats=False):
    """One pass over ``X`` (CUDA float64 ``[n, d]``).  Returns a dict with ``log_prob_norm_sum`` and, as requested,
    ``resp [n, C]``, ``stats [C, 1 + d + d(d+1)/2]``, ``comp_class_weight [C, K]`` (needs ``labels``),
    ``y_prob [n, K]`` / ``y_pred [n]`` (needs ``comp_class_prob``).  See ``pinn_gmm_pass``."""
    if not torch.is_tensor(X) or not X.is_cuda or X.dtype != torch.float64 or X.dim() != 2:
        raise RuntimeError("b200pinn.gmm: X must be a CUDA float64 [n, d] tensor -- there is no CPU path")
    X = X.contiguous()
    dev = X.device
    n, d = X.shape
    w, m, pc = _dev64(weights, dev), _dev64(means, dev), _dev64(prec_chol, dev)
    nc = m.shape[0]
    if m.shape != (nc, d) or pc.shape != (nc, d, d) or w.shape != (nc,):
        raise RuntimeError("b200pinn.gmm: parameter shapes do not match X")
    out = {}
    lab = None
    if labels is not None:
        lab = (labels if torch.is_tensor(labels) else torch.as_tensor(np.asarray(labels))).to(device=dev, dtype=torch.int32).contiguous()
        if lab.shape != (n,):
            raise RuntimeError("b200pinn.gmm: labels must have one entry per row")
        out["comp_class_weight"] = torch.empty(nc, n_classes, device=dev, dtype=torch.float64)
    P = None
    if comp_class_prob is not None:
        P = _dev64(comp_class_prob, dev)
        n_classes = P.shape[1]
        out["y_prob"] = torch.empty(n, n_classes, device=dev, dtype=torch.float64)
        out["y_pred"] = torch.empty(n, device=dev, dtype=torch.int32)
    if want_resp:
        out["resp"] = torch.empty(n, nc, device=dev, dtype=torch.float64)
    if want_stats:
        out["stats"] = torch.empty(nc, 1 + d + d * (d + 1) // 2, device=dev, dtype=torch.float64)
    lpn = torch.empty(1, device=dev, dtype=torch.float64)
    L = _abi.lib()
    nb = L.pinn_gmm_workspace_bytes(d, nc, int(n_classes))
    if nb == 0:
        raise RuntimeError("b200pinn.gmm: unsupported shape (d <= 8, components <= 32, classes <= 16)")
    ws = K._workspace("gmm", nb, dev)
    with torch.cuda.device(dev):
        check(L.pinn_gmm_pass(ptr(X), n, d, nc, ptr(w), ptr(m), ptr(pc), ptr(lab), int(n_classes), ptr(P), ptr(out.get("resp")),
                              ptr(out.get("y_prob")), ptr(out.get("y_pred")), ptr(out.get("stats")),
                              ptr(out.get("comp_class_weight")), ptr(lpn), ptr(ws), nb, K._stream()), "pinn_gmm_pass")
    K.LAUNCHES += 2
    out["log_prob_norm_sum"] = lpn
    return out


def m_step_from_stats(stats, means_old, reg_covar):
    """Finish sklearn's M-step (``_estimate_gaussian_parameters`` + ``_m_step``) from the device statistics, which are
    taken around the OLD means: with delta = mean_new - mean_old,
    sum r (x - mean_new)(x - mean_new)^T = Sxx - Sx delta^T - delta Sx^T + N delta delta^T."""
    stats = np.asarray(stats, np.float64)
    nc, d = means_old.shape
    n0 = stats[:, 0]
    nk = n0 + 10 * np.finfo(np.float64).eps
    sx = stats[:, 1:1 + d]
    means = (sx + means_old * n0[:, None]) / nk[:, None]
    delta = means - means_old
    iu = np.triu_indices(d)
    cov = np.empty((nc, d, d))
    for c in range(nc):
        S = np.zeros((d, d))
        S[iu] = stats[c, 1 + d:]
        S = S + S.T - np.diag(np.diag(S))
        M = S - np.outer(sx[c], delta[c]) - np.outer(delta[c], sx[c]) + n0[c] * np.outer(delta[c], delta[c])
        cov[c] = M / nk[c]
        cov[c].flat[:: d + 1] += reg_covar
    return nk / nk.sum(), means, cov


def fit_gmm_device(gmm, X, init=None):
    """``gmm.fit(X)`` with every pass over the samples on the device (n_init = 1, covariance_type "full").
    ``X``: CUDA float64 tensor or host array.  ``init``: optional ``(weights, means, precisions_cholesky)`` to start
    from; default is sklearn's own initialisation (k-means on the host, same random stream as ``gmm.fit``)."""
    from sklearn.mixture._gaussian_mixture import _compute_precision_cholesky
    from sklearn.utils import check_random_state

    if gmm.covariance_type != "full" or gmm.n_init != 1:
        raise NotImplementedError("b200pinn.gmm: covariance_type='full' and n_init=1 (what 03:381-385 uses)")
    dev = _cuda()
    Xd = X if torch.is_tensor(X) and X.is_cuda else _dev64(X, dev)
    n = Xd.shape[0]
    if init is None:
        Xh = np.asarray(X, np.float64) if not torch.is_tensor(X) else X.detach().cpu().numpy().astype(np.float64)
        gmm._initialize_parameters(Xh, check_random_state(gmm.random_state))
        weights, means, pc = gmm.weights_, gmm.means_, gmm.precisions_cholesky_
    else:
        weights, means, pc = (np.asarray(t, np.float64) for t in init)
    cov = None
    lower, converged, n_iter, bounds = -np.inf, False, 0, []
    for n_iter in range(1, gmm.max_iter + 1):
        prev = lower
        r = gmm_pass(Xd, weights, means, pc, want_stats=True)                 # E-step + sufficient statistics
        lower = float(r["log_prob_norm_sum"].item()) / n
        weights, means, cov = m_step_from_stats(r["stats"].cpu().numpy(), means, gmm.reg_covar)
        pc = _compute_precision_cholesky(cov, "full")
        bounds.append(lower)
        if abs(lower - prev) < gmm.tol:
            converged = True
            break
    gmm.weights_, gmm.means_, gmm.precisions_cholesky_ = weights, means, pc
    if cov is not None:
        gmm.covariances_ = cov
        gmm.precisions_ = np.stack([p @ p.T for p in pc])
    gmm.converged_, gmm.n_iter_, gmm.lower_bound_, gmm.lower_bounds_ = converged, n_iter, lower, bounds
    return gmm


def comp_fault_prob_from_weights(W, n_classes):
    """Row-normalise the calibration sums with the reference's fall-backs (03:397-412)."""
    W = np.asarray(W, np.float64)
    P = np.zeros_like(W)
    for c in range(W.shape[0]):
        s = W[c].sum()
        P[c] = W[c] / s if s > 0 else 1.0 / n_classes
    return P


def fit_gmm_and_get_probabilities(X_tr, y_tr, X_te, n_classes, random_state=42, n_components=None):
    """03:360-426 -- same arguments and return tuple ``(y_prob, y_pred, gmm, comp_fault_prob)``."""
    from sklearn.mixture import GaussianMixture

    if n_components is None:
        n_components = n_classes
    dev = _cuda()
    X_tr = np.asarray(X_tr, np.float64)
    Xtr_d, Xte_d = _dev64(X_tr, dev), _dev64(np.asarray(X_te, np.float64), dev)
    gmm = GaussianMixture(n_components=n_components, covariance_type="full", random_state=random_state)
    fit_gmm_device(gmm, X_tr)
    params = (gmm.weights_, gmm.means_, gmm.precisions_cholesky_)
    cal = gmm_pass(Xtr_d, *params, labels=np.asarray(y_tr), n_classes=n_classes)
    comp_fault_prob = comp_fault_prob_from_weights(cal["comp_class_weight"].cpu().numpy(), n_classes)
    te = gmm_pass(Xte_d, *params, comp_class_prob=comp_fault_prob)
    return te["y_prob"].cpu().numpy(), te["y_pred"].cpu().numpy().astype(np.int64), gmm, comp_fault_prob
