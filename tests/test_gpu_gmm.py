"""GPU parity of the GMM diagnosis pass (pinn_gmm_pass) against goldens recorded from the reference's
``fit_gmm_and_get_probabilities`` (03:360-426, tests/golden/make_golden_gmm.py) and against the numpy oracle.
Everything is float64; tolerances are 1e-9 (reduction order and fused multiply-adds are the only differences)."""
import os

import numpy as np
import pytest
import torch

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gmm4.npz")
TOL = 1e-9


def dev():
    return torch.device("cuda", 0)


def d64(a):
    return torch.tensor(np.asarray(a, np.float64), device=dev())


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLDEN))


def test_posterior_calibration_and_class_probabilities_golden(g):
    from b200pinn import gmm

    params = (g["weights"], g["means"], g["prec_chol"])
    te = gmm.gmm_pass(d64(g["X_te"]), *params, comp_class_prob=g["comp_fault_prob"], want_resp=True)
    assert np.abs(te["resp"].cpu().numpy() - g["resp_te"]).max() < TOL
    assert np.abs(te["y_prob"].cpu().numpy() - g["y_prob"]).max() < TOL
    assert np.array_equal(te["y_pred"].cpu().numpy(), g["y_pred"])
    K = int(g["n_classes"])
    tr = gmm.gmm_pass(d64(g["X_tr"]), *params, labels=g["y_tr"], n_classes=K, want_resp=True)
    assert np.abs(tr["resp"].cpu().numpy() - g["resp_tr"]).max() < TOL
    P = gmm.comp_fault_prob_from_weights(tr["comp_class_weight"].cpu().numpy(), K)
    assert np.abs(P - g["comp_fault_prob"]).max() < TOL
    assert abs(tr["log_prob_norm_sum"].item() / g["X_tr"].shape[0] - float(g["score_tr"])) < TOL


def test_em_from_sklearn_init_reproduces_reference_fit(g):
    """EM on the device from sklearn's own initial parameters: same iteration count, same fitted mixture."""
    from sklearn.mixture import GaussianMixture
    from b200pinn import gmm

    m = GaussianMixture(n_components=int(g["n_components"]), covariance_type="full", random_state=int(g["random_state"]))
    gmm.fit_gmm_device(m, g["X_tr"], init=(g["init_weights"], g["init_means"], g["init_prec_chol"]))
    assert m.converged_ and m.n_iter_ == int(g["n_iter"])
    assert abs(m.lower_bound_ - float(g["lower_bound"])) < 1e-10
    assert np.abs(m.weights_ - g["weights"]).max() < 1e-10
    assert np.abs(m.means_ - g["means"]).max() < 1e-8
    assert np.abs(m.covariances_ - g["covariances"]).max() < 1e-8
    assert np.abs(m.predict_proba(g["X_te"]) - g["resp_te"]).max() < 1e-8       # the returned object is a working sklearn model


def test_drop_in_function_matches_reference_outputs(g):
    """The whole of 03:360-426 through the public function (k-means initialisation by sklearn on this box)."""
    import b200pinn

    y_prob, y_pred, m, P = b200pinn.fit_gmm_and_get_probabilities(g["X_tr"], g["y_tr"], g["X_te"], int(g["n_classes"]),
                                                                  random_state=int(g["random_state"]),
                                                                  n_components=int(g["n_components"]))
    assert y_prob.shape == g["y_prob"].shape and y_pred.dtype == np.int64 and P.shape == g["comp_fault_prob"].shape
    assert np.allclose(y_prob.sum(axis=1), 1.0, atol=1e-12)
    assert (y_pred == g["y_pred"]).mean() >= 0.999
    assert np.abs(y_prob - g["y_prob"]).max() < 1e-5
    assert np.abs(P - g["comp_fault_prob"]).max() < 1e-5


@pytest.mark.parametrize("n,d,C,K", [(1, 4, 3, 2), (33, 2, 32, 16), (1000, 8, 5, 3), (4097, 1, 2, 1), (70000, 4, 20, 13), (600, 8, 32, 16)])
def test_pass_vs_oracle_shapes_and_edges(n, d, C, K):
    """Ragged row counts, the extreme d / components / classes, labels outside [0, K) skipped, every output at once."""
    from b200pinn import gmm

    rng = np.random.default_rng(n + d)
    X = rng.normal(size=(n, d)) * 2.0
    means = rng.normal(size=(C, d)) * 2.0
    A = rng.normal(size=(C, d, d)) * 0.3
    cov = A @ np.transpose(A, (0, 2, 1)) + np.eye(d) * 0.5
    pc = np.stack([np.linalg.inv(np.linalg.cholesky(cv)).T for cv in cov])       # sklearn's precisions_cholesky_
    w = rng.uniform(0.5, 1.5, C)
    w /= w.sum()
    y = rng.integers(-1, K + 1, n)                                                # -1 and K are out of range
    P = rng.uniform(size=(C, K))
    P /= P.sum(axis=1, keepdims=True)
    r = gmm.gmm_pass(d64(X), w, means, pc, labels=y, n_classes=K, comp_class_prob=P, want_resp=True, want_stats=True)
    lpn, resp = O.gmm_resp(X, w, means, pc)
    assert np.abs(r["resp"].cpu().numpy() - resp).max() < TOL
    assert abs(r["log_prob_norm_sum"].item() - lpn.sum()) < TOL * max(1.0, abs(lpn.sum()))
    yp, ypred = O.gmm_class_prob(resp, P)
    assert np.abs(r["y_prob"].cpu().numpy() - yp).max() < TOL
    got = r["y_pred"].cpu().numpy()
    clear = np.sort(yp, axis=1)[:, -1] - (np.sort(yp, axis=1)[:, -2] if K > 1 else 0.0) > 1e-9 if K > 1 else np.ones(n, bool)
    assert np.array_equal(got[clear], ypred[clear])
    Wref = np.stack([[resp[y == k, c].sum() for k in range(K)] for c in range(C)])
    assert np.abs(r["comp_class_weight"].cpu().numpy() - Wref).max() < TOL * max(1.0, n)
    wn, mn, cn = gmm.m_step_from_stats(r["stats"].cpu().numpy(), means, 1e-6)
    wo, mo, co = O.gmm_m_step(X, resp, 1e-6)
    assert np.abs(wn - wo).max() < 1e-10 and np.abs(mn - mo).max() < 1e-8 and np.abs(cn - co).max() < 1e-7


def test_argument_errors():
    from b200pinn import gmm

    X = d64(np.zeros((10, 4)))
    w, m, pc = np.ones(2) / 2, np.zeros((2, 4)), np.stack([np.eye(4)] * 2)
    with pytest.raises(RuntimeError, match="CUDA float64"):
        gmm.gmm_pass(torch.zeros(10, 4, dtype=torch.float64), w, m, pc)
    with pytest.raises(RuntimeError, match="shapes"):
        gmm.gmm_pass(X, w, np.zeros((2, 3)), pc)
    with pytest.raises(RuntimeError, match="unsupported shape"):
        gmm.gmm_pass(d64(np.zeros((10, 9))), w, np.zeros((2, 9)), np.stack([np.eye(9)] * 2))
    with pytest.raises(RuntimeError, match="unsupported shape"):
        gmm.gmm_pass(X, np.ones(33) / 33, np.zeros((33, 4)), np.stack([np.eye(4)] * 33))
