"""Golden vectors for the GMM fault diagnosis (03:360-426), produced by the UNMODIFIED reference function
``fit_gmm_and_get_probabilities`` of 03_unsupervised_gmm_fault_diagnosis.py.py in the build container (sklearn 1.9.0).

    python tests/golden/make_golden_gmm.py      ->  tests/golden/gmm4.npz

The features imitate the script's inputs (pV, pT, pH, pO: per-sample physics residual scores, 03:29,548): four fault
classes, each a blob or a pair of blobs in 4-D with different spreads and some overlap, so that 8 components do not
map one-to-one onto classes.  Also stored: sklearn's initial parameters for the same random_state (what ``fit`` starts
EM from), so that the device EM loop can be checked iteration for iteration without depending on k-means.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_golden_export import load_script  # noqa: E402


def make_features(n, seed):
    rng = np.random.default_rng(seed)
    centers = {0: [(0.0, 0.0, 0.0, 0.0), (0.6, -0.2, 0.1, 0.0)],
               1: [(3.0, 0.5, 0.0, 0.2)],
               2: [(0.3, 2.5, 0.4, 0.0), (0.2, 4.0, 0.6, 0.3)],
               3: [(0.4, 0.3, 2.2, 2.0)]}
    scales = {0: 0.45, 1: 0.8, 2: 0.6, 3: 0.9}
    y = rng.integers(0, 4, n)
    X = np.empty((n, 4))
    for i, k in enumerate(y):
        c = centers[k][rng.integers(0, len(centers[k]))]
        X[i] = np.asarray(c) + rng.normal(0.0, scales[k], 4) * np.array([1.0, 0.7, 1.3, 0.5])
    X[:, 3] += 0.3 * X[:, 2]                       # correlated features: full covariances matter
    return X, y


def main():
    from sklearn.mixture import GaussianMixture
    from sklearn.utils import check_random_state

    ref03 = load_script("ref03", "/root/reference/03_unsupervised_gmm_fault_diagnosis.py.py")
    X_tr, y_tr = make_features(3000, 31)
    X_te, y_te = make_features(1200, 32)
    n_classes, n_comp, rs = 4, 8, 42
    y_prob, y_pred, gmm, P = ref03.fit_gmm_and_get_probabilities(X_tr, y_tr, X_te, n_classes, random_state=rs, n_components=n_comp)
    g0 = GaussianMixture(n_components=n_comp, covariance_type="full", random_state=rs)
    g0._initialize_parameters(X_tr, check_random_state(rs))
    out = dict(X_tr=X_tr, y_tr=y_tr, X_te=X_te, y_te=y_te, n_classes=n_classes, n_components=n_comp, random_state=rs,
               y_prob=y_prob, y_pred=y_pred, comp_fault_prob=P,
               weights=gmm.weights_, means=gmm.means_, covariances=gmm.covariances_, prec_chol=gmm.precisions_cholesky_,
               n_iter=gmm.n_iter_, lower_bound=gmm.lower_bound_, converged=gmm.converged_,
               resp_te=gmm.predict_proba(X_te), resp_tr=gmm.predict_proba(X_tr), score_tr=gmm.score(X_tr),
               init_weights=g0.weights_, init_means=g0.means_, init_prec_chol=g0.precisions_cholesky_)
    path = os.path.join(HERE, "gmm4.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes; n_iter", gmm.n_iter_, "acc", float((y_pred == y_te).mean()))


if __name__ == "__main__":
    main()
