"""Timing of the GMM pass (pinn_gmm_pass) at fleet scale: n rows x d = 4 features, 20 components, 13 classes (03:29,548).
usage: python profiles/gmm_time.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from b200pinn import gmm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
d, C, K = 4, 20, 13
rng = np.random.default_rng(0)
X = torch.tensor(rng.normal(size=(n, d)) * 2.0, device="cuda")
means = rng.normal(size=(C, d)) * 2.0
A = rng.normal(size=(C, d, d)) * 0.3
cov = A @ np.transpose(A, (0, 2, 1)) + np.eye(d) * 0.5
pc = np.stack([np.linalg.inv(np.linalg.cholesky(cv)).T for cv in cov])
w = np.ones(C) / C
y = torch.tensor(rng.integers(0, K, n), device="cuda", dtype=torch.int32)
P = rng.uniform(size=(C, K)); P /= P.sum(axis=1, keepdims=True)
wd, md, pd, Pd = (torch.tensor(t, device="cuda") for t in (w, means, pc, P))


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for name, fn in (("EM iteration (E-step + sufficient statistics)", lambda: gmm.gmm_pass(X, wd, md, pd, want_stats=True)),
                 ("label calibration", lambda: gmm.gmm_pass(X, wd, md, pd, labels=y, n_classes=K)),
                 ("class probabilities (y_prob, y_pred)", lambda: gmm.gmm_pass(X, wd, md, pd, comp_class_prob=Pd))):
    ms = timed(fn)
    print(f"n={n} {name:48s} {ms:8.3f} ms  {n / ms / 1e6:8.2f} G rows/s  X read at {n * d * 8 / ms / 1e6:7.1f} GB/s")
