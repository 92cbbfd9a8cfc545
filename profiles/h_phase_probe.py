"""Probe: per-launch wall time of the hydrogen phase inside the configs[0] schedule (it was erratic: 0.03-0.12 s per 8001 epochs)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200pinn
from b200pinn import kernels as K
from b200pinn.synthetic import make_scaled_dataset

x, y, sx, sy = make_scaled_dataset(20000, seed=1)
torch.manual_seed(0)
m = b200pinn.PhysicsInformedNN(torch.tensor(x), torch.tensor(y), [8, 64, 64, 64, 1], sx, sy, 0.2, True)
m.train_dnn(100, verbose=False)
m.train_thermal(10001, verbose=False)
orig = K.scalar_phase
times = []
def timed_phase(*a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    orig(*a, **k)
    torch.cuda.synchronize(); times.append((a[-2] if not k else None, time.perf_counter() - t0))
K.scalar_phase = timed_phase
for rep in range(3):
    times.clear()
    t0 = time.perf_counter()
    m.train_hydrogen(8001, verbose=False)
    torch.cuda.synchronize()
    tot = time.perf_counter() - t0
    print(f"rep {rep}: total {tot * 1e3:.1f} ms; per launch (steps, ms): " + ", ".join(f"({s}, {1e3 * t:.2f})" for s, t in times))
    print("   lambda_H:", m._lam[9:13].cpu().numpy())
