"""Per-step time of the four scalar phases (persistent pinn_scalar_phase launches vs the launch-per-step loop).
usage: python profiles/c1_phase_times.py [n] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200pinn
from b200pinn.synthetic import make_scaled_dataset

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2001
x, y, sx, sy = make_scaled_dataset(n, seed=1)
torch.manual_seed(0)
m = b200pinn.PhysicsInformedNN(torch.tensor(x), torch.tensor(y), [8, 64, 64, 64, 1], sx, sy, 0.2, True)
phases = [("train_lambda(A)", lambda k: m.train_lambda(k, False, verbose=False)),
          ("train_lambda(B)", lambda k: m.train_lambda(k, True, verbose=False)),
          ("train_thermal", lambda k: m.train_thermal(k, verbose=False)),
          ("train_hydrogen", lambda k: m.train_hydrogen(k, verbose=False)),
          ("train_oxygen", lambda k: m.train_oxygen(k, verbose=False))]
from b200pinn import kernels as K
for mode, cluster in (("1", True), ("1", False), ("0", True)):
    os.environ["B200PINN_PHASE_KERNEL"] = mode
    K.set_default_path_flags(no_phase_cluster=not cluster)
    for name, fn in phases:
        fn(3)
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn(steps)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        print(f"n={n} {('persistent/cluster' if cluster else 'persistent/grid   ') if mode == '1' else 'per-step          '} {name:16s} {1e6 * dt / steps:7.2f} us/step")
