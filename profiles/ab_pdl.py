"""A/B of the dependent-launch modes of the train_dnn step inside one process (alternating, several repetitions).
usage: python profiles/ab_pdl.py [n] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200pinn
from b200pinn import kernels as K
from b200pinn.synthetic import make_scaled_dataset

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
x, y, sx, sy = make_scaled_dataset(n, seed=1)
torch.manual_seed(0)
m = b200pinn.PhysicsInformedNN(torch.tensor(x), torch.tensor(y), [8, 64, 64, 64, 1], sx, sy, 0.2, True)
m.train_dnn(5, verbose=False)
for rep in range(3):
    for mode in (0, 2):
        K.set_dependent_launch(mode)
        m.train_dnn(3, verbose=False)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        m.train_dnn(steps, verbose=False)
        b.record()
        torch.cuda.synchronize()
        print(f"n={n} rep {rep} pdl mode {mode}: {1e3 * a.elapsed_time(b) / steps:.1f} us/step")
