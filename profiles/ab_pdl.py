"""A/B of the dependent-launch modes of the train_dnn step inside one process (alternating, several repetitions).
usage: python profiles/ab_pdl.py [n] [steps] [width] [hidden]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200pinn
from b200pinn import kernels as K
from b200pinn.synthetic import make_scaled_dataset

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
width = int(sys.argv[3]) if len(sys.argv) > 3 else 64
hidden = int(sys.argv[4]) if len(sys.argv) > 4 else 3
x, y, sx, sy = make_scaled_dataset(n, seed=1)
torch.manual_seed(0)
m = b200pinn.PhysicsInformedNN(torch.tensor(x), torch.tensor(y), [8] + [width] * hidden + [1], sx, sy, 0.2, True)
m.train_dnn(5, verbose=False)
for rep in range(3):
    for mode in (0, 2):
        K.set_default_path_flags(dependent_launch=mode)
        m.train_dnn(3, verbose=False)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        m.train_dnn(steps, verbose=False)
        b.record()
        torch.cuda.synchronize()
        t_train = 1e3 * a.elapsed_time(b) / steps
        m.dnn.eval()
        xd = m.x.detach()
        b200pinn.mc_dropout_device(m.dnn, xd, 20, 0.4, seed=1)
        a.record()
        b200pinn.mc_dropout_device(m.dnn, xd, 200, 0.4, seed=1)
        b.record()
        torch.cuda.synchronize()
        print(f"{hidden}x{width} n={n} rep {rep} pdl mode {mode}: train {t_train:.1f} us/step, MC sweep {1e3 * a.elapsed_time(b) / 200:.1f} us/pass")
