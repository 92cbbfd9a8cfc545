"""One warm-up + two train_dnn steps of the 6x256 net at N = 262144 (for `ncu --metrics gpu__time_duration.sum`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, b200pinn
from bench import build_problem
layers = [8, 256, 256, 256, 256, 256, 256, 1] if len(sys.argv) < 2 else [8] + [256] * int(sys.argv[1]) + [1]
X, Y, sx, sy = build_problem(262144, 2)
torch.manual_seed(0)
m = b200pinn.PhysicsInformedNN(X, Y, layers, sx, sy, 0.2, True)
m.train_dnn(3, verbose=False)
torch.cuda.synchronize()
