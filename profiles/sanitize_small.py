"""Small end-to-end exercise of every CUDA path (for compute-sanitizer): 64-wide TC kernels, wide GEMM path, FFMA fallbacks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, b200pinn
from b200pinn import kernels as K
from b200pinn.synthetic import make_scaled_dataset
for layers, n in (([8, 64, 64, 64, 1], 333), ([8, 256, 256, 256, 1], 300), ([8, 128, 128, 1], 200), ([8, 32, 32, 1], 100)):
    x, y, sx, sy = make_scaled_dataset(max(n, 64), seed=3)
    x, y = x[:n], y[:n]
    torch.manual_seed(0)
    m = b200pinn.PhysicsInformedNN(torch.tensor(x), torch.tensor(y), layers, sx, sy, 0.2, True)
    m.dnn.eval()
    xd = m.x.detach()
    out = b200pinn.mc_dropout_device(m.dnn, xd, 3, 0.4, seed=5)
    u, s = K.mlp_forward(K.net_from_module(m.dnn), xd)
    m.train_dnn(2, verbose=False)
    pm, au, eu = b200pinn.get_MC_samples(m, torch.tensor(x), sx, mc_times=2, dropout=0.3)
    torch.cuda.synchronize()
    print(layers, float(out["e_u"].mean()), float(u.mean()), float(np.mean(eu)))
print("sanitize run complete")
