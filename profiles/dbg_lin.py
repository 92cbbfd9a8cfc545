import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import b200pinn
from b200pinn import kernels as K
from b200pinn.synthetic import make_scaled_dataset
rel = lambda p, q: float((p - q).abs().max() / q.abs().max())
x, y, sx, sy = make_scaled_dataset(40000, seed=100)
torch.manual_seed(1000)
m = b200pinn.PhysicsInformedNN(torch.tensor(x), torch.tensor(y), [8, 64, 64, 64, 1], sx, sy, 0.2, True)
net = K.net_from_module(m.dnn)
xd, yd = m.x.detach(), m.u.reshape(-1).contiguous()
names, shapes, offs, _ = K.param_layout(64, 3)
for n in (5001, 8000, 9472, 9600, 10000, 11000, 12000, 12032, 13000, 16000, 18944, 19000, 20000, 30000):
    def run():
        g, s = K.mlp_backward(net, xd[:n].contiguous(), K.make_dropout(0.2, seed=7, pass_offset=5), y=yd[:n].contiguous(), n_global=n)
        torch.cuda.synchronize()
        return g.clone()
    with K.path_flags(dependent_launch=0):
        a = run()
    with K.path_flags(no_tc_bwd=True):
        b = run()
    worst = max(((rel(a[o:o + int(np.prod(s))], b[o:o + int(np.prod(s))]), nm) for nm, s, o in zip(names, shapes, offs)))
    print(n, "tiles", -(-n // 128), "tc vs ffma", f"{rel(a, b):.2e}", "worst tensor", worst[1], f"{worst[0]:.2e}")
