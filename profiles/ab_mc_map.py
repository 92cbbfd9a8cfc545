"""A/B of the MC kernel's tile -> (CTA, group) mapping (pair-first build in build/pairfirst, see ab_k2a_map.py).
usage: python profiles/ab_mc_map.py run <n> [pair]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "physics-informed-neural-network-for-explainable-fault-diagnosis-in-fuel-cells_b200")
sys.path.insert(0, ROOT)
import b200pinn._abi as abi
pair = len(sys.argv) > 3 and sys.argv[3] == "pair"
if pair:
    abi.LIB_PATH = os.path.join(PKG, "build", "pairfirst", "libb200pinn.so")
import torch, b200pinn
from b200pinn.synthetic import make_scaled_dataset
n = int(sys.argv[2])
x, y, sx, sy = make_scaled_dataset(n, seed=1)
torch.manual_seed(0)
m = b200pinn.PhysicsInformedNN(torch.tensor(x), torch.tensor(y), [8, 64, 64, 64, 1], sx, sy, 0.2, True)
m.dnn.eval()
xd = m.x.detach()
b200pinn.mc_dropout_device(m.dnn, xd, 50, 0.4, seed=1)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); out = b200pinn.mc_dropout_device(m.dnn, xd, 2000, 0.4, seed=1); b.record(); torch.cuda.synchronize()
t = a.elapsed_time(b)
with torch.no_grad():
    a.record()
    for _ in range(20):
        m.net_u(xd)
    b.record(); torch.cuda.synchronize()
print(f"n={n} {'pair-first' if pair else 'spread    '}: MC sweep T=2000 {t:.2f} ms ({n * 2000 / t / 1e6:.2f} G sample*passes/s), eval forward {1e3 * a.elapsed_time(b) / 20:.1f} us, checksum {float(out['e_u'].double().sum()):.9e}")
