"""Ablation builds of K2a: bit 8 = no act/del stores, bit 16 = no activation re-loads in dgrad (results are wrong; timing only)."""
import importlib.util, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "physics-informed-neural-network-for-explainable-fault-diagnosis-in-fuel-cells_b200")
VARIANTS = [0, 8, 16, 24]
if sys.argv[1] == "build":
    spec = importlib.util.spec_from_file_location("b", os.path.join(PKG, "build.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    for k in VARIANTS:
        d = os.path.join(PKG, "build", f"abl_{k}")
        os.makedirs(d, exist_ok=True)
        print(m.build(force=True, extra_flags=[f"-DPINN_ABL={k}"], out=os.path.join(d, "libb200pinn.so"), objdir=d))
else:
    for k in VARIANTS:
        code = f"""
import sys; sys.path.insert(0, {ROOT!r})
import b200pinn._abi as abi
abi.LIB_PATH = {os.path.join(PKG, 'build', f'abl_{k}', 'libb200pinn.so')!r}
import torch, b200pinn
from bench import build_problem, LAYERS, P_TRAIN
X, Y, sx, sy = build_problem(1_000_000, 2)
torch.manual_seed(0)
model = b200pinn.PhysicsInformedNN(X, Y, LAYERS, sx, sy, P_TRAIN, True)
model.train_dnn(3, verbose=False)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); model.train_dnn(10, verbose=False); b.record(); torch.cuda.synchronize()
print('ablation', {k}, 'train_ms', a.elapsed_time(b) / 10)
"""
        subprocess.run([sys.executable, "-c", code])
