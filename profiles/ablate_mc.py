"""Ablation builds of the MC kernel (what bounds it?): bit 1 = no MUFU tanh, bit 2 = no Philox rounds,
bit 4 = one MMA per product instead of 24.  `python profiles/ablate_mc.py build` (CPU box) compiles
build/abl_<k>/libb200pinn.so; `python profiles/ablate_mc.py run` (GPU box) times each."""
import importlib.util, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "physics-informed-neural-network-for-explainable-fault-diagnosis-in-fuel-cells_b200")
VARIANTS = [0, 1, 2, 3, 4, 7]
if sys.argv[1] == "build":
    spec = importlib.util.spec_from_file_location("b", os.path.join(PKG, "build.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    for k in VARIANTS:
        d = os.path.join(PKG, "build", f"abl_{k}")
        os.makedirs(d, exist_ok=True)
        print(m.build(force=True, extra_flags=[f"-DPINN_ABL={k}"], out=os.path.join(d, "libb200pinn.so"), objdir=d))
elif sys.argv[1] == "run":
    for k in VARIANTS:
        code = f"""
import sys; sys.path.insert(0, {ROOT!r})
import b200pinn._abi as abi
abi.LIB_PATH = {os.path.join(PKG, 'build', f'abl_{k}', 'libb200pinn.so')!r}
import torch, b200pinn
from bench import build_problem, LAYERS, P_TRAIN, P_MC, T_PASSES
X, Y, sx, sy = build_problem(1_000_000, 2)
torch.manual_seed(0)
model = b200pinn.PhysicsInformedNN(X, Y, LAYERS, sx, sy, P_TRAIN, True)
model.dnn.eval(); xd = model.x.detach()
mc = lambda: b200pinn.mc_dropout_device(model.dnn, xd, T_PASSES, P_MC, seed=1234)
for _ in range(2): mc()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): mc()
b.record(); torch.cuda.synchronize()
print('ablation', {k}, 'mc_ms', a.elapsed_time(b) / 5)
"""
        subprocess.run([sys.executable, "-c", code])
