"""Quick A/B timing on one B200: MC sweep (T=50, N=1M) per kernel variant, eval forward, train_dnn step.
usage: python profiles/quick_time.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200pinn
from b200pinn import _abi
from bench import build_problem, LAYERS, P_TRAIN, P_MC, T_PASSES

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
X, Y, sx, sy = build_problem(n, 2)
torch.manual_seed(0)
model = b200pinn.PhysicsInformedNN(X, Y, LAYERS, sx, sy, P_TRAIN, True)
model.dnn.eval()
xd = model.x.detach()
lib = _abi.lib()


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


res = {}
outs = {}
for tpr in (2, 4):
    if hasattr(lib, "pinn_set_tc_threads_per_row"):
        lib.pinn_set_tc_threads_per_row(tpr)
    mc = lambda: b200pinn.mc_dropout_device(model.dnn, xd, T_PASSES, P_MC, seed=1234)
    res[f"mc_tpr{tpr}_ms"] = timed(mc)
    outs[tpr] = {k: v.clone() for k, v in mc().items() if torch.is_tensor(v)}
for k in outs[2]:
    d = (outs[2][k] - outs[4][k]).abs().max().item() / max(outs[2][k].abs().max().item(), 1e-30)
    res[f"tpr2_vs_4_{k}"] = d
with torch.no_grad():
    res["fwd_ms"] = timed(lambda: model.net_u(xd))
model.train_dnn(3, verbose=False)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
model.train_dnn(10, verbose=False)
b.record()
torch.cuda.synchronize()
res["train_ms"] = a.elapsed_time(b) / 10
for k, v in res.items():
    print(f"{k:28s} {v:.6g}")
