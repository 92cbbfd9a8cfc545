"""Run one kernel a few times (for `ncu -k <name> --launch-skip 2 --launch-count 1 --set full ... python profiles/kernel_once.py <what> [n]`).
what: res (V|DATA training form), resx (export form), mc (T=50 sweep), train (one train_dnn step), rf (8 stacks),
wide (3x256 net, T=10 sweep on the resident-activation kernel; n = 262144)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200pinn
from b200pinn import _abi, kernels as K
from bench import build_problem, LAYERS, P_TRAIN, P_MC, T_PASSES

what = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
if what == "wide":
    Xw, Yw, sxw, syw = build_problem(n, 2)
    torch.manual_seed(0)
    mw = b200pinn.PhysicsInformedNN(Xw, Yw, [8, 256, 256, 256, 1], sxw, syw, P_TRAIN, True)
    mw.dnn.eval()
    for _ in range(4):
        b200pinn.mc_dropout_device(mw.dnn, mw.x.detach(), 10, P_MC, seed=1234)
    torch.cuda.synchronize()
    print("done", what, n)
    sys.exit(0)
base = min(n, 1_000_000)
X, Y, sx, sy = build_problem(base, 2)
torch.manual_seed(0)
model = b200pinn.PhysicsInformedNN(X, Y, LAYERS, sx, sy, P_TRAIN, True)
model.dnn.eval()
xd = model.x.detach()
rep = max(1, n // base)
with torch.no_grad():
    u = model.net_u(xd)[0].reshape(-1).contiguous()
yv = model.u.reshape(-1).contiguous()
xb, ub, yb = xd.repeat(rep, 1).contiguous(), u.repeat(rep).contiguous(), yv.repeat(rep).contiguous()
sc, lam = model._scalers(sx), model._lambdas()
sums = torch.empty(_abi.S_COUNT, device=xd.device, dtype=torch.float64)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=xd.device)
for _ in range(4):
    flush.zero_()
    if what == "res":
        K.residuals(xb, ub, yb, sc, lam, _abi.FAM_V | _abi.FAM_DATA, sums=sums)
    elif what == "resx":
        K.residuals(xb, ub, None, sc, lam, _abi.FAM_V | _abi.FAM_TS | _abi.FAM_H | _abi.FAM_O, sums=sums, want_cols=True)
    elif what == "mc":
        b200pinn.mc_dropout_device(model.dnn, xd, T_PASSES, P_MC, seed=1234)
    elif what == "train":
        model.train_dnn(1, verbose=False)
    elif what == "rf":
        from b200pinn.export import export_rows_device
        from b200pinn.rf import rf_device
        seg = [0] + [base * (i + 1) // 13 for i in range(13)]
        rows = export_rows_device(model, xd, yv, seg, 12, 3, P_MC, sx, sy, seed=1)
        rf_device(rows.unsqueeze(0).expand(8, -1, -1).contiguous())
torch.cuda.synchronize()
print("done", what, n)
