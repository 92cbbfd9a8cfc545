"""One-kernel (fused) vs two-kernel (K2a + row table + K2b) backward of the 64-wide net: gradient agreement and
train_dnn step time.  usage: python profiles/ab_fused.py [n,...] [layers: 3|2]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import b200pinn
from b200pinn import kernels as K
from bench import build_problem, P_TRAIN

ns = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [20_000, 1_000_000]
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 3
layers = [8] + [64] * nl + [1]
out = {}
for n in ns:
    X, Y, sx, sy = build_problem(n, 2)
    torch.manual_seed(0)
    model = b200pinn.PhysicsInformedNN(X, Y, layers, sx, sy, P_TRAIN, True)
    net = K.net_from_module(model.dnn)
    xd, yd = model.x.detach(), model.u.reshape(-1).contiguous()
    drop = lambda: K.make_dropout(P_TRAIN, seed=7, pass_offset=1)
    ga, sa = K.mlp_backward(net, xd, drop(), y=yd, n_global=n)
    ga2, _ = K.mlp_backward(net, xd, drop(), y=yd, n_global=n)
    with K.path_flags(no_fused_bwd=True):
        gb, sb = K.mlp_backward(net, xd, drop(), y=yd, n_global=n)
    torch.cuda.synchronize()
    a, b = ga.double().cpu().numpy(), gb.double().cpu().numpy()
    names, shapes, offs, _ = K.param_layout(net.width, net.n_hidden)
    worst = {}
    for nm, shp, o in zip(names, shapes, offs):
        cnt = int(np.prod(shp))
        worst[nm] = float(np.max(np.abs(a[o:o + cnt] - b[o:o + cnt])) / (np.max(np.abs(b[o:o + cnt])) + 1e-300))
    r = {"max_nrel": max(worst.values()), "worst": max(worst, key=worst.get), "bitwise_repeat": bool(torch.equal(ga, ga2)),
         "loss_rel": float(np.max(np.abs(sa.cpu().numpy() - sb.cpu().numpy()) / (np.abs(sb.cpu().numpy()) + 1e-300)))}

    def step_ms(k=20):
        model.train_dnn(3, verbose=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        model.train_dnn(k, verbose=False)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k

    r["fused_ms"] = step_ms(200 if n <= 100_000 else 20)
    with K.path_flags(no_fused_bwd=True):
        r["two_kernel_ms"] = step_ms(200 if n <= 100_000 else 20)
    out[n] = r
    print(n, json.dumps(r), flush=True)
print(json.dumps(out))
