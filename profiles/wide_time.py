import os, sys
sys.path.insert(0, "/root/repo")
import torch, b200pinn
from bench import build_problem
for layers, n in (([8,256,256,256,1], 262144), ([8,256,256,256,256,256,256,1], 262144)):
    X, Y, sx, sy = build_problem(n, 2)
    torch.manual_seed(0)
    model = b200pinn.PhysicsInformedNN(X, Y, layers, sx, sy, 0.2, True)
    model.dnn.eval(); xd = model.x.detach()
    def timed(fn, reps=3, warm=1):
        for _ in range(warm): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    T = 10
    t_mc = timed(lambda: b200pinn.mc_dropout_device(model.dnn, xd, T, 0.4, seed=1))
    L = len(layers) - 2
    flop_pass = {3: 344704, 6: 737920}[L]; flop_train = {3: 1042304, 6: 2221952}[L]
    print(layers, "MC ms", t_mc, "TFLOP/s", n * T * flop_pass / t_mc / 1e9)
    model.train_dnn(2, verbose=False); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); model.train_dnn(3, verbose=False); b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) / 3
    print(layers, "train ms", t, "TFLOP/s", n * flop_train / t / 1e9)
