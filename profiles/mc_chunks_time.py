"""MC sweep at the shard sizes of configs[2] (T = 1000; 1M / 500k / 250k / 125k rows = 1 / 2 / 4 / 8 GPUs) and at configs[0] size
(N = 20 000, T = 2000 = the reference's export sweep 01:2156-2158): ms per sweep and checksum."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200pinn
from bench import build_problem, LAYERS, P_TRAIN, P_MC

X, Y, sx, sy = build_problem(1_000_000, 2)
torch.manual_seed(0)
model = b200pinn.PhysicsInformedNN(X, Y, LAYERS, sx, sy, P_TRAIN, True)
model.dnn.eval()
xd = model.x.detach()


def timed(fn, reps=3, warm=1):
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
for n, T in ((1_000_000, 1000), (500_000, 1000), (250_000, 1000), (125_000, 1000), (20_000, 2000), (20_000, 50), (1_000_000, 50)):
    xs = xd[:n].contiguous()
    f = lambda: b200pinn.mc_dropout_device(model.dnn, xs, T, P_MC, seed=1234)
    ms = timed(f, reps=2 if n * T > 2e8 else 5)
    o = f()
    res[f"n={n},T={T}"] = {"ms": round(ms, 3), "e_u_sum": float(o["e_u"].double().sum()), "a_u_sum": float(o["a_u"].double().sum())}
    print(f"n={n} T={T}: {ms:.3f} ms  ({n * T / ms / 1e6:.3f} G sample*passes/s)", flush=True)
print(json.dumps(res))
