"""configs[0] end to end on one GPU: the reference's own training schedule (01:2143-2153: train_dnn 4001, train_lambda
4001 x2, train_dnn 8001, train_thermal 10001, train_hydrogen 8001, train_oxygen 8001 = 46 007 full-batch steps) and the
MC-dropout sweep, on N = 20 000 synthetic normal-operation samples with the 3x64 net, through the public drop-in classes."""
import contextlib, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200pinn
from b200pinn.synthetic import make_scaled_dataset

n = 20000
x, y, sx, sy = make_scaled_dataset(n, seed=1)
torch.manual_seed(0)
m = b200pinn.PhysicsInformedNN(torch.tensor(x), torch.tensor(y), [8, 64, 64, 64, 1], sx, sy, 0.2, True)
m.train_dnn(10, verbose=False)                       # warm-up (workspace allocation, module load)
torch.cuda.synchronize()
t = {}
def timed(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        fn()
    torch.cuda.synchronize(); t[name] = time.perf_counter() - t0
timed("train_dnn(4001)", lambda: m.train_dnn(4001))
timed("train_lambda(4001, False)", lambda: m.train_lambda(4001, False))
timed("train_lambda(4001, True)", lambda: m.train_lambda(4001, True))
timed("train_dnn(8001)", lambda: m.train_dnn(8001))
timed("train_thermal(10001)", lambda: m.train_thermal(10001))
timed("train_hydrogen(8001)", lambda: m.train_hydrogen(8001))
timed("train_oxygen(8001)", lambda: m.train_oxygen(8001))
X = torch.tensor(x)
timed("get_MC_samples(T=50)", lambda: b200pinn.get_MC_samples(m, X, sx, mc_times=50, dropout=0.4))
timed("get_MC_samples(T=2000)", lambda: b200pinn.get_MC_samples(m, X, sx, mc_times=2000, dropout=0.4))
for k, v in t.items():
    print(f"{k:28s} {v:8.3f} s")
print(f"{'total':28s} {sum(t.values()):8.3f} s   (46 007 training steps + two sweeps)")
