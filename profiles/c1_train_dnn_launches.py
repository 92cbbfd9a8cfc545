"""train_dnn at configs[0] size (N = 20 000, 3x64): a few warm-up steps, then `steps` timed ones.  Run plain for the
per-step wall / device time, or under `ncu --metrics gpu__time_duration.sum` for the per-kernel launch list.
usage: python profiles/c1_train_dnn_launches.py [n] [steps] [width] [hidden layers]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200pinn
if os.environ.get("B200PINN_LIB"):        # A/B builds (profiles/build_variant.py)
    import b200pinn._abi as _abi
    _abi.LIB_PATH = os.environ["B200PINN_LIB"]
from b200pinn.synthetic import make_scaled_dataset

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
width = int(sys.argv[3]) if len(sys.argv) > 3 else 64
hidden = int(sys.argv[4]) if len(sys.argv) > 4 else 3
x, y, sx, sy = make_scaled_dataset(n, seed=1)
torch.manual_seed(0)
m = b200pinn.PhysicsInformedNN(torch.tensor(x), torch.tensor(y), [8] + [width] * hidden + [1], sx, sy, 0.2, True)
if os.environ.get("B200PINN_PDL", "1") == "0":
    b200pinn.kernels.set_default_path_flags(dependent_launch=0)
m.train_dnn(5, verbose=False)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
a.record()
m.train_dnn(steps, verbose=False)
b.record()
torch.cuda.synchronize()
print(f"{hidden}x{width} n={n} steps={steps} pdl={os.environ.get('B200PINN_PDL', '1')}: wall {1e6 * (time.perf_counter() - t0) / steps:.1f} us/step, device {1e3 * a.elapsed_time(b) / steps:.1f} us/step")
