"""A/B of K2a's tile -> (CTA, group) mapping: `build` compiles the former pair-first mapping into build/pairfirst/, `run` times
train_dnn with either library in separate processes.  usage: python profiles/ab_k2a_map.py build | run <n> [pair]"""
import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "physics-informed-neural-network-for-explainable-fault-diagnosis-in-fuel-cells_b200")
D = os.path.join(PKG, "build", "pairfirst")
if sys.argv[1] == "build":
    spec = importlib.util.spec_from_file_location("b", os.path.join(PKG, "build.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    os.makedirs(D, exist_ok=True)
    print(m.build(force=True, extra_flags=["-DPINN_K2A_PAIR_FIRST"], out=os.path.join(D, "libb200pinn.so"), objdir=D))
else:
    sys.path.insert(0, ROOT)
    import b200pinn._abi as abi
    pair = len(sys.argv) > 3 and sys.argv[3] == "pair"
    if pair:
        abi.LIB_PATH = os.path.join(D, "libb200pinn.so")
    import torch, b200pinn
    from b200pinn.synthetic import make_scaled_dataset
    n = int(sys.argv[2])
    x, y, sx, sy = make_scaled_dataset(n, seed=1)
    torch.manual_seed(0)
    m = b200pinn.PhysicsInformedNN(torch.tensor(x), torch.tensor(y), [8, 64, 64, 64, 1], sx, sy, 0.2, True)
    m.train_dnn(5, verbose=False)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); m.train_dnn(300, verbose=False); b.record(); torch.cuda.synchronize()
    print(f"n={n} {'pair-first' if pair else 'spread    '}: {1e3 * a.elapsed_time(b) / 300:.1f} us/step")
