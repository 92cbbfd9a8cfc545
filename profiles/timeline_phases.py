"""Coarse phase stamps of K2a / K2b (CTA 0) in a -DPINN_TIMELINE build: where a train_dnn step spends its time at a given N.
`build` here, `run [n]` on the GPU box."""
import ctypes, importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "physics-informed-neural-network-for-explainable-fault-diagnosis-in-fuel-cells_b200")
D = os.path.join(PKG, "build", "timeline")
if sys.argv[1] == "build":
    spec = importlib.util.spec_from_file_location("b", os.path.join(PKG, "build.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    os.makedirs(D, exist_ok=True)
    print(m.build(force=True, extra_flags=["-DPINN_TIMELINE"], out=os.path.join(D, "libb200pinn.so"), objdir=D))
else:
    sys.path.insert(0, ROOT)
    import b200pinn._abi as abi
    abi.LIB_PATH = os.path.join(D, "libb200pinn.so")
    import numpy as np, torch, b200pinn
    from bench import build_problem, LAYERS, P_TRAIN
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    X, Y, sx, sy = build_problem(n, 2)
    torch.manual_seed(0)
    model = b200pinn.PhysicsInformedNN(X, Y, LAYERS, sx, sy, P_TRAIN, True)
    model.train_dnn(5, verbose=False)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 16)()
    lib = abi.lib()
    lib.pinn_debug_timeline_phases.argtypes = [ctypes.c_void_p]
    assert lib.pinn_debug_timeline_phases(buf) == 0
    t = np.array(buf, dtype=np.int64).reshape(2, 8)
    ghz = 1.965
    for name, row, labels in (("K2a", t[1], ("prologue (weights staged)", "tiles", "loss partials")),
                              ("K2b", t[0], ("prologue", "streaming + MMA", "TMEM -> partial"))):
        d = np.diff(row[:4]) / ghz / 1e3
        print(f"n={n} {name} CTA 0: " + ", ".join(f"{l} {v:.2f} us" for l, v in zip(labels, d)) + f"  (total {d.sum():.2f} us)")
