"""Clock-stamp timeline of one group's hidden-layer phases in the MC kernel (debug build -DPINN_TIMELINE).
`build` on the CPU box, `run` on the GPU box."""
import ctypes, importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "physics-informed-neural-network-for-explainable-fault-diagnosis-in-fuel-cells_b200")
D = os.path.join(PKG, "build", "timeline")
if sys.argv[1] == "build":
    spec = importlib.util.spec_from_file_location("b", os.path.join(PKG, "build.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    os.makedirs(D, exist_ok=True)
    extra = ["-DPINN_TIMELINE"] + ([f"-DPINN_ABL={sys.argv[2]}"] if len(sys.argv) > 2 else [])
    print(m.build(force=True, extra_flags=extra, out=os.path.join(D, "libb200pinn.so"), objdir=D))
else:
    sys.path.insert(0, ROOT)
    import b200pinn._abi as abi
    abi.LIB_PATH = os.path.join(D, "libb200pinn.so")
    import torch, b200pinn
    from bench import build_problem, LAYERS, P_TRAIN, P_MC, T_PASSES
    X, Y, sx, sy = build_problem(1_000_000, 2)
    torch.manual_seed(0)
    model = b200pinn.PhysicsInformedNN(X, Y, LAYERS, sx, sy, P_TRAIN, True)
    model.dnn.eval(); xd = model.x.detach()
    for _ in range(3):
        b200pinn.mc_dropout_device(model.dnn, xd, T_PASSES, P_MC, seed=1234)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (3 * 64 * 8))()
    lib = abi.lib()
    lib.pinn_debug_timeline.argtypes = [ctypes.c_void_p]
    assert lib.pinn_debug_timeline(buf) == 0
    import numpy as np
    t = np.array(buf, dtype=np.int64).reshape(3, 64, 8)
    t0 = t[0, 0, 0]
    names = ["pre-signal", "post-signal", "post-draw", "post-done", "post-ldtm", "post-epi"]
    for half in (0, 1):
        print(f"half {half}:  phase  start   " + "  ".join(f"d({n})" for n in names[1:]) + "   gap-to-next")
        for i in range(2, 26):
            r = t[half, i]
            d = [r[k] - r[k - 1] for k in range(1, 6)]
            gap = t[half, i + 1, 0] - r[5]
            print(f"   {i:3d} {r[0] - t0:9d}   " + "  ".join(f"{x:10d}" for x in d) + f"   {gap:8d}"
                  + (f"   mma: wake@{t[2, i, 0] - t0} (+{t[2, i, 0] - r[1]} after this thread's signal), issue {t[2, i, 1] - t[2, i, 0]}, done seen +{r[3] - t[2, i, 1]}" if half == 0 else ""))
