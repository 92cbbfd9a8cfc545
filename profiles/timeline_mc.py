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
    # three phases per pass: hidden layers 1, 2 (stamps 0..5) and heads + tail (stamps 0..7)
    for half in (0, 1):
        print(f"half {half}: phase start | deltas between consecutive stamps | gap to next phase")
        for i in range(3, 30):
            r = t[half, i]
            ks = 8 if i % 3 == 2 else 6
            d = [r[k] - r[k - 1] for k in range(1, ks)]
            gap = t[half, i + 1, 0] - r[ks - 1]
            kind = "heads" if i % 3 == 2 else f"layer{1 + i % 3}"
            print(f"   {i:3d} {kind:7s} {r[0] - t0:9d} | " + " ".join(f"{x:6d}" for x in d) + f" | {gap:6d}")
    print("hidden: signal, draw, wait-done, ldtm, epilogue.  heads: signal, draws(6 blocks), wait-done, v0 tanh+select, Wv1 dots, hand-over, v1/logvar/Welford; gap = layer-0 staging of the next pass")
