"""Clock-stamp timeline of the tensor-core wgrad kernel (debug build -DPINN_TIMELINE): `build` here, `run` on the GPU box."""
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
    X, Y, sx, sy = build_problem(1_000_000, 2)
    torch.manual_seed(0)
    model = b200pinn.PhysicsInformedNN(X, Y, LAYERS, sx, sy, P_TRAIN, True)
    model.train_dnn(3, verbose=False)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (2 * 64 * 4))()
    lib = abi.lib()
    lib.pinn_debug_timeline_wgrad.argtypes = [ctypes.c_void_p]
    assert lib.pinn_debug_timeline_wgrad(buf) == 0
    t = np.array(buf, dtype=np.int64).reshape(2, 64, 4)
    t0 = t[1, 0, 0]
    print("stage | loader: start  wait_done  store  load+arrive | mma: wake(rel. loader arrive)  issue | stage period")
    for i in range(0, 40):
        L, M = t[1, i], t[0, i]
        print(f"{i + 16:4d} | {L[0] - t0:8d} {L[1] - L[0]:8d} {L[2] - L[1]:8d} {L[3] - L[2]:8d} | {M[0] - L[3]:8d} {M[1] - M[0]:8d} | {t[1, i + 1, 0] - L[0]:8d}")
