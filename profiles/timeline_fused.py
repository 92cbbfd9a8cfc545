"""Phase stamps of the fused K2 kernel (CTA 0: compute thread 0 and the MMA warp) in a -DPINN_TIMELINE build.
`build` here, `run [n]` on the GPU box."""
import ctypes, importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "physics-informed-neural-network-for-explainable-fault-diagnosis-in-fuel-cells_b200")
D = os.path.join(PKG, "build", "timeline")
if sys.argv[1] == "build":
    spec = importlib.util.spec_from_file_location("b", os.path.join(PKG, "build.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    os.makedirs(D, exist_ok=True)
    print(m.build(force=True, extra_flags=["-DPINN_TIMELINE"] + sys.argv[2:], out=os.path.join(D, "libb200pinn.so"), objdir=D))
else:
    sys.path.insert(0, ROOT)
    import b200pinn._abi as abi
    abi.LIB_PATH = os.path.join(D, "libb200pinn.so")
    import numpy as np, torch, b200pinn
    from bench import build_problem, LAYERS, P_TRAIN
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    X, Y, sx, sy = build_problem(n, 2)
    torch.manual_seed(0)
    model = b200pinn.PhysicsInformedNN(X, Y, LAYERS, sx, sy, P_TRAIN, True)
    model.train_dnn(3, verbose=False)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 512)()
    lib = abi.lib()
    lib.pinn_debug_timeline_fused.argtypes = [ctypes.c_void_p]
    assert lib.pinn_debug_timeline_fused(buf) == 0
    t = np.array(buf, dtype=np.int64).reshape(2, 8, 32)
    names = {0: "tile start", 1: "x parked / ready", 2: "kb0 drawn", 3: "chain L0 done", 4: "kb1 drawn", 5: "chain F1 done", 6: "kb2 drawn",
             7: "chain F2 done", 8: "kbv drawn", 9: "chain FH done", 10: "av0 parked / ready", 11: "32->16 product done", 12: "v1 handed over", 13: "scalar part / ready",
             14: "16->32 product done", 15: "heads^T operand ready", 16: "wg(0T) waited", 17: "batch H staged", 18: "chain BH done", 19: "dz2 / ready",
             20: "wg(H) waited", 21: "batch 2 staged", 22: "chain B2 done", 23: "dz1 / ready", 24: "wg(2) waited", 25: "batch 1 staged",
             26: "chain B1 done", 27: "dz0 done", 28: "wg(1) waited", 29: "batch 0T staged",
             30: "(av0 computed)", 31: "(av0 stores issued)"}
    for tile in (1, 2):
        base = t[0, tile, 0]
        if base == 0:
            continue
        print(f"--- n={n} tile #{tile} of CTA 0, compute thread 0 (clk since tile start; delta)")
        prev = base
        for i in list(range(10)) + [30, 31] + list(range(10, 30)):
            v = t[0, tile, i]
            if v:
                print(f"  {names.get(i, i):>18}: {v - base:7d}  (+{v - prev})")
                prev = v
        print(f"  next tile start   : {t[0, tile + 1, 0] - base:7d}")
        print("    MMA warp (per phase p: ready seen, weights full, chain issued, wgrad issued) relative to the same origin")
        for p in range(7):
            row = t[1, tile, 4 * p:4 * p + 4]
            print("    p=%d " % p + " ".join(f"{(v - base) if v else -1:7d}" for v in row))
