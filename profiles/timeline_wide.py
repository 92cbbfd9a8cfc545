"""Clock-stamp timeline of the resident-activation 256-wide kernel (debug build -DPINN_TIMELINE): CTA 0, compute warps 0
(column quarter 0) and 12 (quarter 3), and the MMA warp.  `build` on the CPU box, `run` on the GPU box."""
import ctypes, importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "physics-informed-neural-network-for-explainable-fault-diagnosis-in-fuel-cells_b200")
D = os.path.join(PKG, "build", "timeline_wide")
if sys.argv[1] == "build":
    spec = importlib.util.spec_from_file_location("b", os.path.join(PKG, "build.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    os.makedirs(D, exist_ok=True)
    print(m.build(force=True, extra_flags=["-DPINN_TIMELINE"] + sys.argv[2:], out=os.path.join(D, "libb200pinn.so"), objdir=D))
else:
    sys.path.insert(0, ROOT)
    import b200pinn._abi as abi
    abi.LIB_PATH = os.path.join(D, "libb200pinn.so")
    import numpy as np
    import torch, b200pinn
    from bench import build_problem, P_TRAIN, P_MC
    layers = [8, 256, 256, 256, 1] if len(sys.argv) < 3 else [8] + [256] * int(sys.argv[2]) + [1]
    X, Y, sx, sy = build_problem(262144, 2)
    torch.manual_seed(0)
    model = b200pinn.PhysicsInformedNN(X, Y, layers, sx, sy, P_TRAIN, True)
    model.dnn.eval(); xd = model.x.detach()
    for _ in range(2):
        b200pinn.mc_dropout_device(model.dnn, xd, 10, P_MC, seed=1234)
    torch.cuda.synchronize()
    cb = (ctypes.c_longlong * (2 * 512 * 8))()
    mb = (ctypes.c_longlong * (512 * 20))()
    lib = abi.lib()
    lib.pinn_debug_wide_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    assert lib.pinn_debug_wide_timeline(cb, mb) == 0
    c = np.array(cb, dtype=np.int64).reshape(2, 512, 8)
    m = np.array(mb, dtype=np.int64).reshape(512, 20)
    L = len(layers) - 2
    per_pass_c, per_pass_m = L + 2, L + 1          # compute records: staging, L-1 hidden, heads, v1;  MMA phases
    p0 = 3                                         # a pass in the middle of the first tile
    t0 = c[0, p0 * per_pass_c, 1]
    names = {100: "stage"}
    print("compute warps: kind | t(before wait) | wait | ld+release | slab deltas ... (clk)")
    for w in (0, 1):
        print(f" warp {'0 (q0)' if w == 0 else '12 (q3)'}")
        for i in range(p0 * per_pass_c, (p0 + 2) * per_pass_c):
            r = c[w, i]
            kind = {100: "stage", 200: "heads", 300: "v1"}.get(int(r[0]), f"hid{int(r[0])}")
            ns = {"stage": 4, "heads": 2, "v1": 1}.get(kind, 4)
            d = [int(r[k] - r[k - 1]) for k in range(2, 4 + ns)]
            print(f"   {kind:6s} {int(r[1] - t0):8d} | wait {d[0]:6d} | ld {d[1]:5d} | " + " ".join(f"{x:5d}" for x in d[2:]))
    print("MMA warp: phase | t(accfree passed) | issue time of each slab relative to it | last commit")
    for i in range(p0 * per_pass_m, (p0 + 2) * per_pass_m):
        r = m[i]
        ph = int(r[0]); ns = 8 if ph == L else 16
        print(f"   ph{ph} {int(r[1] - t0):8d} | " + " ".join(f"{int(r[2 + k] - r[1]):5d}" for k in range(ns)) + f" | {int(r[18] - r[1]):6d}")
    # fine stamps inside the 8 column steps of one hidden epilogue (record 17 of each stamped warp): tanh | select (Philox) |
    # fp16 split | tcgen05.st issue | wait::st | fence + arrive
    sb = (ctypes.c_longlong * (4 * 8 * 8))()
    lib.pinn_debug_wide_steps.argtypes = [ctypes.c_void_p]
    if lib.pinn_debug_wide_steps(sb) == 0:
        st = np.array(sb, dtype=np.int64).reshape(4, 8, 8)
        for w in range(4):
            if st[w, 0, 0] == 0:
                continue
            print(f" steps of warp {'0' if w < 2 else '12'}, hidden layer {1 + w % 2}: start | tanh select split st wait arrive | gap to next step")
            for k in range(8):
                r = st[w, k]
                d = [int(r[j + 1] - r[j]) for j in range(6)]
                gap = int(st[w, k + 1, 0] - r[6]) if k < 7 else 0
                print(f"   step {k}: {int(r[0] - st[w, 0, 0]):6d} | " + " ".join(f"{x:5d}" for x in d) + f" | {gap:5d}")

