"""Quick timing on one B200: MC sweep (T=50, N=1M), eval forward, train_dnn step, K3 at 1M / 8M rows, RF(t).
usage: python profiles/quick_time.py [n] [what,...]   what in {mc, fwd, train, res, rf, wide}"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200pinn
from b200pinn import _abi, kernels as K
if os.environ.get("B200PINN_LIB"):        # A/B builds: python -c "...build(extra_flags=[...], out=..., objdir=...)"
    _abi.LIB_PATH = os.environ["B200PINN_LIB"]
from bench import build_problem, LAYERS, P_TRAIN, P_MC, T_PASSES
if os.environ.get("B200PINN_NO_TC3"):         # A/B: the two-group 3xTF32 form of the 64-wide forward / MC kernel
    K.set_default_path_flags(no_tc3=True)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
what = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else {"mc", "fwd", "train", "res"}
X, Y, sx, sy = build_problem(n, 2)
torch.manual_seed(0)
model = b200pinn.PhysicsInformedNN(X, Y, LAYERS, sx, sy, P_TRAIN, True)
model.dnn.eval()
xd = model.x.detach()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=xd.device)


def timed(fn, reps=5, warm=2, do_flush=False):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if do_flush:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps


res = {}
if "mc" in what:
    mc = lambda: b200pinn.mc_dropout_device(model.dnn, xd, T_PASSES, P_MC, seed=1234)
    res["mc_ms"] = timed(mc, do_flush=True)
    o = mc()
    res["mc_checksum_e_u"] = float(o["e_u"].double().sum())
    res["mc_checksum_a_u"] = float(o["a_u"].double().sum())
    res["mc_T1000_ms"] = timed(lambda: b200pinn.mc_dropout_device(model.dnn, xd, 1000, P_MC, seed=1234), reps=2, warm=1)
    xs = xd[:125_000].contiguous()
    res["mc_T1000_125k_ms"] = timed(lambda: b200pinn.mc_dropout_device(model.dnn, xs, 1000, P_MC, seed=1234), reps=3, warm=1)
if "fwd" in what:
    with torch.no_grad():
        res["fwd_ms"] = timed(lambda: model.net_u(xd))
if "train" in what:
    model.train_dnn(3, verbose=False)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    model.train_dnn(10, verbose=False)
    b.record()
    torch.cuda.synchronize()
    res["train_ms"] = a.elapsed_time(b) / 10
    model.dnn.eval()
if "res" in what:
    with torch.no_grad():
        u = model.net_u(xd)[0].reshape(-1).contiguous()
    yv = model.u.reshape(-1).contiguous()
    sc, lam = model._scalers(sx), model._lambdas()
    sums = torch.empty(_abi.S_COUNT, device=xd.device, dtype=torch.float64)
    fam = _abi.FAM_V | _abi.FAM_DATA
    def back_to_back(bufs, reps=30):
        """`reps` launches inside ONE event pair, rotating over input copies that together exceed the 126 MB L2: the GPU
        never waits for the host between launches, and no launch finds its input in L2."""
        for xb_, ub_, yb_ in bufs:
            K.residuals(xb_, ub_, yb_, sc, lam, fam, sums=sums)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()              # one graph of `reps` launches: no host latency between them
        with torch.cuda.graph(g):
            for i in range(reps):
                xb_, ub_, yb_ = bufs[i % len(bufs)]
                K.residuals(xb_, ub_, yb_, sc, lam, fam, sums=sums)
        g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    small = [(xd.clone(), u.clone(), yv.clone()) for _ in range(6)]          # 6 x 40 MB = 240 MB > L2
    t = back_to_back(small)
    res["res_1M_us"] = 1e3 * t
    res["res_1M_gbs"] = n * 40 / t / 1e6
    xb, ub, yb = xd.repeat(8, 1).contiguous(), u.repeat(8).contiguous(), yv.repeat(8).contiguous()
    t = back_to_back([(xb, ub, yb)], reps=20)                                # 320 MB per launch > L2
    res["res_8M_us"] = 1e3 * t
    res["res_8M_gbs"] = 8 * n * 40 / t / 1e6
    ref = torch.empty_like(sums)
    K.residuals(xb, ub, yb, sc, lam, fam, sums=ref)
    res["res_8M_checksum_FV2"] = float(ref[_abi.S["FV2"]])
    res["res_8M_checksum_GA2"] = float(ref[_abi.S["GA2"]])
if "rf" in what:
    from b200pinn.export import export_rows_device
    from b200pinn.rf import rf_device
    seg = [0] + [n * (i + 1) // 13 for i in range(13)]
    rows, cc = export_rows_device(model, xd, model.u.reshape(-1).contiguous(), seg, 12, 5, P_MC, sx, sy, seed=1, want_rf_cols=True)
    fleet = rows.unsqueeze(0).expand(8, -1, -1).contiguous()
    t = timed(lambda: rf_device(fleet), reps=5, do_flush=True)
    res["rf_8x1M_ms"] = t
    res["rf_hbm_gbs"] = 8 * n * (22 * 8 + 16) / t / 1e6
    fc = cc.unsqueeze(0).expand(8, -1, -1).contiguous()
    t = timed(lambda: rf_device(fc), reps=5, do_flush=True)
    res["rf_compact_8x1M_ms"] = t
    res["rf_compact_equiv_gbs"] = 8 * n * (22 * 8 + 16) / t / 1e6
    res["rf_compact_actual_gbs"] = 8 * n * (48 + 48 + 8 + 8 + 8 + 8 + 8) / t / 1e6
for k, v in res.items():
    print(f"{k:28s} {v:.6g}")
