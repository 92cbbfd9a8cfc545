"""Resident-activation 256-wide kernel (csrc/mlp_wide_res.cu) vs the per-layer GEMM path and the FFMA path on the same
Philox stream / injected masks, then the timing of the MC sweep on both tensor-core paths.

    python profiles/wide_res_check.py [quick]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

if os.environ.get("B200PINN_LIB"):        # A/B builds (profiles/build_variant.py)
    from b200pinn import _abi
    _abi.LIB_PATH = os.environ["B200PINN_LIB"]
import b200pinn
from b200pinn import kernels as K
from b200pinn.synthetic import make_scaled_dataset

dev = torch.device("cuda:0")


def nrel(a, b):
    a, b = a.double().cpu().numpy(), b.double().cpu().numpy()
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def random_net(layers, seed):
    torch.manual_seed(seed)
    dnn = b200pinn.DNN(0.25, True, layers)
    with torch.no_grad():
        dnn.var_layers[5].bias.fill_(0.3)
    return dnn.to(dev).eval()


def run(dnn, xd, T, p, masks=None):
    net = K.net_from_module(dnn)
    u0, s0 = K.mlp_forward(net, xd)
    u1, s1 = K.mlp_forward(net, xd, K.make_dropout(p, seed=9, pass_offset=4))
    mc = b200pinn.mc_dropout_device(dnn, xd, T, p, seed=77, masks=masks, raw=True)
    torch.cuda.synchronize()
    return dict(u0=u0, s0=s0, u1=u1, s1=s1, pm=mc["pred_mean"], a_u=mc["a_u"], e_u=mc["e_u"], mean=mc["mean"])


MODE = sys.argv[1] if len(sys.argv) > 1 else "all"
worst = 0.0
for layers, n, T in () if MODE == "time" else (([8, 256, 256, 256, 1], 1, 2), ([8, 256, 256, 256, 1], 129, 3), ([8, 256, 256, 1], 1000, 2),
                     ([8, 256, 256, 256, 256, 256, 256, 1], 300, 2), ([8, 256, 256, 256, 1], 40000, 5), ([8, 256, 256, 256, 1], 700, 50)):
    x, _, _, _ = make_scaled_dataset(max(n, 64), seed=41)
    xd = torch.tensor(x[:n], device=dev)
    dnn = random_net(layers, 12)
    a = run(dnn, xd, T, 0.3)
    with K.path_flags(no_wide_resident=True):
        b = run(dnn, xd, T, 0.3)
    with K.path_flags(no_wide_tc=True):
        c = run(dnn, xd, T, 0.3)
    L, H = len(layers) - 2, 256
    if n <= 1000:
        mk = torch.tensor((np.random.default_rng(4).random((T, n, L * H + H // 2)) >= 0.3).astype(np.uint8), device=dev)
        ai = run(dnn, xd, T, 0.3, masks=mk)
        with K.path_flags(no_wide_tc=True):
            ci = run(dnn, xd, T, 0.3, masks=mk)
    else:
        ai = ci = None
    for k in a:
        e1, e2 = nrel(a[k], b[k]), nrel(a[k], c[k])
        e3 = nrel(ai[k], ci[k]) if ai is not None else 0.0
        worst = max(worst, e1, e2, e3)
        print(f"L={L} n={n} T={T} {k:5s} resident vs gemm {e1:.2e}  vs ffma {e2:.2e}  injected vs ffma {e3:.2e}  gemm vs ffma {nrel(b[k], c[k]):.2e}")
print("WORST", worst, "PASS" if worst < 1e-5 else "FAIL")

if MODE == "quick":
    sys.exit(0)


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


CASES = (([8, 256, 256, 256, 1], 262144, 10), ([8, 256, 256, 256, 256, 256, 256, 1], 262144, 10),
         ([8, 256, 256, 256, 1], 20000, 50), ([8, 256, 256, 256, 1], 20000, 2000), ([8, 256, 256, 256, 1], 1000000, 50))
for layers, n, T in CASES[:2] if MODE == "time" else CASES:
    x, _, _, _ = make_scaled_dataset(n, seed=2)
    xd = torch.tensor(x, device=dev)
    dnn = random_net(layers, 3)
    L = len(layers) - 2
    flop_pass = {3: 344704, 6: 737920}[L]
    reps = 1 if n * T > 2e7 else 3
    t_res = timed(lambda: b200pinn.mc_dropout_device(dnn, xd, T, 0.4, seed=1), reps=reps)
    t_fwd = timed(lambda: K.mlp_forward(K.net_from_module(dnn), xd), reps=reps)
    if MODE == "time":
        print(f"L={L} n={n} T={T}: MC sweep resident {t_res:.3f} ms ({n * T * flop_pass / t_res / 1e9:.1f} TFLOP/s) | eval forward {t_fwd:.3f} ms", flush=True)
        continue
    with K.path_flags(no_wide_resident=True):
        t_gemm = timed(lambda: b200pinn.mc_dropout_device(dnn, xd, T, 0.4, seed=1), reps=reps)
        t_fwd_g = timed(lambda: K.mlp_forward(K.net_from_module(dnn), xd), reps=reps)
    print(f"L={L} n={n} T={T}: MC sweep resident {t_res:.3f} ms ({n * T * flop_pass / t_res / 1e9:.1f} TFLOP/s)  per-layer GEMM {t_gemm:.3f} ms"
          f"  | eval forward resident {t_fwd:.3f} ms, GEMM {t_fwd_g:.3f} ms")
