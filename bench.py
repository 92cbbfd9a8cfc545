#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the PINN hot path.

    python bench.py --gpus N --steps K --warmup W            (our CUDA path, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...   (reference algorithm on host cores)

Workload (BASELINE.json configs[1]): the 3x64 stack-voltage PINN on N = 1M synthetic
normal-operation samples per GPU.  One "step" is one MC-dropout sweep of T = 50 stochastic
passes over that batch (kernel K4) -- `metric` is MC-dropout sample*passes/s; the same JSON
line also carries PINN train steps/s (K2 + reduce + Adam) and the lambda-phase step (K3).
`value` is timed with inputs resident in HBM; `e2e` is the same sweep through the public
`get_MC_samples` with HOST tensors (H2D of X and D2H of the three result vectors inside).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LAYERS = [8, 64, 64, 64, 1]
N_PER_GPU = 1_000_000
T_PASSES = 50
P_MC = 0.4           # 01:2158 uses dropout=0.4 for the export sweep
P_TRAIN = 0.2        # 01:2141
FLOP_PER_SAMPLE_PASS = 21_664      # SURVEY 8d, 3x64, layer 0 hoisted
FLOP_PER_TRAIN_SAMPLE = 67_040     # SURVEY 8d, fwd + dgrad + wgrad
RES_BYTES_PER_SAMPLE = 40          # x row 32 B + u 4 B + y 4 B (train_lambda form)
METRIC = "mc_dropout_sample_passes_per_s"
UNIT = "sample*passes/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), bf16=float(p["bf16_tflops"]), src="measured (MEASURED_PEAKS.json, burst)")
    return dict(hbm=6650.0, bf16=1590.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [t.strip() for t in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_problem(n, seed):
    import torch
    from b200pinn.synthetic import make_scaled_dataset

    x, y, sx, sy = make_scaled_dataset(n, seed)
    return torch.tensor(x), torch.tensor(y), sx, sy


# ------------------------------------------------------------------------- reference arm
def cpu_port_rates(n_mc, n_train, threads):
    """Time the reference algorithm (oracle/torch_port.py) on the host cores: one
    get_MC_samples with T'=1 and one train_dnn step."""
    import torch
    from oracle.torch_port import PortPINN, get_MC_samples_port

    torch.set_num_threads(threads)
    X, Y, sx, sy = build_problem(max(n_mc, n_train), 2)
    torch.manual_seed(0)
    net = PortPINN(X[:n_mc], Y[:n_mc], LAYERS, sx, sy, P_TRAIN)
    get_MC_samples_port(net, X[:2000], sx, 1, P_MC)            # warm-up
    t0 = time.perf_counter()
    get_MC_samples_port(net, X[:n_mc], sx, 1, P_MC)
    t_mc = time.perf_counter() - t0
    tr = PortPINN(X[:n_train], Y[:n_train], LAYERS, sx, sy, P_TRAIN)
    step = tr.make_dnn_trainer()
    t0 = time.perf_counter()
    step()
    t_tr = time.perf_counter() - t0
    return n_mc / t_mc, (n_train / N_PER_GPU) / t_tr * 1.0, t_mc, t_tr


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import io
    import contextlib
    import torch
    from oracle.torch_port import PortPINN, get_MC_samples_port

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    n = 200_000
    X, Y, sx, sy = build_problem(n, 2)
    torch.manual_seed(0)
    net = PortPINN(X, Y, LAYERS, sx, sy, P_TRAIN)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            get_MC_samples_port(net, X, sx, 1, P_MC)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = n * 1 * len(times) / total
    sample = f"N={n} x T'=1 per step (1 eval pass + 1 dropout pass, each with the discarded 2nd forward, 01:1407)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: 3x64 PINN, N=1M/GPU, MC-dropout sweep T=50 (CPU arm: bounded sample)",
                       "layers": LAYERS, "dropout": P_MC},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import b200pinn
    from b200pinn import _abi, kernels as K

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's "NCCL version ..." banner (printed to stdout at communicator
        # creation when the image sets NCCL_DEBUG) goes to stderr instead
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.n
    X, Y, sx, sy = build_problem(n, 2 + rank)
    torch.manual_seed(0)
    model = b200pinn.PhysicsInformedNN(X, Y, LAYERS, sx, sy, P_TRAIN, True)
    model.dnn.eval()
    xd = model.x.detach()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    K_, W_ = args.steps, args.warmup

    def timed(fn, steps, warm, do_flush=True):
        for _ in range(warm):
            fn()
        barrier()
        evs = []
        launches0 = K.LAUNCHES
        for _ in range(steps):
            if do_flush:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        barrier()
        tot = sum(a.elapsed_time(b) for a, b in evs) / 1e3
        return max_over_ranks(tot), K.LAUNCHES - launches0

    sampler = ClockSampler(local) if rank == 0 else None
    # --- headline: MC-dropout sweep, inputs resident in HBM
    seed = 1234
    mc = lambda: b200pinn.mc_dropout_device(model.dnn, xd, T_PASSES, P_MC, seed=seed, sample_offset=rank * n)
    t_mc, launches = timed(mc, K_, W_)
    clocks = sampler.stop() if sampler else None
    value = world * n * T_PASSES * K_ / t_mc
    # --- train steps (K2 + reduce + [all-reduce] + Adam), back to back
    steps_tr = max(K_, 5)
    model.train_dnn(max(W_, 1), verbose=False)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    model.train_dnn(steps_tr, verbose=False)
    b.record()
    barrier()
    t_tr = max_over_ranks(a.elapsed_time(b) / 1e3)
    # --- lambda-phase step: residual kernel K3 (HBM-bound), L2 flushed every iteration.
    # (i) at the workload's N = 1M the kernel is a ~10 us launch (40 MB); (ii) the roofline
    # fraction is taken at 8M rows = config 5's per-GPU share (64 stacks x 1M / 8 GPUs), 320 MB.
    model.dnn.eval()
    with torch.no_grad():
        u = model.net_u(xd)[0].reshape(-1).contiguous()
    yv = model.u.reshape(-1).contiguous()
    sc, lam = model._scalers(sx), model._lambdas()
    sums = torch.empty(_abi.S_COUNT, device=dev, dtype=torch.float64)
    fam = _abi.FAM_V | _abi.FAM_DATA
    res = lambda: K.residuals(xd, u, yv, sc, lam, fam, sums=sums)
    t_res, _ = timed(res, K_, W_)
    rep = 8
    xb, ub, yb = xd.repeat(rep, 1).contiguous(), u.repeat(rep).contiguous(), yv.repeat(rep).contiguous()
    nb = xb.shape[0]
    resb = lambda: K.residuals(xb, ub, yb, sc, lam, fam, sums=sums)
    t_resb, _ = timed(resb, K_, W_)
    resb_acc = lambda: K.residuals(xb, ub, yb, sc, lam, fam, flags=_abi.RES_ACCURATE_MATH, sums=sums)
    t_resb_acc, _ = timed(resb_acc, K_, W_)
    fam_all = _abi.FAM_V | _abi.FAM_TS | _abi.FAM_H | _abi.FAM_O
    cols = torch.empty(_abi.C_COUNT, nb, device=dev, dtype=torch.float32)
    res_exp = lambda: K.residuals(xb, ub, None, sc, lam, fam_all, sums=sums, cols=cols, want_cols=True)
    t_exp, _ = timed(res_exp, K_, W_)
    del xb, ub, yb, cols
    # --- config 5 share (fleet export): one stack of n timesteps -> 22-column float64 rows (K4 sweep at T_PASSES,
    # eval forward, export-form K3, row writer K5) and the RF(t) risk series of 8 such stacks (K5 scans)
    from b200pinn.export import export_rows_device
    from b200pinn.rf import rf_device
    seg = [0] + [n * (i + 1) // 13 for i in range(13)]                   # normal segment + 12 labelled fault segments (04:75-80)
    yv32 = model.u.reshape(-1).contiguous()
    exp = lambda: export_rows_device(model, xd, yv32, seg, 12, T_PASSES, P_MC, sx, sy, seed=seed)
    t_export, _ = timed(exp, max(3, K_ // 2), 2)
    rows = exp()
    fleet = rows.unsqueeze(0).expand(8, -1, -1).contiguous()
    rfk = lambda: rf_device(fleet)
    t_rf, _ = timed(rfk, max(3, K_ // 2), 2)
    n_exp = max(3, K_ // 2)
    # GMM diagnosis (03:360-426) over the same 8 stacks: one EM iteration = one float64 pass over 8 x n rows x 4 features
    # (the residual-score columns pV, pT, pH, pO) with 20 components
    from b200pinn import gmm as G
    import numpy as np
    rng = np.random.default_rng(0)
    feats = fleet[:, :, 13:17].reshape(-1, 4).contiguous()                 # the four residual columns of the export rows
    mu0 = feats[:: max(1, feats.shape[0] // 20)][:20].clone()
    sd = feats.std(dim=0).clamp_min(1e-6).cpu().numpy()
    gw = np.ones(20) / 20
    gpc = np.stack([np.diag(1.0 / sd)] * 20)
    gmm_it = lambda: G.gmm_pass(feats, gw, mu0, gpc, want_stats=True)
    t_gmm, _ = timed(gmm_it, n_exp, 2, do_flush=False)
    n_gmm = feats.shape[0]
    del fleet, rows, feats
    # --- wide nets (the reference's own Layers = [8,256,256,256,1], 01:2139; config 4 is 6x256): MC sweep on the per-layer
    # tcgen05 GEMM path (mlp_wide_tc.cu): MC sweep and train step, N = 262144
    wide = {}
    n_w, T_w = 262144, 10
    for tag, lay_w, fl_pass, fl_train in (("3x256", [8, 256, 256, 256, 1], 344_704, 1_042_304),
                                          ("6x256", [8, 256, 256, 256, 256, 256, 256, 1], 737_920, 2_221_952)):
        torch.manual_seed(0)
        mw = b200pinn.PhysicsInformedNN(X[:n_w], Y[:n_w], lay_w, sx, sy, P_TRAIN, True)
        mw.dnn.eval()
        xw = mw.x.detach()
        t_w, _ = timed(lambda: b200pinn.mc_dropout_device(mw.dnn, xw, T_w, P_MC, seed=seed), 3, 1)
        mw.train_dnn(1, verbose=False)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        mw.train_dnn(2, verbose=False)
        b.record()
        barrier()
        t_wt = a.elapsed_time(b) / 2e3
        wide[tag] = {"mc_ms": 1e3 * t_w / 3, "mc_sample_passes_per_s": n_w * T_w * 3 / t_w,
                     "mc_tflops": n_w * T_w * 3 * fl_pass / t_w / 1e12, "train_ms": 1e3 * t_wt,
                     "train_tflops": n_w * fl_train / t_wt / 1e12}
        del mw, xw
    wide["what"] = ("N=262144 per GPU; MC sweep T=10 and train_dnn step on the per-layer tcgen05 3xTF32 GEMM path (operands as "
                    "pre-split tf32 planes, TMA bulk copies; dgrad = same kernel on transposed weight planes, weight gradients "
                    "= split-K GEMMs over sample-contiguous copies)")
    # --- configs[0] scale (N = 20 000, the reference's own CPU-runnable case): every step is latency-bound here.
    # train_dnn = K2a + K2b + (gradient reduce + Adam) per step; the scalar phases run as persistent launches
    # (pinn_scalar_phase: 1000 optimiser steps per launch, one grid barrier per step)
    c1 = None
    if world == 1 and not args.no_c1:
        n1 = 20_000
        torch.manual_seed(0)
        m1 = b200pinn.PhysicsInformedNN(X[:n1], Y[:n1], LAYERS, sx, sy, P_TRAIN, True)

        def wall(fn, k):
            fn(3)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn(k)
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / k

        c1 = {"n": n1,
              "train_dnn_us_per_step": 1e6 * wall(lambda k: m1.train_dnn(k, verbose=False), 500),
              "train_lambda_us_per_step": 1e6 * wall(lambda k: m1.train_lambda(k, True, verbose=False), 2001),
              "train_thermal_us_per_step": 1e6 * wall(lambda k: m1.train_thermal(k, verbose=False), 2001),
              "train_hydrogen_us_per_step": 1e6 * wall(lambda k: m1.train_hydrogen(k, verbose=False), 2001),
              "train_oxygen_us_per_step": 1e6 * wall(lambda k: m1.train_oxygen(k, verbose=False), 2001),
              "what": "configs[0] size, wall clock per optimiser step through the drop-in trainers (includes the 1-in-1000 "
                      "progress read-back); the reference's schedule 01:2143-2153 is 12 002 train_dnn + 34 005 scalar-phase "
                      "steps (profiles/c1_pipeline.py runs it end to end)"}
        c1["schedule_s_estimate"] = 1e-6 * (12002 * c1["train_dnn_us_per_step"] + 8002 * c1["train_lambda_us_per_step"]
                                            + 10001 * c1["train_thermal_us_per_step"] + 8001 * c1["train_hydrogen_us_per_step"]
                                            + 8001 * c1["train_oxygen_us_per_step"])
        del m1
    # --- e2e: public API, host tensors in pinned memory, results back on the host
    Xp = X.pin_memory()
    import contextlib
    import io

    def e2e():
        with contextlib.redirect_stdout(io.StringIO()):
            b200pinn.get_MC_samples(model, Xp, sx, mc_times=T_PASSES, dropout=P_MC)

    for _ in range(max(1, W_)):
        e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K_):
        e2e()
    barrier()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * n * T_PASSES * K_ / t_e2e

    if rank != 0:
        if world > 1:
            dist.barrier()           # keep every rank alive until rank 0 has printed its line
            dist.destroy_process_group()
        return
    pk = peaks()
    mc_tflops = n * T_PASSES * FLOP_PER_SAMPLE_PASS / (t_mc / K_) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K_, "warmup": W_,
        "ms_per_step": 1e3 * t_mc / K_, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: 3x64 PINN (layers [8,64,64,64,1]), N=1M samples per GPU; step = MC-dropout "
                               "sweep of T=50 passes (kernel K4)", "n_per_gpu": n, "T": T_PASSES, "dropout": P_MC,
                   "sharding": "samples" if world > 1 else "none",
                   "l2": "256 MB buffer written between timed steps (inputs 32 MB < 126 MB L2)"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 8 * 4, "d2h_bytes_per_step": n * 3 * 4},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": mc_tflops, "peak": pk["bf16"], "unit": "TFLOP/s",
                     "frac": mc_tflops / pk["bf16"], "traffic": 32.1e6 * n / 1e6,
                     "kernel": "mlp_tc_kernel<MC> (tcgen05.mma kind::tf32 with A in tensor memory, 3xTF32 split, "
                               "one MMA warp per 128-sample group)",
                     "peak_source": pk["src"],
                     "ncu": {"source": "profiles/r1_final4_mc.summary.txt (ncu --set full of this build's kernel, 10.58 ms capture)",
                             "sm__pipe_tensor_cycles_active_pct": 29.6, "smsp__issue_active_pct": 57.4,
                             "sm__inst_executed_pipe_alu_pct": 43.9, "sm__inst_executed_pipe_xu_pct": 40.1,
                             "dram_bytes_per_launch": 32.1e6},
                     "note": "achieved = algorithmic FLOPs (21 664 per sample*pass) / CUDA-event time of the launch. The "
                             "contractions run as 3 TF32 MMAs per product (fp32 parity), and TF32 dense peak is half the "
                             "bf16 figure used as `peak`, so the tensor pipe does 6x this fraction of its own peak "
                             "(ncu: sm__pipe_tensor_cycles_active, profiles/r1_final4_mc.summary.txt); the kernel is "
                             "bounded by the CUDA-core epilogue (2 MUFU + ~12 ALU/FMA ops per activation, 5.6 Philox "
                             "instructions per draw) and by the ~1000-clk latency of each 24-MMA batch, see DESIGN.md "
                             "section 4 (tensor pipe active: 29 % here, 43 % in the weight-gradient kernel K2b, 47-51 % in the "
                             "wide-net GEMMs). `traffic` is the ncu dram read+write of one launch at N=1M, scaled by n; "
                             f"vs the fp32 FFMA peak (74.5 TFLOP/s) the kernel stands at {mc_tflops / 74.5:.2f}x"},
        "train": {"steps_per_s": steps_tr / t_tr, "ms_per_step": 1e3 * t_tr / steps_tr,
                  "global_batch": world * n, "tflops": world * n * FLOP_PER_TRAIN_SAMPLE / (t_tr / steps_tr) / 1e12,
                  # K2a writes and K2b reads a 496-row x 512 B table per 128-sample tile (3x64 net): 2 x 1.98 KB per sample
                  "hbm": {"bytes_per_sample": 2 * 496 * 4 + 36, "unit": "GB/s",
                          "achieved": n * (2 * 496 * 4 + 36) / (t_tr / steps_tr) / 1e9,
                          "frac_of_peak": n * (2 * 496 * 4 + 36) / (t_tr / steps_tr) / 1e9 / pk["hbm"]},
                  "what": "train_dnn step: K2a (tcgen05 fwd+loss+dgrad, writes a transposed 2 KB/sample row table) + K2b "
                          "(tcgen05 3xTF32 weight gradients, HBM-bound on that table: see profiles/) + "
                          + ("partial reduce + gradient sum over NVLink peer memory fused into the Adam/StepLR launch"
                             if world > 1 else "partial reduce with Adam/StepLR applied in the same launch")},
        "roofline_residual": {"bound": "hbm", "achieved": nb * RES_BYTES_PER_SAMPLE / (t_resb / K_) / 1e9, "peak": pk["hbm"],
                              "unit": "GB/s", "frac": nb * RES_BYTES_PER_SAMPLE / (t_resb / K_) / 1e9 / pk["hbm"],
                              "traffic": None, "kernel": "residual_kernel<V|DATA, fast math>", "rows": nb,
                              "ms": 1e3 * t_resb / K_, "accurate_math_ms": 1e3 * t_resb_acc / K_,
                              "ms_at_1M_rows": 1e3 * t_res / K_, "lambda_steps_per_s_at_1M": K_ / t_res,
                              "export_form": {"bytes_per_sample": 36 + 4 * 21, "ms": 1e3 * t_exp / K_,
                                              "gbs": nb * (36 + 4 * 21) / (t_exp / K_) / 1e9}},
    }
    line["fleet"] = {"what": "config 5 per-GPU share: export of one 1M-timestep stack (MC sweep T=50 + eval forward + export-form "
                             "residuals + float64 22-column row writer with segment smoothing) and RF(t) for 8 stacks "
                             "(mu/sigma, dead-zone norms, C(t) scan, logistic, EMA, first alarm)",
                     "export_rows_per_s": world * n * n_exp / t_export, "export_ms_per_stack": 1e3 * t_export / n_exp,
                     "rf_rows_per_s": world * 8 * n * n_exp / t_rf, "rf_ms_per_8_stacks": 1e3 * t_rf / n_exp,
                     "rf_hbm_gbs": 8 * n * (22 * 8 + 2 * 8) * n_exp / t_rf / 1e9,
                     "gmm_em_iteration_ms": 1e3 * t_gmm / n_exp, "gmm_rows_per_s": world * n_gmm * n_exp / t_gmm,
                     "gmm_what": "one EM iteration (E-step + M-step statistics, float64) of a 20-component full-covariance "
                                 "GaussianMixture over the 4 residual-score columns of 8 stacks (03:360-426)"}
    line["wide"] = wide
    if c1 is not None:
        line["c1"] = c1
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        mc_rate, tr_rate, t1, t2 = cpu_port_rates(1_000_000, 200_000, threads)
        line["cpu_baseline"] = {"value": mc_rate, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"get_MC_samples port, N=1000000 x T'=1 ({t1:.1f} s); train_dnn port 1 step at "
                                          f"N=200000 ({t2:.1f} s)",
                                "train_steps_per_s_at_1M": tr_rate}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=N_PER_GPU, help="samples per GPU (default: configs[1], 1M)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-c1", action="store_true", help="skip the configs[0]-size step timings (thousands of launches: for ncu runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
