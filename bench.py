#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the PINN hot path.

    python bench.py --gpus N --steps K --warmup W            (our CUDA path, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...   (the reference's own classes on host cores)

Headline (`value`): MC-dropout sample*passes/s of the 3x64 stack-voltage PINN on N = 1M synthetic
normal-operation samples per GPU (the batch of BASELINE.json configs[1]), one "step" = one sweep of
T = 50 stochastic passes (configs[0]'s T) = one launch of kernel K4.  This is not itself one of the
BASELINE configs; the configs are carried by blocks of the same JSON line:

    c1     configs[0]  N = 20 000, every trainer's step time + the reference's whole schedule
    train  configs[1]  N = 1M/GPU full-batch train_dnn step (K2 + reduce + Adam), steps/s, weak efficiency
    c3     configs[2]  MC sweep T = 1000 x N = 1M TOTAL, sample-sharded over --gpus (strong scaling), gathered
    c4     configs[3]  6x256 data-parallel train step at 512k samples per GPU (batch 4M on 8 GPUs)
    fleet  configs[4]  per-GPU share of the 64-stack export + RF(t)

`value` is timed with inputs resident in HBM; `e2e` is the same sweep through the public
`get_MC_samples` with HOST tensors (H2D of X and D2H of the three result vectors inside the timed
region).  `cpu_baseline` / `--impl reference` run the UNMODIFIED reference (oracle/_ref, staged by
__graft_entry__.build()) on the host cores; `gpu_eager_baseline` runs the same unmodified classes on
eager PyTorch-CUDA on this B200 (the reference's own GPU path, 01:21-24).
"""
import argparse
import contextlib
import io
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
NCU_FILE = os.path.join(ROOT, "profiles", "ncu_current.json")


def workload_config(n, world):
    """The `config` block: identical for the product arm and the reference arm (same workload, same sizes)."""
    return {"workload": "MC-dropout sweep, T=50 passes over N=1M samples per GPU, 3x64 PINN (layers [8,64,64,64,1]): the batch of "
                        "BASELINE configs[1] swept with configs[0]'s T -- not itself a BASELINE config (configs[2] = block c3, "
                        "configs[1] training = block train, configs[3] = c4, configs[4] = fleet, configs[0] = c1)",
            "layers": LAYERS, "n_per_gpu": n, "T": T_PASSES, "dropout": P_MC,
            "sharding": "samples" if world > 1 else "none",
            "l2": "256 MB buffer written between timed steps (inputs 32 MB < 126 MB L2)"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), bf16=float(p["bf16_tflops"]), src="measured (MEASURED_PEAKS.json, burst)")
    return dict(hbm=6650.0, bf16=1590.0, src="fallback (B200_PROFILING.md)")


def ncu_record(kernel):
    """Counters of `kernel` from profiles/ncu_current.json (written by profiles/ncu_to_json.py from an `ncu --set full`
    capture of THIS build); None when no capture is on file -- nothing here is a literal."""
    try:
        with open(NCU_FILE) as f:
            return json.load(f).get(kernel)
    except (OSError, ValueError):
        return None


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


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


# ------------------------------------------------------------------------- reference classes
def reference_classes(device):
    """(module-like namespace, kind): the unmodified reference (oracle/_ref) when staged, else the op-for-op port."""
    from oracle import ref_loader

    if ref_loader.available():
        with quiet():
            return ref_loader.load("01", device=device), "reference"
    if device != "cpu":
        return None, "unavailable"
    from oracle import torch_port as P
    import types

    ns = types.SimpleNamespace(
        PhysicsInformedNN=lambda X, u, layers, xs, us, p, logvar: P.PortPINN(X, u, layers, xs, us, p),
        get_MC_samples=P.get_MC_samples_port)
    return ns, "port"


def cpu_reference_rates(X, Y, sx, sy, n_mc, n_train, threads):
    """The reference's own hot loops on the host cores: one get_MC_samples with T'=1 at n_mc rows and train_dnn steps
    at n_train rows."""
    import torch

    ref, kind = reference_classes("cpu")
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    with quiet():
        net = ref.PhysicsInformedNN(X[:n_mc], Y[:n_mc], LAYERS, sx, sy, P_TRAIN, True)
        ref.get_MC_samples(net, X[:2000], sx, 1, P_MC)            # warm-up
        t0 = time.perf_counter()
        ref.get_MC_samples(net, X[:n_mc], sx, 1, P_MC)
        t_mc = time.perf_counter() - t0
        tr = ref.PhysicsInformedNN(X[:n_train], Y[:n_train], LAYERS, sx, sy, P_TRAIN, True)
        if kind == "reference":
            tr.train_dnn(1)
            t0 = time.perf_counter()
            tr.train_dnn(2)
            t_tr = (time.perf_counter() - t0) / 2
        else:
            step = tr.make_dnn_trainer()
            step()
            t0 = time.perf_counter()
            step()
            t_tr = time.perf_counter() - t0
    return kind, n_mc / t_mc, 1.0 / t_tr, t_mc, t_tr


def gpu_eager_rates(X, Y, sx, sy, n):
    """The unmodified reference on eager PyTorch-CUDA on this GPU (its own device selection, 01:21-24): the MC sweep
    with T'=4 and the train_dnn step at the same N.  Host tensors in, numpy out -- exactly how 01 calls it."""
    import torch

    ref, kind = reference_classes("cuda")
    if ref is None:
        return {"unavailable": "oracle/_ref not staged (run __graft_entry__.build() where /root/reference is mounted)"}
    torch.manual_seed(0)
    Tq = 4
    with quiet():
        model = ref.PhysicsInformedNN(X[:n], Y[:n], LAYERS, sx, sy, P_TRAIN, True)
        ref.get_MC_samples(model, X[:20000], sx, 1, P_MC)         # warm-up (cuBLAS handles, allocator)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ref.get_MC_samples(model, X[:n], sx, Tq, P_MC)
        torch.cuda.synchronize()
        t_mc = time.perf_counter() - t0
        model.train_dnn(3)
        torch.cuda.synchronize()
        k = 10
        t0 = time.perf_counter()
        model.train_dnn(k)
        torch.cuda.synchronize()
        t_tr = (time.perf_counter() - t0) / k
        t0 = time.perf_counter()
        model.train_lambda(k, True)
        torch.cuda.synchronize()
        t_lam = (time.perf_counter() - t0) / k
    del model
    torch.cuda.empty_cache()
    return {"kind": kind, "what": "the unmodified reference classes (oracle/_ref) on eager PyTorch-CUDA, same B200, same N, host "
                                  "tensors in / numpy out as 01:2141-2158 calls them",
            "n": n, "mc_sample_passes_per_s": n * Tq / t_mc, "mc_sample": f"get_MC_samples T'={Tq} ({t_mc:.2f} s)",
            "train_dnn_steps_per_s": 1.0 / t_tr, "train_dnn_ms_per_step": 1e3 * t_tr,
            "train_lambda_ms_per_step": 1e3 * t_lam}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    n = args.n
    X, Y, sx, sy = build_problem(n, 2)
    ref, kind = reference_classes("cpu")
    torch.manual_seed(0)
    with quiet():
        net = ref.PhysicsInformedNN(X, Y, LAYERS, sx, sy, P_TRAIN, True)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        with quiet():
            ref.get_MC_samples(net, X, sx, 1, P_MC)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = n * 1 * len(times) / total
    sample = (f"one step = get_MC_samples(mc_times=1) over the full N={n} batch: T'=1 of the workload's T={T_PASSES} passes (cost is "
              "linear in T; 1 eval pass + 1 dropout pass, each with the discarded 2nd forward of predict(), 01:1407)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(n, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import b200pinn
    from b200pinn import _abi, kernels as K
    from b200pinn.dist import gather_rows, shard_range

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

    def timed_block(fn, steps):
        """`steps` back-to-back steps inside ONE pair of events (for loops whose launches are enqueued by one call)."""
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(steps)
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b) / 1e3)

    def graph_timed(calls, reps):
        """Per-launch time of short kernels: `reps` launches (cycling through `calls`) captured in ONE CUDA graph and replayed
        inside one event pair -- the GPU never waits for the host between launches (a ~50 us kernel timed launch by launch
        from Python measures the wrapper, not the kernel).  `calls` must rotate over inputs that together exceed the L2."""
        for c in calls:
            c()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(reps):
                calls[i % len(calls)]()
        g.replay()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b) / 1e3 / reps)

    sampler = ClockSampler(local) if rank == 0 else None
    # --- headline: MC-dropout sweep, inputs resident in HBM
    seed = 1234
    mc = lambda: b200pinn.mc_dropout_device(model.dnn, xd, T_PASSES, P_MC, seed=seed, sample_offset=rank * n)
    t_mc, launches = timed(mc, K_, W_)
    clocks = sampler.stop() if sampler else None
    value = world * n * T_PASSES * K_ / t_mc

    # --- configs[1]: full-batch train_dnn steps (K2 + reduce + [gradient sum over NVLink fused into Adam] + Adam), back to back.
    # Single-GPU step time is measured in the same job (data_parallel = False) so the weak-scaling efficiency of the ONE
    # step that communicates is stated here, not only for the collective-free sweep.
    steps_tr = max(K_, 40)       # one call enqueues them all; long enough that the call's fixed cost (progress-line read-back) vanishes
    model.data_parallel = False
    model.train_dnn(max(W_, 1), verbose=False)
    t_tr_local = timed_block(lambda k: model.train_dnn(k, verbose=False), steps_tr)
    model.data_parallel = True
    if world > 1:
        model.train_dnn(max(W_, 1), verbose=False)
        t_tr = timed_block(lambda k: model.train_dnn(k, verbose=False), steps_tr)
    else:
        t_tr = t_tr_local

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
    # 1M rows = 40 MB per launch: six distinct copies (240 MB > 126 MB L2) visited in turn; 8M rows = 320 MB per launch > L2
    small = [(xd.clone(), u.clone(), yv.clone()) for _ in range(6)]
    t_res = graph_timed([(lambda q=q: K.residuals(q[0], q[1], q[2], sc, lam, fam, sums=sums)) for q in small], 30)
    del small
    rep = 8
    xb, ub, yb = xd.repeat(rep, 1).contiguous(), u.repeat(rep).contiguous(), yv.repeat(rep).contiguous()
    nb = xb.shape[0]
    t_resb = graph_timed([lambda: K.residuals(xb, ub, yb, sc, lam, fam, sums=sums)], 20)
    t_resb_acc = graph_timed([lambda: K.residuals(xb, ub, yb, sc, lam, fam, flags=_abi.RES_ACCURATE_MATH, sums=sums)], 10)
    fam_all = _abi.FAM_V | _abi.FAM_TS | _abi.FAM_H | _abi.FAM_O
    cols = torch.empty(_abi.C_COUNT, nb, device=dev, dtype=torch.float32)
    t_exp = graph_timed([lambda: K.residuals(xb, ub, None, sc, lam, fam_all, sums=sums, cols=cols, want_cols=True)], 10)
    del xb, ub, yb, cols

    # --- configs[2]: MC sweep T = 1000 over N = 1M samples IN TOTAL, sample-sharded over the ranks (strong scaling), the three
    # result vectors gathered on every rank inside the timed region.  Every rank also times the unsharded sweep so that the
    # strong-scaling efficiency t(1 GPU) / (world * t(world GPUs)) comes from one job.
    T3, n3 = 1000, 1_000_000
    X3 = X if (rank == 0 and n == n3) else build_problem(n3, 2)[0]
    lo3, hi3 = shard_range(n3, rank, world)
    x3_full = X3.to(dev)
    x3 = x3_full[lo3:hi3].contiguous()

    def sweep3(xs, off, gather):
        o = b200pinn.mc_dropout_device(model.dnn, xs, T3, P_MC, seed=seed, sample_offset=off)
        r = torch.stack([o["pred_mean"], o["a_u"], o["e_u"]], dim=1)
        return gather_rows(r, n3) if gather else r

    n_c3 = 3
    t_c3, _ = timed(lambda: sweep3(x3, lo3, world > 1), n_c3, 1)
    t_c3_one = t_c3
    if world > 1:
        t_c3_one, _ = timed(lambda: sweep3(x3_full, 0, False), n_c3, 1)
    tiles3 = -(-(hi3 - lo3) // 128)
    sm = _abi.lib().pinn_device_sm_count()
    c3 = {"what": "configs[2]: MC-dropout sweep T=1000 x N=1M samples in total, sample-sharded over the GPUs (strong scaling); no "
                  "data-path collective, the three result vectors are all-gathered inside the timed region; Philox counters are "
                  "keyed on global (sample, pass), so every sharding gives the same numbers",
          "T": T3, "n_total": n3, "rows_per_gpu": hi3 - lo3, "ms_per_sweep": 1e3 * t_c3 / n_c3,
          "sample_passes_per_s": n3 * T3 * n_c3 / t_c3, "tflops": n3 * T3 * FLOP_PER_SAMPLE_PASS * n_c3 / t_c3 / 1e12,
          "one_gpu_ms_per_sweep": 1e3 * t_c3_one / n_c3, "strong_scaling_efficiency": t_c3_one / (world * t_c3),
          "tiles_per_gpu": tiles3, "tile_slots_per_gpu": 3 * sm, "pass_chunks": 4,
          "wave_quantisation": 4 * tiles3 / (3 * sm * -(-4 * tiles3 // (3 * sm))),
          "wave_note": "a GPU runs 3 x SMs work items at a time (three tile groups per CTA, mlp_tc3.cu); an item is a 128-row tile x one run of 250 consecutive passes (T = 1000 "
                       "is cut into 4 runs per tile -- a function of T alone, so the numbers do not depend on the sharding -- whose "
                       "Welford triples a merge launch folds in order); wave_quantisation = items / (slots x waves) is the ceiling "
                       "of the static item->SM map at this shard size"}
    del x3, x3_full, X3

    # --- config 5 share (fleet export): one stack of n timesteps -> 22-column float64 rows (K4 sweep at T_PASSES,
    # eval forward, export-form K3, row writer K5) and the RF(t) risk series of 8 such stacks (K5 scans)
    from b200pinn.export import export_rows_device
    from b200pinn.rf import rf_device
    seg = [0] + [n * (i + 1) // 13 for i in range(13)]                   # normal segment + 12 labelled fault segments (04:75-80)
    yv32 = model.u.reshape(-1).contiguous()
    exp = lambda: export_rows_device(model, xd, yv32, seg, 12, T_PASSES, P_MC, sx, sy, seed=seed, want_rf_cols=True)
    t_export, _ = timed(exp, max(3, K_ // 2), 2)
    rows, rf_cols = exp()
    fleet = rows.unsqueeze(0).expand(8, -1, -1).contiguous()
    rfk = lambda: rf_device(fleet)
    t_rf, _ = timed(rfk, max(3, K_ // 2), 2)
    fleet_c = rf_cols.unsqueeze(0).expand(8, -1, -1).contiguous()         # the dense [n, 6] copy of columns 12..17 the row writer emits
    t_rfc, _ = timed(lambda: rf_device(fleet_c), max(3, K_ // 2), 2)
    del fleet_c, rf_cols
    n_exp = max(3, K_ // 2)
    # GMM diagnosis (03:360-426) over the same 8 stacks: one EM iteration = one float64 pass over 8 x n rows x 4 features
    # (the residual-score columns pV, pT, pH, pO) with 20 components
    from b200pinn import gmm as G
    import numpy as np
    feats = fleet[:, :, 13:17].reshape(-1, 4).contiguous()                 # the four residual columns of the export rows
    mu0 = feats[:: max(1, feats.shape[0] // 20)][:20].clone()
    sd = feats.std(dim=0).clamp_min(1e-6).cpu().numpy()
    gw = np.ones(20) / 20
    gpc = np.stack([np.diag(1.0 / sd)] * 20)
    gmm_it = lambda: G.gmm_pass(feats, gw, mu0, gpc, want_stats=True)
    t_gmm, _ = timed(gmm_it, n_exp, 2, do_flush=False)
    n_gmm = feats.shape[0]
    del fleet, rows, feats

    # --- wide nets: the reference's own Layers = [8,256,256,256,1] (01:2139) at N = 262144 (MC sweep T=10 + train step), and
    # configs[3] = 6x256 data-parallel training at its stated size: batch 4M over 8 GPUs = 524288 samples per GPU
    wide = {}
    T_w = 10
    for tag, lay_w, n_w, fl_pass, fl_train in (("3x256", [8, 256, 256, 256, 1], 262144, 344_704, 1_042_304),
                                               ("6x256", [8, 256, 256, 256, 256, 256, 256, 1], 524288, 737_920, 2_221_952)):
        torch.manual_seed(0)
        mw = b200pinn.PhysicsInformedNN(X[:n_w], Y[:n_w], lay_w, sx, sy, P_TRAIN, True)
        mw.dnn.eval()
        xw = mw.x.detach()
        t_w, _ = timed(lambda: b200pinn.mc_dropout_device(mw.dnn, xw, T_w, P_MC, seed=seed), 3, 1)
        mw.train_dnn(1, verbose=False)
        t_wt = timed_block(lambda k: mw.train_dnn(k, verbose=False), 3) / 3
        wide[tag] = {"n_per_gpu": n_w, "mc_ms": 1e3 * t_w / 3, "mc_sample_passes_per_s": world * n_w * T_w * 3 / t_w,
                     "mc_tflops_per_gpu": n_w * T_w * 3 * fl_pass / t_w / 1e12, "train_ms": 1e3 * t_wt,
                     "train_steps_per_s": 1.0 / t_wt, "global_batch": world * n_w,
                     "train_tflops_per_gpu": n_w * fl_train / t_wt / 1e12}
        del mw, xw
        torch.cuda.empty_cache()
    # the reference's own production sweep (01:2139, 01:2156-2158): Layers 3x256, mc_times = 2000, at configs[0]'s N = 20 000
    torch.manual_seed(0)
    n_r = 20000
    mr = b200pinn.PhysicsInformedNN(X[:n_r], Y[:n_r], [8, 256, 256, 256, 1], sx, sy, P_TRAIN, True)
    mr.dnn.eval()
    xr_ = mr.x.detach()
    t_r, _ = timed(lambda: b200pinn.mc_dropout_device(mr.dnn, xr_, 2000, P_MC, seed=seed), 2, 1)
    wide["3x256"]["mc_T2000_n20000_ms"] = 1e3 * t_r / 2
    wide["3x256"]["mc_T2000_n20000_sample_passes_per_s"] = n_r * 2000 * 2 / t_r
    del mr, xr_
    c4 = dict(wide.pop("6x256"))
    c4["what"] = ("configs[3]: 6x256 PINN, data-parallel train_dnn step at 524288 samples per GPU (= batch 4M on 8 GPUs; global batch "
                  "here = n_gpus x 524288), one gradient-bucket exchange (1.49 MB) per step fused into the Adam launch over NVLink "
                  "peer memory; the MC numbers are the same net's sweep at T=10")
    wide["what"] = ("the reference's own Layers (01:2139): MC sweep T=10 on the resident-activation kernel (csrc/mlp_wide_res.cu: one "
                    "persistent CTA per SM, activations as fp16 hi/lo pairs in tensor memory across layers and passes, weights "
                    "streamed through a TMA ring); train_dnn step on the per-layer tcgen05 3xTF32 GEMM path")
    # the sweep kernel against the tensor roofline: algorithmic FLOPs, and the FLOPs the tensor pipe actually executes (three
    # fp16 products per contraction, heads padded to N = 144)
    rec_w = ncu_record("wide_res_ts_kernel")
    pkw = peaks()
    w3 = wide["3x256"]
    exec_per_pass = 6.0 * (2 * 256 * 256 + 144 * 256 + 64 * 128)
    wide["roofline"] = {"bound": "tensor", "kernel": "wide_res_ts_kernel (3x256, N = 262144, T = 10 + the eval pass)",
                        "achieved": w3["mc_tflops_per_gpu"], "peak": pkw["bf16"], "unit": "TFLOP/s",
                        "frac": w3["mc_tflops_per_gpu"] / pkw["bf16"],
                        "executed_tflops": w3["n_per_gpu"] * (T_w + 1) * exec_per_pass / (w3["mc_ms"] * 1e-3) / 1e12,
                        "executed_frac": w3["n_per_gpu"] * (T_w + 1) * exec_per_pass / (w3["mc_ms"] * 1e-3) / 1e12 / pkw["bf16"],
                        "algorithmic_bytes": 44 * w3["n_per_gpu"],
                        "traffic": (rec_w["dram_bytes"] * w3["n_per_gpu"] / rec_w["n"]) if rec_w else None, "ncu": rec_w,
                        "note": "achieved = algorithmic FLOPs (344 704 per sample*pass, T passes) / CUDA-event time of the sweep call; "
                                "executed = what the tensor pipe runs for fp32 parity (a_l*w_h + a_h*w_l + a_h*w_h as kind::f16 "
                                "products, T + 1 passes); algorithmic bytes = x in (32 B) + three result vectors out"}

    # --- configs[0] scale (N = 20 000, the reference's own CPU-runnable case): every step is latency-bound here.
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

    def e2e():
        with quiet():
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
    rec_mc, rec_res = ncu_record("mlp_tc3_kernel<MC>"), ncu_record("residual_v_fast_kernel")
    rec_tr = ncu_record("mlp_tc_fused_kernel")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K_, "warmup": W_,
        "ms_per_step": 1e3 * t_mc / K_, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(n, world),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 8 * 4, "d2h_bytes_per_step": n * 3 * 4},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": mc_tflops, "peak": pk["bf16"], "unit": "TFLOP/s",
                     "frac": mc_tflops / pk["bf16"],
                     "traffic": (rec_mc["dram_bytes"] * n / rec_mc["n"]) if rec_mc else None,
                     "kernel": "mlp_tc3_kernel<MC> (three 128-sample groups per CTA, tcgen05.mma kind::f16 on fp16 hi/lo pairs with A in "
                               "tensor memory, resident K-permuted weight images, one MMA warp per group)",
                     "peak_source": pk["src"], "ncu": rec_mc,
                     "note": "achieved = algorithmic FLOPs (21 664 per sample*pass) / CUDA-event time of the launch. The "
                             "contractions run as 3 fp16 MMAs per product (a_l*w_h + a_h*w_l + a_h*w_h: fp32 parity), so the "
                             "tensor pipe does 3x this fraction of its own peak; the kernel is bounded by the CUDA-core "
                             "epilogue (tanh, Philox, fp16 split: issue slots), see DESIGN.md "
                             "section 4. `traffic` / `ncu` are read from profiles/ncu_current.json (ncu --set full of this "
                             f"kernel), scaled by n; vs the fp32 FFMA peak (74.5 TFLOP/s) the kernel stands at {mc_tflops / 74.5:.2f}x"},
        "train": {"steps_per_s": steps_tr / t_tr, "ms_per_step": 1e3 * t_tr / steps_tr,
                  "global_batch": world * n, "tflops": world * n * FLOP_PER_TRAIN_SAMPLE / (t_tr / steps_tr) / 1e12,
                  "single_gpu_ms_per_step": 1e3 * t_tr_local / steps_tr, "weak_scaling_efficiency": t_tr_local / t_tr,
                  "frac_of_bf16_peak_per_gpu": n * FLOP_PER_TRAIN_SAMPLE / (t_tr / steps_tr) / 1e12 / pk["bf16"],
                  "what": "configs[1]: full-batch train_dnn step at N=1M per GPU (01:948-955): ONE fused K2 launch (forward + loss + dgrad "
                          "+ all weight gradients on tcgen05, on chip) + the fixed-order gradient reduce, "
                          + ("with the gradient sum over the ranks (NVLink peer memory) and Adam/StepLR inside that reduce launch; "
                             if world > 1 else "Adam/StepLR in the reduce launch; ")
                          + "single_gpu_ms_per_step is the same model stepping without the exchange, timed in this job"},
        "roofline_train": {"bound": "tensor", "kernel": "mlp_tc_fused_kernel<3> (one launch per step: forward, dgrad and every weight gradient of a "
                                                        "tile on chip; accumulators resident in tensor memory, TMA-streamed weight images)",
                           "achieved": n * FLOP_PER_TRAIN_SAMPLE / (t_tr_local / steps_tr) / 1e12, "peak": pk["bf16"], "unit": "TFLOP/s",
                           "frac": n * FLOP_PER_TRAIN_SAMPLE / (t_tr_local / steps_tr) / 1e12 / pk["bf16"],
                           "algorithmic_bytes_per_step": 40 * n,
                           "traffic": (rec_tr["dram_bytes"] * n / rec_tr["n"]) if rec_tr else None, "ncu": rec_tr,
                           "note": "achieved = algorithmic FLOPs (67 040 per sample) / step time of the single-GPU step (fused K2 + reduce/Adam); "
                                   "3xTF32 / hi-lo stacked products: the tensor pipe does ~6x this fraction of the TF32 peak.  `traffic` = "
                                   "dram__bytes of the fused launch from profiles/ncu_current.json (round 1's two-kernel form moved 4.0 GB per "
                                   "step at N = 1M)"},
        "roofline_residual": {"bound": "hbm", "achieved": nb * RES_BYTES_PER_SAMPLE / t_resb / 1e9, "peak": pk["hbm"],
                              "unit": "GB/s", "frac": nb * RES_BYTES_PER_SAMPLE / t_resb / 1e9 / pk["hbm"],
                              "timing": "per launch, from a CUDA-graph replay of 20-30 back-to-back launches (no host latency between "
                                        "launches); 8M rows = 320 MB per launch > 126 MB L2, 1M rows rotate over six 40 MB input copies",
                              "traffic": (rec_res["dram_bytes"] * nb / rec_res["n"]) if rec_res else None, "ncu": rec_res,
                              "kernel": "residual_v_fast_kernel (V|DATA, MUFU math)", "rows": nb,
                              "ms": 1e3 * t_resb, "accurate_math_ms": 1e3 * t_resb_acc,
                              "ms_at_1M_rows": 1e3 * t_res, "gbs_at_1M_rows": n * RES_BYTES_PER_SAMPLE / t_res / 1e9,
                              "frac_at_1M_rows": n * RES_BYTES_PER_SAMPLE / t_res / 1e9 / pk["hbm"],
                              "lambda_steps_per_s_at_1M": 1.0 / t_res,
                              "export_form": {"bytes_per_sample": 36 + 4 * 21, "ms": 1e3 * t_exp,
                                              "gbs": nb * (36 + 4 * 21) / t_exp / 1e9}},
    }
    line["c3"] = c3
    line["c4"] = c4
    line["fleet"] = {"what": "config 5 per-GPU share: export of one 1M-timestep stack (MC sweep T=50 + eval forward + export-form "
                             "residuals + float64 22-column row writer with segment smoothing) and RF(t) for 8 stacks "
                             "(mu/sigma, dead-zone norms, C(t) scan, logistic, EMA, first alarm)",
                     "export_rows_per_s": world * n * n_exp / t_export, "export_ms_per_stack": 1e3 * t_export / n_exp,
                     "rf_rows_per_s": world * 8 * n * n_exp / t_rf, "rf_ms_per_8_stacks": 1e3 * t_rf / n_exp,
                     "rf_hbm_gbs": 8 * n * (22 * 8 + 2 * 8) * n_exp / t_rf / 1e9,
                     "rf_compact": {"ms_per_8_stacks": 1e3 * t_rfc / n_exp, "rows_per_s": world * 8 * n * n_exp / t_rfc,
                                    "equivalent_row_gbs": 8 * n * (22 * 8 + 2 * 8) * n_exp / t_rfc / 1e9,
                                    "actual_gbs": 8 * n * (48 + 48 + 5 * 8) * n_exp / t_rfc / 1e9,
                                    "what": "same series from the dense [n,6] copy of columns 12..17 that the row writer emits on request: "
                                            "two passes over 48-byte rows + S, RF_inst, RF_smooth (136 B/row actually moved); "
                                            "equivalent_row_gbs uses rf_hbm_gbs's accounting (a 22-column row + two outputs)"},
                     "gmm_em_iteration_ms": 1e3 * t_gmm / n_exp, "gmm_rows_per_s": world * n_gmm * n_exp / t_gmm,
                     "gmm_what": "one EM iteration (E-step + M-step statistics, float64) of a 20-component full-covariance "
                                 "GaussianMixture over the 4 residual-score columns of 8 stacks (03:360-426)"}
    line["wide"] = wide
    if c1 is not None:
        line["c1"] = c1
    if world == 1 and not args.no_eager:
        line["gpu_eager_baseline"] = gpu_eager_rates(X, Y, sx, sy, n)
        ge = line["gpu_eager_baseline"]
        if "mc_sample_passes_per_s" in ge:
            ge["ours_over_eager_mc_e2e"] = e2e_value / ge["mc_sample_passes_per_s"]
            ge["ours_over_eager_train"] = (steps_tr / t_tr) / ge["train_dnn_steps_per_s"]
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        kind, mc_rate, tr_rate, t1, t2 = cpu_reference_rates(X, Y, sx, sy, n, min(n, 200_000), threads)
        line["cpu_baseline"] = {"value": mc_rate, "unit": UNIT, "cores": threads, "kind": kind,
                                "sample": f"get_MC_samples, N={n} x T'=1 ({t1:.1f} s); train_dnn steps at "
                                          f"N={min(n, 200_000)} ({t2:.2f} s per step)",
                                "train_steps_per_s_at_200k": tr_rate}
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
    ap.add_argument("--n", type=int, default=N_PER_GPU, help="samples per GPU (default: 1M, the batch of configs[1])")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-eager", action="store_true", help="skip the eager-PyTorch-CUDA reference leg")
    ap.add_argument("--no-c1", action="store_true", help="skip the configs[0]-size step timings (thousands of launches: for ncu runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
