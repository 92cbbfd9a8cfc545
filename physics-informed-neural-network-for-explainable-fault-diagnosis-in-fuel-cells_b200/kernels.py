"""Tensor-level wrappers over the C ABI: torch is used only for device memory and
streams; every computation below is a call into ``libb200pinn.so``."""
from __future__ import annotations

import ctypes as C
from contextlib import contextmanager
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _abi
from ._abi import PinnDropout, PinnNet, PinnScalers, check, ptr

LAUNCHES = 0  # kernels launched through this module (bench.py reports it as gpu_launches)


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"b200pinn: `{name}` must live on a CUDA device -- this build has no CPU path")
    if t.dtype != torch.float32:
        raise RuntimeError(f"b200pinn: `{name}` must be float32, got {t.dtype}")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_WS: dict = {}


def _workspace(kind: str, nbytes: int, device, zero: bool = False) -> Optional[torch.Tensor]:
    """Cached per-(kind, device) scratch, grown on demand; kernels never allocate."""
    if nbytes == 0:
        return None
    key = (kind, device)
    t = _WS.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.zeros(nbytes, dtype=torch.uint8, device=device) if zero else \
            torch.empty(nbytes, dtype=torch.uint8, device=device)
        _WS[key] = t
    return t


@dataclass
class Net:
    """Device view of a DNN (01:389-438): ctypes descriptor + the tensors it points at."""
    desc: PinnNet
    width: int
    n_hidden: int
    tensors: list
    base_flags: int = 0

    def ref(self):
        """``byref`` of the descriptor with this call's option bits (model flags | active ``path_flags``)."""
        self.desc.flags = self.base_flags | _EXTRA_NET_FLAGS
        return C.byref(self.desc)

    @property
    def flags(self) -> int:
        return self.base_flags | _EXTRA_NET_FLAGS

    @property
    def mask_width(self) -> int:
        return self.n_hidden * self.width + self.width // 2


# Per-call option bits (``pinn_net_t.flags``) OR-ed into every descriptor built while a ``path_flags`` block is
# active: ablation runs and tests route a shape through another kernel family without any state in the library.
_EXTRA_NET_FLAGS = 0
_PHASE_FLAGS = 0
_DEFAULT_CM = None


@contextmanager
def path_flags(no_tc_fwd=False, no_tc_bwd=False, no_wide_tc=False, dependent_launch=None, no_phase_cluster=False,
               no_fused_bwd=False, no_wide_resident=False, no_tma_input=False, no_tc3=False):
    """``with path_flags(no_tc_fwd=True): ...`` -- inside the block the 64-wide forward / MC sweep run on the FFMA
    kernels (likewise ``no_tc_bwd``, ``no_wide_tc``); ``dependent_launch`` = 0 never / 2 always chain a step's launches
    with programmatic dependent launch (default 1: small batches only); ``no_phase_cluster`` keeps ``scalar_phase`` on
    the cooperative-grid form; ``no_fused_bwd`` runs the 64-wide backward as the two-kernel form (K2a + row table +
    K2b) instead of the one-kernel form; ``no_wide_resident`` runs the 256-wide forward / MC sweep as one GEMM launch
    per layer instead of the resident-activation kernel; ``no_tma_input`` makes the tensor-core forward / MC kernels load
    their input tiles with plain global loads instead of TMA tensor-map copies; ``no_tc3`` keeps the 64-wide forward / MC
    sweep on the two-group 3xTF32 kernel instead of the three-group fp16-pair kernel."""
    global _EXTRA_NET_FLAGS, _PHASE_FLAGS
    prev = (_EXTRA_NET_FLAGS, _PHASE_FLAGS)
    f = (_abi.NET_NO_TC_FWD if no_tc_fwd else 0) | (_abi.NET_NO_TC_BWD if no_tc_bwd else 0) | \
        (_abi.NET_NO_WIDE_TC if no_wide_tc else 0) | (_abi.NET_NO_FUSED_BWD if no_fused_bwd else 0) | \
        (_abi.NET_NO_WIDE_RESIDENT if no_wide_resident else 0) | (_abi.NET_NO_TMA_INPUT if no_tma_input else 0) | \
        (_abi.NET_NO_TC3 if no_tc3 else 0)
    if dependent_launch is not None:
        f |= {0: _abi.NET_PDL_NEVER, 1: 0, 2: _abi.NET_PDL_ALWAYS}[int(dependent_launch)]
    _EXTRA_NET_FLAGS |= f
    if no_phase_cluster:
        _PHASE_FLAGS |= _abi.RES_NO_CLUSTER
    try:
        yield
    finally:
        _EXTRA_NET_FLAGS, _PHASE_FLAGS = prev


def set_default_path_flags(**kw):
    """Non-scoped form of ``path_flags`` for ablation scripts (``profiles/``): replaces the process defaults held in
    this Python module -- the shared library itself has no switches."""
    global _EXTRA_NET_FLAGS, _PHASE_FLAGS, _DEFAULT_CM
    _DEFAULT_CM = None
    _EXTRA_NET_FLAGS = _PHASE_FLAGS = 0
    cm = path_flags(**kw)
    cm.__enter__()          # never exited: the flags stay until the next call ...
    _DEFAULT_CM = cm        # ... so the generator must stay alive (collecting it would run its `finally` and undo them)


def net_from_module(dnn) -> Net:
    """Build the descriptor from any module with the reference's DNN structure
    (``layers.layer_i``, ``predict``, ``var_layers.{0,3,5}``) -- ours or the reference's.  A module
    constructed with ``logvar=False`` (01:436) gets ``NET_NO_LOGVAR``: the kernels then treat the
    log-variance as identically zero (loss, gradients, a_u)."""
    L = dnn.depth - 1
    lin = [getattr(dnn.layers, f"layer_{i}") for i in range(L)]
    H = lin[0].out_features
    heads = [dnn.predict, dnn.var_layers[0], dnn.var_layers[3], dnn.var_layers[5]]
    tensors = []
    for m in lin + heads:
        for t in (m.weight, m.bias):
            _require_cuda(t, "DNN parameter")
            if not t.is_contiguous():
                raise RuntimeError("b200pinn: DNN parameters must be contiguous")
            tensors.append(t)
    d = PinnNet()
    d.n_in, d.width, d.n_hidden = lin[0].in_features, H, L
    for i, m in enumerate(lin):
        d.W[i] = m.weight.data_ptr()
        d.b[i] = m.bias.data_ptr()
    d.Wp, d.bp = dnn.predict.weight.data_ptr(), dnn.predict.bias.data_ptr()
    d.Wv0, d.bv0 = heads[1].weight.data_ptr(), heads[1].bias.data_ptr()
    d.Wv1, d.bv1 = heads[2].weight.data_ptr(), heads[2].bias.data_ptr()
    d.Wv2, d.bv2 = heads[3].weight.data_ptr(), heads[3].bias.data_ptr()
    return Net(d, H, L, tensors, _abi.NET_NO_LOGVAR if getattr(dnn, "logvar", True) is False else 0)


def param_layout(width: int, n_hidden: int):
    """Offsets (in floats) of every tensor in the padded flat bucket used by
    ``mlp_backward`` / ``adam_step``: each tensor starts on a 16-byte boundary."""
    pad4 = lambda v: (v + 3) & ~3
    names, shapes = [], []
    for l in range(n_hidden):
        names += [f"layers.layer_{l}.weight", f"layers.layer_{l}.bias"]
        shapes += [(width, _abi.N_IN if l == 0 else width), (width,)]
    names += ["predict.weight", "predict.bias", "var_layers.0.weight", "var_layers.0.bias",
              "var_layers.3.weight", "var_layers.3.bias", "var_layers.5.weight", "var_layers.5.bias"]
    shapes += [(1, width), (1,), (width // 2, width), (width // 2,), (width // 4, width // 2), (width // 4,),
               (1, width // 4), (1,)]
    offs, o = [], 0
    for s in shapes:
        offs.append(o)
        o = pad4(o + int(np.prod(s)))
    assert o == _abi.lib().pinn_param_count(width, n_hidden), "layout mismatch with libb200pinn"
    return names, shapes, offs, o


def make_dropout(p: float, seed: int = 0, sample_offset: int = 0, pass_offset: int = 0,
                 masks: Optional[torch.Tensor] = None, mask_rows: int = 0) -> PinnDropout:
    d = PinnDropout()
    d.p = float(p)
    d.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    d.sample_offset, d.pass_offset = int(sample_offset), int(pass_offset)
    if masks is not None:
        if not masks.is_cuda or masks.dtype != torch.uint8 or not masks.is_contiguous():
            raise RuntimeError("b200pinn: injected masks must be a contiguous CUDA uint8 tensor")
        d.masks = masks.data_ptr()
        d.mask_sample_stride_n = int(mask_rows)
        d._keepalive = masks        # the struct only holds a raw pointer: pin the tensor's lifetime to it
    return d


def mlp_forward(net: Net, x: torch.Tensor, drop: Optional[PinnDropout] = None):
    """K1: ``DNN.forward`` (01:421-438) -> ``(u[n], logvar[n])``."""
    global LAUNCHES
    _require_cuda(x, "x")
    x = x.contiguous()
    n = x.shape[0]
    u = torch.empty(n, device=x.device, dtype=torch.float32)
    s = torch.empty(n, device=x.device, dtype=torch.float32)
    L = _abi.lib()
    nb = L.pinn_mlp_fwd_workspace_bytes_flags(net.width, net.n_hidden, n, net.flags)
    ws = _workspace("fwd", nb, x.device)
    with torch.cuda.device(x.device):
        check(L.pinn_mlp_fwd(net.ref(), ptr(x), n, C.byref(drop) if drop is not None else None,
                             ptr(u), ptr(s), ptr(ws), nb, _stream()), "pinn_mlp_fwd")
    LAUNCHES += 1
    return u, s


def mlp_backward(net: Net, x: torch.Tensor, drop: Optional[PinnDropout], grad_u=None, grad_logvar=None,
                 y=None, n_global: int = 0, grad_flat=None, loss_sums=None):
    """K2: parameter gradients in the padded flat layout.  Either upstream grads
    (autograd path) or ``y`` (fused aleatoric loss, 01:916-927) must be given."""
    global LAUNCHES
    _require_cuda(x, "x")
    x = x.contiguous()
    n = x.shape[0]
    L = _abi.lib()
    total = L.pinn_param_count(net.width, net.n_hidden)
    if grad_flat is None:       # zeros: the alignment padding of the bucket is never written by the 128/256-wide path
        grad_flat = torch.zeros(total, device=x.device, dtype=torch.float32)
    if loss_sums is None:
        loss_sums = torch.zeros(4, device=x.device, dtype=torch.float64)
    nb = L.pinn_mlp_bwd_workspace_bytes_flags(net.width, net.n_hidden, n, net.flags)
    ws = _workspace("bwd", nb, x.device)
    for t, nm in ((grad_u, "grad_u"), (grad_logvar, "grad_logvar"), (y, "y")):
        if t is not None:
            _require_cuda(t, nm)
            if not t.is_contiguous():
                raise RuntimeError(f"b200pinn: `{nm}` must be contiguous")
    with torch.cuda.device(x.device):
        check(L.pinn_mlp_bwd(net.ref(), ptr(x), n, C.byref(drop) if drop is not None else None,
                             ptr(grad_u), ptr(grad_logvar), ptr(y), int(n_global), ptr(grad_flat),
                             ptr(loss_sums), ptr(ws), nb, _stream()), "pinn_mlp_bwd")
    LAUNCHES += 2
    return grad_flat, loss_sums


def p_h2o_f32() -> float:
    """``10 ** x`` at the constant Tc = 55 in fp32, op for op as 01:745,752-753."""
    Tc = np.float32(55)
    x = np.float32(-2.1794) + np.float32(0.02953) * Tc - np.float32(9.1837e-5) * (Tc ** 2) \
        + np.float32(1.4454e-7) * (Tc ** 3)
    return float(np.float32(10) ** np.float32(x))


def make_scalers(x_scal, u_scal) -> PinnScalers:
    """Fold two sklearn ``MinMaxScaler``s (duck-typed: ``min_``, ``scale_``,
    ``data_min_``, ``data_max_``, ``feature_range``) into the kernel's affine form."""
    s = PinnScalers()
    inv = 1.0 / np.asarray(x_scal.scale_, np.float64)
    off = np.asarray(x_scal.min_, np.float64) * inv
    for j in range(_abi.N_IN):
        s.x_inv_scale[j] = float(inv[j])
        s.x_off[j] = float(off[j])
    yinv = 1.0 / float(np.asarray(u_scal.scale_).reshape(-1)[0])
    s.y_inv_scale = yinv
    s.y_off = float(np.asarray(u_scal.min_).reshape(-1)[0]) * yinv
    lo, hi = float(u_scal.feature_range[0]), float(u_scal.feature_range[1])
    dmin = np.float32(np.asarray(u_scal.data_min_).reshape(-1)[0])
    dmax = np.float32(np.asarray(u_scal.data_max_).reshape(-1)[0])
    scale_y = np.float32(hi - lo) / (dmax - dmin + np.float32(1e-12))      # 01:1021
    s.scale_y = float(scale_y)
    s.min_y = float(np.float32(lo) - dmin * scale_y)                       # 01:1022
    s.p_h2o = p_h2o_f32()
    return s


def residuals(x, u, y, scalers: PinnScalers, lambdas: torch.Tensor, families: int, flags: int = 0,
              want_cols: bool = False, halo_x=None, halo_u=None, sums=None, cols=None):
    """K3: returns ``(sums float64[S_COUNT], cols float32[C_COUNT, n] or None)``."""
    global LAUNCHES
    _require_cuda(x, "x")
    x = x.contiguous()
    n = x.shape[0]
    dev = x.device
    L = _abi.lib()
    if sums is None:
        sums = torch.empty(_abi.S_COUNT, device=dev, dtype=torch.float64)
    if want_cols and cols is None:
        cols = torch.zeros(_abi.C_COUNT, n, device=dev, dtype=torch.float32)
    nb = L.pinn_residuals_workspace_bytes(n)
    ws = _workspace("res", nb, dev, zero=True)
    for t, nm in ((u, "u"), (y, "y"), (lambdas, "lambdas"), (halo_x, "halo_x"), (halo_u, "halo_u")):
        if t is not None:
            _require_cuda(t, nm)
            if not t.is_contiguous():
                raise RuntimeError(f"b200pinn: `{nm}` must be contiguous")
    with torch.cuda.device(dev):
        check(L.pinn_residuals(ptr(x), ptr(u), ptr(y), n, C.byref(scalers), ptr(lambdas), families, flags,
                               ptr(halo_x), ptr(halo_u), ptr(cols), ptr(sums), ptr(ws), nb, _stream()),
              "pinn_residuals")
    LAUNCHES += 1
    return sums, cols


def mc_dropout(net: Net, x: torch.Tensor, T: int, drop: PinnDropout, finalize: bool = True, raw: bool = False):
    """K4: one eval forward + ``T`` dropout passes with in-kernel Welford statistics.
    Returns a dict with ``pred_mean, a_u, e_u`` (finalize) and/or ``mean, m2, sum_logvar`` (raw)."""
    global LAUNCHES
    _require_cuda(x, "x")
    x = x.contiguous()
    n, dev = x.shape[0], x.device
    new = lambda: torch.empty(n, device=dev, dtype=torch.float32)
    out = {"pred_mean": new()}
    if finalize:
        out["a_u"], out["e_u"] = new(), new()
    if raw:
        out["mean"], out["m2"], out["sum_logvar"] = new(), new(), new()
    L = _abi.lib()
    nb = L.pinn_mc_workspace_bytes_flags(net.width, net.n_hidden, n, net.flags)
    ws = _workspace("mc", nb, dev)
    with torch.cuda.device(dev):
        check(L.pinn_mc_dropout(net.ref(), ptr(x), n, int(T), C.byref(drop), ptr(out["pred_mean"]),
                                ptr(out.get("a_u")), ptr(out.get("e_u")), ptr(out.get("mean")),
                                ptr(out.get("m2")), ptr(out.get("sum_logvar")), ptr(ws), nb, _stream()),
              "pinn_mc_dropout")
    LAUNCHES += 1
    return out


def new_step_counter(device) -> torch.Tensor:
    return torch.zeros(2, device=device, dtype=torch.int64)


def adam_step(params, grads, exp_avg, exp_avg_sq, step_counter, lr0, gamma, step_size, grad_scale=1.0,
              active=None, lo=None, hi=None, advance=True):
    """f3: fused Adam + StepLR (+ clamp) over a flat fp32 bucket; state stays on device."""
    global LAUNCHES
    L = _abi.lib()
    with torch.cuda.device(params.device):
        check(L.pinn_adam_step(ptr(params), ptr(grads), ptr(exp_avg), ptr(exp_avg_sq), params.numel(),
                               ptr(step_counter), float(lr0), float(gamma), int(step_size), float(grad_scale),
                               ptr(active), ptr(lo), ptr(hi), 1 if advance else 0, _stream()), "pinn_adam_step")
    LAUNCHES += 1


P2P_FLAG_WORDS = 64


def adam_step_p2p(params, peer_ptrs, rank, world, slot, step_tag, exp_avg, exp_avg_sq, step_counter, lr0, gamma, step_size):
    """Fused gradient all-reduce (NVLink peer loads, rank-ordered sum) + Adam + StepLR; see ``pinn_adam_step_p2p``."""
    global LAUNCHES
    L = _abi.lib()
    with torch.cuda.device(params.device):
        check(L.pinn_adam_step_p2p(ptr(params), ptr(peer_ptrs), int(rank), int(world), int(slot), int(step_tag) & 0xFFFFFFFF,
                                   ptr(exp_avg), ptr(exp_avg_sq), params.numel(), ptr(step_counter), float(lr0), float(gamma),
                                   int(step_size), _stream()), "pinn_adam_step_p2p")
    LAUNCHES += 1


def dp_bucket_words(width: int, n_hidden: int, world: int) -> int:
    """32-bit words of the symmetric buffer ``train_dnn_steps_dp`` exchanges the gradient bucket through."""
    return int(_abi.lib().pinn_dp_bucket_words(int(width), int(n_hidden), int(world)))


def train_dnn_steps_dp(net: Net, x, drop: Optional[PinnDropout], y, n_global: int, params_flat, exp_avg, exp_avg_sq,
                       step_counter, lr0, gamma, step_size, n_steps: int, peer_ptrs, rank: int, world: int, first_tag: int,
                       loss_sums, grad_flat=None):
    """``n_steps`` data-parallel ``train_dnn`` steps from one call: the gradient sum over the ranks runs inside the
    gradient-reduce launch over NVLink peer memory, Adam + StepLR in the same launch (``pinn_train_dnn_steps_dp``)."""
    global LAUNCHES
    _require_cuda(x, "x")
    _require_cuda(y, "y")
    if not x.is_contiguous() or not y.is_contiguous():
        raise RuntimeError("b200pinn: `x` and `y` must be contiguous")
    n = x.shape[0]
    L = _abi.lib()
    nb = L.pinn_mlp_bwd_workspace_bytes_flags(net.width, net.n_hidden, n, net.flags)
    ws = _workspace("bwd", nb, x.device)
    with torch.cuda.device(x.device):
        check(L.pinn_train_dnn_steps_dp(net.ref(), ptr(x), n, C.byref(drop) if drop is not None else None, ptr(y), int(n_global),
                                        ptr(params_flat), ptr(exp_avg), ptr(exp_avg_sq), ptr(step_counter), float(lr0),
                                        float(gamma), int(step_size), int(n_steps), ptr(peer_ptrs), int(rank), int(world),
                                        int(first_tag) & 0xFFFFFFFF, ptr(grad_flat), ptr(loss_sums), ptr(ws), nb, _stream()),
              "pinn_train_dnn_steps_dp")
    LAUNCHES += 2 * int(n_steps) + 1


def adam_step_from_sums(params, sums, grad_slot, exp_avg, exp_avg_sq, step_counter, lr0, gamma, step_size,
                        lo=None, hi=None):
    """f3 for the physics scalars: gradient i = sums[grad_slot[i]] / sums[N]."""
    global LAUNCHES
    L = _abi.lib()
    with torch.cuda.device(params.device):
        check(L.pinn_adam_step_from_sums(ptr(params), ptr(sums), ptr(grad_slot), ptr(exp_avg), ptr(exp_avg_sq),
                                         params.numel(), ptr(step_counter), float(lr0), float(gamma),
                                         int(step_size), ptr(lo), ptr(hi), _stream()),
              "pinn_adam_step_from_sums")
    LAUNCHES += 1


def scalar_phase(x, u, y, scalers: PinnScalers, lambdas, families: int, flags: int, first: int, slots, bounds,
                 exp_avg, exp_avg_sq, step_counter, lr0, gamma, step_size, n_steps: int, sums):
    """f3, persistent form: ``n_steps`` optimiser steps of one scalar phase (residual sums -> Adam -> clamp)
    in one cooperative launch; ``sums`` receives the totals of the last step.  See ``pinn_scalar_phase``."""
    global LAUNCHES
    _require_cuda(x, "x")
    if not x.is_contiguous():
        raise RuntimeError("b200pinn: `x` must be contiguous")
    for t, nm in ((u, "u"), (y, "y"), (lambdas, "lambdas")):
        if t is not None:
            _require_cuda(t, nm)
            if not t.is_contiguous():
                raise RuntimeError(f"b200pinn: `{nm}` must be contiguous")
    cnt = len(slots)
    c_slot = (C.c_int32 * cnt)(*[int(s) for s in slots])
    c_lo = (C.c_float * cnt)(*[float(b[0]) for b in bounds])
    c_hi = (C.c_float * cnt)(*[float(b[1]) for b in bounds])
    L = _abi.lib()
    nb = L.pinn_scalar_phase_workspace_bytes()
    ws = _workspace("phase", nb, x.device, zero=True)
    with torch.cuda.device(x.device):
        check(L.pinn_scalar_phase(ptr(x), ptr(u), ptr(y), x.shape[0], C.byref(scalers), ptr(lambdas), families,
                                  int(flags) | _PHASE_FLAGS,
                                  int(first), cnt, c_slot, c_lo, c_hi, ptr(exp_avg), ptr(exp_avg_sq), ptr(step_counter),
                                  float(lr0), float(gamma), int(step_size), int(n_steps), ptr(sums), ptr(ws), nb,
                                  _stream()), "pinn_scalar_phase")
    LAUNCHES += 1


def train_dnn_step(net: Net, x, drop: Optional[PinnDropout], y, n_global: int, params_flat, exp_avg, exp_avg_sq,
                   step_counter, lr0, gamma, step_size, grad_flat, loss_sums, n_steps: int = 1):
    """One ``train_dnn`` step (01:948-955) in one call: K2a, K2b and the gradient reduce with Adam + StepLR fused
    into it; ``n_steps`` > 1 enqueues that many consecutive steps from one call (Philox masks only: step i uses the
    dropout descriptor's ``pass_offset + i``).  ``net``'s tensors must be views into ``params_flat``.
    See ``pinn_train_dnn_step`` / ``pinn_train_dnn_steps``."""
    global LAUNCHES
    _require_cuda(x, "x")
    if not x.is_contiguous() or not y.is_contiguous():
        raise RuntimeError("b200pinn: `x` and `y` must be contiguous")
    _require_cuda(y, "y")
    n = x.shape[0]
    L = _abi.lib()
    nb = L.pinn_mlp_bwd_workspace_bytes_flags(net.width, net.n_hidden, n, net.flags)
    ws = _workspace("bwd", nb, x.device)
    with torch.cuda.device(x.device):
        check(L.pinn_train_dnn_steps(net.ref(), ptr(x), n, C.byref(drop) if drop is not None else None, ptr(y),
                                     int(n_global), ptr(params_flat), ptr(exp_avg), ptr(exp_avg_sq), ptr(step_counter),
                                     float(lr0), float(gamma), int(step_size), int(n_steps), ptr(grad_flat), ptr(loss_sums),
                                     ptr(ws), nb, _stream()), "pinn_train_dnn_steps")
    LAUNCHES += 3 * int(n_steps)
