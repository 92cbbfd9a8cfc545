"""``get_MC_samples`` -- drop-in for the reference's MC-dropout routine (01:1413-1491).

The reference runs ``2*mc_times`` Python-level ``predict`` calls (each with a discarded
second forward and two host scaler round trips), stacks three ``(T, N, 1)`` arrays on the
host and reduces them with numpy.  Here the whole sweep is ONE launch of kernel K4:
Philox masks drawn in registers, Welford mean/variance per sample in registers, 44 bytes
of HBM traffic per sample per sweep.  Return contract is unchanged: three squeezed 1-D
fp32 numpy arrays in the normalised domain; every ``nn.Dropout.p`` is restored and the
network is left in eval mode (01:1468-1473).
"""
from __future__ import annotations

import torch

from . import kernels as K


def mc_dropout_device(dnn, x: torch.Tensor, mc_times: int, dropout: float, seed=None, sample_offset: int = 0,
                      pass_offset: int = 0, masks=None, raw: bool = False):
    """Device-level sweep: ``x`` is a CUDA tensor ``[n, 8]``; returns CUDA tensors (dict of
    ``pred_mean, a_u, e_u`` [+ ``mean, m2, sum_logvar`` when ``raw``]).  ``masks``: optional
    uint8 keep bits ``[T, n, L*H + H/2]`` (parity injection)."""
    net = K.net_from_module(dnn)
    if seed is None:
        seed = getattr(dnn, "_drop_seed", None) or torch.initial_seed()
    drop = K.make_dropout(dropout, seed=seed, sample_offset=sample_offset, pass_offset=pass_offset,
                          masks=masks, mask_rows=x.shape[0] if masks is not None else 0)
    return K.mc_dropout(net, x.detach().float(), int(mc_times), drop, finalize=True, raw=raw)


_SIDE_STREAMS: dict = {}
PIPELINE_MIN_ROWS = 1 << 16      # host inputs at least this long are swept in chunks (copy / compute overlap)
PIPELINE_CHUNKS = 4


def _pipeline_chunks(n: int, wave: int):
    """Row ranges of the host pipeline.  A chunk is a whole number of "waves" (three 128-row tiles per SM: the MC kernel keeps three in flight) so that no chunk
    but the last ends on a partial wave; long inputs start with a short chunk (2 waves, then 4) -- the first chunk's upload
    and the last chunk's download are the only copies that are not hidden behind a sweep."""
    waves = -(-n // wave)
    if waves <= 2 * PIPELINE_CHUNKS:
        sizes = [max(1, -(-waves // PIPELINE_CHUNKS))] * PIPELINE_CHUNKS
    else:
        rest = waves - 6
        k = PIPELINE_CHUNKS - 1
        sizes = [2, 4] + [rest // k + (1 if i < rest % k else 0) for i in range(k)]
    out, lo = [], 0
    for w in sizes:
        if lo >= n:
            break
        hi = min(n, lo + w * wave)
        out.append((lo, hi))
        lo = hi
    return out


def _mc_host_pipelined(dnn, X, mc_times, dropout, pass_offset, dev):
    """Host tensor in, host arrays out, for long inputs: the rows are cut into a few chunks whose
    H2D copy, sweep (K4, alternating between two streams) and D2H copy overlap, so only the first chunk's
    upload and the last chunk's download are exposed.  Philox counters are keyed on the global row index
    (``sample_offset``), so the result is identical to the single-launch sweep."""
    n = X.shape[0]
    Xc = X.detach()
    if Xc.dtype != torch.float32 or not Xc.is_contiguous():
        Xc = Xc.float().contiguous()
    host = torch.empty(3, n, dtype=torch.float32, pin_memory=True)
    xd = torch.empty(n, Xc.shape[1], device=dev, dtype=torch.float32)
    cur = torch.cuda.current_stream(dev)
    if dev not in _SIDE_STREAMS:
        _SIDE_STREAMS[dev] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev))
    s_in, s_out, s_alt = _SIDE_STREAMS[dev]
    s_in.wait_stream(cur)
    s_alt.wait_stream(cur)
    for k, (lo, hi) in enumerate(_pipeline_chunks(n, 384 * torch.cuda.get_device_properties(dev).multi_processor_count)):
        with torch.cuda.stream(s_in):
            xd[lo:hi].copy_(Xc[lo:hi], non_blocking=True)
            up = torch.cuda.Event()
            up.record(s_in)
        # consecutive sweeps alternate between two streams: nothing orders them, so the next chunk's CTAs take over the SMs
        # the current one's last wave leaves idle (a sweep's CTAs do not finish together)
        s_k = cur if k % 2 == 0 else s_alt
        with torch.cuda.stream(s_k):
            s_k.wait_event(up)
            out = mc_dropout_device(dnn, xd[lo:hi], mc_times, dropout, sample_offset=lo, pass_offset=pass_offset)
            done = torch.cuda.Event()
            done.record(s_k)
        with torch.cuda.stream(s_out):
            s_out.wait_event(done)
            for i, name in enumerate(("pred_mean", "a_u", "e_u")):
                host[i, lo:hi].copy_(out[name], non_blocking=True)
                out[name].record_stream(s_out)
    xd.record_stream(s_in)
    xd.record_stream(s_alt)
    cur.wait_stream(s_alt)
    s_out.synchronize()
    return host[0].numpy(), host[1].numpy(), host[2].numpy()


def get_MC_samples(network, X, x_scal, mc_times=64, dropout=0.6):
    dnn = network.dnn
    original = {}
    for name, module in dnn.named_modules():
        if isinstance(module, torch.nn.Dropout):
            original[name] = module.p
    print(f"MC-dropout sweep: {mc_times} passes, dropout {dropout}")
    dnn.eval()
    for name, module in dnn.named_modules():                # 01:1449-1454
        if isinstance(module, torch.nn.Dropout):
            module.p = dropout
    try:
        dev = next(dnn.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("b200pinn.get_MC_samples: the network is on the CPU; there is no CPU path")
        masks = getattr(dnn, "_injected_mc", None)
        calls = getattr(dnn, "_drop_calls", 0)
        host_result = None
        if masks is None and X.device.type == "cpu" and X.shape[0] >= PIPELINE_MIN_ROWS:
            host_result = _mc_host_pipelined(dnn, X[:, 0:], int(mc_times), float(dropout), calls, dev)
        else:
            x = X[:, 0:].detach().to(dev, torch.float32).contiguous()
            out = mc_dropout_device(dnn, x, mc_times, float(dropout), pass_offset=calls, masks=masks)
        if hasattr(dnn, "_drop_calls"):
            dnn._drop_calls = calls + int(mc_times)
    finally:
        for name, module in dnn.named_modules():            # 01:1468-1470
            if isinstance(module, torch.nn.Dropout):
                module.p = original[name]
        dnn.eval()                                           # 01:1473
    if host_result is not None:
        return host_result
    pm = out["pred_mean"].cpu().numpy()
    au = out["a_u"].cpu().numpy()
    eu = out["e_u"].cpu().numpy()
    return pm.squeeze(), au.squeeze(), eu.squeeze()
