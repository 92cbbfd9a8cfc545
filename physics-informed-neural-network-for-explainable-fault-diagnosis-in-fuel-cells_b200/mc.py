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
        x = X[:, 0:].detach().to(dev, torch.float32).contiguous()
        masks = getattr(dnn, "_injected_mc", None)
        calls = getattr(dnn, "_drop_calls", 0)
        out = mc_dropout_device(dnn, x, mc_times, float(dropout), pass_offset=calls, masks=masks)
        if hasattr(dnn, "_drop_calls"):
            dnn._drop_calls = calls + int(mc_times)
    finally:
        for name, module in dnn.named_modules():            # 01:1468-1470
            if isinstance(module, torch.nn.Dropout):
                module.p = original[name]
        dnn.eval()                                           # 01:1473
    pm = out["pred_mean"].cpu().numpy()
    au = out["a_u"].cpu().numpy()
    eu = out["e_u"].cpu().numpy()
    return pm.squeeze(), au.squeeze(), eu.squeeze()
