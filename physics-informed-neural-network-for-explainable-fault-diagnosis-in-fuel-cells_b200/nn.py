"""``DNN`` -- drop-in for the reference's network class (01:389-438).

Same constructor, attributes, ``state_dict`` keys and ``forward`` contract; real
``torch.nn.Dropout`` submodules are kept because ``get_MC_samples`` rewrites their
``.p`` and toggles ``train()/eval()`` from outside (01:1432-1473) -- both
are read at call time.  The arithmetic runs in ``libb200pinn.so``: ``forward`` is
kernel K1, its autograd backward is kernel K2 (forward recomputed in-kernel, no
activations saved).  CPU tensors raise: there is no CPU path.
"""
from __future__ import annotations

from collections import OrderedDict
from contextlib import contextmanager

import torch

from . import kernels as K


class _DNNFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dnn, drop_cfg, *params):
        net = K.net_from_module(dnn)
        drop = K.make_dropout(**drop_cfg) if drop_cfg is not None else None
        u, s = K.mlp_forward(net, x.detach(), drop)
        ctx.dnn, ctx.drop_cfg = dnn, drop_cfg
        ctx.save_for_backward(x.detach())
        return u.view(-1, 1), s.view(-1, 1)

    @staticmethod
    def backward(ctx, grad_u, grad_s):
        if not any(ctx.needs_input_grad[3:]):
            return (None,) * len(ctx.needs_input_grad)
        (x,) = ctx.saved_tensors
        dnn = ctx.dnn
        net = K.net_from_module(dnn)
        drop = K.make_dropout(**ctx.drop_cfg) if ctx.drop_cfg is not None else None
        n = x.shape[0]
        gu = (grad_u if grad_u is not None else torch.zeros(n, 1, device=x.device)).reshape(-1).contiguous().float()
        gs = (grad_s if grad_s is not None else torch.zeros(n, 1, device=x.device)).reshape(-1).contiguous().float()
        flat, _ = K.mlp_backward(net, x, drop, grad_u=gu, grad_logvar=gs)
        names, shapes, offs, _ = K.param_layout(net.width, net.n_hidden)
        grads = []
        for shp, off in zip(shapes, offs):
            cnt = 1
            for d in shp:
                cnt *= d
            grads.append(flat[off:off + cnt].view(shp))
        # dL/dx is deliberately not formed (the reference computes and discards it, 01:446)
        return (None, None, None, *grads)


class DNN(torch.nn.Module):
    def __init__(self, p, logvar, layers):
        super().__init__()
        if len(layers) < 3 or layers[-1] != 1:
            raise ValueError("b200pinn.DNN: layers must be [n_in, H, ..., H, 1]")
        if any(h != layers[1] for h in layers[1:-1]):
            raise ValueError("b200pinn.DNN: all hidden layers must share one width (kernel restriction)")
        self.depth = len(layers) - 1
        self.p = p
        self.logvar = logvar
        self.activation = torch.nn.Tanh
        seq = []
        for i in range(self.depth - 1):
            seq.append((f"layer_{i}", torch.nn.Linear(layers[i], layers[i + 1])))
            seq.append((f"activation_{i}", self.activation()))
            seq.append((f"dropout_{i}", torch.nn.Dropout(p=self.p)))
        self.layers = torch.nn.Sequential(OrderedDict(seq))
        h = layers[-2]
        self.predict = torch.nn.Linear(h, layers[-1])
        self.var_layers = torch.nn.Sequential(
            torch.nn.Linear(h, h // 2), torch.nn.Tanh(), torch.nn.Dropout(p=self.p),
            torch.nn.Linear(h // 2, h // 4), torch.nn.Tanh(), torch.nn.Linear(h // 4, layers[-1]))
        # dropout stream: Philox key + a counter advanced once per stochastic forward
        self._drop_seed = None
        self._drop_calls = 0
        self._row_offset = 0     # global index of this shard's first row (data-parallel training: Philox counters are keyed on
                                 # GLOBAL rows, so an N-GPU run draws the masks of the 1-GPU run)
        self._injected = None  # uint8 keep bits [n, D] (or [calls, n, D]) for the next stochastic forwards
        self._inj_idx = 0

    # ------------------------------------------------------------------ helpers
    def kernel_params(self):
        """Parameters in the kernel's canonical order (== the first 2L+8 of ``parameters()``)."""
        mods = [getattr(self.layers, f"layer_{i}") for i in range(self.depth - 1)]
        mods += [self.predict, self.var_layers[0], self.var_layers[3], self.var_layers[5]]
        out = []
        for m in mods:
            out += [m.weight, m.bias]
        return out

    def active_dropout_p(self) -> float:
        """Drop probability in force right now: 0 in eval mode, else the (common) ``.p`` of
        the ``nn.Dropout`` submodules as last written by the caller (01:1449-1454)."""
        drops = [m for m in self.modules() if isinstance(m, torch.nn.Dropout)]
        live = {float(m.p) for m in drops if m.training}
        if not live:
            return 0.0
        if len(live) != 1 or any(not m.training for m in drops):
            raise NotImplementedError("b200pinn.DNN: the fused kernels need one dropout rate/mode for all layers")
        return live.pop()

    def next_dropout_cfg(self, n, p):
        if p <= 0.0:
            return None
        if self._drop_seed is None:
            self._drop_seed = torch.initial_seed()
        cfg = dict(p=p, seed=self._drop_seed, sample_offset=self._row_offset, pass_offset=self._drop_calls)
        self._drop_calls += 1
        if self._injected is not None:
            m = self._injected
            if m.dim() == 3:                       # one mask set per stochastic call, consumed in order
                m = m[self._inj_idx % m.shape[0]]
                self._inj_idx += 1
            if m.shape[0] != n:
                raise RuntimeError("b200pinn.DNN: injected masks have the wrong number of rows")
            cfg.update(masks=m, mask_rows=n)
        return cfg

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("b200pinn.DNN.forward: input is on the CPU; this implementation runs only on "
                               "CUDA (sm_100a) -- no CPU fallback")
        x32 = x if x.dtype == torch.float32 else x.float()
        cfg = self.next_dropout_cfg(x32.shape[0], self.active_dropout_p())
        out, logvar = _DNNFunction.apply(x32, self, cfg, *self.kernel_params())
        if not self.logvar:
            logvar = torch.zeros(out.size()).to(out.device)      # 01:436
        return out, logvar


@contextmanager
def inject_masks(dnn: DNN, masks: torch.Tensor):
    """Parity hook: use ``masks`` (uint8 keep bits ``[n, L*H + H/2]``) instead of Philox
    for every stochastic forward inside the block."""
    prev = dnn._injected
    dnn._injected = masks.contiguous()
    dnn._inj_idx = 0
    try:
        yield dnn
    finally:
        dnn._injected = prev
