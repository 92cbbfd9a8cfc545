"""Generate golden vectors by running the UNMODIFIED reference in the build container.

Usage (build container only -- /root/reference does not exist on the GPU box):

    python tests/golden/make_golden.py

Imports ``/root/reference/01_train_pinn_multiphysics_model.py`` by path with the two
out-of-tree shims of SURVEY.md section 8c (matplotlib stub; ``StepLR(verbose=)`` kwarg
dropped), feeds it seeded synthetic data from ``b200pinn.synthetic`` and records
inputs, parameters, dropout masks and the reference's own outputs into
``tests/golden/*.npz``.  Nothing here is used at test time except the ``.npz`` files.
"""
import importlib.util
import os
import sys
from unittest.mock import MagicMock

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/01_train_pinn_multiphysics_model.py"


def load_reference():
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.font_manager", "matplotlib.lines"):
        sys.modules.setdefault(m, MagicMock())
    spec = importlib.util.spec_from_file_location("ref01", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    real = ref.StepLR

    def step_lr(opt, step_size, gamma=0.1, **kw):
        kw.pop("verbose", None)
        return real(opt, step_size=step_size, gamma=gamma, **kw)

    ref.StepLR = step_lr
    return ref


class MaskTap:
    """Record the scaled mask of every active nn.Dropout call without disturbing
    the RNG stream (SURVEY 8c recipe)."""

    def __init__(self, dnn):
        self.masks, self._state, self.handles = [], {}, []
        for mod in dnn.modules():
            if isinstance(mod, torch.nn.Dropout):
                self.handles.append(mod.register_forward_pre_hook(self._pre))
                self.handles.append(mod.register_forward_hook(self._post))

    def _pre(self, mod, inp):
        if mod.training:
            self._state[id(mod)] = torch.get_rng_state()

    def _post(self, mod, inp, out):
        if mod.training:
            after = torch.get_rng_state()
            torch.set_rng_state(self._state.pop(id(mod)))
            m = F.dropout(torch.ones_like(inp[0]), mod.p, True)
            assert torch.equal(inp[0] * m, out)
            torch.set_rng_state(after)
            self.masks.append((m != 0).numpy())

    def take(self):
        out, self.masks = self.masks, []
        return out

    def close(self):
        for h in self.handles:
            h.remove()


def scaler_arrays(prefix, s):
    return {f"{prefix}_min": s.min_, f"{prefix}_scale": s.scale_,
            f"{prefix}_data_min": s.data_min_, f"{prefix}_data_max": s.data_max_}


def sd_arrays(dnn):
    return {"P:" + k: v.detach().numpy().copy() for k, v in dnn.state_dict().items()}


def lam_names():
    return (["lambda_1", "lambda_2", "lambda_3", "lambda_4"] + [f"lambda_T{i}" for i in range(1, 6)]
            + [f"lambda_H{i}" for i in range(1, 5)] + [f"lambda_O{i}" for i in range(1, 5)])


def lam_vector(model):
    return np.array([getattr(model, n).item() for n in lam_names()], np.float64)


def pack(masks):
    """list of bool arrays [N,w] -> one packed uint8 array + widths."""
    cat = np.concatenate([m.reshape(m.shape[0], -1) for m in masks], axis=1)
    return np.packbits(cat, axis=1), cat.shape[1]


def build(ref, layers, n, seed, p, logvar=True):
    from b200pinn.synthetic import make_scaled_dataset

    x, y, sx, sy = make_scaled_dataset(n, seed)
    torch.manual_seed(0)
    model = ref.PhysicsInformedNN(torch.tensor(x), torch.tensor(y), layers, sx, sy, p, logvar)
    # Move lambda_1..3 off the values the synthetic voltages were generated with:
    # at the exact optimum the mode-A gradient is pure cancellation noise (the
    # reference's own fp32 result differs from fp64 by 1 %), useless as a ruler.
    with torch.no_grad():
        model.lambda_1.mul_(1.2)
        model.lambda_2.mul_(1.5)
        model.lambda_3.mul_(1.1)
    return model, x, y, sx, sy


def golden_net(ref, name, layers, n, seed, p=0.2, mc_T=6, mc_p=0.4, logvar=True):
    model, x, y, sx, sy = build(ref, layers, n, seed, p, logvar)
    g = dict(layers=np.array(layers), p=p, x=x, y=y, logvar=logvar, **scaler_arrays("sx", sx), **scaler_arrays("sy", sy))
    g.update(sd_arrays(model.dnn))
    g["lam0"] = lam_vector(model)
    X = torch.tensor(x)

    # --- DNN forward, eval + train (01:421-438)
    model.dnn.eval()
    out, lv = model.dnn(X)
    g["eval_out"], g["eval_logvar"] = out.detach().numpy(), lv.detach().numpy()
    tap = MaskTap(model.dnn)
    model.dnn.train()
    torch.manual_seed(11)
    for q in model.dnn.parameters():
        q.requires_grad = True
        q.grad = None
    out, lv = model.net_u(model.x)
    masks = tap.take()
    g["train_masks"], g["mask_width"] = pack(masks)
    g["train_out"], g["train_logvar"] = out.detach().numpy(), lv.detach().numpy()
    # --- aleatoric loss + autograd grads (01:950-953)
    loss = model.aleatoric_loss(model.u, out, lv)
    loss.backward()
    g["aleatoric_loss"] = loss.item()
    for k, q in model.dnn.named_parameters():
        if q.grad is not None and not k.startswith("lambda"):
            g["G:" + k] = q.grad.numpy().copy()
    model.dnn.eval()

    # --- residual tuples (01:535-914)
    names = {"V": ["f", "V_act", "V_ohm", "V_conc", "E", "V_est5", "i", "il", "V_out5"],
             "Ts": ["f", "T_pred", "T_real"], "T": ["f", "T_pred", "T_real"],
             "H": ["f", "actual", "target", "I_total", "I_thr"],
             "O": ["f", "actual", "target", "Q", "o2"]}
    fns = {"V": model.net_f_V, "Ts": model.net_f_T_simple, "T": model.net_f_T,
           "H": model.net_f_H, "O": model.net_f_O}
    for fam, fn in fns.items():
        res = fn(X, sx)
        for nm, t in zip(names[fam], res):
            g[f"R:{fam}:{nm}"] = t.detach().numpy().copy()

    # --- one-step losses + lambda grads of every phase trainer
    def grads_of(loss_t, plist):
        for q in plist:
            q.grad = None
            q.requires_grad_(True)
        loss_t.backward()
        return np.array([0.0 if q.grad is None else q.grad.item() for q in plist])

    lamV = [model.lambda_1, model.lambda_2, model.lambda_3]
    for mode in (False, True):
        u_pred, _ = model.net_u(model.x)
        f_pred, _, _, _, _, V5, _, _, _ = model.net_f_V(model.X, model.x_scal)
        scale_y = 2.0 / (torch.tensor(sy.data_max_, dtype=torch.float32) - torch.tensor(sy.data_min_, dtype=torch.float32) + 1e-12)
        min_y = -1.0 - torch.tensor(sy.data_min_, dtype=torch.float32) * scale_y
        phys = torch.mean(f_pred ** 2) if mode else torch.mean((model.u - (V5 * scale_y + min_y)) ** 2)
        data = torch.mean((model.u - u_pred) ** 2)
        tot = phys + data
        g[f"L:lambda:{int(mode)}"] = np.array([tot.item(), phys.item(), data.item()])
        g[f"LG:lambda:{int(mode)}"] = grads_of(tot, lamV)
    fT, _, _ = model.net_f_T_simple(model.X, sx)
    lT = torch.mean(fT ** 2)
    g["L:thermal"] = np.array([lT.item(), torch.mean(torch.abs(fT)).item()])
    g["LG:thermal"] = grads_of(lT, [model.lambda_T1, model.lambda_T2, model.lambda_T3, model.lambda_T4, model.lambda_T5])
    fH, *_ = model.net_f_H(model.X, sx)
    lH = torch.mean(fH ** 2)
    g["L:hydrogen"] = np.array([lH.item()])
    g["LG:hydrogen"] = grads_of(lH, [model.lambda_H1, model.lambda_H2, model.lambda_H3, model.lambda_H4])
    fO, *_ = model.net_f_O(model.X, sx)
    lO = torch.mean(fO ** 2)
    g["L:oxygen"] = np.array([lO.item()])
    g["LG:oxygen"] = grads_of(lO, [model.lambda_O1, model.lambda_O2, model.lambda_O3, model.lambda_O4])

    # --- short trajectories of the reference's own trainers (Adam+StepLR+clamp)
    K = 5
    model.train_lambda(K, False)
    g["traj:lambda0"] = lam_vector(model)
    model.train_lambda(K, True)
    g["traj:lambda1"] = lam_vector(model)
    model.train_thermal(K)
    g["traj:thermal"] = lam_vector(model)
    model.train_hydrogen(K)
    g["traj:hydrogen"] = lam_vector(model)
    model.train_oxygen(K)
    g["traj:oxygen"] = lam_vector(model)
    # train_dnn with captured masks (3 steps) -> parameters afterwards
    torch.manual_seed(21)
    model.train_dnn(3)
    m3 = tap.take()
    per = len(m3) // 3
    for s in range(3):
        g[f"traj:dnn_masks{s}"], _ = pack(m3[s * per:(s + 1) * per])
    for k, v in model.dnn.state_dict().items():
        if not k.startswith("lambda"):
            g["traj:dnn:" + k] = v.detach().numpy().copy()

    # --- MC dropout (01:1413-1491); keep the first L+1 masks of every 2(L+1)
    torch.manual_seed(31)
    tap.take()
    pm, au, eu = ref.get_MC_samples(model, X, sx, mc_times=mc_T, dropout=mc_p)
    mm = tap.take()
    nd = len(layers) - 2 + (1 if logvar else 0)        # logvar=False never runs var_layers (01:428-436): no mask for its dropout
    assert len(mm) == mc_T * 2 * nd, (len(mm), mc_T, nd)
    for t in range(mc_T):
        g[f"mc_masks{t}"], _ = pack(mm[t * 2 * nd: t * 2 * nd + nd])
    g["mc_T"], g["mc_p"] = mc_T, mc_p
    g["mc_pred_mean"], g["mc_a_u"], g["mc_e_u"] = pm, au, eu
    for k, v in model.dnn.state_dict().items():
        if not k.startswith("lambda"):
            g["mcP:" + k] = v.detach().numpy().copy()
    tap.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **g)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in list(g.items())[:6]})


if __name__ == "__main__":
    ref = load_reference()
    which = sys.argv[1:] or ["net64", "net32", "net256", "net64nl"]
    if "net64" in which:
        golden_net(ref, "net64", [8, 64, 64, 64, 1], n=320, seed=1)
    if "net32" in which:
        golden_net(ref, "net32", [8, 32, 32, 1], n=97, seed=7, p=0.1, mc_T=4, mc_p=0.5)
    if "net256" in which:        # the reference's own Layers (01:2139)
        golden_net(ref, "net256", [8, 256, 256, 256, 1], n=160, seed=3, mc_T=4)
    if "net64nl" in which:       # DNN(logvar=False): log-variance identically zero (01:436)
        golden_net(ref, "net64nl", [8, 64, 64, 64, 1], n=200, seed=5, mc_T=4, logvar=False)
