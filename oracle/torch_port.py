"""CPU baseline: PyTorch-eager restatement of the reference's hot loops, op for op.

TEST/BENCH INFRASTRUCTURE ONLY (never imported by the product package).  The reference is a
PyTorch program, so its honest CPU baseline is PyTorch eager on the host cores with the
reference's own op sequence -- including the costs it chooses to pay: a second, discarded
forward plus two scaler round trips inside every ``predict`` (01:1407), T redundant eval
passes (01:1442-1445), per-pass host copies and the (T,N,1) numpy reduction (01:1475-1486).
/root/reference cannot travel to the GPU box, hence this port (``cpu_baseline.kind`` =
"port") -- used only where oracle/_ref (the staged unmodified scripts) is absent.
tests/test_oracle_golden.py pins its forward, residual, ``get_MC_samples_port`` (same torch RNG stream as the
reference) and ``make_dnn_trainer`` to the golden vectors.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F


class PortDNN(torch.nn.Module):
    """01:389-438."""

    def __init__(self, p, logvar, layers):
        super().__init__()
        self.depth = len(layers) - 1
        self.p, self.logvar = p, logvar
        seq = []
        for i in range(self.depth - 1):
            seq += [(f"layer_{i}", torch.nn.Linear(layers[i], layers[i + 1])), (f"activation_{i}", torch.nn.Tanh()),
                    (f"dropout_{i}", torch.nn.Dropout(p=p))]
        self.layers = torch.nn.Sequential(OrderedDict(seq))
        h = layers[-2]
        self.predict = torch.nn.Linear(h, layers[-1])
        self.var_layers = torch.nn.Sequential(torch.nn.Linear(h, h // 2), torch.nn.Tanh(), torch.nn.Dropout(p=p),
                                              torch.nn.Linear(h // 2, h // 4), torch.nn.Tanh(),
                                              torch.nn.Linear(h // 4, layers[-1]))

    def forward(self, x):
        feats = self.layers(x)
        out = self.predict(feats)
        if self.logvar:
            logvar = torch.log(F.softplus(self.var_layers(feats)) + 1e-6)
        else:
            logvar = torch.zeros(out.size()).to(out.device)            # 01:436
        return out, logvar


class PortPINN:
    """The parts of ``PhysicsInformedNN`` (01:441-1410) the benchmarks drive."""

    def __init__(self, X, u, layers, x_scal, u_scal, p, logvar=True):
        self.x = X.clone().detach().requires_grad_(True).float()
        self.u = u.clone().detach().float()
        self.X, self.x_scal, self.u_scal = X, x_scal, u_scal
        self.lambda_1 = torch.nn.Parameter(torch.tensor([0.167897923477715]))
        self.lambda_2 = torch.nn.Parameter(torch.tensor([2.36682075851268e-06]))
        self.lambda_3 = torch.nn.Parameter(torch.tensor([2.43414469188443]))
        self.dnn = PortDNN(p, logvar, layers)

    def net_u(self, x):
        return self.dnn(x)

    def net_f_V(self, X, x_scal):
        """01:724-765 (host round trips through the sklearn scalers included)."""
        x_in = X[:, 0:].clone().detach().requires_grad_(True).float()
        real = torch.tensor(x_scal.inverse_transform(X.detach().cpu().numpy()))
        A = torch.tensor([270.0])
        i = real[:, 0:1] / A + 1e-5
        T_out = real[:, 5:6]
        u, _ = self.net_u(x_in)
        V_out = torch.tensor(self.u_scal.inverse_transform(u.detach().cpu().numpy())) / torch.tensor([5.0])
        R, Fc, Tc = torch.tensor([8.314]), torch.tensor([96485.0]), torch.tensor([55.0])
        P_H2 = real[:, 3:4] / 101 + 1
        P_air = real[:, 4:5] / 101 + 1
        alpha, Gf = torch.tensor([0.5]), torch.tensor([-220170.0])
        Tk = T_out + torch.tensor([273.15])
        P_H2O = 10 ** (-2.1794 + 0.02953 * Tc - 9.1837e-5 * (Tc ** 2) + 1.4454e-7 * (Tc ** 3))
        pp_H2 = 0.5 * (P_H2 / torch.exp(1.653 * i / (Tk ** 1.334)) - P_H2O)
        pp_O2 = P_air / torch.exp(4.192 * i / (Tk ** 1.334)) - P_H2O
        b = R * Tk / (2.0 * alpha * Fc)
        V_act = -b * torch.log(i / self.lambda_2)
        V_ohm = -(i * self.lambda_1)
        V_conc = alpha * b * torch.log(1 - i / self.lambda_3)
        E = -Gf / (2 * Fc) - (R * Tk) * torch.log(P_H2O / (pp_H2 * pp_O2 ** 0.5)) / (2 * Fc)
        V_est = E + V_act + V_ohm + V_conc
        return V_est - V_out, V_est * 5

    def aleatoric_loss(self, gt, pred, logvar):
        loss = torch.mean(0.5 * torch.exp(-logvar) * (gt - pred) ** 2 + 0.5 * logvar)
        return loss + 0.01 * torch.mean(torch.abs(logvar))

    def predict(self, X, x_scal):
        """01:1401-1410."""
        u, log_var = self.net_u(X[:, 0:])
        self.net_f_V(X, x_scal)
        return u.detach().cpu().numpy(), log_var.detach().cpu().numpy()

    def make_dnn_trainer(self):
        """Returns a closure running one ``train_dnn`` step (01:948-955)."""
        opt = torch.optim.Adam(self.dnn.parameters(), lr=0.01)
        sch = torch.optim.lr_scheduler.StepLR(opt, step_size=1000, gamma=0.8)
        self.dnn.train()

        def step():
            u_pred, log_var = self.net_u(self.x)
            loss = self.aleatoric_loss(self.u, u_pred, log_var)
            opt.zero_grad()
            loss.backward()
            opt.step()
            sch.step()
            return loss

        return step


def get_MC_samples_port(network, X, x_scal, mc_times=64, dropout=0.6):
    """01:1413-1491."""
    drops = [m for m in network.dnn.modules() if isinstance(m, torch.nn.Dropout)]
    orig = [m.p for m in drops]
    pe, pd, au = [], [], []
    network.dnn.eval()
    for _ in range(mc_times):
        pe.append(network.predict(X, x_scal)[0])
    for m in drops:
        m.p = dropout
    for _ in range(mc_times):
        network.dnn.train()
        u, lv = network.predict(X, x_scal)
        pd.append(u)
        au.append(lv)
    for m, p in zip(drops, orig):
        m.p = p
    network.dnn.eval()
    pd, pe, au = np.array(pd), np.array(pe), np.array(au)
    return (np.mean(pe, axis=0).squeeze(), np.sqrt(np.exp(np.mean(au, axis=0))).squeeze(),
            np.sqrt(np.var(pd, axis=0)).squeeze())
