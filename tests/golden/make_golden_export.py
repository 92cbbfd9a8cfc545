"""Golden vectors for the export writer (01:1877-2047) and the RF(t) risk series (04:181-300),
produced by the UNMODIFIED reference scripts in the build container.

    python tests/golden/make_golden_export.py      ->  tests/golden/export64.npz
"""
import importlib.util
import os
import sys
from unittest.mock import MagicMock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_golden import MaskTap, load_reference, pack, scaler_arrays  # noqa: E402


def load_script(name, path):
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.font_manager", "matplotlib.lines"):
        sys.modules.setdefault(m, MagicMock())
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    from sklearn.preprocessing import MinMaxScaler
    from b200pinn.synthetic import make_stack_data

    ref = load_reference()
    ref04 = load_script("ref04", "/root/reference/04_risk_function_early_warning_index.py.py")
    n_norm, n_f = 600, 250
    Xn, Un = make_stack_data(n_norm, seed=21)
    Xa, Ua = make_stack_data(n_f, seed=22)
    Xa[:, 6] *= np.linspace(1.0, 0.45, n_f)            # hydrogen flow collapses: H residual drifts
    Ua -= np.linspace(0.0, 0.25, n_f)[:, None]
    Xb, Ub = make_stack_data(n_f, seed=23)
    Xb[:, 5] += np.linspace(0.0, 9.0, n_f)             # outlet temperature runs away: T residual drifts
    Xb[:, 7] *= np.linspace(1.0, 0.5, n_f)
    sx, sy = MinMaxScaler(feature_range=(-1, 1)).fit(Xn), MinMaxScaler(feature_range=(-1, 1)).fit(Un)
    x_train, y_train = sx.transform(Xn).astype(np.float32), sy.transform(Un).astype(np.float32)
    x_test = sx.transform(np.vstack([Xn, Xa, Xb])).astype(np.float32)
    y_test = sy.transform(np.vstack([Un, Ua, Ub])).astype(np.float32)
    boundaries = [n_norm, n_norm + n_f, n_norm + 2 * n_f]
    info = {"boundary_lines": list(boundaries), "fault_data_list": [(None, None, "h2"), (None, None, "thermal")]}
    layers = [8, 64, 64, 64, 1]
    torch.manual_seed(0)
    model = ref.PhysicsInformedNN(torch.tensor(x_train), torch.tensor(y_train), layers, sx, sy, 0.2, True)
    with torch.no_grad():
        model.lambda_T1.fill_(0.012); model.lambda_T3.fill_(-3.0); model.lambda_T5.fill_(30.0)
        model.lambda_H1.fill_(1.4); model.lambda_H2.fill_(0.05)
        model.lambda_O1.fill_(2.4); model.lambda_O2.fill_(0.02)
    g = dict(layers=np.array(layers), p=0.2, x=x_train, y=y_train, x_test=x_test, y_test=y_test,
             boundaries=np.array(boundaries), **scaler_arrays("sx", sx), **scaler_arrays("sy", sy))
    g.update({"P:" + k: v.detach().numpy().copy() for k, v in model.dnn.state_dict().items()})
    g["lam0"] = np.array([getattr(model, n).item() for n in
                          (["lambda_1", "lambda_2", "lambda_3", "lambda_4"] + [f"lambda_T{i}" for i in range(1, 6)]
                           + [f"lambda_H{i}" for i in range(1, 5)] + [f"lambda_O{i}" for i in range(1, 5)])])
    T, p = 5, 0.4
    tap = MaskTap(model.dnn)
    torch.manual_seed(41)
    dataset = (torch.tensor(x_train), torch.tensor(y_train), torch.tensor(x_test), torch.tensor(y_test), sx, sy, info)
    res = ref.create_comprehensive_results_array_v2(model, dataset, mc_times=T, dropout=p)
    mm = tap.take()
    tap.close()
    nd = len(layers) - 2 + 1
    mc = [m for m in mm if m.shape[0] == x_test.shape[0]]
    assert len(mc) >= T * 2 * nd
    for t in range(T):
        g[f"mc_masks{t}"], _ = pack(mc[t * 2 * nd: t * 2 * nd + nd])
    g["mc_T"], g["mc_p"] = T, p
    g["results"] = res
    mu, sigma = ref04.estimate_mu_sigma_normal(res)
    rf_inst, rf_smooth, extra = ref04.compute_rf_time_series(res, mu, sigma)
    g["rf_mu"], g["rf_sigma"] = mu, sigma
    g["rf_inst"], g["rf_smooth"], g["rf_C"], g["rf_S"] = rf_inst, rf_smooth, extra["C"], extra["S_tot"]
    a = ref04.find_first_alarm_index(rf_smooth, ref04.RF_WARN_THRESHOLD)
    g["rf_alarm"] = -1 if a is None else a
    np.savez_compressed(os.path.join(HERE, "export64.npz"), **g)
    print("wrote export64", res.shape, "RF max", rf_smooth.max(), "alarm", g["rf_alarm"], "C max", extra["C"].max())


if __name__ == "__main__":
    main()
