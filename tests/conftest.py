import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


GOLDEN = os.path.join(ROOT, "tests", "golden")


class Scaler:
    """Duck-typed sklearn MinMaxScaler rebuilt from a golden file (the reference
    only touches these attributes: 01:542,1017-1020,1920-1925)."""

    def __init__(self, g, prefix):
        self.min_ = g[prefix + "_min"]
        self.scale_ = g[prefix + "_scale"]
        self.data_min_ = g[prefix + "_data_min"]
        self.data_max_ = g[prefix + "_data_max"]
        self.feature_range = (-1, 1)

    def inverse_transform(self, X):
        X = np.array(X, copy=True)
        X -= self.min_
        X /= self.scale_
        return X

    def transform(self, X):
        X = np.array(X, copy=True)
        X *= self.scale_
        X += self.min_
        return X


def unpack_masks(packed, layers, p, dtype=np.float32):
    """packed uint8 [N, ceil(D/8)] -> list of L+1 scaled masks ({0, 1/(1-p)})."""
    L, H = len(layers) - 2, int(layers[1])
    widths = [H] * L + [H // 2]
    bits = np.unpackbits(packed, axis=1)
    bits = bits[:, : sum(widths)] if bits.shape[1] >= sum(widths) else bits[:, : L * H]
    keep = dtype(1.0 - p)
    scale = dtype(1.0) / keep
    out, o = [], 0
    for w in widths:
        m = bits[:, o:o + w]
        if m.shape[1] < w:           # logvar=False goldens carry no variance-head mask (01:428-436 never runs that dropout)
            m = np.ones((bits.shape[0], w), np.uint8)
        out.append(m.astype(dtype) * scale)
        o += w
    return out


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    g["params"] = {k[2:]: v for k, v in g.items() if k.startswith("P:") and not k[2:].startswith("lambda")}
    g["layers"] = [int(v) for v in g["layers"]]
    g["p"] = float(g["p"])
    g["logvar"] = bool(g["logvar"]) if "logvar" in g else True
    g["sx"] = Scaler(g, "sx")
    g["sy"] = Scaler(g, "sy")
    return g


# net64 / net32: the headline and the narrow net; net256: the reference's own Layers (01:2139) on the wide tcgen05 path;
# net64nl: DNN(logvar=False) (01:436) -- all four recorded from the unmodified reference by tests/golden/make_golden.py
@pytest.fixture(params=["net64", "net32", "net256", "net64nl"])
def golden(request):
    return load_golden(request.param)


def nrel(a, b):
    """Norm-relative error max|a-b| / max|b| (element-wise relative error
    is meaningless on near-zero residuals)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def masks_u8(packed, layers):
    """packed golden bits -> uint8 keep matrix [N, L*H + H/2] (the C-ABI injection layout)."""
    L, H = len(layers) - 2, int(layers[1])
    D = L * H + H // 2
    bits = np.unpackbits(packed, axis=1)[:, :D]
    if packed.shape[1] * 8 < D:      # logvar=False goldens: the reference never ran the variance head's dropout (01:428-436)
        bits = np.concatenate([bits[:, :L * H], np.ones((bits.shape[0], D - L * H), np.uint8)], axis=1)
    return np.ascontiguousarray(bits)


def make_model(g, params_prefix="P:"):
    """Our PhysicsInformedNN on cuda:0 holding the golden file's data, weights and lambdas."""
    import torch
    import b200pinn

    model = b200pinn.PhysicsInformedNN(torch.tensor(g["x"]), torch.tensor(g["y"]), g["layers"], g["sx"], g["sy"],
                                       g["p"], g["logvar"])
    sd = {k[len(params_prefix):]: torch.tensor(v) for k, v in g.items()
          if k.startswith(params_prefix) and not k[len(params_prefix):].startswith("lambda")}
    missing, unexpected = model.dnn.load_state_dict(sd, strict=False)
    assert not unexpected and all(m.startswith("lambda") for m in missing)
    with torch.no_grad():
        for name, v in zip(b200pinn.LAMBDA_NAMES, g["lam0"]):
            getattr(model, name).fill_(float(v))
    return model
