"""API-surface parity with the reference source itself (build container only: skipped where
/root/reference is absent, e.g. on the GPU box)."""
import importlib.util
import inspect
import os
import sys
from unittest.mock import MagicMock

import pytest

REF = "/root/reference/01_train_pinn_multiphysics_model.py"
REF04 = "/root/reference/04_risk_function_early_warning_index.py.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference checkout not mounted")


def _load(name, path):
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.font_manager", "matplotlib.lines"):
        sys.modules.setdefault(m, MagicMock())
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def ref():
    return _load("ref01_surface", REF)


def params(fn):
    return [(p.name, p.default) for p in inspect.signature(fn).parameters.values()]


def test_physics_informed_nn_methods_and_signatures(ref):
    import b200pinn

    ours, theirs = b200pinn.PhysicsInformedNN, ref.PhysicsInformedNN
    ref_methods = {n for n, f in inspect.getmembers(theirs, inspect.isfunction)}
    our_methods = {n for n, f in inspect.getmembers(ours, inspect.isfunction)}
    assert ref_methods <= our_methods, ref_methods - our_methods
    for name in sorted(ref_methods):
        rp, op = params(getattr(theirs, name)), params(getattr(ours, name))
        assert op[:len(rp)] == rp, (name, rp, op)          # same leading parameters and defaults
        assert all(d is not inspect.Parameter.empty for _, d in op[len(rp):]), name   # extras are optional


def test_dnn_and_functions_signatures(ref):
    import b200pinn

    assert params(b200pinn.DNN.__init__) == params(ref.DNN.__init__)
    assert params(b200pinn.get_MC_samples) == params(ref.get_MC_samples)
    assert params(b200pinn.create_comprehensive_results_array_v2) == params(ref.create_comprehensive_results_array_v2)
    assert params(b200pinn.create_fault_labels) == params(ref.create_fault_labels)
    theirs, ours = ref.DNN(0.2, True, [8, 64, 64, 64, 1]), b200pinn.DNN(0.2, True, [8, 64, 64, 64, 1])
    assert list(theirs.state_dict().keys()) == list(ours.state_dict().keys())
    assert [tuple(v.shape) for v in theirs.state_dict().values()] == [tuple(v.shape) for v in ours.state_dict().values()]
    assert [n for n, _ in theirs.named_modules()] == [n for n, _ in ours.named_modules()]
    for attr in ("depth", "p", "logvar", "activation"):
        assert getattr(theirs, attr) == getattr(ours, attr)


def test_lambda_initial_values_match_reference_source(ref):
    """LAMBDA_INIT must equal what the reference constructor assigns (01:453-517)."""
    import torch
    from b200pinn.pinn import LAMBDA_INIT, LAMBDA_NAMES

    m = ref.PhysicsInformedNN(torch.zeros(4, 8), torch.zeros(4, 1), [8, 32, 32, 1], None, None, 0.1, True)
    for name, v in zip(LAMBDA_NAMES, LAMBDA_INIT):
        assert abs(getattr(m, name).item() - v) <= 1e-7 * max(1.0, abs(v)), name
    keys = [k for k in m.dnn.state_dict() if k.startswith("lambda")]
    assert keys == ["lambda_1", "lambda_2", "lambda_3"] + [f"lambda_T{i}" for i in range(1, 6)] + \
        [f"lambda_H{i}" for i in range(1, 5)] + [f"lambda_O{i}" for i in range(1, 5)]
    assert m.dnn.state_dict()["lambda_3"].item() == 1.0      # the key holds lambda_4 (01:468)


def test_rf_signatures_and_constants():
    from b200pinn import rf

    r4 = _load("ref04_surface", REF04)
    for name in ("estimate_mu_sigma_normal", "find_first_alarm_index", "compute_rf_time_series"):
        assert [p for p, _ in params(getattr(rf, name))] == [p for p, _ in params(getattr(r4, name))], name
    rp = dict(params(r4.compute_rf_time_series))
    op = dict(params(rf.compute_rf_time_series))
    for k in ("p_layer", "z_safe", "lambda_decay", "k_logistic", "C0_logistic", "C_max", "alpha_smooth"):
        assert op[k] == rp[k], k
    assert op["layer_config"] == rp["layer_config"] and op["layer_weights"] == rp["layer_weights"]
    assert tuple(op["res_keys"]) == tuple(rp["res_keys"]) and list(op["feature_weights"]) == list(rp["feature_weights"])
    with pytest.raises(NotImplementedError):
        rf.compute_rf_time_series(None, None, None, p_layer=3.0)
    assert rf.RF_WARN_THRESHOLD == r4.RF_WARN_THRESHOLD
