"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol the header
declares, the Python constants agree with the header, host logic behaves."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200pinn.h")


def header_text():
    with open(HEADER) as f:
        return f.read()


def declared_functions():
    txt = re.sub(r"/\*.*?\*/", "", header_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(pinn_[a-z0-9_]+)\s*\(", txt)))


def enum_values(prefix):
    """Parse `enum { A = 1, B, C }` blocks of the header into {name: value}."""
    txt = re.sub(r"/\*.*?\*/", "", header_text(), flags=re.S)
    out = {}
    for body in re.findall(r"enum\s*\{(.*?)\}", txt, flags=re.S):
        val = -1
        for item in [i.strip() for i in body.split(",") if i.strip()]:
            if "=" in item:
                name, v = [t.strip() for t in item.split("=")]
                val = int(v, 0)
            else:
                name, val = item, val + 1
            if name.startswith(prefix):
                out[name] = val
    return out


def test_library_exports_every_declared_symbol():
    import b200pinn._abi as abi

    lib = ctypes.CDLL(abi.LIB_PATH)
    fns = declared_functions()
    assert len(fns) >= 12
    for fn in fns:
        assert hasattr(lib, fn), f"{fn} declared in b200pinn.h but not exported"
    assert sorted(abi.EXPORTS) == fns, "ctypes signature table out of sync with the header"
    assert abi.lib().pinn_abi_version() == abi.ABI_VERSION == int(re.search(r"#define PINN_ABI_VERSION (\d+)", header_text()).group(1))
    # the library keeps no process-global switches: path selection travels in pinn_net_t.flags / `flags`
    assert not [f for f in fns if f.startswith("pinn_set_")]


def test_python_constants_match_header():
    import b200pinn._abi as abi

    s = enum_values("PINN_S_")
    assert s.pop("PINN_S_COUNT") == abi.S_COUNT
    assert {k[len("PINN_S_"):]: v for k, v in s.items()} == abi.S
    c = enum_values("PINN_C_")
    assert c.pop("PINN_C_COUNT") == abi.C_COUNT
    assert {k[len("PINN_C_"):]: v for k, v in c.items()} == abi.COL
    r = enum_values("PINN_RES_")
    assert (r["PINN_RES_ACCURATE_MATH"], r["PINN_RES_NO_MODE_A"], r["PINN_RES_NO_MODE_B"], r["PINN_RES_NO_CLUSTER"]) == \
        (abi.RES_ACCURATE_MATH, abi.RES_NO_MODE_A, abi.RES_NO_MODE_B, abi.RES_NO_CLUSTER)
    nf = enum_values("PINN_NET_")
    assert nf == {"PINN_NET_NO_TC_FWD": abi.NET_NO_TC_FWD, "PINN_NET_NO_TC_BWD": abi.NET_NO_TC_BWD,
                  "PINN_NET_NO_WIDE_TC": abi.NET_NO_WIDE_TC, "PINN_NET_PDL_NEVER": abi.NET_PDL_NEVER,
                  "PINN_NET_PDL_ALWAYS": abi.NET_PDL_ALWAYS, "PINN_NET_NO_LOGVAR": abi.NET_NO_LOGVAR,
                  "PINN_NET_NO_FUSED_BWD": abi.NET_NO_FUSED_BWD, "PINN_NET_NO_WIDE_RESIDENT": abi.NET_NO_WIDE_RESIDENT,
                  "PINN_NET_NO_TMA_INPUT": abi.NET_NO_TMA_INPUT, "PINN_NET_NO_TC3": abi.NET_NO_TC3}
    f = enum_values("PINN_FAM_")
    assert (f["PINN_FAM_V"], f["PINN_FAM_TS"], f["PINN_FAM_T"], f["PINN_FAM_H"], f["PINN_FAM_O"],
            f["PINN_FAM_DATA"]) == (abi.FAM_V, abi.FAM_TS, abi.FAM_T, abi.FAM_H, abi.FAM_O, abi.FAM_DATA)


def test_struct_sizes():
    import b200pinn._abi as abi

    assert ctypes.sizeof(abi.PinnNet) == 16 + 8 * 8 * 2 + 8 * 8
    assert ctypes.sizeof(abi.PinnDropout) == 48
    assert ctypes.sizeof(abi.PinnScalers) == 4 * 22


@pytest.mark.parametrize("H,L,raw", [(64, 3, 11586), (256, 3, 175362), (256, 6, 372738)])
def test_param_layout_matches_reference_counts(H, L, raw):
    """SURVEY 8a [probe]: 3x64 -> 11 586, 3x256 -> 175 362, 6x256 -> 372 738 parameters."""
    from b200pinn import kernels as K

    names, shapes, offs, total = K.param_layout(H, L)
    assert sum(int(np.prod(s)) for s in shapes) == raw
    assert all(o % 4 == 0 for o in offs) and total >= raw
    assert names[0] == "layers.layer_0.weight" and names[-1] == "var_layers.5.bias"


def test_dnn_surface_matches_reference_state_dict():
    import b200pinn

    dnn = b200pinn.DNN(0.2, True, [8, 64, 64, 64, 1])
    keys = list(dnn.state_dict().keys())
    want = [f"layers.layer_{i}.{t}" for i in range(3) for t in ("weight", "bias")] + \
           ["predict.weight", "predict.bias"] + [f"var_layers.{i}.{t}" for i in (0, 3, 5) for t in ("weight", "bias")]
    assert keys == want
    drops = [n for n, m in dnn.named_modules() if isinstance(m, torch.nn.Dropout)]
    assert drops == ["layers.dropout_0", "layers.dropout_1", "layers.dropout_2", "var_layers.2"]
    assert dnn.depth == 4 and dnn.p == 0.2 and dnn.logvar is True
    dnn.eval()
    assert dnn.active_dropout_p() == 0.0
    dnn.train()
    for m in dnn.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.4                      # what get_MC_samples does (01:1449-1454)
    assert dnn.active_dropout_p() == 0.4


def test_no_cpu_fallback():
    import b200pinn

    dnn = b200pinn.DNN(0.2, True, [8, 32, 32, 1])
    with pytest.raises(RuntimeError, match="no CPU"):
        dnn(torch.zeros(4, 8))


def test_scalers_fold():
    from sklearn.preprocessing import MinMaxScaler
    from b200pinn import kernels as K
    from b200pinn.synthetic import make_stack_data
    from oracle import np_oracle as O

    X, U = make_stack_data(500, 3)
    sx, sy = MinMaxScaler((-1, 1)).fit(X), MinMaxScaler((-1, 1)).fit(U)
    s = K.make_scalers(sx, sy)
    xn = sx.transform(X).astype(np.float32)
    r = xn * np.array(list(s.x_inv_scale), np.float32) - np.array(list(s.x_off), np.float32)
    assert np.allclose(r, O.inverse_transform(sx, xn), rtol=3e-6, atol=1e-6)
    sc, mn = O.y_affine(sy)
    assert np.isclose(s.scale_y, sc[0], rtol=1e-7) and np.isclose(s.min_y, mn[0], rtol=1e-6)
    assert np.isclose(s.p_h2o, O.p_h2o(np.float32), rtol=1e-7)


def test_shard_range_partition():
    from b200pinn.dist import shard_range

    for n in (0, 1, 7, 1000003):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_chan_merge_equals_single_pass():
    from b200pinn.dist import chan_merge, finalize

    rng = np.random.default_rng(0)
    u = torch.tensor(rng.normal(size=(37, 50)))
    s = torch.tensor(rng.normal(size=(37, 50)))
    parts = []
    for lo, hi in ((0, 9), (9, 30), (30, 37)):
        uu, ss = u[lo:hi], s[lo:hi]
        parts.append((hi - lo, uu.mean(0), ((uu - uu.mean(0)) ** 2).sum(0), ss.sum(0)))
    acc = (0, None, None, None)
    for p in parts:
        acc = chan_merge(*acc, *p)
    a_u, e_u = finalize(acc[0], acc[2], acc[3])
    assert torch.allclose(acc[1], u.mean(0)) and torch.allclose(e_u, u.var(0, unbiased=False).sqrt())
    assert torch.allclose(a_u, torch.sqrt(torch.exp(s.mean(0))))


def test_abi_error_codes_without_gpu():
    """Argument / shape validation precedes any CUDA call, so the error contract is testable on CPU:
    negative PINN_E_* codes, readable messages, Python wrappers raising RuntimeError."""
    import ctypes as C
    import b200pinn._abi as abi

    L = abi.lib()
    e = enum_values("PINN_E_")
    assert e == {"PINN_E_ARG": -1, "PINN_E_SHAPE": -2, "PINN_E_WORKSPACE": -3, "PINN_E_ALIGN": -4}
    for code in e.values():
        assert L.pinn_error_string(code).decode().startswith("b200pinn:")
    assert L.pinn_error_string(0) == b"success"
    assert L.pinn_mlp_fwd(None, None, 0, None, None, None, None, 0, None) == e["PINN_E_ARG"]
    net = abi.PinnNet()
    net.n_in, net.width, net.n_hidden = 8, 48, 3                 # unsupported width
    assert L.pinn_mlp_fwd(C.byref(net), None, 0, None, None, None, None, 0, None) == e["PINN_E_SHAPE"]
    net.n_in, net.width, net.n_hidden = 7, 64, 3                 # wrong feature count
    assert L.pinn_mc_dropout(C.byref(net), None, 0, 1, None, None, None, None, None, None, None, None, 0, None) == e["PINN_E_SHAPE"]
    net.n_in, net.width, net.n_hidden = 8, 64, 9                 # too deep
    assert L.pinn_mlp_bwd(C.byref(net), None, 0, None, None, None, None, 0, None, None, None, 0, None) == e["PINN_E_SHAPE"]
    net.n_in, net.width, net.n_hidden = 8, 64, 3                 # right shape, null weight pointers
    assert L.pinn_mlp_fwd(C.byref(net), None, 0, None, None, None, None, 0, None) == e["PINN_E_ARG"]
    assert L.pinn_residuals(None, None, None, 5, None, None, 1, 0, None, None, None, None, None, 0, None) == e["PINN_E_ARG"]
    assert L.pinn_rf_series(None, 0, 1, 22, 12, None, None, None, None, None, None, None, None, 0, None) == e["PINN_E_ARG"]
    assert L.pinn_export_rows(None, None, None, None, None, None, None, 0, 0, 0, None, 5, None, None, None) == e["PINN_E_ARG"]
    assert L.pinn_param_count(64, 0) == e["PINN_E_SHAPE"]
    # the two whole-step entry points validate before touching the device as well
    assert L.pinn_train_dnn_step(None, None, 0, None, None, 0, None, None, None, None, 1e-2, 0.8, 1000, None, None, None, 0,
                                 None) == e["PINN_E_ARG"]
    net.n_in, net.width, net.n_hidden = 8, 48, 3
    assert L.pinn_train_dnn_step(C.byref(net), None, 0, None, None, 0, None, None, None, None, 1e-2, 0.8, 1000, None, None,
                                 None, 0, None) == e["PINN_E_SHAPE"]
    assert L.pinn_scalar_phase(None, None, None, 5, None, None, abi.FAM_TS, 0, 4, 5, None, None, None, None, None, None,
                               1.0, 0.8, 1000, 10, None, None, 0, None) == e["PINN_E_ARG"]
    with pytest.raises(RuntimeError, match="null or inconsistent"):
        abi.check(e["PINN_E_ARG"], "unit test")


def test_dropin_install_rebinds_reference_names():
    import types
    import b200pinn

    ref = types.ModuleType("ref01")
    ref.DNN = ref.get_MC_samples = ref.PhysicsInformedNN = ref.create_comprehensive_results_array_v2 = None
    b200pinn.install(ref, "A")
    assert ref.DNN is b200pinn.DNN and ref.get_MC_samples is None
    b200pinn.install(ref, "B")
    assert ref.get_MC_samples is b200pinn.get_MC_samples and ref.PhysicsInformedNN is None
    b200pinn.install(ref, "C")
    assert ref.PhysicsInformedNN is b200pinn.PhysicsInformedNN
    assert ref.create_comprehensive_results_array_v2 is b200pinn.create_comprehensive_results_array_v2
    with pytest.raises(ValueError):
        b200pinn.install(ref, "Z")


@pytest.mark.parametrize("n", [1, 37887, 37888, 65536, 100000, 300000, 1000000, 8000001])
def test_host_pipeline_chunks_cover_the_rows_in_whole_waves(n):
    """get_MC_samples' host pipeline (mc._pipeline_chunks): the chunks tile [0, n) without gaps, every chunk but the last
    is a whole number of waves, and long inputs start with a short chunk so that little of the first upload is exposed."""
    from b200pinn.mc import _pipeline_chunks

    wave = 256 * 148
    c = _pipeline_chunks(n, wave)
    assert c[0][0] == 0 and c[-1][1] == n and all(a[1] == b[0] for a, b in zip(c, c[1:]))
    assert all((hi - lo) % wave == 0 and hi > lo for lo, hi in c[:-1]) and c[-1][1] > c[-1][0]
    if n >= 9 * wave:
        assert c[0][1] - c[0][0] == 2 * wave and len(c) <= 5
