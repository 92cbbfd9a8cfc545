"""GPU parity for the downstream rows: the device export writer vs the
reference's create_comprehensive_results_array_v2, and the RF(t) kernels vs script 04."""
import numpy as np
import pytest
import torch

from conftest import load_golden, make_model, masks_u8, nrel
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda", 0)


def export_model(g):
    g2 = dict(g)
    g2["x"], g2["y"] = g["x"], g["y"]
    return make_model(g2)


def test_export_rows_golden():
    """22-column comprehensive_results through the public drop-in, reference masks injected."""
    import b200pinn

    g = load_golden("export64")
    m = export_model(g)
    T, p = int(g["mc_T"]), float(g["mc_p"])
    m.dnn._injected_mc = torch.tensor(np.stack([masks_u8(g[f"mc_masks{t}"], g["layers"]) for t in range(T)]), device=dev())
    bl = [int(b) for b in g["boundaries"]]
    info = {"boundary_lines": bl, "fault_data_list": [(None, None, "h2"), (None, None, "thermal")]}
    dataset = (torch.tensor(g["x"]), torch.tensor(g["y"]), torch.tensor(g["x_test"]), torch.tensor(g["y_test"]),
               g["sx"], g["sy"], info)
    out = b200pinn.create_comprehensive_results_array_v2(m, dataset, mc_times=T, dropout=p)
    ref = g["results"]
    assert out.shape == ref.shape and out.dtype == np.float64
    for c in range(22):
        assert nrel(out[:, c], ref[:, c]) < (2e-5 if c in (10, 11) else 1e-5), c
    assert np.array_equal(out[:, 17], ref[:, 17])
    assert np.array_equal(b200pinn.create_fault_labels(ref.shape[0], info), ref[:, 17])


def test_segment_smoothing_edges():
    """Window clipping at segment borders, a trailing segment, windows larger than a segment."""
    import b200pinn
    from b200pinn.export import export_rows_device

    g = load_golden("export64")
    m = export_model(g)
    n = 777
    x = torch.tensor(g["x_test"][:n], device=dev())
    y = torch.tensor(g["y_test"][:n], device=dev()).reshape(-1).contiguous()
    for bl, win in (([100, 130, 777], 200), ([5, 776, 777], 7), ([777], 1), ([300, 500], 50)):
        out = export_rows_device(m, x, y, bl, len(bl) - 1, 3, 0.3, g["sx"], g["sy"], seed=5, window=win).cpu().numpy()
        mc = b200pinn.mc_dropout_device(m.dnn, x, 3, 0.3, seed=5)
        den = (2.0 / (g["sy"].data_max_.astype(np.float64) - g["sy"].data_min_.astype(np.float64) + 1e-12)) + 1e-12
        ale = mc["a_u"].cpu().numpy().astype(np.float64) / den
        bounds = bl if bl[-1] == n else bl + [n]
        want = O.smooth_by_segments(ale, bounds, win)
        assert nrel(out[:, 10], want) < 1e-12, (bl, win)


def test_rf_golden():
    from b200pinn import rf

    g = load_golden("export64")
    res = g["results"]
    mu, sigma = rf.estimate_mu_sigma_normal(res)
    assert np.allclose(mu, g["rf_mu"], rtol=1e-12, atol=1e-15) and np.allclose(sigma, g["rf_sigma"], rtol=1e-12)
    inst, smooth, extra = rf.compute_rf_time_series(res, g["rf_mu"], g["rf_sigma"])
    assert np.allclose(extra["S_tot"], g["rf_S"], rtol=1e-12, atol=1e-12)
    assert np.allclose(extra["C"], g["rf_C"], rtol=1e-11, atol=1e-11)
    assert np.allclose(inst, g["rf_inst"], rtol=1e-11, atol=1e-13)
    assert np.allclose(smooth, g["rf_smooth"], rtol=1e-11, atol=1e-13)
    assert rf.find_first_alarm_index(smooth, 0.3) == int(g["rf_alarm"])
    o = rf.rf_device(torch.tensor(res[None], device=dev()))
    assert int(o["first_alarm"][0]) == int(g["rf_alarm"])


def test_rf_fleet_long_series_vs_oracle():
    """Many independent stacks, series spanning many scan chunks (carry propagation), NaNs in
    the normal block, one stack that never alarms."""
    from b200pinn import rf

    rng = np.random.default_rng(0)
    S, n = 5, 20011
    res = np.zeros((S, n, 22))
    res[:, :, 12:17] = rng.normal(size=(S, n, 5))
    res[:, 6000:, 17] = 1
    for s in range(1, S):                                   # stack 0 stays quiet
        drift = np.linspace(0, 3.0 + s, n - 9000)
        res[s, 9000:, 12 + (s % 5)] += drift
    res[2, 17, 13] = np.nan
    o = rf.rf_device(torch.tensor(res, device=dev()), want_extra=True)
    for s in range(S):
        mu, sigma = O.rf_mu_sigma(res[s])
        ms = o["mu_sigma"][s].cpu().numpy()
        assert np.allclose(ms[:5], mu, rtol=1e-12) and np.allclose(ms[5:], sigma, rtol=1e-12)
        rfi, sm, Sv, C = O.rf_series(np.nan_to_num(res[s]) if False else res[s], mu, sigma)
        ok = ~np.isnan(Sv)
        assert np.allclose(o["S_tot"][s].cpu().numpy()[ok], Sv[ok], rtol=1e-12, atol=1e-12)
        if s != 2:                                          # the NaN row poisons C downstream in numpy too
            assert np.allclose(o["C"][s].cpu().numpy(), C, rtol=1e-10, atol=1e-10)
            assert np.allclose(o["rf_smooth"][s].cpu().numpy(), sm, rtol=1e-10, atol=1e-12)
            assert int(o["first_alarm"][s]) == O.first_alarm(sm, 0.3)
    assert int(o["first_alarm"][0]) == -1


def test_rf_compact_columns_match_full_rows():
    """The row writer's dense [n, 6] copy of columns 12..17 equals those columns of the 22-column rows bit for bit, and RF(t) over
    the compact form (config 5's fleet pipeline) gives the same series as over the full rows."""
    from b200pinn import rf
    from b200pinn.export import export_rows_device

    g = load_golden("export64")
    m = export_model(g)
    n = 1100
    x = torch.tensor(g["x_test"][:n], device=dev())
    y = torch.tensor(g["y_test"][:n], device=dev()).reshape(-1).contiguous()
    bl = [int(b) for b in g["boundaries"]]
    rows, compact = export_rows_device(m, x, y, bl, len(bl) - 1, 4, 0.4, g["sx"], g["sy"], seed=3, want_rf_cols=True)
    assert compact.shape == (n, 6) and torch.equal(compact, rows[:, 12:18])
    a = rf.rf_device(rows.unsqueeze(0).contiguous(), want_extra=True)
    b = rf.rf_device(compact.unsqueeze(0).contiguous(), want_extra=True)
    for k in ("mu_sigma", "rf_inst", "rf_smooth", "first_alarm", "C", "S_tot"):
        assert torch.equal(a[k], b[k]), k
