"""Seeded synthetic "normal-operation stack" data in the reference's column order.

The reference's raw data is private (Zenodo, ``README_DATA.docx``); its loader
``load_data_normal_raw`` (01:115-160) yields ``X[N,8] = [I, m_W, T_W_in, P_H_in,
P_O_in, T_W_out, m_H2, m_O2]`` and ``Y[N,1] = U`` after filtering ``50 < I < 800``
(01:143).  This module generates data of that shape:
polarisation staircases 0.1..1.5 A/cm^2 x 270 cm^2, 90 samples per level, the
reference's own voltage model (01:729-765, at the initial lambda values
01:453-455) plus noise for ``U``, and stoichiometry-consistent gas flows.

Scaling follows ``combine_and_normalize_datasets`` (01:271-282): sklearn
``MinMaxScaler(feature_range=(-1, 1))`` fitted on the same (normal) data, fp32.
"""
from __future__ import annotations

import numpy as np

A_CELL = 270.0
FARADAY = 96485.0
R_GAS = 8.314
N_CELLS = 5.0
LAMBDA_INIT = (0.167897923477715, 2.36682075851268e-06, 2.43414469188443)


def _stack_voltage(I, T_out, P_H, P_O, lam=LAMBDA_INIT):
    """Stack voltage 5*V_est of the reference's model (01:729-765), float64."""
    r, io, il = lam
    i = I / A_CELL + 1e-5
    Tk = T_out + 273.15
    P_H2 = P_H / 101.0 + 1.0
    P_air = P_O / 101.0 + 1.0
    Tc = 55.0
    P_H2O = 10.0 ** (-2.1794 + 0.02953 * Tc - 9.1837e-5 * Tc ** 2 + 1.4454e-7 * Tc ** 3)
    pp_H2 = 0.5 * (P_H2 / np.exp(1.653 * i / Tk ** 1.334) - P_H2O)
    pp_O2 = P_air / np.exp(4.192 * i / Tk ** 1.334) - P_H2O
    b = R_GAS * Tk / (2.0 * 0.5 * FARADAY)
    V_act = -b * np.log(i / io)
    V_ohm = -i * r
    V_conc = 0.5 * b * np.log(1.0 - i / il)
    E = 220170.0 / (2 * FARADAY) - R_GAS * Tk * np.log(P_H2O / (pp_H2 * np.sqrt(pp_O2))) / (2 * FARADAY)
    return N_CELLS * (E + V_act + V_ohm + V_conc)


def make_stack_data(n: int, seed: int = 1):
    """Return physical-domain ``(X[n,8] float64, U[n,1] float64)``."""
    rng = np.random.default_rng(seed)
    levels = np.arange(1, 16) * 0.1 * A_CELL           # 27 .. 405 A
    stair = np.repeat(levels, 90)
    stair = stair[stair > 50.0]                          # loader filter 01:143
    reps = -(-n // stair.size)
    I = np.tile(stair, reps)[:n] + rng.normal(0.0, 1.0, n)
    I = np.clip(I, 50.5, 500.0)                          # keep i < il
    m_W = rng.uniform(0.1, 0.5, n)
    T_in = rng.uniform(55.0, 65.0, n)
    T_out = T_in + rng.uniform(1.0, 8.0, n)
    P_H = rng.uniform(40.0, 120.0, n)
    P_O = rng.uniform(30.0, 110.0, n)
    q_h2 = I / (2 * FARADAY) * N_CELLS * 22.4 * 60.0
    q_o2 = I * N_CELLS / (4 * FARADAY) * 22.4 * 60.0
    m_H2 = 1.5 * q_h2 * (1.0 + rng.normal(0.0, 0.03, n))
    m_O2 = 2.5 * q_o2 / 0.21 * (1.0 + rng.normal(0.0, 0.03, n))
    U = _stack_voltage(I, T_out, P_H, P_O) + rng.normal(0.0, 0.01, n)
    X = np.stack([I, m_W, T_in, P_H, P_O, T_out, m_H2, m_O2], axis=1)
    return X, U[:, None]


def make_scaled_dataset(n: int, seed: int = 1):
    """``(x_norm f32 [n,8], y_norm f32 [n,1], scaler_X, scaler_Y)`` as 01:271-282."""
    from sklearn.preprocessing import MinMaxScaler

    X, U = make_stack_data(n, seed)
    sx = MinMaxScaler(feature_range=(-1, 1)).fit(X)
    sy = MinMaxScaler(feature_range=(-1, 1)).fit(U)
    return (sx.transform(X).astype(np.float32), sy.transform(U).astype(np.float32), sx, sy)
