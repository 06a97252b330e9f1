"""Helpers shared by the parity tests: drive the CPU oracle and the CUDA engine side by side."""
from __future__ import annotations

import numpy as np

from samsim_b200 import api

# arrays / scalars compared bit for bit between the oracle's "det" build and the GPU
CMP_ARRAYS = ["m", "S_abs", "H_abs", "thick", "T", "phi", "S_bu", "psi_s", "psi_l", "psi_g", "ray", "perm", "flush_v",
              "flush_h", "fl_Q"]
CMP_SCALARS = [s for s in api.SCALAR_IDS if s != "S_total"] + ["S_total"]
CMP_INTS = ["N_active", "status", "styropor_flag"]


def config_from_oracle(col) -> api.Config:
    st = {n: col.int(n) for n in ["testcase", "Nlayer", "N_top", "N_middle", "N_bottom", "atmoflux_flag", "grav_flag",
                                   "prescribe_flag", "grav_heat_flag", "flush_heat_flag", "turb_flag", "salt_flag",
                                   "boundflux_flag", "flush_flag", "flood_flag", "bottom_flag", "precip_flag",
                                   "harmonic_flag", "tank_flag", "albedo_flag", "lab_snow_flag", "freeboard_snow_flag",
                                   "snow_flush_flag", "snow_precip_flag", "i_time_out", "bgc_flag", "N_bgc"]}
    st.update({n: col.scalar(n) for n in ["dt", "thick_0", "thick_min", "time_out", "alpha_flux_instable",
                                          "alpha_flux_stable", "m_total", "max_flux_plate", "k_snow_flush", "k_styropor"]})
    return api.Config.from_state(st)


def same_bits(a, b) -> np.ndarray:
    """elementwise: identical doubles, treating +0/-0 and NaN/NaN as equal"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return (a == b) | (np.isnan(a) & np.isnan(b))


def compare_column(oracle_col, eng, col: int = 0, rtol: float = 0.0, label: str = "") -> list[str]:
    """Return a list of human-readable mismatches between an oracle column and engine column `col`."""
    bad = []
    ost = oracle_col.state()
    for n in CMP_INTS:
        g = int(eng.get_int(n, col, 1)[0])
        if g != int(ost[n]):
            bad.append(f"{label}{n}: oracle {ost[n]} gpu {g}")
    for n in CMP_ARRAYS + [f"bgc_abs{t}" for t in range(1, eng.cfg.N_bgc + 1)]:
        g = eng.get_array(n, col, 1)[0]
        o = np.asarray(ost[n])[: len(g)]
        ok = same_bits(o, g) if rtol == 0.0 else np.isclose(o, g, rtol=rtol, atol=0.0) | same_bits(o, g)
        if not ok.all():
            k = int(np.argmin(ok))
            rel = abs(o[k] - g[k]) / max(abs(o[k]), 1e-300)
            bad.append(f"{label}{n}[{k + 1}]: oracle {o[k]!r} gpu {g[k]!r} rel {rel:.3e} ({(~ok).sum()} of {len(ok)} differ)")
    for n in CMP_SCALARS:
        g = float(eng.get_scalar(n, col, 1)[0])
        o = float(ost[n])
        ok = same_bits(o, g) if rtol == 0.0 else (np.isclose(o, g, rtol=rtol, atol=0.0) | same_bits(o, g))
        if not ok:
            bad.append(f"{label}{n}: oracle {o!r} gpu {g!r}")
    clk = eng.get_clock()
    for n in ["time", "i", "n_time_out"] + (["time_counter"] if oracle_col.int("atmoflux_flag") == 2 else []):
        if clk[n] != ost[n]:
            bad.append(f"{label}clock {n}: oracle {ost[n]} gpu {clk[n]}")
    return bad


def engine_from_oracle(oracle_col, ncol: int = 1, device: int = 0) -> api.Engine:
    """Create an engine whose every column equals the oracle column's current state."""
    cfg = config_from_oracle(oracle_col)
    eng = api.Engine(cfg, ncol, device)
    eng.load_column_state(oracle_col.state(), 0)
    if ncol > 1:
        eng.broadcast_column(0, 0, ncol)
    return eng
