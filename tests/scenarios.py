"""Constructed parity scenarios for the branches of the path that the SHEBA year and the testcases from `init` do
not enter (flood, flood_simple, flush4, the styropor factor, snow_thermo with snow_flush_flag 0, the
top_grow / top_melt sub-cases, bottom_melt_simple with a full grid, the warm branches of snow_coupling, sub_melt_snow
with all the snow flooded, flush_flag 4).

Each scenario prepares an ORACLE column (restart state + edits + flags), names the forcing it needs, the launch chunks
to advance, and the branch events (samsim_event_id, include/samsim_b200.h) that the run MUST execute.  The same list
drives tests/test_kernel_on_host_cpu.py (device code compiled for the host, no GPU) and tests/test_parity_gpu.py
(CUDA path through the C ABI): both compare every array / scalar / integer bit for bit with the oracle and assert
that the oracle's branch counters are > 0 AND that the device's event bits are set for the required events.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"

# event bits that only the device sets (which code path a step took), not branches of the reference
DEVICE_ONLY = {"two_pass_step"}

# events every SHEBA-flag column produces all the time; not interesting for "which branch ran" reports
BACKGROUND = {"turb", "getT_saltfree", "getT_liquid", "gas_refill", "scrub", "snow_precip", "snow_thermo_meltwater"}


def sheba_state(rec: int) -> dict:
    z = np.load(GOLDEN / "sheba_oracle_states.npz")
    p = f"state{rec}_"
    return {k[len(p):]: (np.array(z[k]) if z[k].ndim else z[k].item()) for k in z.files if k.startswith(p)}


def sheba_forcing() -> np.ndarray:
    return np.load(GOLDEN / "forcing_era.npz")["sheba"]


def lab_series(n: int, styropor_hours=(4, 5)) -> np.ndarray:
    """Synthetic per-second lab inputs (the reference's 2017_input files are not shipped): Tice, snowfall, heat, styropor."""
    t = np.arange(n, dtype=np.float64)
    Tice = -5.0 - 10.0 * (1.0 - np.cos(2.0 * np.pi * t / 86400.0))
    snow = np.where((t >= 2 * 3600) & (t < 3 * 3600), 1e-7, 0.0)
    heat = np.full(n, 5.0)
    sty = np.where((t >= styropor_hours[0] * 3600) & (t < styropor_hours[1] * 3600), 1.0, 0.0)
    return np.stack([Tice, snow, heat, sty])


@dataclass
class Scenario:
    name: str
    col: object                       # prepared oracle column (det build)
    chunks: tuple                     # launch sizes
    require: set                      # events that must have run on both sides
    forcing: np.ndarray | None = None  # [4, nrec] ERA-style series (atmoflux_flag 2)
    lab: np.ndarray | None = None      # [4, nrec] lab series
    cite: str = ""                    # reference lines the scenario is about
    notes: dict = field(default_factory=dict)


def _scale_snow(st: dict, f: float) -> dict:
    for k in ("m_snow", "thick_snow", "H_abs_snow"):
        st[k] = st[k] * f
    return st


def _scale_top_layer(st: dict, f: float) -> dict:
    # a thicker / thinner top layer of the same material: rho, S_bu, H of layer 1 unchanged
    for k in ("thick", "m", "S_abs", "H_abs"):
        a = np.array(st[k], dtype=np.float64)
        a[0] *= f
        st[k] = a
    return st


def _sheba(oracle_mod, st: dict, ints: dict | None = None, scalars: dict | None = None):
    F = sheba_forcing()
    col = oracle_mod.Column(4, "det")
    col.set_forcing(*F)
    col.load_state(st)
    for k, v in (ints or {}).items():
        col.set_int(k, v)
    for k, v in (scalars or {}).items():
        col.set_scalar(k, v)
    return col, F


TC7_FLAGS = dict(flood_flag=3, flush_flag=6, grav_flag=3, albedo_flag=1)  # mo_init.f90 testcase 7: the simple parametrisations

NAMES = [
    "flood_with_instant_flooding", "flood_simple", "flush4", "flush_flag4_inline", "snow_thermo_flag0",
    "top_grow_middle", "top_melt_middle", "top_grow_few_layers", "top_melt_few_layers", "bottom_melt_full_grid",
    "snow_coupling_warm1_melt_snow_all", "snow_coupling_warm2", "styropor", "salt_clamp",
    "flush3_heat_flag1", "flush3_snow_flush_flag0", "flush3_with_tracers",
]


def build(oracle_mod, name: str) -> Scenario:
    if name == "flood_with_instant_flooding":
        # mid-winter column under six times its snow load: freeboard < 0 every step, and beyond neg_free
        col, F = _sheba(oracle_mod, _scale_snow(sheba_state(200), 6.0))
        return Scenario(name, col, (1, 2, 120, 277), {"flood", "flood_neg_free", "top_grow_c"}, forcing=F,
                        cite="mo_flood.f90:55-153 incl. :117-138; mo_layer_dynamics.f90:679")
    if name == "flood_simple":
        col, F = _sheba(oracle_mod, _scale_snow(sheba_state(330), 5.0), ints=TC7_FLAGS)
        return Scenario(name, col, (1, 2, 597), {"flood_simple", "grav_drain_simple", "top_grow_c"}, forcing=F,
                        cite="mo_flood.f90:167-210, mo_grav_drain.f90:218")
    if name == "flush4":
        # bare melting ice (no snow: thick_snow < thick_0, mo_grotz.f90:729-733) with the testcase-7 flags
        col, F = _sheba(oracle_mod, sheba_state(360), ints=TC7_FLAGS)
        return Scenario(name, col, (1, 2, 797), {"flush4", "grav_drain_simple", "getT_Tfr_fallback"}, forcing=F,
                        cite="mo_flush.f90:253-296")
    if name == "flush_flag4_inline":
        col, F = _sheba(oracle_mod, sheba_state(345), ints=dict(flush_flag=4))
        return Scenario(name, col, (1, 2, 597), {"flush_inline", "snow_meltwater_to_ice", "snow_wet"}, forcing=F,
                        cite="mo_grotz.f90:704-713")
    if name == "flush3_heat_flag1":
        # bare-ice melt with the heat of the flushed brine staying in the lowest layer (flush_heat_flag 1: mo_flush.f90
        # :185-187 and :211-213 are skipped); SHEBA runs with flag 2
        col, F = _sheba(oracle_mod, sheba_state(345), ints=dict(flush_heat_flag=1))
        return Scenario(name, col, (1, 2, 597), {"flush3", "melt_thick"}, forcing=F, cite="mo_flush.f90:185-187, :211-213")
    if name == "flush3_with_tracers":
        # the melt season with two passive tracers: the general flush3 (tracer replay through the brine-flux matrix,
        # mo_flush.f90:166-178) and bgc_advection under flushing; the reference ships no such case (tracers: testcases 1, 2, 6)
        col, F = _sheba(oracle_mod, sheba_state(345), ints=dict(bgc_flag=2, N_bgc=2))
        m = col.array("m")
        for q, conc in ((1, 400.0), (2, 500.0)):
            col.set_array(f"bgc_abs{q}", m * conc)
            col.set_scalar(f"bgc_bottom{q}", conc)
            col.set_scalar(f"bgc_total{q}", float((m * conc).sum()))
        return Scenario(name, col, (1, 2, 597), {"flush3", "flush3_clamp", "grav_drained"}, forcing=F, cite="mo_flush.f90:166-178, mo_grotz.f90:742-747")
    if name == "flush3_snow_flush_flag0":
        # flush3 with the permeability of snow_flush_flag 0 (mo_flush.f90:114-130: inactive layers fully permeable)
        # (state 360: bare melting ice; with this flag a snow-covered column does not flush)
        col, F = _sheba(oracle_mod, sheba_state(360), ints=dict(snow_flush_flag=0))
        return Scenario(name, col, (1, 2, 597), {"flush3"}, forcing=F, cite="mo_flush.f90:114-130")
    if name == "snow_thermo_flag0":
        col, F = _sheba(oracle_mod, sheba_state(330), ints=dict(snow_flush_flag=0))
        return Scenario(name, col, (1, 2, 1497, 1500), {"snow_thermo", "snow_wet", "snow_compaction"}, forcing=F,
                        cite="mo_snow.f90:212-319")
    if name == "top_grow_middle":
        col, F = _sheba(oracle_mod, _scale_top_layer(sheba_state(80), 2.0))
        return Scenario(name, col, (1, 2, 47), {"top_grow_b"}, forcing=F, cite="mo_layer_dynamics.f90:665-677")
    if name == "top_melt_middle":
        col, F = _sheba(oracle_mod, _scale_top_layer(sheba_state(80), 0.3))
        return Scenario(name, col, (1, 2, 47), {"top_melt_b"}, forcing=F, cite="mo_layer_dynamics.f90:253-270")
    if name in ("top_grow_few_layers", "top_melt_few_layers"):
        # testcase 1 after 400 s: N_active = 4 <= N_top = 5
        c0 = oracle_mod.Column(1, "det")
        assert c0.step(400) == 0 and c0.int("N_active") <= c0.int("N_top")
        col = oracle_mod.Column(1, "det")
        col.load_state(_scale_top_layer(c0.state(), 2.0 if name == "top_grow_few_layers" else 0.3))
        ev = {"top_grow_a"} if name == "top_grow_few_layers" else {"top_melt_a"}
        return Scenario(name, col, (1, 2, 27), ev,
                        cite="mo_layer_dynamics.f90:656-663" if "grow" in name else "mo_layer_dynamics.f90:244-251")
    if name == "bottom_melt_full_grid":
        # a just-filled grid (middle layers ~ thick_0) over a hot ocean: bottom_melt, then bottom_melt_simple in both forms
        col, F = _sheba(oracle_mod, sheba_state(150), scalars=dict(oflux_amp=1000.0))
        return Scenario(name, col, (1, 2, 9997, 10000, 10000), {"bottom_melt", "bottom_melt_simple_a", "bottom_melt_simple_b"},
                        forcing=F, cite="mo_layer_dynamics.f90:85-112, :341-420, :573-590")
    if name in ("snow_coupling_warm1_melt_snow_all", "snow_coupling_warm2"):
        # thin ice in autumn whose top layer has been warmed above 0 degC under a thin cold snow cover
        st = sheba_state(80)
        warm1 = name.startswith("snow_coupling_warm1")
        m_snow, T1 = (0.5, 0.5) if warm1 else (0.03, 2.0)
        H = np.array(st["H_abs"], dtype=np.float64)
        H[0] = st["m"][0] * 3400.0 * T1
        st["H_abs"] = H
        st.update(thick_snow=m_snow / 330.0, m_snow=m_snow, H_abs_snow=-m_snow * 333500.0 - m_snow * 2020.0 * 2.0,
                  psi_s_snow=0.36, psi_g_snow=0.64, psi_l_snow=0.0, T_snow=-2.0)
        col, F = _sheba(oracle_mod, st)
        ev = {"snow_coupling_warm1", "snow_coupling_iter", "melt_snow_all"} if warm1 else {"snow_coupling_warm2"}
        return Scenario(name, col, (1, 2, 17), ev, forcing=F, cite="mo_snow.f90:76-85, mo_functions.f90:453-460")
    if name == "styropor":
        # the styropor plate goes on in hour 1-2, BEFORE the first snow (the factor applies only to bare ice,
        # mo_heat_fluxes.f90:202-222); snowfall in hour 2-3 as in the config-3 series
        n = 9000
        series = lab_series(n, styropor_hours=(1, 2))
        col = oracle_mod.Column(101, "det")
        col.set_lab_forcing(*series)
        return Scenario(name, col, (1, 2, 3598, 3, 5000), {"styropor", "tank"}, lab=series,
                        cite="mo_thermo_functions.f90:276-287 via mo_heat_fluxes.f90:202-222")
    if name == "salt_clamp":
        # Reachable only where nothing repairs a negative S_abs before S24: in ice, mass_transfer's MAX(fl*S_br, -S_abs)
        # (mo_mass.f90:83) zeroes it and fl_grav_drain STOPs 1337 on it.  Open water (N_active = 1, no expulsion,
        # no gravity drainage) keeps it until the health check clamps it.
        st = sheba_state(60)
        a = np.array(st["S_abs"], dtype=np.float64)
        a[0] = -1.0
        st["S_abs"] = a
        col, F = _sheba(oracle_mod, st)
        return Scenario(name, col, (1, 2, 17), {"salt_clamp", "snow_precip_0"}, forcing=F, cite="mo_grotz.f90:812-818")
    raise KeyError(name)


def run(sc: Scenario, dev, compare) -> list[str]:
    """Advance the oracle column and `dev` (HostKernel or api.Engine built from sc.col) chunk by chunk; returns the
    list of mismatches (bitwise state, status, and the required events on both sides)."""
    bad = []
    for n in sc.chunks:
        rc_o, rc_d = sc.col.step(n), dev.step(n)
        rc_d = 0 if rc_d is None else rc_d
        if rc_o != 0:
            bad.append(f"{sc.name}: oracle STOP {rc_o}")
            break
        bad += compare(sc.col, dev, 0, label=f"{sc.name} +{n}: ")
        if bad:
            break
    counts = sc.col.event_counts()
    dev_events = dev.events(0) - DEVICE_ONLY
    for ev in sorted(sc.require):
        if counts.get(ev, 0) <= 0:
            bad.append(f"{sc.name}: the oracle never executed '{ev}' ({sc.cite})")
        if ev not in dev_events:
            bad.append(f"{sc.name}: the device never executed '{ev}' ({sc.cite})")
    # every branch event must agree between the two sides, not only the required ones
    o_events = {k for k, v in counts.items() if v > 0}
    if o_events != dev_events:
        bad.append(f"{sc.name}: branch events differ: oracle-only {sorted(o_events - dev_events)}, device-only {sorted(dev_events - o_events)}")
    return bad
