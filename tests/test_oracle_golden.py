"""Pin the CPU oracle to the reference's own artefacts (runs on CPU, no GPU needed).

Golden vectors: reference_output/Reference_testcase1_with_Version_2 and Reference_SHEBA_with_Version_2, committed
as tests/golden/{tc1,sheba}_reference.npz by tools/make_fixtures.py.  The per-layer files are printed with F9.3
(F9.5 for thick, ES14.7 for perm/flush/melt) so agreement means |oracle - golden| <= half a unit of the last
printed digit; N_active (count of non-zero thick) must be exact; dat_T2m_T_top.dat carries T_top with 17 digits.

SHEBA finding (DESIGN.md "Oracle"): that golden run ("testing snow_precip change") was produced by an
intermediate revision of snow_precip (mo_snow.f90:147-148, `dt*T2m*solid_precip*rho_l*c_s` instead of the final
`min(T2m,-1._wp)`); the `*_shebagold` oracle builds revert that one line and reproduce the golden files through
output record 347 (one full year: freeze-up, winter, melt onset, snow melt and flushing).
"""
import os

import numpy as np
import pytest

HALF = {"T": 0.5e-3, "psi_s": 0.5e-3, "psi_l": 0.5e-3, "psi_g": 0.5e-3, "S_bu": 0.5e-3, "ray": 0.5e-3, "thick": 0.5e-5}
EPS = 2e-9  # slack for values that sit on a rounding boundary of the printed decimal


def _state(z, j):
    p = f"state{j}_"
    return {k[len(p):]: (z[k] if z[k].ndim else z[k].item()) for k in z.files if k.startswith(p)}


def _check_record(rec, gold, r, fields=HALF, label=""):
    for k, tol in fields.items():
        g = gold[k][r]
        m = np.asarray(rec[k])[: g.shape[0]]
        d = np.abs(m - g)
        # Half a unit of the last printed digit.  Isolated entries may sit up to one unit off: a getT
        # Newton-exit flip (|f| within rounding of 1 J/kg, mo_thermo_functions.f90:99) moves T by <= 1e-4 K, and the
        # skeletal bottom layers (S_bu changing by 10 g/kg per layer) amplify last-bit libm differences to ~3e-5
        # relative.  At most 5 % of a record's entries may exceed the half unit, none one unit + 1e-4 relative.
        # ray(k) is proportional to S_br(k) - S_br(N_active): the sensitive bottom layer enters every entry, so
        # ray gets a relative allowance on top of the print precision.
        lim = 2.0 * tol + 1e-4 * np.abs(g)
        soft = tol + EPS + (3e-4 * np.abs(g) if k == "ray" else 0.0)
        assert np.all(d <= lim + (soft - tol)), f"{label}{k}: max |diff| {d.max():.3e} at layer {int(d.argmax()) + 1}"
        assert (d > soft).mean() <= 0.05, f"{label}{k}: {(d > soft).sum()} entries beyond print precision"


def test_testcase1_full_run_matches_reference_output(oracle_mod, golden_dir):
    """Config 1: all 72 records x 90 layers of the reference's testcase-1 output."""
    gold = np.load(golden_dir / "tc1_reference.npz")
    bgc = np.load(golden_dir / "tc1_bgc_reference.npz")
    col = oracle_mod.Column(1, "libm")
    col.record_outputs()
    assert col.step(col.int("i_time")) == 0
    assert len(col.records) == 72
    for r, rec in enumerate(col.records):
        assert rec["N_active"] == gold["N_active"][r], f"record {r}"
        # passive tracers (bgc_flag 2): dat_bgc0{1,2}.{bu,br}.dat are printed with F16.8, so these four
        # files pin the brine fluxes of expulsion and gravity drainage and the layer shifts to 8 decimals
        for k in ("bgc1_bu", "bgc1_br", "bgc2_bu", "bgc2_br"):
            assert np.abs(rec[k] - bgc[k][r]).max() <= 0.5e-8 + 1e-9, (k, r)
        _check_record(rec, gold, r, label=f"record {r} ")
        vs = gold["vital_signs"][r]
        mine = [rec["energy_stored"], rec["freshwater"], rec["total_resist"], rec["thickness"], rec["bulk_salin"]]
        assert abs(mine[0] - vs[0]) <= 0.05 + 1e-6 and np.all(np.abs(np.array(mine[1:]) - vs[1:]) <= 0.5e-5 + EPS)
        gd = gold["grav_drain"][r]
        assert abs(rec["grav_drain"] - gd[0]) <= 0.5e-6 + EPS and abs(rec["grav_salt"] - gd[1]) <= 0.5e-5 + EPS
        assert abs(rec["grav_temp"] - gd[2]) <= 0.5e-3 + EPS
        assert abs(rec["freeboard"] - gold["freeboard"][r]) <= 0.5e-3 + EPS
    assert gold["N_active"].max() == 75


def test_det_and_libm_backends_agree(oracle_mod):
    """The deterministic math (shared with the GPU) and glibc differ by < 1 ulp per call: trajectories stay
    together to ~1e-10 over 20000 steps of testcase 1."""
    a, b = oracle_mod.Column(1, "libm"), oracle_mod.Column(1, "det")
    a.step(20000)
    b.step(20000)
    assert a.int("N_active") == b.int("N_active")
    for k in ("H_abs", "S_abs", "m", "thick", "T"):
        x, y = a.array(k), b.array(k)
        assert np.allclose(x, y, rtol=1e-9, atol=1e-12), k


def _sheba(oracle_mod, golden_dir, backend):
    F = np.load(golden_dir / "forcing_era.npz")["sheba"]
    col = oracle_mod.Column(4, backend)
    col.set_forcing(*F)
    return col


def test_sheba_first_days_from_init(oracle_mod, golden_dir):
    """Config 2 from the reference's initial state (open water, 1 July): records 0, 6, 12."""
    gold = np.load(golden_dir / "sheba_reference.npz")
    sel = {int(r): j for j, r in enumerate(gold["records"])}
    col = _sheba(oracle_mod, golden_dir, "libm_shebagold")
    col.record_outputs()
    period = col.int("i_time_out") + 1
    assert period == 8641
    assert col.step(12 * period + 1) == 0
    tt = gold["T2m_T_top"]
    for r in (0, 6, 12):
        rec = col.records[r]
        assert rec["N_active"] == gold["N_active"][r]
        g = {k: gold[k] for k in HALF}
        _check_record(rec, g, sel[r], label=f"record {r} ")
        assert rec["T2m"] == tt[r, 0]                      # forcing interpolation, all printed digits
        assert abs(rec["T_top"] - tt[r, 1]) <= 1e-9 * max(1.0, abs(tt[r, 1]))


@pytest.mark.parametrize("start,last", [(200, 205), (300, 346)])
def test_sheba_windows_match_reference_output(oracle_mod, golden_dir, start, last):
    """Restart from a committed oracle state and follow the golden run: mid winter (grid full), and the whole
    first melt onset incl. wet snow, melt-water flushing (ES14.7 files pin 8 digits) and top-layer dynamics."""
    gold = np.load(golden_dir / "sheba_reference.npz")
    sel = {int(r): j for j, r in enumerate(gold["records"])}
    z = np.load(golden_dir / "sheba_oracle_states.npz")
    col = _sheba(oracle_mod, golden_dir, "libm_shebagold")
    col.load_state(_state(z, start))   # state before the step that writes 1-based record `start`
    col.record_outputs()
    nrec = last - (start - 1) + 1
    assert col.step((nrec - 1) * 8641 + 1) == 0
    assert len(col.records) == nrec
    tt, melt = gold["T2m_T_top"], gold["melt"]
    checked = 0
    for q, rec in enumerate(col.records):
        r = start - 1 + q  # 0-based record index
        assert rec["N_active"] == gold["N_active"][r], f"record {r}"
        assert rec["T2m"] == tt[r, 0]
        assert abs(rec["T_top"] - tt[r, 1]) <= 1e-5, f"T_top record {r}: {rec['T_top']} vs {tt[r, 1]}"
        for j, name in enumerate(("melt_thick_output1", "melt_thick_output2", "melt_thick_output3")):
            assert abs(rec[name] - melt[r, j]) <= 1e-5 * abs(melt[r, j]) + 1e-12, f"{name} record {r}"
        if r in sel:
            g = {k: gold[k] for k in HALF}
            _check_record(rec, g, sel[r], label=f"record {r} ")
            for k in ("perm", "flush_v", "flush_h"):
                gg = gold[k][sel[r]]
                mm = np.asarray(rec[k])
                rel = np.full(gg.shape, 2e-4)   # 8 printed digits; the HEAD source is 5e-3 off here
                rel[-3:] = 2e-3                 # skeletal bottom layers: most sensitive entries of the column
                assert np.all(np.abs(mm - gg) <= rel * np.abs(gg) + 1e-30), f"{k} record {r}"
            checked += 1
    assert checked >= 1


@pytest.mark.skipif(os.environ.get("SAMSIM_SLOW") != "1", reason="3.0 M oracle steps (~3 min): set SAMSIM_SLOW=1; the result "
                    "of the full 14.2 M-step version is committed as profiles/r2_oracle_sheba_pin_*.json (tools/pin_oracle_sheba.py)")
def test_sheba_one_year_from_init_follows_the_golden_run(oracle_mod, golden_dir):
    """From the reference's initial state (open water, 1 July) through record 347 -- freeze-up, winter with the grid
    full, melt onset, wet snow, flushing -- without any restart state: N_active exact in all 347 records, T2m in all
    printed digits, T_top to 1e-6, melt / flushing totals to 1e-4 relative."""
    gold = np.load(golden_dir / "sheba_reference.npz")
    col = _sheba(oracle_mod, golden_dir, "libm_shebagold")
    col.record_outputs()
    assert col.step(346 * 8641 + 1) == 0 and len(col.records) == 347
    Na = np.array([r["N_active"] for r in col.records])
    assert np.array_equal(Na, gold["N_active"][:347])
    tt = gold["T2m_T_top"][:347]
    assert np.array_equal(np.array([r["T2m"] for r in col.records]), tt[:, 0])
    assert np.abs(np.array([r["T_top"] for r in col.records]) - tt[:, 1]).max() <= 1e-6
    melt = np.array([[r["melt_thick_output1"], r["melt_thick_output2"], r["melt_thick_output3"]] for r in col.records])
    assert (np.abs(melt - gold["melt"][:347]) <= 1e-4 * np.abs(gold["melt"][:347]) + 1e-12).all()


def test_sheba_head_source_differs_only_after_spring(oracle_mod, golden_dir):
    """The HEAD source (min(T2m,-1)) and the golden run agree to 1e-9 in T_top all winter; they separate when
    snow first falls at -1 C < T2m <= 0 C (spring).  Guards the attribution of the drift to that one line."""
    gold = np.load(golden_dir / "sheba_reference.npz")
    z = np.load(golden_dir / "sheba_oracle_states.npz")
    col = _sheba(oracle_mod, golden_dir, "libm")
    col.load_state(_state(z, 300))
    col.record_outputs()
    assert col.step(31 * 8641 + 1) == 0
    tt, melt = gold["T2m_T_top"], gold["melt"]
    d = [abs(rec["T_top"] - tt[299 + q, 1]) for q, rec in enumerate(col.records)]
    assert max(d[:25]) < 1e-8
    rel = abs(col.records[31]["melt_thick_output2"] - melt[330, 1]) / melt[330, 1]
    assert 1e-4 < rel < 2e-3   # 7.3e-4 with the HEAD source, 2.5e-7 with the shebagold build


def test_bare_ice_melt_onset_amplifies_one_ulp(oracle_mod, golden_dir):
    """Why the SHEBA golden run cannot be followed bit-level past record 347: from the SAME restart state the
    libm and det back-ends (pow/exp/sin equal to < 1 ulp) stay together to 1e-13 K for days, but the hours in
    which the last snow vanishes and bare-ice melt starts amplify that last-bit difference by ~1e9
    (melt_thick = MIN(psi_l*thick, ...) with psi_l(1) ~ 7e-4 obtained by cancellation, mo_functions.f90:398).
    The golden file differs from either back-end by the same 5e-5 ... 9e-5 there."""
    z = np.load(golden_dir / "sheba_oracle_states.npz")
    recs = {}
    for be in ("libm", "det"):
        col = _sheba(oracle_mod, golden_dir, be)
        col.load_state(_state(z, 345))
        col.record_outputs()
        assert col.step(4 * 8641 + 1) == 0
        recs[be] = col.records
    dT = [np.abs(np.asarray(a["T"]) - np.asarray(b["T"])).max() for a, b in zip(recs["libm"], recs["det"])]
    assert max(dT[:3]) < 1e-12                       # three days of identical trajectories
    a, b = recs["libm"][3]["melt_thick_output1"], recs["det"][3]["melt_thick_output1"]
    assert a > 0.005 and b > 0.005                    # bare-ice melt has started
    assert 1e-6 < abs(a - b) / a < 1e-3               # 8.9e-5: last-bit differences are now visible at 1e-4
    assert dT[3] > 1e-8
