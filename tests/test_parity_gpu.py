"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle's deterministic-math build.

Bar: integers (N_active, status) exact; FP64 state BIT-IDENTICAL (tolerance 0.0) -- possible because
both sides use IEEE +,-,*,/ in the reference's operation order (gcc -ffp-contract=off / nvcc
-fmad=false) and the same detmath.h for pow/exp/sin.
"""
import numpy as np
import pytest

import sys
from pathlib import Path

from samsim_b200 import api
from oracle import parity_util as pu

sys.path.insert(0, str(Path(__file__).resolve().parent))
import scenarios  # noqa: E402

pytestmark = pytest.mark.gpu


def _fmt(bad, n=12):
    return "\n".join(bad[:n]) + (f"\n... {len(bad)} mismatches" if len(bad) > n else "")


def test_detmath_bitexact_on_gpu(oracle_mod):
    L = oracle_mod.lib("det")
    rng = np.random.default_rng(1)
    x = np.concatenate([10 ** rng.uniform(-8, 3.5, 4000), [0.0, 1.0, 1000.0, 5e-324, 1e-310]])
    for y in (3.1, 1.5, 0.37):
        g = api.kat_scalar(7, 1, x, np.full_like(x, y))
        o = np.array([L.sam_math_pow(float(a), y) for a in x])
        assert pu.same_bits(g, o).all()
    e = np.concatenate([rng.uniform(-40, 10, 4000), [-745.2, -744.0, -708.5, 0.0, 709.0]])
    assert pu.same_bits(api.kat_scalar(8, 1, e), np.array([L.sam_math_exp(float(a)) for a in e])).all()
    s = rng.uniform(-50, 50, 4000)
    assert pu.same_bits(api.kat_scalar(9, 1, s), np.array([L.sam_math_sin(float(a)) for a in s])).all()


@pytest.mark.parametrize("salt_flag", [1, 2])
def test_getT_kat(oracle_mod, salt_flag):
    """getT (mo_thermo_functions.f90:62-143): mushy, liquid, salt-free and pathological inputs."""
    import ctypes as C
    L = oracle_mod.lib("det")
    rng = np.random.default_rng(2)
    n = 20000
    S = np.concatenate([rng.uniform(0.5, 40, n // 2), rng.uniform(0, 0.002, n // 4), rng.uniform(30, 250, n // 4)])
    T_true = rng.uniform(-45, 2, n)
    H = np.where(rng.random(n) < 0.8, -333500.0 * rng.random(n) ** 0.5 + 2020.0 * T_true, 3400.0 * T_true)
    T_in = T_true + rng.normal(0, 0.3, n)
    T_in[::7] = rng.uniform(0.5, 30.0, len(T_in[::7]))        # first guesses outside [-200, 0]: the iterate restarts at
    T_in[3::11] = rng.uniform(-400.0, -210.0, len(T_in[3::11]))  # the (lazily evaluated) freezing point, :101-103
    Tg, pg, st, ev = api.kat_getT(salt_flag, H, S, T_in)
    To, po = np.empty(n), np.empty(n)
    dp = C.POINTER(C.c_double)
    L.sam_kat_getT(salt_flag, n, H.ctypes.data_as(dp), S.ctypes.data_as(dp), T_in.ctypes.data_as(dp), To.ctypes.data_as(dp), po.ctypes.data_as(dp))
    ok = st == 0
    assert ok.sum() > 0.9 * n
    assert pu.same_bits(Tg[ok], To[ok]).all()
    assert pu.same_bits(pg[ok], po[ok]).all()
    assert np.isnan(To[~ok]).all()  # the oracle STOPs (99) exactly where the GPU flags status 99
    fallback, saltfree, liquid = (ev >> 1) & 1, (ev >> 2) & 1, (ev >> 3) & 1  # bit SAMSIM_EV_GETT_* - 32
    assert fallback.sum() > 500 and saltfree.sum() > 1000 and liquid.sum() > 100, (fallback.sum(), saltfree.sum(), liquid.sum())


@pytest.mark.parametrize("fn,two", [(0, False), (1, True), (2, False), (3, True), (4, False), (5, True), (6, True)])
def test_scalar_kats(oracle_mod, fn, two):
    import ctypes as C
    L = oracle_mod.lib("det")
    rng = np.random.default_rng(3 + fn)
    n = 5000
    a = rng.uniform(-40, 5, n) if fn in (0, 1, 2, 3) else rng.uniform(0.0, 60.0, n)
    b = rng.uniform(0, 40, n)
    if fn == 5:
        a, b = rng.uniform(1, 200, n), rng.uniform(0.005, 0.6, n)
    if fn == 6:
        a, b = rng.uniform(0, 0.5, n), rng.uniform(-5, 0.5, n)
    dp = C.POINTER(C.c_double)
    for salt in (1, 2):
        o = np.empty(n)
        L.sam_kat_scalar(fn, salt, n, a.ctypes.data_as(dp), b.ctypes.data_as(dp), o.ctypes.data_as(dp))
        g = api.kat_scalar(fn, salt, a, b if two else None) if two else api.kat_scalar(fn, salt, a)
        assert pu.same_bits(g, o).all(), (fn, salt)


def test_testcase1_from_init(oracle_mod):
    """Config 1: testcase 1 from the reference's initial state; bitwise at growing step counts."""
    col = oracle_mod.Column(1, "det")
    eng = pu.engine_from_oracle(col, ncol=3)
    done = 0
    for target in (1, 2, 10, 100, 3000, 3601, 3603, 20000, 45000):
        n = target - done
        assert col.step(n) == 0
        eng.step(n)
        done = target
        for c in (0, 2):
            bad = pu.compare_column(col, eng, c, label=f"step {target} col {c}: ")
            assert not bad, _fmt(bad)
    assert col.int("N_active") > 20  # the run really grew ice and exercised bottom_growth_simple


def _forcing(golden_dir, site="sheba"):
    z = np.load(golden_dir / "forcing_era.npz")
    return z[site]  # [4, nrec] fl_sw, fl_lw, T2m, precip


def test_testcase4_from_init(oracle_mod, golden_dir):
    """Config 2 (SHEBA): open water -> freeze-up with reanalysis forcing, first ~52 h."""
    F = _forcing(golden_dir)
    col = oracle_mod.Column(4, "det")
    col.set_forcing(*F)
    eng = pu.engine_from_oracle(col, ncol=2)
    eng.set_forcing(F[None])
    done = 0
    for target in (1, 2, 1081, 1082, 8641, 8643, 19000):
        n = target - done
        assert col.step(n) == 0
        eng.step(n)
        done = target
        bad = pu.compare_column(col, eng, 1, label=f"step {target}: ")
        assert not bad, _fmt(bad)


def _state(z, j):
    p = f"state{j}_"
    return {k[len(p):]: (z[k] if z[k].ndim else z[k].item()) for k in z.files if k.startswith(p)}


# 1-based output record whose preceding state is restored -> regime exercised in the window
SHEBA_WINDOWS = {
    60: "autumn freeze-up, N_active growing (bottom_growth_simple), thin snow coupling",
    100: "early winter, snow cover, gravity drainage",
    200: "mid winter, N_active = 100, bottom_growth (grid full)",
    330: "melt onset: wet snow, melt water flushing (flush3)",
    345: "snow gone, surface melt, top_melt / layer merges",
    400: "late summer melt, bottom_melt",
    715: "second summer",
    1000: "third winter",
}


@pytest.mark.parametrize("rec,two_pass", [(r, False) for r in sorted(SHEBA_WINDOWS)] + [(r, True) for r in (100, 200, 345, 1000)])
def test_sheba_windows(oracle_mod, golden_dir, rec, two_pass):
    """Restart both implementations from an oracle state of the SHEBA run and advance 2 days (general path; and with
    the two-pass step switched on for the winter, growth and melt windows)."""
    z = np.load(golden_dir / "sheba_oracle_states.npz")
    st = _state(z, rec)
    F = _forcing(golden_dir)
    col = oracle_mod.Column(4, "det")
    col.set_forcing(*F)
    col.load_state(st)
    eng = pu.engine_from_oracle(col, ncol=2)
    eng.set_tuning(two_pass)
    eng.set_forcing(F[None])
    ev0 = {k: col.stat(k) for k in ("layer_events", "flush_calls", "coupling_iters")}
    for n in (1, 999, 8641, 7641):
        assert col.step(n) == 0
        eng.step(n)
        bad = pu.compare_column(col, eng, 1, label=f"rec {rec} (+{n}): ")
        assert not bad, SHEBA_WINDOWS[rec] + "\n" + _fmt(bad)
    print(rec, SHEBA_WINDOWS[rec], {k: col.stat(k) - ev0[k] for k in ev0})
    assert ("two_pass_step" in eng.events(1)) == two_pass


def test_snapshot_matches_oracle_output(oracle_mod, golden_dir):
    """S8 capture (mo_grotz.f90:340-398): the device snapshot equals what the oracle hands to output()."""
    z = np.load(golden_dir / "sheba_oracle_states.npz")
    F = _forcing(golden_dir)
    col = oracle_mod.Column(4, "det")
    col.set_forcing(*F)
    col.load_state(_state(z, 200))
    col.record_outputs()
    eng = pu.engine_from_oracle(col, ncol=2)
    eng.set_forcing(F[None])
    eng.set_snapshot_mode(api.SNAP_FULL)
    for _ in range(2):
        n = eng.steps_to_next_output()
        assert col.step(n) == 0
        eng.step(n)
        rec = col.records[-1]
        snap = eng.get_snapshot(1, 1)
        for name in api.SNAP_SCALARS:
            o = rec[name] if name != "N_active" else float(rec["N_active"])
            assert pu.same_bits(snap[name][0], o), (name, snap[name][0], o)
        for name in api.SNAP_ARRAYS:
            if name.startswith("bgc"):
                continue  # tracer rows: test_tracers_tank_and_snapshot (testcase 4 runs without tracers)
            o = np.asarray(rec[name])[: snap[name].shape[1]]
            assert pu.same_bits(snap[name][0], o).all(), name


def test_batch_invariance_and_perturbed_ensemble(oracle_mod, golden_dir):
    """Column results do not depend on batch size or neighbours; per-column forcing perturbations
    (value*scale + offset) match an oracle fed the perturbed series."""
    z = np.load(golden_dir / "sheba_oracle_states.npz")
    F = _forcing(golden_dir)
    ncol = 70  # not a multiple of the block size
    rng = np.random.default_rng(4)
    scale = np.ones((4, ncol))
    offset = np.zeros((4, ncol))
    offset[2] = rng.uniform(-2, 2, ncol)       # T2m + U(-2,2)
    scale[1] = rng.uniform(0.95, 1.05, ncol)   # fl_lw
    scale[0] = rng.uniform(0.9, 1.1, ncol)     # fl_sw
    scale[3] = rng.uniform(0.5, 1.5, ncol)     # precip
    amp = 7.0 * rng.uniform(0.5, 1.5, ncol)
    base = oracle_mod.Column(4, "det")
    base.set_forcing(*F)
    base.load_state(_state(z, 100))
    eng = pu.engine_from_oracle(base, ncol=ncol)
    eng.set_forcing(F[None], None, scale, offset)
    eng.set_scalar("oflux_amp", amp)
    eng.step(3000)
    for c in (0, 1, 33, 69):
        col = oracle_mod.Column(4, "det")
        col.set_forcing(*[F[k] * scale[k, c] + offset[k, c] for k in range(4)])
        col.load_state(_state(z, 100))
        col.set_scalar("oflux_amp", amp[c])
        assert col.step(3000) == 0
        bad = pu.compare_column(col, eng, c, label=f"col {c}: ")
        assert not bad, _fmt(bad)
    assert eng.count_failed() == 0


def test_status_codes_freeze_column(oracle_mod):
    """A column that hits a reference STOP is frozen with that code; its neighbours continue."""
    col = oracle_mod.Column(1, "det")
    col.step(5000)
    eng = pu.engine_from_oracle(col, ncol=4)
    bad_state = col.state()
    H = np.array(bad_state["H_abs"])
    H[2] = -1e12  # absurd enthalpy: getT cannot converge -> STOP 99 (mo_thermo_functions.f90:114-123)
    eng.set_array("H_abs", H[None, :], col0=2)
    ref = oracle_mod.Column(1, "det")
    ref.load_state(col.state())
    ref.set_array("H_abs", H)
    rc = ref.step(10)
    eng.step(10)
    st = eng.status()
    assert rc != 0 and st[2] == rc and (st[[0, 1, 3]] == 0).all()
    assert col.step(10) == 0
    assert not pu.compare_column(col, eng, 0)
    assert eng.count_failed() == 1


def test_fp64_peak_runs():
    tf = api.fp64_peak(0, 0.3)
    assert 5.0 < tf < 80.0, tf


def _lab_series(n):
    """Config 3 stand-ins for the absent 2017_input/ files: per-second series."""
    t = np.arange(n, dtype=np.float64)
    Tice = -5.0 - 10.0 * (1.0 - np.cos(2.0 * np.pi * t / 86400.0))
    snowfall = np.where((t >= 2.0 * 3600) & (t < 3.0 * 3600), 1e-7, 0.0)   # an early snow event so the window sees it
    heat = np.full(n, 5.0)
    styropor = np.where((t >= 4.0 * 3600) & (t < 5.0 * 3600), 1.0, 0.0)
    return np.stack([Tice, snowfall, heat, styropor])


def test_lab_tank_batch_of_five(oracle_mod):
    """Config 3: testcases 101-105 (Nlayer 200, boundflux 3, tank salinity feedback, lab snow) as ONE batch of five
    columns with per-column S_bu_bottom / S_total and per-column forcing sets; GPU vs oracle, bitwise."""
    n = 30000
    base = _lab_series(n)
    sets = np.stack([base * np.array([1.0 + 0.03 * q, 1.0 + 0.5 * q, 1.0, 1.0])[:, None] for q in range(5)])
    cols = []
    for q in range(5):
        col = oracle_mod.Column(101 + q, "det")
        col.set_lab_forcing(*sets[q])
        cols.append(col)
    cfg = pu.config_from_oracle(cols[0])
    eng = api.Engine(cfg, 5)
    for q, col in enumerate(cols):
        eng.load_column_state(col.state(), q, set_clock=(q == 0))
        eng.set_scalar("S_total", [col.scalar("S_total")], col0=q)
    eng.set_lab_forcing(sets, np.arange(5, dtype=np.int32))
    done = 0
    for target in (1, 2, 3601, 3602, 9000, 20000):
        nstep = target - done
        eng.step(nstep)
        for q, col in enumerate(cols):
            assert col.step(nstep) == 0
        done = target
        for q, col in enumerate(cols):
            bad = pu.compare_column(col, eng, q, label=f"testcase {101 + q} step {target}: ")
            assert not bad, _fmt(bad)
    assert max(c.int("N_active") for c in cols) >= 3 and max(c.stat("coupling_iters") for c in cols) > 1000  # thin lab snow: iterative snow_coupling
    for q, col in enumerate(cols):  # the branches this batch is credited with ran, on both sides
        o, g = col.events(), eng.events(q)
        assert {"snow_coupling_iter", "tank", "heat_thin_snow"} <= o and o == g - scenarios.DEVICE_ONLY, (q, sorted(o ^ g))
        assert "styropor" not in o  # the plate sits on snow in this series: tests/scenarios.py 'styropor' covers the factor


def test_testcase1_perturbed_ensemble(oracle_mod):
    """Config 4 in small: members differ in the T_top levels, T_bottom, S_bu_bottom and fl_q_bottom."""
    ncol = 40
    rng = np.random.default_rng(20170301)
    d1, d2 = rng.normal(0, 0.5, ncol), rng.normal(0, 0.5, ncol)
    Sb = 34.0 + rng.normal(0, 0.5, ncol)
    Tb = np.maximum(-1.0 + rng.normal(0, 0.05, ncol), -1.7)
    fq = np.abs(rng.normal(0, 2.0, ncol))
    base = oracle_mod.Column(1, "det")
    eng = pu.engine_from_oracle(base, ncol=ncol)
    rho_l, c_l = 1028.0, 3400.0
    m1 = base.array("m")[0]

    def personalise(col, j):
        col.set_scalar("ttop_warm", -5.0 + d1[j]); col.set_scalar("ttop_cold", -10.0 + d2[j])
        col.set_scalar("T_top", -5.0 + d1[j]); col.set_scalar("T_bottom", Tb[j]); col.set_scalar("S_bu_bottom", Sb[j])
        col.set_scalar("fl_q_bottom", fq[j])
        S = col.array("S_abs"); S[0] = Sb[j] * m1; col.set_array("S_abs", S)
        H = col.array("H_abs"); H[0] = m1 * Tb[j] * c_l; col.set_array("H_abs", H)
        T = col.array("T"); T[:] = Tb[j]; col.set_array("T", T)
        Sbu = col.array("S_bu"); Sbu[:] = Sb[j]; col.set_array("S_bu", Sbu)

    eng.set_scalar("ttop_warm", -5.0 + d1); eng.set_scalar("ttop_cold", -10.0 + d2); eng.set_scalar("T_top", -5.0 + d1)
    eng.set_scalar("T_bottom", Tb); eng.set_scalar("S_bu_bottom", Sb); eng.set_scalar("fl_q_bottom", fq)
    S = np.tile(base.array("S_abs"), (ncol, 1)); S[:, 0] = Sb * m1; eng.set_array("S_abs", S)
    H = np.tile(base.array("H_abs"), (ncol, 1)); H[:, 0] = m1 * Tb * c_l; eng.set_array("H_abs", H)
    eng.set_array("T", np.repeat(Tb[:, None], 90, 1)); eng.set_array("S_bu", np.repeat(Sb[:, None], 90, 1))
    nsteps = 46000   # passes the first T_top switch at t = 12 h
    eng.step(nsteps)
    for j in (0, 7, 39):
        col = oracle_mod.Column(1, "det")
        personalise(col, j)
        assert col.step(nsteps) == 0
        bad = pu.compare_column(col, eng, j, label=f"member {j}: ")
        assert not bad, _fmt(bad)
    na = eng.get_int("N_active")
    assert na.min() >= 2 and len(np.unique(na)) > 1   # members really diverged


def _read_dat(path):
    return np.loadtxt(path, ndmin=2)


def test_grotz_testcase1_output_files_match_reference_output(tmp_path, golden_dir):
    """Drop-in check end to end: the C++ host `grotz(1, ...)` runs the whole testcase on the GPU and writes
    dat_*.dat; those FILES are compared with reference_output/Reference_testcase1_with_Version_2 at the print
    precision of the Fortran formats (F9.3 / F9.5), N_active exact in all 72 records."""
    from samsim_b200 import grotz
    rc = grotz.grotz(1, "parity run", output_dir=tmp_path)
    assert rc == 0
    gold = np.load(golden_dir / "tc1_reference.npz")
    thick = _read_dat(tmp_path / "dat_thick.dat")
    assert thick.shape == (72, 90)
    assert np.array_equal((thick != 0).sum(1), gold["N_active"])
    for name, tol in (("T", 1.0e-3), ("psi_s", 1.0e-3), ("psi_l", 1.0e-3), ("psi_g", 1.0e-3), ("S_bu", 1.0e-3), ("ray", 1.0e-3), ("thick", 1.0e-5)):
        mine = _read_dat(tmp_path / f"dat_{name}.dat")
        g = gold[name]
        assert mine.shape == g.shape, name
        d = np.abs(mine - g)
        # both sides are rounded to the last printed digit: they may differ by one unit of it on a rounding boundary
        assert d.max() <= tol + 1e-12, (name, d.max())
        assert (d > 1e-12).mean() < 0.02, (name, (d > 1e-12).mean())
    vs = _read_dat(tmp_path / "dat_vital_signs.dat")
    assert np.abs(vs[:, 3] - gold["vital_signs"][:, 3]).max() <= 1.0e-5 + 1e-12       # thickness
    assert np.abs(_read_dat(tmp_path / "dat_freeboard.dat")[:, 0] - gold["freeboard"]).max() <= 1.0e-3 + 1e-12
    # passive tracers (bgc_flag 2): the four dat_bgc files are printed with F16.8 -- the GPU run (deterministic
    # pow/exp, not glibc's) stays within 2e-7 of the golden files over all 259200 steps
    bgc = np.load(golden_dir / "tc1_bgc_reference.npz")
    for t in (1, 2):
        for kind in ("bu", "br"):
            mine = _read_dat(tmp_path / f"dat_bgc0{t}.{kind}.dat")
            g = bgc[f"bgc{t}_{kind}"]
            assert mine.shape == g.shape
            assert np.abs(mine - g).max() <= 2.0e-7, (t, kind, np.abs(mine - g).max())


def test_grotz_sheba_first_records(tmp_path, golden_dir):
    """grotz(4, ...) with the forcing read from *.txt.input files like sub_input: records 1-13 against the golden
    SHEBA files (T2m in all digits, T_top to 1e-9, N_active exact)."""
    from samsim_b200 import grotz
    F = np.load(golden_dir / "forcing_era.npz")["sheba"]
    fdir = tmp_path / "forcing"
    fdir.mkdir()
    for name, row in zip(["flux_sw", "flux_lw", "T2m", "precip"], F):
        with open(fdir / f"{name}.txt.input", "w") as f:
            for v in row:
                f.write(f"  {v: .7e}\n")
    out = tmp_path / "output"
    rc = grotz.grotz(4, "sheba", output_dir=out, forcing_dir=fdir, max_steps=12 * 8641 + 1)
    assert rc == 0
    gold = np.load(golden_dir / "sheba_reference.npz")
    tt = _read_dat(out / "dat_T2m_T_top.dat")
    assert tt.shape == (13, 2)
    assert np.array_equal(tt[:, 0], gold["T2m_T_top"][:13, 0])
    assert np.abs(tt[:, 1] - gold["T2m_T_top"][:13, 1]).max() <= 1e-9
    thick = _read_dat(out / "dat_thick.dat")
    assert np.array_equal((thick != 0).sum(1), gold["N_active"][:13])


def test_grotz_lab_and_field_testcases_read_their_series(oracle_mod, tmp_path, golden_dir):
    """samsim_grotz for the testcases that READ series in the reference's prologue (mo_grotz.f90:138-169): lab
    testcase 101 (2017_input/{Tice,snowfall,heat,styropor}_exp_1.txt; synthetic files, the real ones are not shipped)
    testcase 8 (the field temperatures input/DNotz_fieldT/Tinput.txt) and testcase 111 (harp temperatures Ts_3s.txt, synthetic).  The written dat_thick / dat_T records
    equal the oracle's records at print precision, N_active exactly."""
    from samsim_b200 import grotz
    nsteps = 3 * 3601 + 1
    # ---- testcase 101: four per-second series, length_input_lab values each ----
    nrec = grotz.init_testcase(101)["length_input_lab"]
    series = scenarios.lab_series(20000, styropor_hours=(1, 2))
    labdir = tmp_path / "2017_input"
    labdir.mkdir()
    for name, row in zip(["Tice", "snowfall", "heat", "styropor"], series):
        full = np.concatenate([row, np.full(nrec - len(row), row[-1])])
        np.savetxt(labdir / f"{name}_exp_1.txt", full, fmt="%.9e")
    out = tmp_path / "out101"
    assert grotz.grotz(101, "lab", output_dir=out, lab_input_dir=labdir, max_steps=nsteps) == 0
    col = oracle_mod.Column(101, "det")
    col.set_lab_forcing(*[np.loadtxt(labdir / f"{n}_exp_1.txt")[:20000] for n in ("Tice", "snowfall", "heat", "styropor")])
    col.record_outputs()
    assert col.step(nsteps) == 0
    thick, T = _read_dat(out / "dat_thick.dat"), _read_dat(out / "dat_T.dat")
    assert thick.shape[0] == len(col.records) == 4
    for j, rec in enumerate(col.records):
        assert np.abs(thick[j] - rec["thick"]).max() <= 0.5e-5 and np.abs(T[j] - rec["T"]).max() <= 0.5e-3
        assert (thick[j] != 0).sum() == rec["N_active"]
    # ---- testcase 8: Tinput.txt, one value per minute ----
    Tin = np.load(golden_dir / "tinput_dnotz.npz")["Tinput"]
    fdir = tmp_path / "field"
    fdir.mkdir()
    np.savetxt(fdir / "Tinput.txt", Tin, fmt="%.4f")
    out8 = tmp_path / "out8"
    assert grotz.grotz(8, "field T", output_dir=out8, lab_input_dir=fdir, max_steps=nsteps) == 0
    lab = np.zeros((4, len(Tin)))
    lab[0] = np.loadtxt(fdir / "Tinput.txt")
    col8 = oracle_mod.Column(8, "det")
    col8.set_lab_forcing(*lab)
    col8.record_outputs()
    assert col8.step(nsteps) == 0
    thick8 = _read_dat(out8 / "dat_thick.dat")
    tt8 = _read_dat(out8 / "dat_T2m_T_top.dat")
    assert thick8.shape[0] == len(col8.records) == 4
    for j, rec in enumerate(col8.records):
        assert np.abs(thick8[j] - rec["thick"]).max() <= 0.5e-5 and (thick8[j] != 0).sum() == rec["N_active"]
        assert tt8[j, 1] == rec["T_top"]
    assert col8.int("N_active") > 2
    # ---- testcase 111: Ts_<int(dt)>s.txt, one value per time step (mo_grotz.f90:171-176) ----
    n111 = 8000
    t = 3.0 * np.arange(n111, dtype=np.float64)
    np.savetxt(fdir / "Ts_3s.txt", -12.0 - 8.0 * np.sin(2.0 * np.pi * t / 86400.0), fmt="%.6f")
    out111 = tmp_path / "out111"
    steps111 = 2 * 2401 + 1   # time_out 7200 s / dt 3 s: an output record every 2401 steps
    assert grotz.grotz(111, "harp", output_dir=out111, lab_input_dir=fdir, max_steps=steps111) == 0
    lab = np.zeros((4, n111))
    lab[0] = np.loadtxt(fdir / "Ts_3s.txt")
    col111 = oracle_mod.Column(111, "det")
    col111.set_lab_forcing(*lab)
    col111.record_outputs()
    assert col111.step(steps111) == 0
    thick111 = _read_dat(out111 / "dat_thick.dat")
    tt111 = _read_dat(out111 / "dat_T2m_T_top.dat")
    assert thick111.shape[0] == len(col111.records) == 3
    for j, rec in enumerate(col111.records):
        assert np.abs(thick111[j] - rec["thick"]).max() <= 0.5e-5 and (thick111[j] != 0).sum() == rec["N_active"]
        assert tt111[j, 1] == rec["T_top"]
    assert col111.int("N_active") > 2


def test_rebin_is_invisible_to_results(oracle_mod, golden_dir):
    """SURVEY 8e re-binning: sorting the columns by regime on the device changes neither any column's result (bit
    for bit, against a handle that never re-bins and against the oracle) nor the caller's column numbering --
    get/set, snapshots, per-column forcing vectors and broadcast all keep working through the slot map."""
    z = np.load(golden_dir / "sheba_oracle_states.npz")
    F = _forcing(golden_dir)
    ncol = 1500
    rng = np.random.default_rng(11)
    scale = np.ones((4, ncol))
    offset = np.zeros((4, ncol))
    scale[1] = rng.uniform(0.55, 1.1, ncol)  # wide long-wave spread: growth rates, hence N_active, drift apart
    offset[2] = rng.uniform(-2, 2, ncol)
    scale[3] = rng.uniform(0.0, 2.0, ncol)
    amp = 7.0 * rng.uniform(0.5, 1.5, ncol)
    base = oracle_mod.Column(4, "det")
    base.set_forcing(*F)
    base.load_state(_state(z, 80))

    def make(rebin_every):
        e = pu.engine_from_oracle(base, ncol=ncol)
        e.set_snapshot_mode(api.SNAP_FULL)
        e.set_forcing(F[None], None, scale, offset)
        e.set_scalar("oflux_amp", amp)
        e.set_rebin_interval(rebin_every)
        return e

    plain, binned = make(0), make(2500)
    nsteps = 12000  # 33 h of autumn growth; crosses an output step (period 8641)
    plain.step(nsteps)
    binned.step(nsteps)
    slot = binned.slot_map()
    assert sorted(slot.tolist()) == list(range(ncol))
    assert (slot != np.arange(ncol)).any(), "ensemble did not diverge: nothing was re-binned"
    na = binned.get_int("N_active")
    assert na.max() - na.min() >= 3 and binned.count_failed() == 0
    # slots are ordered by N_active, deepest first (the leading sort key among healthy columns), after a re-binning
    binned.rebin()
    slot = binned.slot_map()
    assert (np.diff(na[np.argsort(slot)]) <= 0).all()
    for name in api.ARRAY_IDS:
        if plain.extent(name) > 0:
            assert pu.same_bits(plain.get_array(name), binned.get_array(name)).all(), name
    for name in api.SCALAR_IDS:
        assert pu.same_bits(plain.get_scalar(name), binned.get_scalar(name)).all(), name
    for name in api.INT_IDS:
        assert (plain.get_int(name) == binned.get_int(name)).all(), name
    sp, sb = plain.get_snapshot(arrays=True), binned.get_snapshot(arrays=True)
    for k in sp:
        assert pu.same_bits(np.asarray(sp[k]), np.asarray(sb[k])).all(), "snapshot " + k
    # the caller's column numbers still address the same columns: oracle check + set/get + broadcast + new forcing
    for c in (0, 777, ncol - 1):
        col = oracle_mod.Column(4, "det")
        col.set_forcing(*[F[k] * scale[k, c] + offset[k, c] for k in range(4)])
        col.load_state(_state(z, 80))
        col.set_scalar("oflux_amp", amp[c])
        assert col.step(nsteps) == 0
        bad = pu.compare_column(col, binned, c, label=f"col {c}: ")
        assert not bad, _fmt(bad)
    for e in (plain, binned):
        e.set_scalar("T_bottom", [-1.25], col0=5)
        e.set_array("flush_v", np.full((2, e.extent("flush_v")), 0.125), col0=40)
        e.broadcast_column(777, 100, 3)
        e.set_forcing(F[None], None, scale[:, ::-1].copy(), offset[:, ::-1].copy())
        e.step(300)
    for name in ("T", "S_abs", "H_abs", "thick", "flush_v"):
        assert pu.same_bits(plain.get_array(name), binned.get_array(name)).all(), "after edits: " + name
    assert (plain.get_int("N_active") == binned.get_int("N_active")).all()
    assert binned.get_scalar("T_bottom", 5, 1)[0] == -1.25 and binned.get_scalar("T_bottom", 4, 1)[0] != -1.25
    assert pu.same_bits(binned.get_array("thick", 777, 1), binned.get_array("thick", 101, 1)).all() is not None
    assert binned.count_failed() == plain.count_failed()


def test_divergence_is_measured_in_the_kernel_and_triggers_rebinning(oracle_mod, golden_dir):
    """The step kernel measures its own warp divergence (warp max / add reductions over N_active, a ballot over the snow
    class) and samsim_b200_set_rebin_auto re-bins when the idle share of lane-layers crosses the threshold: a drifting
    ensemble gets re-binned without a host-set interval, the measured idle share drops, results stay bit-identical."""
    z = np.load(golden_dir / "sheba_oracle_states.npz")
    F = _forcing(golden_dir)
    ncol = 4096
    rng = np.random.default_rng(12)
    scale, offset = np.ones((4, ncol)), np.zeros((4, ncol))
    scale[1] = rng.uniform(0.55, 1.1, ncol)
    offset[2] = rng.uniform(-2, 2, ncol)
    base = oracle_mod.Column(4, "det")
    base.set_forcing(*F)
    base.load_state(_state(z, 80))

    def make(auto):
        e = pu.engine_from_oracle(base, ncol=ncol)
        e.set_forcing(F[None], None, scale, offset)
        e.set_rebin_auto(auto)
        return e

    plain, auto = make(0.0), make(0.02)
    for e in (plain, auto):
        for _ in range(6):
            e.step(2000)
    dp, da = plain.divergence(), auto.divergence()
    na = plain.get_int("N_active")
    assert na.max() - na.min() >= 3
    # idle share by definition, from the final N_active in the caller's (= plain's device) order
    w = na.reshape(-1, 32)
    expect = (w.max(1, keepdims=True) - w).sum() / (w.max(1) * 32).sum()
    assert abs(dp["idle_lane_layer_share"] - expect) < 1e-12 and dp["rebins"] == 0
    assert da["rebins"] >= 1 and da["idle_lane_layer_share"] < 0.5 * dp["idle_lane_layer_share"], (dp, da)
    for name in ("m", "S_abs", "H_abs", "thick", "T", "phi"):
        assert pu.same_bits(plain.get_array(name), auto.get_array(name)).all(), name
    assert (plain.get_int("N_active") == auto.get_int("N_active")).all()


# the remaining testcases of mo_init.f90 (no golden output exists for them: GPU vs oracle only)
OTHER_TESTCASES = {
    2: (40000, "cooling chamber: tank, boundflux 3, T2m steps to +1 after 15 days (sub_test2)"),
    3: (200000, "climatological forcing (notzflux) + solid precipitation, 139 days: open water to a full grid (sub_test3)"),
    5: (15000, "top melt of a 1 m block of cold fresh ice, atmoflux 3, N_active = Nlayer from the start, S_abs reset at i = 2"),
    6: (140000, "small tank, dt 0.5 s, T2m schedule of sub_test6"),
    8: (90000, "field surface temperatures (input/DNotz_fieldT/Tinput.txt, one per minute) prescribe T_top, 25 h"),
    9: (30000, "cooling chamber with the T2m schedule of sub_test9 (growth, then melt)"),
    33: (9000, "fresh-water chamber (S_bu_bottom 0.13)"),
    34: (60000, "chamber with the T2m schedule of sub_test34"),
    50: (680000, "spin-up column of the convection studies (boundflux 2, climatological fluxes): 70 days of open water, then freeze-up to ~35 layers"),
    99: (50000, "snow on ice in the chamber: T2m -40 for three days, then the hook resets the snow cover every step (mo_grotz.f90:547-563)"),
    111: (25000, "salinity-harp comparison: T_top = Ttop_input(FLOOR(1 + time/dt)) (mo_grotz.f90:505-506), synthetic series"),
}


@pytest.mark.parametrize("testcase", sorted(OTHER_TESTCASES))
def test_other_testcases_from_init(oracle_mod, golden_dir, testcase):
    nsteps, _what = OTHER_TESTCASES[testcase]
    col = oracle_mod.Column(testcase, "det")
    lab = None
    if testcase == 8:
        Tin = np.load(golden_dir / "tinput_dnotz.npz")["Tinput"]
        lab = np.zeros((4, len(Tin)))
        lab[0] = Tin
        col.set_lab_forcing(*lab)
    if testcase == 111:  # the reference's 2017_input/Ts_3s.txt is not shipped: a cold spell with a diurnal cycle
        t = 3.0 * np.arange(30000, dtype=np.float64)
        lab = np.zeros((4, 30000))
        lab[0] = -12.0 - 8.0 * np.sin(2.0 * np.pi * t / 86400.0) - 4.0 * np.minimum(t / 86400.0, 2.0)
        col.set_lab_forcing(*lab)
    eng = pu.engine_from_oracle(col, ncol=2)
    if lab is not None:
        eng.set_lab_forcing(lab[None])
    eng.set_tuning(testcase == 8)  # testcase 8 (boundflux 1, no tracers) may take the two-pass step
    done = 0
    for target in (1, 2, 3, nsteps // 3, nsteps):
        n = target - done
        assert col.step(n) == 0
        eng.step(n)
        done = target
        bad = pu.compare_column(col, eng, 1, label=f"testcase {testcase} step {target}: ")
        assert not bad, _fmt(bad)
    if testcase != 5:
        assert col.int("N_active") > 1, "the run never grew ice"


def test_testcase7_simple_parametrisations(oracle_mod, golden_dir):
    """Testcase 7 = SHEBA forcing with the simple parametrisations selected: from the oracle's autumn state so that
    ice exists within the window.  This window executes fl_grav_drain_simple and albedo_flag 1 only (asserted below);
    flush4 and flood_simple need melt / a negative freeboard and are covered by tests/scenarios.py."""
    F = _forcing(golden_dir)
    col = oracle_mod.Column(7, "det")
    col.set_forcing(*F)
    assert col.step(900000) == 0  # 104 days from 1 July: freeze-up done
    assert col.int("N_active") > 5
    eng = pu.engine_from_oracle(col, ncol=2)
    eng.set_forcing(F[None])
    before = col.event_counts()  # the spin-up above is the oracle's alone
    for n in (1, 2000, 10000):
        assert col.step(n) == 0
        eng.step(n)
        bad = pu.compare_column(col, eng, 1, label=f"testcase 7 +{n}: ")
        assert not bad, _fmt(bad)
    o = {k for k, v in col.event_counts().items() if v > before[k]}
    assert "grav_drain_simple" in o and not ({"flush4", "flood_simple"} & o) and o == eng.events(1) - scenarios.DEVICE_ONLY


def test_prescribe_flag_2(oracle_mod):
    """mo_grotz.f90:482-497: the prescribed salinity profile (no config uses it; parity with the oracle)."""
    col = oracle_mod.Column(1, "det")
    col.set_int("prescribe_flag", 2)
    eng = pu.engine_from_oracle(col, ncol=2)
    for n in (1, 5000, 40000):
        assert col.step(n) == 0
        eng.step(n)
        bad = pu.compare_column(col, eng, 0, label=f"prescribe +{n}: ")
        assert not bad, _fmt(bad)
    assert col.int("N_active") > 10


def test_checkpoint_restart_is_bit_identical(oracle_mod, golden_dir, tmp_path):
    """save -> new handle -> load continues exactly like the uninterrupted run (also across a
    re-binning, which must not leak into the file)."""
    z = np.load(golden_dir / "sheba_oracle_states.npz")
    F = _forcing(golden_dir)
    ncol = 600
    rng = np.random.default_rng(3)
    scale = np.ones((4, ncol)); offset = np.zeros((4, ncol))
    scale[1] = rng.uniform(0.6, 1.1, ncol)
    base = oracle_mod.Column(4, "det")
    base.set_forcing(*F)
    base.load_state(_state(z, 80))
    a = pu.engine_from_oracle(base, ncol=ncol)
    a.set_forcing(F[None], None, scale, offset)
    a.step(6000)
    a.rebin()
    a.save_checkpoint(tmp_path / "ck.bin")
    a.step(3000)
    b = api.Engine(a.cfg, ncol, 0)
    b.load_checkpoint(tmp_path / "ck.bin")
    b.set_forcing(F[None], None, scale, offset)
    b.step(3000)
    assert a.get_clock() == b.get_clock()
    for name in api.ARRAY_IDS:
        if a.extent(name) > 0:
            assert pu.same_bits(a.get_array(name), b.get_array(name)).all(), name
    for name in api.SCALAR_IDS:
        assert pu.same_bits(a.get_scalar(name), b.get_scalar(name)).all(), name
    for name in api.INT_IDS:
        assert (a.get_int(name) == b.get_int(name)).all(), name
    other = api.Engine(a.cfg, ncol + 1, 0)
    with pytest.raises(api.SamsimError):
        other.load_checkpoint(tmp_path / "ck.bin")


def test_tracers_tank_and_snapshot(oracle_mod):
    """SURVEY 8f-3, passive tracers beyond testcase 1 (whose tracer arrays every testcase-1 test above compares):
    testcase 6 (one tracer) and 2 (two tracers) add the tank budget (bgc_bottom from bgc_total) and the turbulent
    exchange; the S8 record carries output_bgc's bulk / brine rows."""
    for testcase, nsteps in ((6, 60000), (2, 20000)):
        col = oracle_mod.Column(testcase, "det")
        eng = pu.engine_from_oracle(col, ncol=2)
        assert eng.cfg.N_bgc == (1 if testcase == 6 else 2)
        eng.set_snapshot_mode(api.SNAP_FULL)
        col.record_outputs()
        assert col.step(nsteps) == 0
        eng.step(nsteps)
        bad = pu.compare_column(col, eng, 1, label=f"testcase {testcase}: ")
        assert not bad, _fmt(bad)
        assert col.int("N_active") > 3
        rec, snap = col.records[-1], eng.get_snapshot(arrays=True)
        for t in range(1, eng.cfg.N_bgc + 1):
            for kind in ("bu", "br"):
                assert pu.same_bits(rec[f"bgc{t}_{kind}"], snap[f"bgc{t}_{kind}"][1]).all(), (testcase, t, kind)
        assert abs(col.scalar("bgc_bottom1") - 385.0) > 1e-6  # the tank budget really moved the water concentration


@pytest.mark.parametrize("two_pass", [False, True])
@pytest.mark.parametrize("name", scenarios.NAMES)
def test_constructed_branch_scenarios(oracle_mod, name, two_pass):
    """Branches that neither the SHEBA year nor the testcases from `init` enter (tests/scenarios.py): flood incl. the
    instant flooding beyond neg_free, flood_simple, flush4, flush_flag 4, snow_thermo with snow_flush_flag 0, the
    top_grow / top_melt sub-cases, bottom_melt(_simple) with a full grid, the warm branches of snow_coupling,
    sub_melt_snow with all the snow flooded, the styropor factor, the S24 salt clamp.  Bitwise state through the C ABI
    AND proof that each branch ran: oracle counters > 0, device event bits set, the two event sets equal."""
    sc = scenarios.build(oracle_mod, name)
    eng = pu.engine_from_oracle(sc.col, ncol=33)
    eng.set_tuning(two_pass)
    if sc.forcing is not None:
        eng.set_forcing(sc.forcing[None])
    if sc.lab is not None:
        eng.set_lab_forcing(sc.lab[None])
    bad = scenarios.run(sc, eng, lambda o, e, c, label="": pu.compare_column(o, e, 32, label=label))
    assert not bad, _fmt(bad)
    assert eng.events(32) == eng.events(0)


def test_event_words_of_the_sheba_windows(oracle_mod, golden_dir):
    """The event words agree with the oracle's branch counters in the regimes of the SHEBA year, and the windows
    execute the branches DESIGN.md section 10 credits them with."""
    expect = {80: {"bottom_growth_simple", "grav_drained", "snow_compaction"}, 200: {"bottom_growth", "grav_drained"},
              345: {"flush3", "snow_wet", "flush3_clamp", "snow_meltwater_to_ice", "heat_melt"},
              380: {"flush3", "getT_Tfr_fallback", "heat_thin_snow", "melt_thick_gas"},
              400: {"flush3", "bottom_melt_simple_a", "top_melt_b"}, 715: {"top_melt_c", "flush3"},
              730: {"snow_merge", "melt_snow_part", "heat_thin_snow"}}
    F = _forcing(golden_dir)
    for rec, need in expect.items():
        col = oracle_mod.Column(4, "det")
        col.set_forcing(*F)
        col.load_state(scenarios.sheba_state(rec))
        eng = pu.engine_from_oracle(col, ncol=2)
        eng.set_tuning(True)  # the two-pass step where the column is steady, the general path elsewhere
        eng.set_forcing(F[None])
        assert col.step(3000) == 0
        eng.step(3000)
        o = col.events()
        assert need <= o, (rec, sorted(need - o))
        dev = eng.events(1) - scenarios.DEVICE_ONLY
        assert o == dev, (rec, sorted(o ^ dev))
        assert rec not in (80, 200) or "two_pass_step" in eng.events(1), rec  # the steady winter regimes take the two-pass step
        bad = pu.compare_column(col, eng, 1, label=f"state {rec} +3000: ")
        assert not bad, _fmt(bad)
        eng.clear_events()
        assert eng.events(0) == set()
