"""The CUDA device functions (samsim_b200/csrc/step.cuh, physics.cuh) compiled for the HOST by
tests/hostbuild/kernel_on_host.cpp and advanced next to the oracle: every state array, scalar and integer must be
bit-identical, exactly like tests/test_parity_gpu.py demands of the GPU build.  This puts the kernel's own logic --
fused sweeps, memoised forward sums, lazy Rayleigh numbers, tracer replay, launch-boundary handling -- under the
no-GPU test stage.  It is a test harness: the product has no CPU path (see test_cabi_cpu.py)."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
from hostbuild import hostkernel as hk  # noqa: E402
from oracle import parity_util as pu  # noqa: E402
import scenarios  # noqa: E402


def _fmt(bad, n=10):
    return "\n".join(bad[:n]) + (f"\n... {len(bad)} mismatches" if len(bad) > n else "")


def _from_oracle(col, two_pass=False):
    k = hk.HostKernel(pu.config_from_oracle(col))
    k.load_state(col.state())
    k.set_tuning(two_pass)
    return k


def _state(z, j):
    p = f"state{j}_"
    return {k[len(p):]: (z[k] if z[k].ndim else z[k].item()) for k in z.files if k.startswith(p)}


def _advance(col, k, chunks):
    for n in chunks:  # every chunk is one "launch": S4 reuse restarts, ray is exact at its last step
        assert col.step(n) == 0
        assert k.step(n) == 0
        bad = pu.compare_column(col, k, 0, label=f"+{n}: ")
        assert not bad, _fmt(bad)


@pytest.mark.parametrize("two_pass", [False, True])
@pytest.mark.parametrize("testcase,chunks", [
    (1, (1, 1, 7, 100, 3493, 3601, 12000)),          # plate cooling, NaCl, two tracers, first output records
    (2, (1, 2, 5000, 15000)),                         # tank + tracers, boundflux 3
    (3, (1, 2, 100000, 60000)),                       # notzflux, snow, 20-layer grid filling up
    (5, (1, 1, 3000, 4000)),                          # full grid from the start, top melt
    (6, (1, 2, 70000)),                               # small tank, one tracer, dt 0.5 s
    (9, (1, 2, 12000)),
    (33, (1, 2, 9000)),                               # fresh-water chamber (S_bu 0.13: the salt-free branch of getT)
    (34, (1, 2, 50000)),                              # chamber with the T2m schedule of sub_test34 (0 -> -15 degC after 2 h)
    (50, (1, 2, 600000, 80000)),                      # spin-up column of the convection studies: 70 days of open water under the climatological fluxes, then freeze-up
    (99, (1, 2, 27000, 20000)),                       # snow on ice in the chamber: the hook resets the snow cover from day 3 on
])
def test_device_code_on_host_equals_oracle_from_init(oracle_mod, testcase, chunks, two_pass):
    col = oracle_mod.Column(testcase, "det")
    if two_pass and testcase == 1:
        col.set_int("bgc_flag", 1)  # without its tracers testcase 1 is eligible for the two-pass step
    _advance(col, _from_oracle(col, two_pass), chunks)


@pytest.mark.parametrize("two_pass", [False, True])
@pytest.mark.parametrize("rec,chunks", [
    (60, (1, 500, 2500)), (100, (1, 700, 1500)), (200, (3, 400, 1200)), (330, (1, 900, 1500)),
    (345, (2, 600, 1500)), (400, (1, 800, 1200)),
])
def test_device_code_on_host_equals_oracle_in_the_sheba_year(oracle_mod, golden_dir, rec, chunks, two_pass):
    """Restart states of the SHEBA run: freeze-up, winter growth, full grid, melt onset with flushing, bare-ice melt
    with layer merges, late-summer bottom melt."""
    z = np.load(golden_dir / "sheba_oracle_states.npz")
    F = np.load(golden_dir / "forcing_era.npz")["sheba"]
    col = oracle_mod.Column(4, "det")
    col.set_forcing(*F)
    col.load_state(_state(z, rec))
    k = _from_oracle(col, two_pass)
    k.set_forcing(F)
    _advance(col, k, chunks)
    took = "two_pass_step" in k.events()
    assert not (took and not two_pass) and (took or not two_pass or rec not in (100, 200)), (rec, two_pass, took)


def test_device_code_on_host_output_record_and_simple_parametrisations(oracle_mod, golden_dir):
    """S8 record captured by the device code = what the oracle hands to output(); testcase 7 switches to
    fl_grav_drain_simple / flush4 / flood_simple."""
    z = np.load(golden_dir / "sheba_oracle_states.npz")
    F = np.load(golden_dir / "forcing_era.npz")["sheba"]
    col = oracle_mod.Column(4, "det")
    col.set_forcing(*F)
    col.load_state(_state(z, 200))
    col.record_outputs()
    k = _from_oracle(col)
    k.set_forcing(F)
    n = 8641 - col.int("n_time_out") + 5  # crosses the next output step
    assert col.step(n) == 0 and k.step(n) == 0 and len(col.records) >= 1
    rec, snap = col.records[-1], k.snapshot()
    for name in ("T", "psi_s", "thick", "S_bu", "ray", "psi_l", "perm", "flush_v", "flush_h", "psi_g"):
        assert pu.same_bits(np.asarray(rec[name])[: len(snap[name])], snap[name]).all(), name
    for name in ("freeboard", "thick_snow", "energy_stored", "total_resist", "thickness", "bulk_salin", "grav_drain", "T_top"):
        assert pu.same_bits(rec[name], snap[name]), name
    col7 = oracle_mod.Column(7, "det")
    col7.set_forcing(*F)
    assert col7.step(880000) == 0 and col7.int("N_active") > 5
    k7 = _from_oracle(col7)
    k7.set_forcing(F)
    _advance(col7, k7, (1, 3000, 6000))


@pytest.mark.parametrize("two_pass", [False, True])
def test_device_code_on_host_testcase8_field_temperatures(oracle_mod, golden_dir, two_pass):
    """Testcase 8 (mo_init.f90:1451-1494, mo_grotz.f90:539-544): T_top follows the field series of
    input/DNotz_fieldT/Tinput.txt (fixture tests/golden/tinput_dnotz.npz), one record per minute."""
    Tin = np.load(golden_dir / "tinput_dnotz.npz")["Tinput"]
    lab = np.zeros((4, len(Tin)))
    lab[0] = Tin
    col = oracle_mod.Column(8, "det")
    col.set_lab_forcing(*lab)
    k = _from_oracle(col, two_pass)
    k.set_lab_forcing(lab)
    _advance(col, k, (1, 2, 3598, 40000, 50000))
    assert col.int("N_active") > 10 and col.scalar("T_top") == Tin[int(1 + (col.scalar("time") - 1.0) / 60) - 1]
    assert ("two_pass_step" in k.events()) == two_pass


def harp_series(n: int) -> np.ndarray:
    """Synthetic uppermost-harp temperatures for testcase 111 (the reference's 2017_input/Ts_3s.txt is not shipped):
    a cold spell with a diurnal cycle, one value per 3 s step."""
    t = 3.0 * np.arange(n, dtype=np.float64)
    return -12.0 - 8.0 * np.sin(2.0 * np.pi * t / 86400.0) - 4.0 * np.minimum(t / 86400.0, 2.0)


@pytest.mark.parametrize("two_pass", [False, True])
def test_device_code_on_host_testcase111_harp_temperatures(oracle_mod, two_pass):
    """Testcase 111 (mo_init.f90:141-221, mo_grotz.f90:171-176, :505-506): T_top = Ttop_input(FLOOR(1 + time/dt))."""
    lab = np.zeros((4, 30000))
    lab[0] = harp_series(30000)
    col = oracle_mod.Column(111, "det")
    col.set_lab_forcing(*lab)
    k = _from_oracle(col, two_pass)
    k.set_lab_forcing(lab)
    _advance(col, k, (1, 2, 2397, 20000))
    assert col.int("N_active") > 5 and col.scalar("T_top") == lab[0][int(col.int("i")) - 1]


def test_device_code_on_host_with_impermeable_layers(oracle_mod, golden_dir):
    """fl_grav_drain's `minval(perm(k:N_active-1)) < 1e-14 -> harmonic_perm = 0` branch (mo_grav_drain.f90:112-113):
    a band of nearly fresh layers in the mid-winter column makes the layers above it non-draining while the layers
    below keep their Rayleigh numbers."""
    z = np.load(golden_dir / "sheba_oracle_states.npz")
    F = np.load(golden_dir / "forcing_era.npz")["sheba"]
    st = _state(z, 200)
    st["S_abs"] = np.array(st["S_abs"], dtype=np.float64)
    st["S_abs"][35:45] *= 0.02
    col = oracle_mod.Column(4, "det")
    col.set_forcing(*F)
    col.load_state(st)
    k = _from_oracle(col)
    k.set_forcing(F)
    _advance(col, k, (1, 2, 50, 400))
    psi_l = col.array("psi_l")[: col.int("N_active") - 1]
    perm = 1e-17 * (1000.0 * np.abs(psi_l)) ** 3.10
    assert (perm < 1e-14).any() and (perm[-10:] >= 1e-14).all(), "the impermeable band is not there"
    ray = col.array("ray")
    j_last = int(np.nonzero(perm < 1e-14)[0].max())
    assert (ray[: j_last + 1] == 0.0).all() and (ray[j_last + 1: col.int("N_active") - 1] > 0.0).any()


def _lab_series(n):
    """Synthetic per-second lab inputs (the reference's 2017_input files are not shipped): Tice, snowfall, heat, styropor."""
    t = np.arange(n, dtype=np.float64)
    Tice = -5.0 - 10.0 * (1.0 - np.cos(2.0 * np.pi * t / 86400.0))
    snow = np.where((t >= 2 * 3600) & (t < 3 * 3600), 1e-7, 0.0)
    heat = np.full(n, 5.0)
    sty = np.where((t >= 4 * 3600) & (t < 5 * 3600), 1.0, 0.0)
    return np.stack([Tice, snow, heat, sty])


@pytest.mark.parametrize("testcase", [101, 104])
def test_device_code_on_host_lab_tank(oracle_mod, testcase):
    """Config 3 (lab tank, Nlayer 200, boundflux 3, tank salinity feedback, lab snow, styropor)."""
    n = 22000
    series = _lab_series(n) * np.array([1.0 + 0.03 * (testcase - 101), 1.5, 1.0, 1.0])[:, None]
    col = oracle_mod.Column(testcase, "det")
    col.set_lab_forcing(*series)
    k = _from_oracle(col)
    k.set_lab_forcing(series)
    _advance(col, k, (1, 2, 3598, 3, 9000, 9000))
    assert col.int("N_active") >= 3


def test_device_code_on_host_perturbed_forcing_and_prescribed_salinity(oracle_mod, golden_dir):
    """value*scale + offset forcing of one column (the device interpolates the base series) against an oracle fed the
    perturbed series; prescribe_flag 2 (mo_grotz.f90:482-497) on testcase 1."""
    z = np.load(golden_dir / "sheba_oracle_states.npz")
    F = np.load(golden_dir / "forcing_era.npz")["sheba"]
    scale, offset = np.array([1.07, 0.96, 1.0, 1.4]), np.array([0.0, 0.0, -1.7, 0.0])
    col = oracle_mod.Column(4, "det")
    col.set_forcing(*[F[q] * scale[q] + offset[q] for q in range(4)])
    col.load_state(_state(z, 100))
    k = _from_oracle(col)
    k.set_forcing(F, scale, offset)
    _advance(col, k, (1, 1081, 1500))
    col1 = oracle_mod.Column(1, "det")
    col1.set_int("prescribe_flag", 2)
    _advance(col1, _from_oracle(col1), (1, 4000, 16000))


@pytest.mark.parametrize("what", ["H_abs layer 3 = -1e12 (getT cannot converge, STOP 99)",
                                  "S_abs layer 5 negative (soft path: PRINT + clamp, mo_grotz.f90:812-818)",
                                  "thick layer 1 = 10 thick_0 (top_grow on consecutive steps)"])
def test_device_code_on_host_stop_codes(oracle_mod, what):
    """Reference STOPs become per-column status codes: the device code must report the oracle's code, and take the
    reference's soft paths (clamp and continue) exactly like it where the reference does not stop."""
    col = oracle_mod.Column(1, "det")
    assert col.step(5000) == 0
    st = col.state()
    if what.startswith("H_abs"):
        a = np.array(st["H_abs"]); a[2] = -1e12; st["H_abs"] = a
    elif what.startswith("S_abs"):
        a = np.array(st["S_abs"]); a[4] = -abs(a[4]); st["S_abs"] = a
    else:
        a = np.array(st["thick"]); a[0] = 10.0 * col.scalar("thick_0"); st["thick"] = a
    ref = oracle_mod.Column(1, "det")
    ref.load_state(st)
    k = hk.HostKernel(pu.config_from_oracle(ref))
    k.load_state(ref.state())
    rc_o, rc_k = ref.step(50), k.step(50)
    assert rc_o == rc_k, (rc_o, rc_k)
    assert int(k.get_int("status")[0]) == rc_o
    if rc_o == 0:
        bad = pu.compare_column(ref, k, 0)
        assert not bad, _fmt(bad)


@pytest.mark.parametrize("two_pass", [False, True])
@pytest.mark.parametrize("name", scenarios.NAMES)
def test_device_code_on_host_constructed_branch_scenarios(oracle_mod, name, two_pass):
    """Branches the SHEBA year never enters (tests/scenarios.py): bitwise state AND proof that the branch ran on both
    sides (oracle branch counters > 0, device event bits set, and the two event sets equal)."""
    sc = scenarios.build(oracle_mod, name)
    k = _from_oracle(sc.col, two_pass)
    if sc.forcing is not None:
        k.set_forcing(sc.forcing)
    if sc.lab is not None:
        k.set_lab_forcing(sc.lab)
    bad = scenarios.run(sc, k, pu.compare_column)
    assert not bad, _fmt(bad)


def test_device_code_on_host_events_of_the_standard_runs(oracle_mod, golden_dir):
    """The event words agree with the oracle's branch counters on the regimes of the SHEBA year, and the runs execute
    the branches DESIGN.md section 10 credits them with."""
    z = np.load(golden_dir / "sheba_oracle_states.npz")
    F = np.load(golden_dir / "forcing_era.npz")["sheba"]
    expect = {80: {"bottom_growth_simple", "grav_drained", "snow_compaction"}, 200: {"bottom_growth", "grav_drained"},
              345: {"flush3", "snow_wet", "flush3_clamp", "snow_meltwater_to_ice", "heat_melt"},
              380: {"flush3", "getT_Tfr_fallback", "heat_thin_snow", "melt_thick_gas"},
              400: {"flush3", "bottom_melt_simple_a", "top_melt_b"}, 715: {"top_melt_c", "flush3"},
              730: {"snow_merge", "melt_snow_part", "heat_thin_snow"}}
    for rec, need in expect.items():
        col = oracle_mod.Column(4, "det")
        col.set_forcing(*F)
        col.load_state(_state(z, rec))
        k = _from_oracle(col, two_pass=True)
        k.set_forcing(F)
        assert col.step(3000) == 0 and k.step(3000) == 0
        o = col.events()
        assert need <= o, (rec, sorted(need - o))
        dev = k.events() - scenarios.DEVICE_ONLY
        assert o == dev, (rec, sorted(o ^ dev))
        assert rec not in (80, 200) or "two_pass_step" in k.events(), rec  # the steady winter regimes take the two-pass step


@pytest.mark.parametrize("salt_flag", [1, 2])
def test_device_getT_on_host_lazy_freezing_point(oracle_mod, salt_flag):
    """getT evaluates the freezing point only when an iterate leaves [-200, 0] degC (mo_thermo_functions.f90:101-103);
    the reference computes it every call.  Same bits on inputs that take the fallback and on inputs that do not."""
    import ctypes as C
    L = oracle_mod.lib("det")
    rng = np.random.default_rng(12)
    n = 20000
    S = np.concatenate([rng.uniform(0.5, 40, n // 2), rng.uniform(0, 0.002, n // 4), rng.uniform(30, 250, n // 4)])
    T_true = rng.uniform(-45, 2, n)
    H = np.where(rng.random(n) < 0.8, -333500.0 * rng.random(n) ** 0.5 + 2020.0 * T_true, 3400.0 * T_true)
    T_in = T_true + rng.normal(0, 0.3, n)
    T_in[::7] = rng.uniform(0.5, 30.0, len(T_in[::7]))      # first guesses above 0 degC: iterates leave the interval
    T_in[3::11] = rng.uniform(-400.0, -210.0, len(T_in[3::11]))
    Tk, pk, st, ev = hk.kat_getT(salt_flag, H, S, T_in)
    To, po = np.empty(n), np.empty(n)
    dp = C.POINTER(C.c_double)
    L.sam_kat_getT(salt_flag, n, H.ctypes.data_as(dp), S.ctypes.data_as(dp), T_in.ctypes.data_as(dp), To.ctypes.data_as(dp), po.ctypes.data_as(dp))
    ok = st == 0
    assert ok.sum() > 0.9 * n
    assert pu.same_bits(Tk[ok], To[ok]).all() and pu.same_bits(pk[ok], po[ok]).all()
    assert np.isnan(To[~ok]).all()
    fallback = (ev >> 1) & 1   # bit (SAMSIM_EV_GETT_TFR_FALLBACK - 32)
    assert fallback.sum() > 500 and (fallback == 0).sum() > 5000, (fallback.sum(), n)
