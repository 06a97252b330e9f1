"""Fuzz the device code (compiled for the host, tests/hostbuild) against the oracle: random restart states of the
SHEBA year, random forcing perturbations, random edits of the column, random launch chunking.  Every state value must
stay bit-identical.  Usage: python tools/fuzz_host_kernel.py [trials] [seed]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from hostbuild import hostkernel as hk
from oracle import oracle, parity_util as pu

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 50
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
EXTREME = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0   # share of trials with absurd edits (error paths)
rng = np.random.default_rng(seed)
z = np.load(ROOT / "tests/golden/sheba_oracle_states.npz")
F = np.load(ROOT / "tests/golden/forcing_era.npz")["sheba"]
recs = sorted({int(k.split("_")[0][5:]) for k in z.files if k.startswith("state") and k.split("_")[0][5:].isdigit()})
nbad = 0


def other_testcase(t):
    """testcases 1, 2, 3, 5, 6, 9, 33, 34, 50, 99, 111 and lab cases from init with random boundary values (and testcase 1's tracers)"""
    tc = int(rng.choice([1, 1, 2, 3, 5, 6, 9, 33, 34, 50, 99, 111, 101, 103, 105]))
    col = oracle.Column(tc, "det")
    edits = []
    lab = None
    if tc == 111:  # synthetic harp temperatures, one per 3 s step
        n = 30000
        tt = 3.0 * np.arange(n, dtype=np.float64)
        amp, mean = float(rng.uniform(1, 10)), float(rng.uniform(-20, -3))
        lab = np.zeros((4, n))
        lab[0] = mean - amp * np.sin(2.0 * np.pi * tt / 86400.0)
        col.set_lab_forcing(*lab)
        edits = [f"harp mean {mean:.1f} amp {amp:.1f}"]
    elif tc > 100:   # synthetic per-second lab inputs (the reference's 2017_input files are not shipped)
        n = 40000
        tt = np.arange(n, dtype=np.float64)
        amp, mean = float(rng.uniform(2, 14)), float(rng.uniform(-12, -2))
        snow0 = int(rng.integers(0, 20000))
        lab = np.stack([mean - amp * (1.0 - np.cos(2.0 * np.pi * tt / 86400.0)),
                        np.where((tt >= snow0) & (tt < snow0 + 3600), float(rng.choice([0.0, 1e-7, 4e-7])), 0.0),
                        np.full(n, float(rng.uniform(0, 15))),
                        np.where((tt >= 9000) & (tt < 15000), float(rng.integers(0, 2)), 0.0)])
        col.set_lab_forcing(*lab)
        edits = [f"lab mean {mean:.1f} amp {amp:.1f}"]
    if tc == 1:
        w, c_ = float(rng.uniform(-9, -1)), float(rng.uniform(-25, -6))
        col.set_scalar("ttop_warm", w); col.set_scalar("ttop_cold", c_); col.set_scalar("T_top", w)
        col.set_scalar("fl_q_bottom", float(rng.uniform(0, 12)))
        if rng.random() < 0.3: col.set_int("prescribe_flag", 2)
        edits = [f"ttop {w:.2f}/{c_:.2f}"]
    elif tc in (2, 6, 9, 33, 34, 99):
        col.set_scalar("fl_q_bottom", float(rng.uniform(0, 40)))
        col.set_scalar("alpha_flux_instable", float(rng.uniform(10, 40)))
    k = hk.HostKernel(pu.config_from_oracle(col))
    k.load_state(col.state())
    if lab is not None:
        k.set_lab_forcing(lab)
    total = int(rng.integers(500, {1: 60000, 2: 30000, 3: 250000, 5: 20000, 6: 150000, 9: 30000, 33: 9000, 34: 80000, 50: 750000, 99: 60000, 111: 29000}.get(tc, 38000)))
    done, ok = 0, True
    while done < total and ok:
        n = int(min(total - done, rng.choice([1, 2, 5, 100, 3601, 20000])))
        rc_o, rc_k = col.step(n), k.step(n)
        done += n
        bad = pu.compare_column(col, k, 0) if (rc_o == 0 and rc_k == 0) else []
        if rc_o != rc_k or bad:
            ok = False
            print(f"MISMATCH trial {t} testcase {tc} {edits} after {done} steps rc {rc_o}/{rc_k}")
            for b_ in bad[:6]: print("   ", b_)
        if rc_o != 0:
            break
    if ok:
        print(f"trial {t}: testcase {tc} {total} steps N_active {col.int('N_active')} status {col.int('status')} {edits} ok", flush=True)
    return ok


for t in range(trials):
    if rng.random() < 0.35:
        nbad += 0 if other_testcase(t) else 1
        continue
    rec = int(rng.choice(recs))
    p = f"state{rec}_"
    st = {k[len(p):]: (z[k] if z[k].ndim else z[k].item()) for k in z.files if k.startswith(p)}
    scale = np.array([rng.uniform(0.5, 1.5), rng.uniform(0.8, 1.15), 1.0, rng.uniform(0.0, 3.0)])
    offset = np.array([0.0, 0.0, rng.uniform(-15, 8), 0.0])
    Na = int(st["N_active"])
    edits = []
    if Na > 4 and rng.random() < 0.6:   # salinity / enthalpy edits in a random band
        a = int(rng.integers(0, Na - 1)); b = int(rng.integers(a + 1, Na))
        f = float(rng.choice([0.02, 0.3, 0.8, 1.3]))
        st["S_abs"] = np.array(st["S_abs"], dtype=np.float64); st["S_abs"][a:b] *= f
        edits.append(f"S_abs[{a}:{b}]*={f}")
    if rng.random() < 0.3:
        st["oflux_amp"] = float(rng.uniform(0, 40)); edits.append(f"oflux {st['oflux_amp']:.1f}")
    if rng.random() < EXTREME:          # absurd edits: the reference STOPs (or clamps) -- the codes must agree too
        kind = int(rng.integers(0, 4)); j = int(rng.integers(0, max(Na, 1)))
        if kind == 0:
            st["H_abs"] = np.array(st["H_abs"], dtype=np.float64); st["H_abs"][j] *= float(rng.choice([-50.0, 40.0, 1e4]))
            edits.append(f"H_abs[{j}] absurd")
        elif kind == 1:
            st["S_abs"] = np.array(st["S_abs"], dtype=np.float64); st["S_abs"][j] = -abs(st["S_abs"][j]) * float(rng.choice([1.0, 1e-3]))
            edits.append(f"S_abs[{j}] negative")
        elif kind == 2:
            st["thick"] = np.array(st["thick"], dtype=np.float64); st["thick"][j] *= float(rng.choice([0.05, 8.0]))
            edits.append(f"thick[{j}] rescaled")
        else:
            st["m_snow"] = float(st.get("m_snow", 0.0)) * 40.0 + 300.0; edits.append("snow load")
    col = oracle.Column(4, "det")
    col.set_forcing(*[F[q] * scale[q] + offset[q] for q in range(4)])
    col.load_state(st)
    if "oflux_amp" in st: col.set_scalar("oflux_amp", st["oflux_amp"])
    k = hk.HostKernel(pu.config_from_oracle(col))
    k.load_state(col.state())
    k.set_forcing(F, scale, offset)
    total = int(rng.integers(200, 6000))
    done = 0
    ok = True
    while done < total and ok:
        n = int(min(total - done, rng.choice([1, 2, 3, 17, 250, 1081, 3000])))
        rc_o, rc_k = col.step(n), k.step(n)
        done += n
        # after a reference STOP only the code is defined (the Fortran process is gone; the device freezes the column)
        bad = pu.compare_column(col, k, 0) if (rc_o == 0 and rc_k == 0) else []
        if rc_o != rc_k or bad:
            ok = False
            nbad += 1
            print(f"MISMATCH trial {t} rec {rec} scale {scale} offset {offset} edits {edits} after {done} steps rc {rc_o}/{rc_k}")
            for b_ in bad[:6]: print("   ", b_)
        if rc_o != 0:
            break
    if ok:
        print(f"trial {t}: rec {rec} {total} steps N_active {col.int('N_active')} status {rc_o} {edits} ok", flush=True)
print("mismatching trials:", nbad)
