"""Build the committed test fixtures from /root/reference (this container only).

  tests/golden/forcing_era.npz      ERA-interim forcing of the 9 sites (sub_input format, mo_functions.f90:304-327):
                                    sheba: 13148 records (full testcase-4 run), other sites: first 2928 records (1 year)
  tests/golden/tc1_reference.npz    reference_output/Reference_testcase1_with_Version_2 (all 72 records)
  tests/golden/tc1_bgc_reference.npz  its dat_bgc0{1,2}.{bu,br}.dat tracer files (72 x 90, F16.8)
  tests/golden/sheba_reference.npz  reference_output/Reference_SHEBA_with_Version_2: every record of the scalar files,
                                    every 6th record (+ the melt-onset window 320-360) of the per-layer files
Usage: python tools/make_fixtures.py            (reference data -> fixtures)
       python tools/make_fixtures.py states <npz from tools/run_oracle_sheba.py det ...>
"""
import sys
from pathlib import Path

import numpy as np

REF = Path('/root/reference')
OUT = Path(__file__).resolve().parent.parent / 'tests' / 'golden'
OUT.mkdir(parents=True, exist_ok=True)

SITES = ['sheba', '70N00W', '75N00W', '75N180E', '80N00E', '80N90E', '85N180E', 'NorthPole', 'barrow']
KINDS = ['flux_sw', 'flux_lw', 'T2m', 'precip']  # order of samsim_forcing_kind


def forcing():
    d = {}
    for s in SITES:
        n = 13148 if s == 'sheba' else 2928
        arr = np.stack([np.loadtxt(REF / 'input' / 'ERA-interim' / f'{s}-p2' / f'{k}.txt.input')[:n] for k in KINDS])
        d[s] = arr
    root = np.stack([np.loadtxt(REF / f'{k}.txt.input')[:13148] for k in KINDS])
    assert np.array_equal(root, d['sheba']), 'root-level forcing is expected to equal sheba-p2'
    np.savez_compressed(OUT / 'forcing_era.npz', sites=np.array(SITES), **d)
    print('forcing_era.npz', (OUT / 'forcing_era.npz').stat().st_size)


LAYER_FILES = ['T', 'psi_s', 'psi_l', 'psi_g', 'S_bu', 'thick', 'ray', 'perm', 'flush_v', 'flush_h']
SCALAR_FILES = ['freeboard', 'snow', 'vital_signs', 'grav_drain', 'T2m_T_top', 'melt']


def golden(dirname, outname, rec_sel):
    p = REF / 'reference_output' / dirname
    d = {}
    for f in LAYER_FILES:
        a = np.loadtxt(p / f'dat_{f}.dat')
        if f == 'thick':
            d['N_active'] = (a != 0).sum(1).astype(np.int32)  # exact: inactive layers print 0.00000
        d[f] = a[rec_sel]
    for f in SCALAR_FILES:
        d[f] = np.loadtxt(p / f'dat_{f}.dat')
    d['records'] = np.asarray(rec_sel)
    d['settings'] = np.array((p / 'dat_settings.dat').read_text())
    np.savez_compressed(OUT / outname, **d)
    print(outname, (OUT / outname).stat().st_size)


def bgc_golden():
    """dat_bgc0{1,2}.{bu,br}.dat of the testcase-1 golden run: 72 records x 90 layers, F16.8 (tight pins on the
    brine fluxes that move the passive tracers)."""
    d = REF / 'reference_output' / 'Reference_testcase1_with_Version_2'
    keep = {}
    for t in (1, 2):
        for kind in ('bu', 'br'):
            a = np.loadtxt(d / f'dat_bgc0{t}.{kind}.dat')
            assert a.shape == (72, 90), a.shape
            keep[f'bgc{t}_{kind}'] = a
    np.savez_compressed(OUT / 'tc1_bgc_reference.npz', **keep)
    print('tc1_bgc_reference.npz', (OUT / 'tc1_bgc_reference.npz').stat().st_size)


def field_temperatures():
    """input/DNotz_fieldT/Tinput.txt: the per-minute surface temperatures testcase 8 prescribes (mo_grotz.f90:539-544)."""
    a = np.loadtxt(REF / 'input' / 'DNotz_fieldT' / 'Tinput.txt', dtype=np.float64)
    np.savez_compressed(OUT / 'tinput_dnotz.npz', Tinput=a)
    print('tinput_dnotz.npz', a.shape, (OUT / 'tinput_dnotz.npz').stat().st_size)


def main_reference():
    forcing()
    field_temperatures()
    bgc_golden()
    golden('Reference_testcase1_with_Version_2', 'tc1_reference.npz', np.arange(72))
    sel = sorted(set(range(0, 1643, 6)) | set(range(320, 361)) | {1642})
    golden('Reference_SHEBA_with_Version_2', 'sheba_reference.npz', np.array(sel))


def oracle_states(npz_path, outname='sheba_oracle_states.npz'):
    """Restartable oracle states (HEAD source, det math) harvested by tools/run_oracle_sheba.py:
    state<j>_* = full mo_data state BEFORE the step that writes 1-based output record j."""
    src = np.load(npz_path)
    keep = {}
    for k in src.files:
        if k.startswith('state') or k in ('stats',):
            keep[k] = src[k]
    np.savez_compressed(OUT / outname, **keep)
    print(outname, (OUT / outname).stat().st_size)


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == 'states':
        oracle_states(sys.argv[2])
    else:
        main_reference()
