"""C++ host layer (include/samsim_b200_host.h) without a GPU: init(testcase) tables and the forcing reader."""
import re
from pathlib import Path

import numpy as np
import pytest

from samsim_b200 import api, build, grotz

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module", autouse=True)
def _built():
    build.build()


@pytest.mark.parametrize("testcase", [1, 2, 3, 4, 5, 6, 7, 8, 9, 33, 34, 50, 99, 101, 102, 103, 104, 105, 111])
def test_init_testcase_equals_reference_init(oracle_mod, testcase):
    """samsim_host_init_testcase against the oracle's restatement of mo_init.f90 (flags, grid, initial column)."""
    st = grotz.init_testcase(testcase)
    ost = oracle_mod.Column(testcase, "det").state()
    checked = 0
    for k, v in st.items():
        if k not in ost:
            continue
        o = ost[k]
        if isinstance(v, np.ndarray):
            assert np.array_equal(np.asarray(o)[: len(v)], v), k
        else:
            assert float(o) == float(v), (k, o, v)
        checked += 1
    assert checked > 60


def test_unknown_testcase_is_an_error():
    for tc in (51, 0, 44):  # 51 is a table of 280 restart values; 44 has no init block
        with pytest.raises(api.SamsimError):
            grotz.init_testcase(tc)


def test_read_forcing_is_sub_input(tmp_path, golden_dir):
    """The four ASCII series, one value per line, first nrec records (mo_functions.f90:304-327)."""
    F = np.load(golden_dir / "forcing_era.npz")["barrow"]
    for name, row in zip(["flux_sw", "flux_lw", "T2m", "precip"], F):
        with open(tmp_path / f"{name}.txt.input", "w") as f:
            for v in row:
                f.write(f"  {v: .7e}\n")
    got = grotz.read_forcing(tmp_path, 2000)
    assert np.allclose(got, F[:, :2000], rtol=1e-7, atol=0)
    with pytest.raises(api.SamsimError):
        grotz.read_forcing(tmp_path / "missing", 10)


def test_read_lab_series_is_the_reference_read(tmp_path):
    """mo_grotz.f90:138-169: READ(1234,*) of 2017_input/{Tice,snowfall,heat,styropor}_exp_<N>.txt, list-directed."""
    rng = np.random.default_rng(5)
    want = rng.normal(size=(4, 500))
    for name, row in zip(["Tice", "snowfall", "heat", "styropor"], want):
        with open(tmp_path / f"{name}_exp_3.txt", "w") as f:
            # list-directed input: blanks, commas and line breaks all separate values
            f.write("\n".join(", ".join(repr(float(v)) for v in row[j:j + 7]) for j in range(0, 500, 7)))
    got = grotz.read_lab_series(tmp_path, 103, 500)
    assert np.array_equal(got, want)
    with pytest.raises(api.SamsimError):
        grotz.read_lab_series(tmp_path, 103, 501)   # shorter than length_input_lab
    with pytest.raises(api.SamsimError):
        grotz.read_lab_series(tmp_path, 104, 10)    # files of another experiment are missing


def test_host_header_symbols_exported():
    L = api.load_library()
    hdr = (ROOT / "include" / "samsim_b200_host.h").read_text()
    for name in sorted(set(re.findall(r"\b(samsim_(?:host_[a-z_]+|grotz))\s*\(", hdr))):
        assert hasattr(L, name), name
