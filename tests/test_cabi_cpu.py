"""The C-ABI shared library: loads, exports every symbol of include/samsim_b200.h, ids in sync, and fails
loudly (SAMSIM_ERR_NO_DEVICE) instead of falling back when there is no GPU."""
import ctypes as C
import re
from pathlib import Path

import pytest

from samsim_b200 import api, build

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "samsim_b200.h").read_text()


@pytest.fixture(scope="module")
def lib():
    build.build()
    return api.load_library()


def test_every_declared_symbol_is_exported(lib):
    declared = sorted(set(re.findall(r"\b(samsim_b200_[a-zA-Z0-9_]+)\s*\(", HEADER)))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(api.EXPORTED_SYMBOLS) == declared


def _enum_names(prefix):
    """names of the enumerators of the typedef enum whose members start with `prefix`, in declaration order"""
    for body, _name in re.findall(r"typedef enum \{(.*?)\}\s*(\w+);", HEADER, flags=re.S):
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = re.findall(r"\b(" + prefix + r"[A-Z0-9_]+)\b", body)
        if names:
            return [m for m in names if not m.endswith("_COUNT")]
    return []


def test_ids_match_header():
    def uniq(seq):
        out = []
        for s in seq:
            if s not in out:
                out.append(s)
        return out
    arr = [n[len("SAMSIM_ARR_"):].lower() for n in uniq(_enum_names("SAMSIM_ARR_"))]
    assert arr == [k.lower() for k in api.ARRAY_IDS]
    sc = [n[len("SAMSIM_SC_"):].lower() for n in uniq(_enum_names("SAMSIM_SC_"))]
    assert sc == [k.lower() for k in api.SCALAR_IDS]
    snap = [n[len("SAMSIM_SNAPSC_"):].lower() for n in uniq(_enum_names("SAMSIM_SNAPSC_"))]
    assert snap == [k.lower() for k in api.SNAP_SCALARS]
    snapa = [n[len("SAMSIM_SNAPARR_"):].lower() for n in uniq(_enum_names("SAMSIM_SNAPARR_"))]
    assert snapa == [k.lower() for k in api.SNAP_ARRAYS]


def test_event_ids_match_header_and_oracle(oracle_mod):
    """samsim_event_id (header) = api.EVENT_NAMES = the oracle's branch counters, same order: the parity tests compare
    the device's event bits with the oracle's counters by name."""
    ev = [n[len("SAMSIM_EV_"):].lower() for n in _enum_names("SAMSIM_EV_")]
    assert ev == [k.lower() for k in api.EVENT_NAMES] and len(ev) <= 64
    assert list(oracle_mod.Column(1, "det").event_counts()) == api.EVENT_NAMES
    ints = [n[len("SAMSIM_INT_"):].lower() for n in _enum_names("SAMSIM_INT_")]
    assert ints == ["n_active", "status", "styropor_flag", "events0", "events1"] and list(api.INT_IDS.values()) == [0, 1, 2, 3, 4]
    assert api.decode_events(1 | (1 << 31), 1 | (1 << 15)) == {"flood", "salt_clamp", "gas_refill", "tank"}


def test_config_struct_layout():
    # 26 int32 + 10 doubles, no padding surprises: the Fortran BIND(C) type in fortran/mo_samsim_b200.f90 mirrors it
    assert C.sizeof(api._CConfig) == 26 * 4 + 10 * 8


def test_no_gpu_means_error_not_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = api.Config(testcase=1, Nlayer=90, N_top=5, N_middle=80, N_bottom=5, dt=1.0, thick_0=0.002, thick_min=0.001,
                     time_out=3600.0, i_time_out=3600, salt_flag=2)
    with pytest.raises(api.SamsimError) as e:
        api.Engine(cfg, 4)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)
    with pytest.raises(api.SamsimError):
        api.kat_scalar(8, 1, [0.0])


def test_product_does_not_import_the_oracle():
    for p in (ROOT / "samsim_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".cpp"):
            txt = p.read_text()
            assert "oracle" not in txt.replace("the oracle", "").replace("CPU oracle", "").replace("oracle's", "").lower() \
                or "import oracle" not in txt and "oracle/" not in txt and "from oracle" not in txt, p
