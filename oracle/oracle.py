"""ctypes front-end of the CPU oracle (oracle/samsim_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under samsim_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_BUILD = _HERE / "_build"

ARRAY_NAMES = [
    "H", "H_abs", "fl_Q", "T", "S_bu", "S_abs", "S_br", "thick", "m", "fl_m", "V_ex", "phi", "psi_s", "psi_l",
    "psi_g", "ray", "perm", "flush_v", "flush_h", "flush_v_old", "flush_h_old", "fl_rad", "bgc_abs1", "bgc_abs2",
]
SCALAR_NAMES = [
    "dt", "thick_0", "time", "freeboard", "T_freeze", "time_out", "time_total", "T_bottom", "T_top", "S_bu_bottom",
    "T2m", "fl_q_bottom", "psi_s_snow", "psi_l_snow", "psi_g_snow", "phi_s", "S_abs_snow", "H_abs_snow", "m_snow",
    "T_snow", "thick_snow", "liquid_precip", "solid_precip", "fl_q_snow", "energy_stored", "total_resist",
    "freshwater", "thickness", "bulk_salin", "thick_min", "albedo", "fl_sw", "fl_lw", "fl_sen", "fl_lat", "fl_rest",
    "grav_drain", "grav_salt", "grav_temp", "melt_thick", "melt_thick_snow", "melt_thick_snow_old",
    "melt_thick_output1", "melt_thick_output2", "melt_thick_output3", "alpha_flux_instable", "alpha_flux_stable",
    "m_total", "S_total", "tank_depth", "melt_err", "max_flux_plate", "k_snow_flush", "k_styropor", "ttop_warm",
    "ttop_cold", "oflux_amp", "bgc_bottom1", "bgc_bottom2", "bgc_total1", "bgc_total2",
]
INT_NAMES = [
    "testcase", "Nlayer", "N_top", "N_middle", "N_bottom", "N_active", "i", "i_time", "i_time_out", "n_time_out",
    "time_counter", "length_input", "styropor_flag", "atmoflux_flag", "grav_flag", "prescribe_flag",
    "grav_heat_flag", "flush_heat_flag", "turb_flag", "salt_flag", "boundflux_flag", "flush_flag", "flood_flag",
    "bottom_flag", "debug_flag", "precip_flag", "harmonic_flag", "tank_flag", "albedo_flag", "lab_snow_flag",
    "freeboard_snow_flag", "snow_flush_flag", "snow_precip_flag", "bgc_flag", "N_bgc", "status",
]


def build(force: bool = False) -> None:
    """Compile both oracle back-ends with oracle/Makefile (gcc, seconds)."""
    if force:
        subprocess.run(["make", "-C", str(_HERE), "clean"], check=True, capture_output=True)
    subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)


_LIBS: dict[str, C.CDLL] = {}
_OUTPUT_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)


def lib(backend: str = "det") -> C.CDLL:
    if backend not in ("det", "libm", "det_shebagold", "libm_shebagold", "count"):
        raise ValueError(backend)
    if backend in _LIBS:
        return _LIBS[backend]
    path = _BUILD / f"liboracle_{backend}.so"
    if not path.exists():
        if backend == "count":  # C++ operation-counting build (oracle/count_real.h), see tools/count_flops.py
            subprocess.run(["make", "-C", str(_HERE), "count"], check=True, capture_output=True)
        else:
            build()
    L = C.CDLL(str(path))
    L.sam_create.restype = C.c_void_p
    L.sam_create.argtypes = [C.c_int]
    L.sam_destroy.argtypes = [C.c_void_p]
    dp = C.POINTER(C.c_double)
    L.sam_set_forcing.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, dp]
    L.sam_set_lab_forcing.argtypes = [C.c_void_p, C.c_long, dp, dp, dp, dp]
    L.sam_step.restype = C.c_int
    L.sam_step.argtypes = [C.c_void_p, C.c_long]
    L.sam_array_len.argtypes = [C.c_void_p, C.c_char_p]
    L.sam_get_array.argtypes = [C.c_void_p, C.c_char_p, dp]
    L.sam_set_array.argtypes = [C.c_void_p, C.c_char_p, dp]
    L.sam_get_scalar.argtypes = [C.c_void_p, C.c_char_p, dp]
    L.sam_set_scalar.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    L.sam_get_int.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
    L.sam_set_int.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.sam_get_stat.restype = C.c_long
    L.sam_get_stat.argtypes = [C.c_void_p, C.c_char_p]
    L.sam_event_name.restype = C.c_char_p
    L.sam_event_name.argtypes = [C.c_int]
    L.sam_set_output_hook.argtypes = [C.c_void_p, _OUTPUT_FN]
    L.sam_math_backend.restype = C.c_char_p
    for f in ("sam_math_exp", "sam_math_sin"):
        getattr(L, f).restype = C.c_double
        getattr(L, f).argtypes = [C.c_double]
    L.sam_math_pow.restype = C.c_double
    L.sam_math_pow.argtypes = [C.c_double, C.c_double]
    L.sam_run_batch.restype = C.c_int
    L.sam_run_batch.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_long, C.c_int]
    L.sam_kat_getT.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp]
    L.sam_kat_scalar.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp]
    if backend == "count":
        L.sam_get_op_counts.argtypes = [C.POINTER(C.c_longlong)]
    _LIBS[backend] = L
    return L


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


# fields captured by an output record, in the order of the reference's output() (mo_output.f90:116-146)
SNAP_ARRAYS = ["T", "psi_s", "thick", "S_bu", "ray", "psi_l", "perm", "flush_v", "flush_h", "psi_g"]
SNAP_SCALARS = [
    "freeboard", "thick_snow", "T_snow", "psi_l_snow", "psi_s_snow", "energy_stored", "freshwater", "total_resist",
    "thickness", "bulk_salin", "grav_drain", "grav_salt", "grav_temp", "T2m", "T_top", "melt_thick_output1",
    "melt_thick_output2", "melt_thick_output3", "time",
]


class Column:
    """One oracle column = one run of the reference's grotz(testcase, ...)."""

    def __init__(self, testcase: int, backend: str = "det"):
        self.L = lib(backend)
        self.backend = backend
        self.h = self.L.sam_create(testcase)
        if not self.h:
            raise ValueError(f"testcase {testcase} is not covered by the oracle")
        self.records: list[dict] = []
        self._cb = None

    def close(self):
        if self.h:
            self.L.sam_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- forcing -----------------------------------------------------------------------
    def set_forcing(self, fl_sw, fl_lw, T2m, precip):
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (fl_sw, fl_lw, T2m, precip)]
        n = len(a[0])
        assert all(len(x) == n for x in a)
        self.L.sam_set_forcing(self.h, n, *[_dp(x) for x in a])

    def set_lab_forcing(self, Tice, snowfall, heat, styropor):
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (Tice, snowfall, heat, styropor)]
        n = len(a[0])
        self.L.sam_set_lab_forcing(self.h, n, *[_dp(x) for x in a])

    # --- stepping ------------------------------------------------------------------------
    def step(self, n: int = 1) -> int:
        return self.L.sam_step(self.h, n)

    def record_outputs(self, enable: bool = True):
        """Capture a snapshot at every S8 output (mo_grotz.f90:340-398) into self.records."""
        if not enable:
            self._cb = _OUTPUT_FN(0)
        else:
            def cb(_c, _u):
                self.records.append(self.snapshot())
            self._cb = _OUTPUT_FN(cb)
        self.L.sam_set_output_hook(self.h, self._cb)

    def snapshot(self) -> dict:
        d = {n: self.array(n) for n in SNAP_ARRAYS}
        d.update({n: self.scalar(n) for n in SNAP_SCALARS})
        d["N_active"] = self.int("N_active")
        if self.int("bgc_flag") == 2:
            d.update(self.bgc_output())
        return d

    def bgc_output(self) -> dict:
        """bgc_bu / bgc_br rows as output_bgc writes them (mo_output.f90:156-188)."""
        rho_l = 1028.0
        Na, N = self.int("N_active"), self.int("Nlayer")
        m, psi_l, thick = self.array("m"), self.array("psi_l"), self.array("thick")
        out = {}
        for t in range(1, self.int("N_bgc") + 1):
            a, bottom = self.array(f"bgc_abs{t}"), self.scalar(f"bgc_bottom{t}")
            bu, br = np.full(N, bottom), np.full(N, bottom)
            for k in range(Na):
                if m[k] != 0.0:
                    bu[k] = a[k] / m[k]
                    br[k] = a[k] / psi_l[k] / thick[k] / rho_l if (psi_l[k] != 0.0 and thick[k] != 0) else 0.0
                else:
                    bu[k] = br[k] = 0.0
            out[f"bgc{t}_bu"], out[f"bgc{t}_br"] = bu, br
        return out

    # --- access --------------------------------------------------------------------------
    def array(self, name: str) -> np.ndarray:
        n = self.L.sam_array_len(self.h, name.encode())
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.float64)
        self.L.sam_get_array(self.h, name.encode(), _dp(out))
        return out

    def set_array(self, name: str, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        n = self.L.sam_array_len(self.h, name.encode())
        assert len(v) == n, (name, len(v), n)
        self.L.sam_set_array(self.h, name.encode(), _dp(v))

    def scalar(self, name: str) -> float:
        v = C.c_double()
        if self.L.sam_get_scalar(self.h, name.encode(), C.byref(v)) != 0:
            raise KeyError(name)
        return v.value

    def set_scalar(self, name: str, v: float):
        if self.L.sam_set_scalar(self.h, name.encode(), float(v)) != 0:
            raise KeyError(name)

    def int(self, name: str) -> int:
        v = C.c_int()
        if self.L.sam_get_int(self.h, name.encode(), C.byref(v)) != 0:
            raise KeyError(name)
        return v.value

    def set_int(self, name: str, v: int):
        if self.L.sam_set_int(self.h, name.encode(), int(v)) != 0:
            raise KeyError(name)

    def stat(self, name: str) -> int:
        return self.L.sam_get_stat(self.h, name.encode())

    def event_counts(self) -> dict:
        """how often each rarely taken branch of the path ran (names = samsim_b200.api.EVENT_NAMES)"""
        out, j = {}, 0
        while True:
            n = self.L.sam_event_name(j)
            if n is None:
                return out
            out[n.decode()] = self.stat("ev_" + n.decode())
            j += 1

    def events(self) -> set:
        return {n for n, v in self.event_counts().items() if v > 0}

    def state(self) -> dict:
        """Everything: arrays, double scalars and ints, by mo_data name."""
        d = {n: self.array(n) for n in ARRAY_NAMES}
        d.update({n: self.scalar(n) for n in SCALAR_NAMES})
        d.update({n: self.int(n) for n in INT_NAMES})
        return d

    def load_state(self, st: dict):
        for n in ARRAY_NAMES:
            if n in st:
                self.set_array(n, st[n])
        for n in SCALAR_NAMES:
            if n in st:
                self.set_scalar(n, st[n])
        for n in INT_NAMES:
            if n in st and n not in ("testcase", "Nlayer"):
                self.set_int(n, st[n])


def run_batch(cols: list[Column], nsteps: int, nthreads: int) -> int:
    """Advance many columns nsteps each, one column per OS thread (pthreads inside the library)."""
    L = cols[0].L
    arr = (C.c_void_p * len(cols))(*[c.h for c in cols])
    return L.sam_run_batch(arr, len(cols), nsteps, nthreads)


def read_forcing_dir(path: str | os.PathLike, n: int = 13148):
    """sub_input (mo_functions.f90:304-327): first n records of the four ASCII series."""
    p = Path(path)
    out = []
    for f in ("flux_sw.txt.input", "flux_lw.txt.input", "T2m.txt.input", "precip.txt.input"):
        out.append(np.loadtxt(p / f, dtype=np.float64)[:n].copy())
    return tuple(out)  # fl_sw, fl_lw, T2m, precip
