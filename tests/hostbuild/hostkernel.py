"""ctypes front-end of tests/hostbuild/kernel_on_host.cpp -- TEST HARNESS (the CUDA device functions compiled for the
host so that the no-GPU test stage can compare them with the oracle).  Not part of the product."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from samsim_b200 import api

HERE = Path(__file__).resolve().parent
SRC = HERE / "kernel_on_host.cpp"
LIB = HERE / "_build" / "libkernel_on_host.so"
CSRC = HERE.parent.parent / "samsim_b200" / "csrc"


def build() -> Path:
    deps = [SRC, CSRC / "step.cuh", CSRC / "physics.cuh", CSRC / "detmath.h", CSRC / "params.cuh",
            HERE.parent.parent / "include" / "samsim_b200.h"]
    if LIB.exists() and all(LIB.stat().st_mtime > d.stat().st_mtime for d in deps):
        return LIB
    LIB.parent.mkdir(exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-Wno-unknown-pragmas",
           "-o", str(LIB), str(SRC)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("host build of the device code failed:\n" + r.stderr)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(str(build()))
        dp = C.POINTER(C.c_double)
        L.hostk_create.restype = C.c_void_p
        L.hostk_create.argtypes = [C.POINTER(api._CConfig)]
        L.hostk_destroy.argtypes = [C.c_void_p]
        L.hostk_set_tuning.argtypes = [C.c_void_p, C.c_int]
        L.hostk_set_array.argtypes = [C.c_void_p, C.c_int, dp, C.c_int]
        L.hostk_get_array.argtypes = [C.c_void_p, C.c_int, dp, C.c_int]
        L.hostk_scalars.restype = dp
        L.hostk_scalars.argtypes = [C.c_void_p]
        L.hostk_set_ints.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.hostk_get_ints.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.hostk_set_clock.argtypes = [C.c_void_p, C.c_double, C.c_longlong, C.c_int, C.c_int]
        L.hostk_get_clock.argtypes = [C.c_void_p, dp, C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.hostk_set_forcing.argtypes = [C.c_void_p, C.c_int, dp, dp, dp]
        L.hostk_set_lab_forcing.argtypes = [C.c_void_p, C.c_longlong, dp]
        L.hostk_snapshot_scalars.restype = dp
        L.hostk_snapshot_scalars.argtypes = [C.c_void_p]
        L.hostk_get_snapshot_array.argtypes = [C.c_void_p, C.c_int, dp, C.c_int]
        ipp = C.POINTER(C.c_int)
        L.hostk_kat_getT.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp, ipp, ipp]
        L.hostk_step.restype = C.c_int
        L.hostk_step.argtypes = [C.c_void_p, C.c_longlong]
        _lib = L
    return _lib


class HostKernel:
    """One column advanced by the device code compiled for the host; the accessors mirror api.Engine's so that
    oracle.parity_util.compare_column can diff it against an oracle column."""

    def __init__(self, cfg: api.Config):
        self.L = lib()
        self.cfg = cfg
        self.ncol = 1
        cc = cfg.to_c()
        self.h = self.L.hostk_create(C.byref(cc))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.hostk_destroy(self.h)
            self.h = None

    def set_tuning(self, two_pass: bool):
        self.L.hostk_set_tuning(self.h, int(bool(two_pass)))

    def extent(self, name: str) -> int:
        N = self.cfg.Nlayer
        if name.startswith("bgc_abs"):
            return N if int(name[-1]) <= self.cfg.N_bgc else -1
        return N - 1 if name == "ray" else (N + 1 if name == "fl_Q" else N)

    def set_array(self, name, v):
        v = np.ascontiguousarray(v, dtype=np.float64).ravel()[: self.extent(name)]
        self.L.hostk_set_array(self.h, api.ARRAY_IDS[name], api._dp(v), len(v))

    def get_array(self, name, col=0, n=1):
        out = np.empty(self.extent(name))
        self.L.hostk_get_array(self.h, api.ARRAY_IDS[name], api._dp(out), len(out))
        return out[None, :]

    def _sc(self):
        return np.ctypeslib.as_array(self.L.hostk_scalars(self.h), shape=(len(api.SCALAR_IDS),))

    def set_scalar(self, name, v):
        self._sc()[api.SCALAR_IDS[name]] = float(np.atleast_1d(v)[0])

    def get_scalar(self, name, col=0, n=1):
        return np.array([self._sc()[api.SCALAR_IDS[name]]])

    def get_int(self, name, col=0, n=1):
        out = (C.c_int * 5)()
        self.L.hostk_get_ints(self.h, out)
        return np.array([out[api.INT_IDS[name]]], dtype=np.int32)

    def events(self, col: int = 0) -> set:
        return api.decode_events(self.get_int("events0")[0], self.get_int("events1")[0])

    def get_clock(self):
        t, i, n, tc = C.c_double(), C.c_longlong(), C.c_int(), C.c_int()
        self.L.hostk_get_clock(self.h, C.byref(t), C.byref(i), C.byref(n), C.byref(tc))
        return {"time": t.value, "i": i.value, "n_time_out": n.value, "time_counter": tc.value}

    def load_state(self, st: dict):
        for name in api.ARRAY_IDS:
            if name in st and self.extent(name) > 0:
                self.set_array(name, st[name])
        for name in api.SCALAR_IDS:
            if name in st:
                self.set_scalar(name, st[name])
        self.L.hostk_set_ints(self.h, int(st["N_active"]), int(st.get("status", 0)), int(st.get("styropor_flag", 0)))
        self.L.hostk_set_clock(self.h, float(st["time"]), int(st["i"]), int(st["n_time_out"]), max(int(st.get("time_counter", 1)), 1))

    def set_forcing(self, series, scale=None, offset=None):
        s = np.ascontiguousarray(series, dtype=np.float64)
        assert s.ndim == 2 and s.shape[0] == 4
        sc = None if scale is None else np.ascontiguousarray(scale, dtype=np.float64)
        of = None if offset is None else np.ascontiguousarray(offset, dtype=np.float64)
        self.L.hostk_set_forcing(self.h, s.shape[1], api._dp(s), api._dp(sc), api._dp(of))

    def set_lab_forcing(self, series):
        s = np.ascontiguousarray(series, dtype=np.float64)
        assert s.ndim == 2 and s.shape[0] == 4
        self.L.hostk_set_lab_forcing(self.h, s.shape[1], api._dp(s))

    def step(self, n: int) -> int:
        return self.L.hostk_step(self.h, int(n))

    def snapshot(self) -> dict:
        sc = np.ctypeslib.as_array(self.L.hostk_snapshot_scalars(self.h), shape=(len(api.SNAP_SCALARS),)).copy()
        out = {name: sc[j] for j, name in enumerate(api.SNAP_SCALARS)}
        for j, name in enumerate(api.SNAP_ARRAYS):
            a = np.empty(self.cfg.Nlayer)
            self.L.hostk_get_snapshot_array(self.h, j, api._dp(a), len(a))
            out[name] = a[: self.cfg.Nlayer - 1] if name == "ray" else a
        return out


def kat_getT(salt_flag: int, H, S_bu, T_in):
    """getT of the device code on the host, elementwise: T, phi, STOP code, word-1 event bits"""
    H, S_bu, T_in = (np.ascontiguousarray(x, dtype=np.float64) for x in (H, S_bu, T_in))
    n = len(H)
    T, phi = np.empty(n), np.empty(n)
    st, ev = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    lib().hostk_kat_getT(salt_flag, n, api._dp(H), api._dp(S_bu), api._dp(T_in), api._dp(T), api._dp(phi), ip(st), ip(ev))
    return T, phi, st, ev
