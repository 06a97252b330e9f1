"""detmath.h (shared by the GPU and the oracle's det build) against glibc: < 1 ulp, exact special cases."""
import numpy as np


def _ulps(a, b):
    return np.abs(a - b) / np.spacing(np.abs(b))


def test_pow_exp_sin_within_one_ulp_of_glibc(oracle_mod):
    Ld, Lm = oracle_mod.lib("det"), oracle_mod.lib("libm")
    assert Ld.sam_math_backend() == b"det" and Lm.sam_math_backend() == b"libm"
    rng = np.random.default_rng(0)
    x = 10 ** rng.uniform(-6, 3.2, 6000)      # 1000*|psi_l| and salinities
    for y in (3.1, 1.5):
        a = np.array([Ld.sam_math_pow(float(v), y) for v in x])
        b = np.array([Lm.sam_math_pow(float(v), y) for v in x])
        assert _ulps(a, b).max() <= 1.0
    e = rng.uniform(-30, 5, 6000)
    a = np.array([Ld.sam_math_exp(float(v)) for v in e]); b = np.array([Lm.sam_math_exp(float(v)) for v in e])
    assert _ulps(a, b).max() <= 1.0
    s = rng.uniform(0, 30, 6000)              # sub_test4 phase: t*2pi/year over 4.5 years
    a = np.array([Ld.sam_math_sin(float(v)) for v in s]); b = np.array([Lm.sam_math_sin(float(v)) for v in s])
    assert np.abs(a - b).max() <= 2.3e-16


def test_special_cases(oracle_mod):
    L = oracle_mod.lib("det")
    assert L.sam_math_pow(0.0, 3.1) == 0.0          # perm of a fully solid layer
    assert L.sam_math_pow(1.0, 3.1) == 1.0
    assert L.sam_math_pow(4.0, 0.5) == 2.0
    assert L.sam_math_pow(2.0, 10.0) == 1024.0
    assert L.sam_math_pow(5e-324, 3.1) == 0.0       # underflow
    assert L.sam_math_exp(0.0) == 1.0
    assert L.sam_math_exp(-800.0) == 0.0
    assert L.sam_math_sin(0.0) == 0.0
