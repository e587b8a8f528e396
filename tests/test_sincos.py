"""Accuracy of the kernel core's sincos (ss_env_core.cuh sincos_d) against 80-bit libm.
The function is pure IEEE arithmetic with explicit FMAs, so this host build returns the
same bits as the sm_100a build."""
import ctypes

import numpy as np

from tests.hostsim.sim import hs


def _sincos(x):
    x = np.ascontiguousarray(x, np.float64)
    s, c = np.empty_like(x), np.empty_like(x)
    f = hs().hs_sincos
    f.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    f(x.ctypes.data, x.size, s.ctypes.data, c.ctypes.data)
    return s, c


def _ulp_err(got, x, fn):
    ref = fn(x.astype(np.longdouble))
    ulp = np.spacing(np.abs(ref.astype(np.float64))).astype(np.longdouble)
    return np.max(np.abs(got.astype(np.longdouble) - ref) / ulp)


def test_sincos_accuracy():
    rng = np.random.default_rng(0)
    xs = np.concatenate([
        rng.uniform(-700, 700, 400000),                    # rotations of a 2000-tick game: |r| <= 500
        rng.uniform(-4, 4, 200000),
        rng.uniform(-9e4, 9e4, 100000),
        np.arange(-2000, 2001) * 0.25,                      # the discrete look steps
        (np.arange(-400, 401) * (np.pi / 2)),               # multiples of pi/2 (worst cancellation)
        np.nextafter(np.arange(1, 400) * (np.pi / 2), 0), rng.uniform(-1e-6, 1e-6, 1000),
    ])
    s, c = _sincos(xs)
    es, ec = _ulp_err(s, xs, np.sin), _ulp_err(c, xs, np.cos)
    assert es < 1.6 and ec < 1.6, (es, ec)      # CUDA's own sincos is documented at <= 2 ulp


def test_sincos_exact_special_values_and_fallback():
    s, c = _sincos(np.array([0.0, -0.0, 2.5e5, -3.0e7, np.inf, np.nan]))
    assert s[0] == 0.0 and c[0] == 1.0 and s[1] == 0.0 and c[1] == 1.0
    assert s[2] == np.sin(2.5e5) and c[3] == np.cos(-3.0e7)          # library fallback path
    assert np.isnan(s[4]) and np.isnan(c[5])
