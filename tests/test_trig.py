"""CPU: accuracy of the rollout's sine/cosine routines (csrc/qt_trig.cuh compiled for the host)."""
import ctypes
import pathlib
import subprocess

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = tmp_path_factory.mktemp("trig") / "trig_shim.so"
    subprocess.run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so),
                    str(ROOT / "tests" / "helpers" / "trig_host_shim.cpp")], check=True)
    return ctypes.CDLL(str(so))


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _exact(x):
    import mpmath as mp

    mp.mp.dps = 40
    return (np.array([float(mp.sin(mp.mpf(float(v)))) for v in x]), np.array([float(mp.cos(mp.mpf(float(v)))) for v in x]))


@pytest.mark.parametrize("lo,hi", [(-3.5, 3.5), (-200.0, 200.0), (-1e5, 1e5)])
def test_full_range_sincos_is_accurate_to_an_ulp(shim, lo, hi):
    rng = np.random.default_rng(1)
    x = rng.uniform(lo, hi, 3000)
    s, c = np.zeros_like(x), np.zeros_like(x)
    shim.trig_full_host(_ptr(x), ctypes.c_long(len(x)), _ptr(s), _ptr(c))
    es, ec = _exact(x)
    assert np.abs(s - es).max() < 1.7e-16 and np.abs(c - ec).max() < 1.7e-16


def test_special_arguments(shim):
    x = np.array([0.0, -0.0, np.pi, -np.pi, np.pi / 2, 1e-300, 1e9, np.nan, np.inf])
    s, c = np.zeros_like(x), np.zeros_like(x)
    shim.trig_full_host(_ptr(x), ctypes.c_long(len(x)), _ptr(s), _ptr(c))
    np.testing.assert_allclose(s[:7], np.sin(x[:7]), rtol=0, atol=2e-16)
    np.testing.assert_allclose(c[:7], np.cos(x[:7]), rtol=0, atol=2e-16)
    assert np.isnan(s[7:]).all() and np.isnan(c[7:]).all()


@pytest.mark.parametrize("small,bound", [(2, 0.0009765625), (1, 0.0078125), (3, 0.03125), (0, 0.26)])
def test_shifted_evaluation_matches_direct_evaluation(shim, small, bound):
    """sin/cos(fl(base + delta)) from sin/cos(base): the RK4 stage angles of qt_step.cuh."""
    rng = np.random.default_rng(2)
    base = rng.uniform(-3.3, 3.3, 4000)
    arg = base + rng.uniform(-bound, bound, 4000) * 0.999
    s, c = np.zeros_like(arg), np.zeros_like(arg)
    shim.trig_shifted_host(_ptr(base), _ptr(arg), ctypes.c_long(len(arg)), small, _ptr(s), _ptr(c))
    es, ec = _exact(arg)
    assert np.abs(s - es).max() < 2.5e-16 and np.abs(c - ec).max() < 2.5e-16
