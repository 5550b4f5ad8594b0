"""CPU: the device generator's algorithm (csrc/qt_rng.cuh, compiled for the host) against numpy."""
import ctypes
import pathlib
import subprocess

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = tmp_path_factory.mktemp("rng") / "rng_shim.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so),
                    str(ROOT / "tests" / "helpers" / "rng_host_shim.cpp")], check=True)
    return ctypes.CDLL(str(so))


SEEDS = np.concatenate([np.arange(0, 2000, dtype=np.uint64),
                        np.array([2**31 - 1, 2**32 - 1, 2**32, 2**40 + 17, 2**63 + 5, 2**64 - 1], dtype=np.uint64)])


def test_seedsequence_pcg64_uniform_and_ziggurat_are_bit_exact(shim):
    E, nu, nn = len(SEEDS), 5, 40
    out = np.zeros((nu + nn, E))
    shim.rng_probe_host(SEEDS.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(E), nu, nn,
                        out.ctypes.data_as(ctypes.c_void_p))
    ref = np.zeros_like(out)
    for j, s in enumerate(SEEDS):
        g = np.random.default_rng(int(s))
        ref[:nu, j] = g.random(nu)
        ref[nu:, j] = g.standard_normal(nn)
    assert (out[:nu] != ref[:nu]).sum() == 0
    assert (out[nu:] != ref[nu:]).sum() == 0  # includes ~1 % wedge and tail draws


def test_jump_ahead_resumes_the_stream(shim):
    for seed, skip in [(42, 0), (42, 3), (7, 1000), (123456789012, 12345)]:
        out = np.zeros(5)
        shim.rng_resume_host(ctypes.c_uint64(seed), ctypes.c_uint64(skip), 5, out.ctypes.data_as(ctypes.c_void_p))
        np.testing.assert_array_equal(out, np.random.default_rng(seed).random(skip + 5)[skip:])


def test_numpy_norm_of_three_vector_is_the_fma_chain():
    """qt_env.cu normalises the linear-motion direction with fma(z,z,fma(y,y,x*x)); that is what
    np.linalg.norm computes for a 3-vector with the BLAS numpy ships (see oracle/qt_oracle.c)."""
    import math

    libm = ctypes.CDLL("libm.so.6")
    libm.fma.restype = ctypes.c_double
    libm.fma.argtypes = [ctypes.c_double] * 3
    rng = np.random.default_rng(5)
    X = rng.standard_normal((5000, 3))
    got = np.array([math.sqrt(libm.fma(x[2], x[2], libm.fma(x[1], x[1], x[0] * x[0]))) for x in X])
    ref = np.array([np.linalg.norm(x) for x in X])
    mism = int((got != ref).sum())
    if mism:
        pytest.skip(f"this host's BLAS orders the 3-term dot differently ({mism}/5000 differ by an ulp); "
                    "linear-motion directions may differ from numpy in the last bit here")
