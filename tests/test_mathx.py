"""Accuracy of the constant-bank FP64 elementary functions (pfmds_b200/csrc/mathx.cuh), host build."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("mx") / "libmx.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "tests", "emu"), "-o", out, os.path.join(ROOT, "tests", "mathx_host.cpp")], check=True)
    L = C.CDLL(out)
    L.mx_exp.restype = C.c_double
    L.mx_exp.argtypes = [C.c_double]
    L.mx_rsqrt.restype = C.c_double
    L.mx_rsqrt.argtypes = [C.c_double]
    L.mx_sincos.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    return L


def test_exp(lib):
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.uniform(-700, 700, 20000), rng.uniform(-20, 20, 20000), rng.uniform(-1, 1, 5000), [0.0, -0.0, 1e-300, 700.0, -700.0]])
    got = np.array([lib.mx_exp(float(x)) for x in xs])
    ref = np.exp(xs.astype(np.longdouble)).astype(np.float64)
    assert np.max(np.abs(got - ref) / ref) < 3e-14


def test_sincos(lib):
    rng = np.random.default_rng(2)
    xs = np.concatenate([rng.uniform(0, np.pi, 40000), [0.0, np.pi / 4, np.pi / 2, 3 * np.pi / 4, np.pi, 3.14159265358979]])
    s, c = C.c_double(), C.c_double()
    err = 0.0
    for x in xs:
        lib.mx_sincos(float(x), C.byref(s), C.byref(c))
        err = max(err, abs(s.value - float(np.sin(np.longdouble(x)))), abs(c.value - float(np.cos(np.longdouble(x)))))
    assert err < 4e-16


def test_cos_switch(lib):
    lib.mx_cos_switch.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    rng = np.random.default_rng(4)
    xs = np.concatenate([rng.uniform(0, np.pi, 40000), [0.0, np.pi / 2, np.pi, 3.14159265358979, 1e-9, np.pi - 1e-9]])
    f, s = C.c_double(), C.c_double()
    err = 0.0
    for x in xs:
        lib.mx_cos_switch(float(x), C.byref(f), C.byref(s))
        xl = np.longdouble(x)
        err = max(err, abs(f.value - float((1 + np.cos(xl)) / 2)), abs(s.value - float(np.sin(xl))))
    assert err < 5e-16


def test_exp_nc(lib):
    lib.mx_exp_nc.restype = C.c_double
    lib.mx_exp_nc.argtypes = [C.c_double]
    xs = np.random.default_rng(5).uniform(-600, 600, 20000)
    got = np.array([lib.mx_exp_nc(float(x)) for x in xs])
    ref = np.exp(xs.astype(np.longdouble)).astype(np.float64)
    assert np.max(np.abs(got - ref) / ref) < 3e-14


def test_log_pow_rcp(lib):
    for f in (lib.mx_log, lib.mx_rcp):
        f.restype = C.c_double
        f.argtypes = [C.c_double]
    lib.mx_pow.restype = C.c_double
    lib.mx_pow.argtypes = [C.c_double, C.c_double]
    rng = np.random.default_rng(8)
    xs = np.concatenate([10.0 ** rng.uniform(-12, 12, 20000), rng.uniform(0.5, 2.0, 20000), 1 + rng.uniform(-1e-6, 1e-6, 2000)])
    got = np.array([lib.mx_log(float(x)) for x in xs])
    ref = np.log(xs.astype(np.longdouble)).astype(np.float64)
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3)) < 1e-15 * 1e3 and np.max(np.abs(got - ref)) < 1e-14
    got = np.array([lib.mx_rcp(float(x)) * x for x in xs])
    assert np.max(np.abs(got - 1)) < 5e-16
    xs = rng.uniform(1e-6, 3.0, 20000)
    ys = rng.uniform(-3, 3, 20000)
    got = np.array([lib.mx_pow(float(x), float(y)) for x, y in zip(xs, ys)])
    ref = (xs.astype(np.longdouble) ** ys.astype(np.longdouble)).astype(np.float64)
    assert np.max(np.abs(got - ref) / ref) < 1e-13
    assert lib.mx_pow(0.0, 2.0) == 0.0


def test_rsqrt(lib):
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.uniform(0.01, 200, 20000), 10.0 ** rng.uniform(-10, 10, 5000)])
    got = np.array([lib.mx_rsqrt(float(x)) for x in xs])
    ref = (1 / np.sqrt(xs.astype(np.longdouble))).astype(np.float64)
    assert np.max(np.abs(got - ref) / ref) < 5e-16


def test_short_forms(lib):
    """The reduced-instruction variants used by the second-generation rjl kernels: exp_m, rsqrt_q, cos_switch_m, half_switch."""
    for f in (lib.mx_exp_m, lib.mx_rsqrt_q, lib.mx_half_switch):
        f.restype = C.c_double
        f.argtypes = [C.c_double]
    lib.mx_cos_switch_m.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    rng = np.random.default_rng(11)
    xs = np.concatenate([rng.uniform(-40, 40, 40000), [0.0, -0.0, 0.34657359, -0.34657359, 0.34657360, 1e-300]])
    got = np.array([lib.mx_exp_m(float(x)) for x in xs])
    ref = np.exp(xs.astype(np.longdouble)).astype(np.float64)
    assert np.max(np.abs(got - ref) / ref) < 8e-15
    xs = rng.uniform(-690, 690, 20000)
    got = np.array([lib.mx_exp_m(float(x)) for x in xs])
    ref = np.exp(xs.astype(np.longdouble)).astype(np.float64)
    assert np.max(np.abs(got - ref) / ref) < 3e-14
    xs = np.concatenate([rng.uniform(0.01, 200, 20000), 10.0 ** rng.uniform(-10, 10, 5000)])
    got = np.array([lib.mx_rsqrt_q(float(x)) for x in xs])
    ref = (1 / np.sqrt(xs.astype(np.longdouble))).astype(np.float64)
    assert np.max(np.abs(got - ref) / ref) < 1e-11   # host stand-in seed is up to 1.9e-6 off (device: 9e-7 -> 1.3e-12)
    xs = np.concatenate([rng.uniform(0, np.pi, 40000), [0.0, np.pi / 2, np.pi, 3.14159265358979, 1e-9, np.pi - 1e-9]])
    f, s = C.c_double(), C.c_double()
    err = errh = 0.0
    for x in xs:
        lib.mx_cos_switch_m(float(x), C.byref(f), C.byref(s))
        xl = np.longdouble(x)
        err = max(err, abs(f.value - float((1 + np.cos(xl)) / 2)), abs(s.value - float(np.sin(xl))))
        errh = max(errh, abs(lib.mx_half_switch(float(x)) - float((1 + np.cos(xl)) / 2)))
    assert err < 2e-14 and errh < 1e-15, (err, errh)
