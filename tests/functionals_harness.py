"""Host build of the order-dependent functionals (tests/native/functionals_host.cpp = opensmile_b200/csrc/functionals_seq.cuh compiled
with g++): the statements lane 0 of a contour's warp executes, callable on the CPU.  Test infrastructure."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_L = None


def lib():
    global _L
    if _L is None:
        so = "/tmp/osm_functionals_host_%d.so" % os.getuid()
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                               os.path.join(ROOT, "tests", "native", "functionals_host.cpp")])
        _L = C.CDLL(so)
    return _L


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _stats(x):
    x = np.ascontiguousarray(x, np.float32)
    return x, np.float32(x.min()), np.float32(x.max()), np.float32(x.astype(np.float64).sum() / len(x))


def segments(spec, x, period, norm):
    x, mn, mx, _ = _stats(x)
    out = np.zeros(8, np.float32)
    n = lib().fsh_segments(C.byref(spec), _fp(x), C.c_long(len(x)), C.c_float(mn), C.c_float(mx), C.c_float(period), norm, _fp(out))
    return out[:n]


def peaks2(spec, x, period, norm):
    x, mn, mx, mean = _stats(x)
    out = np.zeros(32, np.float32)
    n = lib().fsh_peaks2(C.byref(spec), _fp(x), C.c_long(len(x)), C.c_float(mn), C.c_float(mx), C.c_float(mean), C.c_float(period), norm, _fp(out))
    return out[:n]


def lpc(spec, x):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(20, np.float32)
    n = lib().fsh_lpc(C.byref(spec), _fp(x), C.c_long(len(x)), _fp(out))
    return out[:n]


def onset(spec, x, period, norm):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(8, np.float32)
    n = lib().fsh_onset(C.byref(spec), _fp(x), C.c_long(len(x)), C.c_float(period), norm, _fp(out))
    return out[:n]


def peaks(spec, x, period, norm):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(8, np.float32)
    n = lib().fsh_peaks(C.byref(spec), _fp(x), C.c_long(len(x)), C.c_float(period), norm, _fp(out))
    return out[:n]


def crossings(spec, x):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(4, np.float32)
    n = lib().fsh_crossings(C.byref(spec), _fp(x), C.c_long(len(x)), _fp(out))
    return out[:n]
