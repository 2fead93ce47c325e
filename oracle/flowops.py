"""TEST INFRASTRUCTURE: ctypes face of oracle/flowops_oracle.c (numpy in / numpy out)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libflowops_oracle.so')
_lib = None


def build():
    subprocess.check_call(['make', '-s', '-C', _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, 'flowops_oracle.c')):
            build()
        _lib = C.CDLL(_SO)
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def correlation(in1, in2, pad=20, k=1, max_disp=20, s1=1, s2=2):
    n, c, h, w = in1.shape
    oc, oh, ow = C.c_int(), C.c_int(), C.c_int()
    lib().corr_out_shape(h, w, pad, k, max_disp, s1, s2, C.byref(oc), C.byref(oh), C.byref(ow))
    out = np.zeros((n, oc.value, oh.value, ow.value), np.float32)
    a, pa = _f(in1)
    b, pb = _f(in2)
    assert lib().correlation_forward(pa, pb, out.ctypes.data_as(C.POINTER(C.c_float)), n, c, h, w, pad, k, max_disp, s1, s2)
    return out


def resample2d(in1, flow, kernel_size=1):
    n, c, ih, iw = in1.shape
    _, _, h, w = flow.shape
    out = np.zeros((n, c, h, w), np.float32)
    a, pa = _f(in1)
    b, pb = _f(flow)
    assert lib().resample2d_forward(pa, pb, out.ctypes.data_as(C.POINTER(C.c_float)), n, c, h, w, ih, iw, kernel_size)
    return out


def channelnorm(x, norm_deg=2):
    n, c, h, w = x.shape
    out = np.zeros((n, 1, h, w), np.float32)
    a, pa = _f(x)
    assert lib().channelnorm_forward(pa, out.ctypes.data_as(C.POINTER(C.c_float)), n, c, h, w, norm_deg)
    return out
