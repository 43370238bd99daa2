"""oracle/c_oracle.py -- TEST INFRASTRUCTURE.  numpy front-end of oracle/msda_oracle.c."""
import ctypes
import os

import numpy as np

from .build import C_LIB, build_c_oracle

_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(C_LIB):
            build_c_oracle()
        _lib = ctypes.CDLL(C_LIB)
        assert _lib.msda_oracle_abi_version() == 1
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _prep(value, shapes, lsi, loc, aw):
    dt = value.dtype
    assert dt in (np.float32, np.float64)
    value = np.ascontiguousarray(value, dt)
    loc = np.ascontiguousarray(loc, dt)
    aw = np.ascontiguousarray(aw, dt)
    shapes = np.ascontiguousarray(shapes, np.int64)
    lsi = np.ascontiguousarray(lsi, np.int64)
    n, s, m, d = value.shape
    _, lq, _, nl, p, _ = loc.shape
    dims = [ctypes.c_int(x) for x in (n, s, m, d, nl, lq, p)]
    return value, shapes, lsi, loc, aw, dims, "f32" if dt == np.float32 else "f64"


def forward(value, shapes, lsi, loc, aw):
    value, shapes, lsi, loc, aw, dims, sfx = _prep(value, shapes, lsi, loc, aw)
    n, s, m, d = value.shape
    lq = loc.shape[1]
    out = np.empty((n, lq, m * d), value.dtype)
    getattr(_load(), "msda_oracle_forward_" + sfx)(_ptr(value), _ptr(shapes), _ptr(lsi), _ptr(loc),
                                                   _ptr(aw), *dims, _ptr(out))
    return out


def backward(value, shapes, lsi, loc, aw, grad_out):
    value, shapes, lsi, loc, aw, dims, sfx = _prep(value, shapes, lsi, loc, aw)
    grad_out = np.ascontiguousarray(grad_out, value.dtype)
    gv, gl, ga = np.empty_like(value), np.empty_like(loc), np.empty_like(aw)
    getattr(_load(), "msda_oracle_backward_" + sfx)(_ptr(value), _ptr(shapes), _ptr(lsi), _ptr(loc),
                                                    _ptr(aw), _ptr(grad_out), *dims, _ptr(gv),
                                                    _ptr(gl), _ptr(ga))
    return gv, gl, ga
