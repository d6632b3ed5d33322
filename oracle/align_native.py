"""ctypes binding of oracle/align_c.c (CPU ORACLE, test infrastructure only)."""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.LIB
        if not os.path.exists(path):
            path = _build.build()
        L = ctypes.CDLL(path)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        L.oracle_pair_cost.argtypes = [fp, fp] + [ctypes.c_int] * 4 + [fp]
        L.oracle_pair_cost.restype = None
        L.oracle_align_batch.argtypes = [fp, fp] + [ctypes.c_int] * 5 + [fp, ip, ip, ctypes.c_int]
        L.oracle_align_batch.restype = None
        u8p = ctypes.POINTER(ctypes.c_uint8)
        L.oracle_align_phase_batch.argtypes = ([fp, fp, u8p, u8p, ctypes.c_float] + [ctypes.c_int] * 5 +
                                               [fp, ip, ip, ctypes.c_int])
        L.oracle_align_phase_batch.restype = None
        _lib = L
    return _lib


def _fp(x):
    return x.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _ip(x):
    return x.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def pair_cost_c(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    Ta, V, Cc = a.shape
    Tb = b.shape[0]
    out = np.empty((Ta, Tb), dtype=np.float32)
    lib().oracle_pair_cost(_fp(a), _fp(b), Ta, Tb, V, Cc, _fp(out))
    return out


def align_batch_c(a, b, num_threads: int = 1):
    """a [N,Ta,V,Cc], b [N,Tb,V,Cc] -> cost [N] f32, path [N,Ta+Tb-1,2] i32 (-1 padded), path_len [N]."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    N, Ta, V, Cc = a.shape
    Tb = b.shape[1]
    cost = np.empty(N, dtype=np.float32)
    path = np.empty((N, Ta + Tb - 1, 2), dtype=np.int32)
    plen = np.empty(N, dtype=np.int32)
    lib().oracle_align_batch(_fp(a), _fp(b), N, Ta, Tb, V, Cc, _fp(cost), _ip(path), _ip(plen),
                             int(num_threads))
    return cost, path, plen


def align_phase_batch_c(a, b, la, lb, penalty: float, num_threads: int = 1):
    """align_batch_c with phase labels la [N,Ta], lb [N,Tb] (u8) and the mismatch penalty."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    la = np.ascontiguousarray(la, dtype=np.uint8)
    lb = np.ascontiguousarray(lb, dtype=np.uint8)
    N, Ta, V, Cc = a.shape
    Tb = b.shape[1]
    assert la.shape == (N, Ta) and lb.shape == (N, Tb)
    cost = np.empty(N, dtype=np.float32)
    path = np.empty((N, Ta + Tb - 1, 2), dtype=np.int32)
    plen = np.empty(N, dtype=np.int32)
    u8p = ctypes.POINTER(ctypes.c_uint8)
    lib().oracle_align_phase_batch(_fp(a), _fp(b), la.ctypes.data_as(u8p), lb.ctypes.data_as(u8p),
                                   ctypes.c_float(penalty), N, Ta, Tb, V, Cc, _fp(cost), _ip(path), _ip(plen),
                                   int(num_threads))
    return cost, path, plen
