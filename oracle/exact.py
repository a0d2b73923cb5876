"""TEST INFRASTRUCTURE — ctypes binding of oracle/qmf_exact.c (see that file's header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblrf_oracle.so")
_lib = None


class TieStats(ctypes.Structure):
    _fields_ = [("near_ties", ctypes.c_long), ("min_margin", ctypes.c_double)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "qmf_exact.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.lrfo_sse_u8.restype = ctypes.c_uint64
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def plan(H, W, patch=(8, 8), ycbcr=True, scale=(0.5, 0.5)):
    osz = (ctypes.c_int * 6)()
    psz = (ctypes.c_int * 6)()
    rows = (ctypes.c_int * 3)()
    lib().lrfo_plan(H, W, patch[0], patch[1], int(ycbcr), ctypes.c_double(scale[0]),
                    ctypes.c_double(scale[1]), osz, psz, rows)
    return ([(osz[2 * i], osz[2 * i + 1]) for i in range(3)],
            [(psz[2 * i], psz[2 * i + 1]) for i in range(3)], list(rows))


def frontend(rgb: np.ndarray, patch=(8, 8), ycbcr=True, scale=(0.5, 0.5)):
    """rgb (3,H,W) uint8 or float32 → list of patch matrices (M,N) float32."""
    _, H, W = rgb.shape
    _, _, rows = plan(H, W, patch, ycbcr, scale)
    n = patch[0] * patch[1] * (1 if ycbcr else 3)
    xs = [np.zeros((rows[i], n), np.float32) for i in range(3 if ycbcr else 1)]
    ptrs = [_p(x, ctypes.c_float) for x in xs] + [None] * (3 - len(xs))
    rgb = np.ascontiguousarray(rgb)
    if rgb.dtype == np.uint8:
        rc = lib().lrfo_frontend_u8(_p(rgb, ctypes.c_uint8), H, W, patch[0], patch[1], int(ycbcr),
                                    ctypes.c_double(scale[0]), ctypes.c_double(scale[1]), *ptrs)
    else:
        rgb = rgb.astype(np.float32, copy=False)
        rc = lib().lrfo_frontend_f32(_p(rgb, ctypes.c_float), H, W, patch[0], patch[1], int(ycbcr),
                                     ctypes.c_double(scale[0]), ctypes.c_double(scale[1]), *ptrs)
    assert rc == 0
    return xs


def bcd(x: np.ndarray, u0: np.ndarray, v0: np.ndarray, bounds=(-16, 15), num_iters=10,
        tie_window=1e-5):
    """Exact-arithmetic BCD from an injected init → (U, V) float32 integer-valued, TieStats."""
    x = np.ascontiguousarray(x, np.float32)
    u = np.ascontiguousarray(u0, np.float32).copy()
    v = np.ascontiguousarray(v0, np.float32).copy()
    M, N = x.shape
    R = u.shape[1]
    st = TieStats(0, 1.0)
    rc = lib().lrfo_bcd(_p(x, ctypes.c_float), M, N, R, _p(u, ctypes.c_float), _p(v, ctypes.c_float),
                        ctypes.c_float(bounds[0]), ctypes.c_float(bounds[1]), num_iters,
                        ctypes.byref(st), ctypes.c_double(tie_window))
    assert rc == 0
    return u, v, st


def half_sweep(x, u, v, which, bounds=(-16, 15), tie_window=1e-5):
    x = np.ascontiguousarray(x, np.float32)
    u = np.ascontiguousarray(u, np.float32).copy()
    v = np.ascontiguousarray(v, np.float32).copy()
    M, N = x.shape
    st = TieStats(0, 1.0)
    rc = lib().lrfo_half_sweep(_p(x, ctypes.c_float), M, N, u.shape[1], _p(u, ctypes.c_float),
                               _p(v, ctypes.c_float), ctypes.c_float(bounds[0]),
                               ctypes.c_float(bounds[1]), which, ctypes.byref(st),
                               ctypes.c_double(tie_window))
    assert rc == 0
    return u, v, st


def decode(factors, ranks, H, W, patch=(8, 8), ycbcr=True, scale=(0.5, 0.5)):
    """factors: list of int8 (rows,R) arrays [U0,V0,(U1,V1,U2,V2)] → uint8 (3,H,W)."""
    fs = [np.ascontiguousarray(f, np.int8) for f in factors]
    arr = (ctypes.POINTER(ctypes.c_int8) * 6)()
    for i, f in enumerate(fs):
        arr[i] = _p(f, ctypes.c_int8)
    rk = (ctypes.c_int * 3)(*(list(ranks) + [0, 0, 0])[:3])
    out = np.zeros((3, H, W), np.uint8)
    rc = lib().lrfo_decode_u8(arr, rk, H, W, patch[0], patch[1], int(ycbcr), ctypes.c_double(scale[0]),
                              ctypes.c_double(scale[1]), _p(out, ctypes.c_uint8))
    assert rc == 0
    return out


def sse_u8(a: np.ndarray, b: np.ndarray) -> int:
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return int(lib().lrfo_sse_u8(_p(a, ctypes.c_uint8), _p(b, ctypes.c_uint8), ctypes.c_size_t(a.size)))
