"""TEST TOOLING — builds lrf_b200/csrc with g++ -DLRFB_SIM (cuda_sim.h) and drives it through the
same C ABI with numpy arrays standing in for device memory.  It checks the kernels' index math and
arithmetic against the oracle on the CPU-only box; it is not a fallback (lrf_b200 never loads it)."""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

import numpy as np

from lrf_b200 import _cabi

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_CSRC = os.path.join(_ROOT, "lrf_b200", "csrc")
_SO = os.path.join(_HERE, "_build", "liblrfb_sim.so")
_lib = None


def build() -> str:
    srcs = glob.glob(os.path.join(_CSRC, "*.cu*")) + [os.path.join(_HERE, "cuda_sim.h"),
                                                       os.path.join(_ROOT, "include", "lrfb.h")]
    if os.path.exists(_SO) and os.path.getmtime(_SO) >= max(os.path.getmtime(s) for s in srcs):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    subprocess.check_call([
        "g++", "-std=c++20", "-O2", "-pthread", "-DLRFB_SIM", "-DLRFB_DEV", "-ffp-contract=off", "-fvisibility=hidden",
        "-fPIC", "-shared", "-I", _HERE, "-I", _CSRC, "-x", "c++", os.path.join(_CSRC, "lrfb_api.cu"),
        "-o", _SO, "-lz",
    ])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = _cabi.bind(C.CDLL(build()))
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def layout(cfg):
    L = _cabi.QmfLayout()
    _cabi.check(lib().lrfb_qmf_layout_query(C.byref(cfg), C.byref(L)), "layout", lib())
    return L


def encode(images: np.ndarray, cfg, inits=None, sign_flip=None, stop_after=0):
    """images (B,3,H,W) uint8/float32 → (factors (B, record_bytes) int8, workspace bytes, map, layout)."""
    images = np.ascontiguousarray(images)
    B = images.shape[0]
    L = layout(cfg)
    m = _cabi.QmfWorkspaceMap()
    _cabi.check(lib().lrfb_qmf_workspace_query(C.byref(cfg), B, C.byref(m)), "ws query", lib())
    ws = np.zeros(m.total_bytes + 256, np.uint8)
    fac = np.zeros((B, L.record_bytes), np.int8)
    dbg = _cabi.QmfDebug()
    keep = []
    dbg.stop_after = stop_after
    if inits is not None:
        for pl, (u0, v0) in enumerate(inits):
            u0 = np.ascontiguousarray(u0, np.float32)
            v0 = np.ascontiguousarray(v0, np.float32)
            keep += [u0, v0]
            dbg.d_init_u[pl] = u0.ctypes.data
            dbg.d_init_v[pl] = v0.ctypes.data
    if sign_flip is not None:
        for pl, s in enumerate(sign_flip):
            s = np.ascontiguousarray(s, np.int32)
            keep.append(s)
            dbg.d_sign_flip[pl] = s.ctypes.data
    rc = lib().lrfb_qmf_encode(C.byref(cfg), B, _ptr(images), _ptr(fac), _ptr(ws), m.total_bytes,
                               C.byref(dbg), None)
    _cabi.check(rc, "encode", lib())
    return fac, ws, m, L


def ws_view(ws, offset, dtype, shape):
    n = int(np.prod(shape))
    return np.frombuffer(ws, dtype=dtype, count=n, offset=int(offset)).reshape(shape)


def split_record(rec: np.ndarray, L):
    """One image's int8 record → [U0, V0, ...] as (rows, R) matrices (undoing the fiber-major layout)."""
    out = []
    for pl in range(L.n_planes):
        r = L.rank[pl]
        u = rec[L.u_offset[pl]: L.u_offset[pl] + L.rows[pl] * r].reshape(r, L.rows[pl]).T
        v = rec[L.v_offset[pl]: L.v_offset[pl] + L.cols * r].reshape(r, L.cols).T
        out += [u, v]
    return out


def decode(factors: np.ndarray, cfg):
    B = factors.shape[0]
    out = np.zeros((B, 3, cfg.height, cfg.width), np.uint8)
    factors = np.ascontiguousarray(factors)
    _cabi.check(lib().lrfb_qmf_decode(C.byref(cfg), B, _ptr(factors), _ptr(out), None), "decode", lib())
    return out


def factorize(x: np.ndarray, R, bounds=(-16, 15), num_iters=10, init=None, sign_flip=None):
    x = np.ascontiguousarray(x, np.float32)
    n_mat, M, N = x.shape
    wsb = lib().lrfb_factorize_workspace_bytes(n_mat, M, N, R)
    ws = np.zeros(wsb + 256, np.uint8)
    u = np.zeros((n_mat, M, R), np.float32)
    v = np.zeros((n_mat, N, R), np.float32)
    iu = iv = None
    if init is not None:
        iu = np.ascontiguousarray(init[0], np.float32)
        iv = np.ascontiguousarray(init[1], np.float32)
    sf = None if sign_flip is None else np.ascontiguousarray(sign_flip, np.int32)
    rc = lib().lrfb_factorize(_ptr(x), n_mat, M, N, R, bounds[0], bounds[1], num_iters, _ptr(u), _ptr(v),
                              _ptr(iu), _ptr(iv), _ptr(sf), _ptr(ws), wsb, None)
    _cabi.check(rc, "factorize", lib())
    return u, v


def sse(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    B = a.shape[0]
    out = np.zeros(B, np.uint64)
    per = a.size // B
    _cabi.check(lib().lrfb_sse_u8(_ptr(a), _ptr(b), per, B, _ptr(out), None), "sse", lib())
    return out


def svd_encode(images: np.ndarray, cfg, sign_flip=None):
    images = np.ascontiguousarray(images)
    B = images.shape[0]
    L = layout(cfg)
    m = _cabi.QmfWorkspaceMap()
    _cabi.check(lib().lrfb_qmf_workspace_query(C.byref(cfg), B, C.byref(m)), "ws query", lib())
    ws = np.zeros(m.total_bytes + 256, np.uint8)
    codes = np.zeros((B, L.record_bytes), np.uint8)
    qp = np.zeros((B, 4), np.float32)
    dbg = _cabi.QmfDebug()
    keep = None
    if sign_flip is not None:
        keep = np.ascontiguousarray(sign_flip, np.int32)
        dbg.d_sign_flip[0] = keep.ctypes.data
    rc = lib().lrfb_svd_encode(C.byref(cfg), B, _ptr(images), _ptr(codes), _ptr(qp), _ptr(ws), m.total_bytes,
                               C.byref(dbg), None)
    _cabi.check(rc, "svd_encode", lib())
    return codes, qp, ws, m, L


def svd_decode(codes: np.ndarray, qp6: np.ndarray, cfg):
    B = codes.shape[0]
    out = np.zeros((B, 3, cfg.height, cfg.width), np.uint8)
    codes = np.ascontiguousarray(codes)
    qp6 = np.ascontiguousarray(qp6, np.float32)
    _cabi.check(lib().lrfb_svd_decode(C.byref(cfg), B, _ptr(codes), _ptr(qp6), _ptr(out), None), "svd_decode", lib())
    return out


def pack_device(records: np.ndarray, cfg, meta: bytes):
    """lrfb_qmf_pack_device on the shim: records (B, record_bytes) int8 -> list of framed byte strings."""
    records = np.ascontiguousarray(records, np.int8)
    B = records.shape[0]
    wsb = lib().lrfb_qmf_pack_device_workspace(C.byref(cfg), B)
    assert wsb > 0, lib().lrfb_last_error()
    ws = np.zeros(wsb + 256, np.uint8)
    cap = B * lib().lrfb_qmf_pack_bound(C.byref(cfg), len(meta))
    blob = np.zeros(cap, np.uint8)
    offs = np.zeros(B + 1, np.int64)
    rc = lib().lrfb_qmf_pack_device(C.byref(cfg), B, _ptr(records), meta, len(meta), _ptr(blob), cap, _ptr(offs),
                                    _ptr(ws), wsb, None)
    _cabi.check(rc, "pack_device", lib())
    return [blob[offs[i]:offs[i + 1]].tobytes() for i in range(B)]


def deflate9_serial(data: bytes) -> bytes:
    fn = lib().lrfb_sim_deflate9_serial
    fn.restype, fn.argtypes = C.c_int64, [C.c_char_p, C.c_int32, C.c_void_p]
    out = np.zeros(len(data) + 64, np.uint8)
    n = fn(data, len(data), _ptr(out))
    return out[:n].tobytes()


def unpack_device(encoded, cfg):
    """lrfb_qmf_unpack_device on the shim: list of encoded streams -> (B, record_bytes) int8 records."""
    B = len(encoded)
    L = layout(cfg)
    offs = np.zeros(B + 1, np.int64)
    np.cumsum([len(e) for e in encoded], out=offs[1:])
    blob = np.frombuffer(b"".join(encoded), np.uint8).copy()
    wsb = lib().lrfb_qmf_unpack_device_workspace(C.byref(cfg), B)
    ws = np.zeros(wsb + 256, np.uint8)
    rec = np.zeros((B, L.record_bytes), np.int8)
    rc = lib().lrfb_qmf_unpack_device(C.byref(cfg), B, _ptr(blob), _ptr(offs), _ptr(rec), _ptr(ws), wsb, None)
    _cabi.check(rc, "unpack_device", lib())
    return rec
