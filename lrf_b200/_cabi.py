"""ctypes binding of liblrfb.so (include/lrfb.h) — the only door from Python into the CUDA kernels.

There is no CPU fallback: if the shared library is missing or no CUDA device is present, every
compute entry point raises.  Build the library with ``python -m lrf_b200.build`` (or
``__graft_entry__.build()``); it is kept in-tree at ``lrf_b200/csrc/build/liblrfb.so``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "build", "liblrfb.so")

LRFB_RGB, LRFB_YCBCR = 0, 1
LRFB_U8, LRFB_F32 = 0, 1


class QmfConfig(C.Structure):
    _fields_ = [
        ("height", C.c_int32), ("width", C.c_int32),
        ("patch_h", C.c_int32), ("patch_w", C.c_int32),
        ("color_space", C.c_int32), ("input_dtype", C.c_int32),
        ("scale_h", C.c_double), ("scale_w", C.c_double),
        ("rank", C.c_int32 * 3),
        ("bound_lo", C.c_float), ("bound_hi", C.c_float),
        ("num_iters", C.c_int32),
    ]


class QmfLayout(C.Structure):
    _fields_ = [
        ("n_planes", C.c_int32), ("cols", C.c_int32),
        ("orig_h", C.c_int32 * 3), ("orig_w", C.c_int32 * 3),
        ("pad_h", C.c_int32 * 3), ("pad_w", C.c_int32 * 3),
        ("rows", C.c_int32 * 3), ("rank", C.c_int32 * 3),
        ("u_offset", C.c_int64 * 3), ("v_offset", C.c_int64 * 3),
        ("record_bytes", C.c_int64), ("x_floats", C.c_int64),
    ]


class QmfWorkspaceMap(C.Structure):
    _fields_ = [
        ("x", C.c_int64 * 3), ("u", C.c_int64 * 3), ("v", C.c_int64 * 3),
        ("gram", C.c_int64 * 3), ("evec", C.c_int64 * 3), ("sigma", C.c_int64 * 3),
        ("total_bytes", C.c_int64),
    ]


class QmfDebug(C.Structure):
    _fields_ = [
        ("d_init_u", C.c_void_p * 3), ("d_init_v", C.c_void_p * 3),
        ("d_sign_flip", C.c_void_p * 3), ("stop_after", C.c_int32),
    ]


# name -> (restype, argtypes); every symbol include/lrfb.h declares
PROTOTYPES = {
    "lrfb_abi_version": (C.c_int32, []),
    "lrfb_last_error": (C.c_char_p, []),
    "lrfb_device_count": (C.c_int32, []),
    "lrfb_qmf_layout_query": (C.c_int32, [C.POINTER(QmfConfig), C.POINTER(QmfLayout)]),
    "lrfb_qmf_workspace_query": (C.c_int32, [C.POINTER(QmfConfig), C.c_int32, C.POINTER(QmfWorkspaceMap)]),
    "lrfb_qmf_encode": (C.c_int32, [C.POINTER(QmfConfig), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int64, C.POINTER(QmfDebug), C.c_void_p]),
    "lrfb_qmf_decode": (C.c_int32, [C.POINTER(QmfConfig), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lrfb_qmf_frontend": (C.c_int32, [C.POINTER(QmfConfig), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lrfb_factorize_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "lrfb_factorize": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                   C.c_float, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "lrfb_bcd": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float,
                             C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int64,
                             C.c_void_p]),
    "lrfb_launch_count": (C.c_int64, []),
    "lrfb_ffma_probe": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p]),
    "lrfb_svd_encode": (C.c_int32, [C.POINTER(QmfConfig), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int64, C.POINTER(QmfDebug), C.c_void_p]),
    "lrfb_svd_decode": (C.c_int32, [C.POINTER(QmfConfig), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lrfb_sse_u8": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "lrfb_ctx_create": (C.c_int32, [C.c_int32, C.POINTER(C.c_void_p)]),
    "lrfb_ctx_destroy": (None, [C.c_void_p]),
    "lrfb_qmf_encode_host": (C.c_int32, [C.c_void_p, C.POINTER(QmfConfig), C.c_int32, C.c_void_p, C.c_void_p]),
    "lrfb_qmf_decode_host": (C.c_int32, [C.c_void_p, C.POINTER(QmfConfig), C.c_int32, C.c_void_p, C.c_void_p]),
    "lrfb_ctx_set_chunk_bytes": (C.c_int32, [C.c_void_p, C.c_int64]),
    "lrfb_qmf_pack_bound": (C.c_int64, [C.POINTER(QmfConfig), C.c_int64]),
    "lrfb_qmf_pack_host": (C.c_int32, [C.POINTER(QmfConfig), C.c_int32, C.c_void_p, C.c_char_p, C.c_int64, C.c_void_p,
                                       C.c_int64, C.c_void_p, C.c_int32]),
    "lrfb_qmf_pack_device_workspace": (C.c_int64, [C.POINTER(QmfConfig), C.c_int32]),
    "lrfb_qmf_pack_device": (C.c_int32, [C.POINTER(QmfConfig), C.c_int32, C.c_void_p, C.c_char_p, C.c_int64, C.c_void_p,
                                         C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "lrfb_qmf_encode_bytes_host": (C.c_int32, [C.c_void_p, C.POINTER(QmfConfig), C.c_int32, C.c_void_p, C.c_char_p,
                                               C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "lrfb_qmf_unpack_device_workspace": (C.c_int64, [C.POINTER(QmfConfig), C.c_int32]),
    "lrfb_qmf_unpack_device": (C.c_int32, [C.POINTER(QmfConfig), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_int64, C.c_void_p]),
    "lrfb_qmf_decode_bytes_host": (C.c_int32, [C.c_void_p, C.POINTER(QmfConfig), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lrfb_debug_set": (C.c_int32, [C.c_char_p, C.c_int32]),
    "lrfb_qmf_decode_planes": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32),
                                           C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p,
                                           C.c_void_p]),
}


class LrfbError(RuntimeError):
    pass


def bind(lib: C.CDLL) -> C.CDLL:
    """Attach prototypes; raises AttributeError if a declared symbol is missing."""
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def lib() -> C.CDLL:
    """The product library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LrfbError(
                f"{LIB_PATH} not found: the CUDA extension is not built "
                "(run `python -m lrf_b200.build`); lrf_b200 has no CPU fallback"
            )
        _lib = bind(C.CDLL(LIB_PATH))
        if _lib.lrfb_abi_version() != 1:
            raise LrfbError("liblrfb.so ABI version mismatch")
    return _lib


def check(rc: int, what: str, library=None) -> None:
    if rc != 0:
        msg = (library or lib()).lrfb_last_error()
        raise LrfbError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def make_config(height, width, patch_size, color_space, input_dtype, scale_factor, ranks, bounds,
                num_iters) -> QmfConfig:
    cfg = QmfConfig()
    cfg.height, cfg.width = int(height), int(width)
    cfg.patch_h, cfg.patch_w = int(patch_size[0]), int(patch_size[1])
    cfg.color_space = LRFB_YCBCR if color_space == "YCbCr" else LRFB_RGB
    cfg.input_dtype = input_dtype
    cfg.scale_h, cfg.scale_w = float(scale_factor[0]), float(scale_factor[1])
    rk = list(ranks) + [0] * (3 - len(ranks))
    for i in range(3):
        cfg.rank[i] = int(rk[i])
    cfg.bound_lo, cfg.bound_hi = float(bounds[0]), float(bounds[1])
    cfg.num_iters = int(num_iters)
    return cfg
