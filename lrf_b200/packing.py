"""Host-side lossless packing: the encoded-object layout of the reference, byte for byte.

    combine(a, b)       = BE32(len(a)) ‖ a ‖ b                      lrf/compression/utils.py:246-265
    combine_bytes([..]) = left fold of combine                      :290-300
    E(matrix)           = combine(json{"num_fibers","mode","dtype"}, combine_bytes(zlib9 per column))  :354-390
    image               = combine(json(metadata), combine_bytes([E(U_y), E(V_y), ...]))  compression/qmf.py:288-290

This module is the host form (north_star: lossless packing on the host, timed separately — bench.py `host_pack`); the
device hands over factors already fiber-major, so each column is a contiguous slice and is compressed without a
transpose.  The same bytes also come from the GPU: `lrfb_qmf_pack_device` (lrf_b200/csrc/deflate9.cuh, zlib level 9
restated as a kernel; bench.py `pack`), which `compression.qmf_encode_batch` uses when the shape allows.
"""
from __future__ import annotations

import json
import zlib
from functools import reduce

import numpy as np


def _join2(a: bytes, b: bytes) -> bytes:
    if not isinstance(a, bytes) or not isinstance(b, bytes):
        raise TypeError("Both payload1 and payload2 must be bytes objects.")
    if len(a) > 0xFFFFFFFF:
        raise ValueError("payload1 is too large to encode.")
    return len(a).to_bytes(4, byteorder="big") + a + b


def combine_bytes(payloads) -> bytes:
    return reduce(_join2, payloads)


def separate_bytes(combined: bytes, num_payloads: int = 2):
    parts = []
    head = combined
    for _ in range(num_payloads - 1):
        if not isinstance(head, bytes):
            raise TypeError("Combined must be a bytes object.")
        if len(head) < 4:
            raise ValueError("Combined data is too short to decode.")
        n = int.from_bytes(head[:4], byteorder="big")
        head, tail = head[4 : 4 + n], head[4 + n :]
        parts.insert(0, tail)
    parts.insert(0, head)
    return tuple(parts)


def dict_to_bytes(d: dict) -> bytes:
    return json.dumps(d).encode("utf-8")


def bytes_to_dict(b: bytes) -> dict:
    return json.loads(b.decode("utf-8"))


def encode_fibers(fibers: np.ndarray, dtype_name: str = "int8", level: int = 9) -> bytes:
    """fibers: (R, rows) array whose row r holds column r of the (rows, R) factor matrix."""
    cols = [zlib.compress(np.ascontiguousarray(fibers[r]).tobytes(), level) for r in range(fibers.shape[0])]
    meta = {"num_fibers": int(fibers.shape[0]), "mode": "col", "dtype": dtype_name}
    return combine_bytes([dict_to_bytes(meta), combine_bytes(cols)])


def decode_fibers(blob: bytes) -> np.ndarray:
    """Inverse of encode_fibers → (R, rows) array (fiber-major)."""
    meta_b, body = separate_bytes(blob)
    meta = bytes_to_dict(meta_b)
    if meta.get("mode", "col") != "col":
        raise NotImplementedError("only mode='col' matrices occur on the codec path")
    cols = separate_bytes(body, meta["num_fibers"])
    return np.stack([np.frombuffer(zlib.decompress(c), dtype=np.dtype(meta["dtype"])) for c in cols], axis=0)


def encode_tensor_nd(arr: np.ndarray, level: int = 9) -> bytes:
    """The N-D branch of encode_tensor (lrf/compression/utils.py:429-455): one zlib stream of the raw buffer behind a
    {"shape", "dtype"} header — what the patch=False branches of qmf_encode emit for their 3-D factors."""
    arr = np.ascontiguousarray(arr)
    meta = {"shape": list(arr.shape), "dtype": str(arr.dtype)}
    return combine_bytes([dict_to_bytes(meta), zlib.compress(arr.tobytes(), level)])


def decode_tensor(blob: bytes) -> np.ndarray:
    """decode_tensor (lrf/compression/utils.py:458-490): column-wise matrices come back as (rows, R), N-D tensors in
    their stored shape."""
    meta_b, body = separate_bytes(blob)
    meta = bytes_to_dict(meta_b)
    if "num_fibers" in meta:
        return np.ascontiguousarray(decode_fibers(blob).T)
    return np.frombuffer(zlib.decompress(body), dtype=np.dtype(meta["dtype"])).reshape(meta["shape"])


def pack_qmf_record(record: np.ndarray, layout, metadata: dict) -> bytes:
    """One image's int8 factor record (lrfb_qmf_layout order) + metadata → encoded bytes."""
    blobs = []
    for pl in range(layout.n_planes):
        r, m, n = layout.rank[pl], layout.rows[pl], layout.cols
        u = record[layout.u_offset[pl] : layout.u_offset[pl] + r * m].reshape(r, m)
        v = record[layout.v_offset[pl] : layout.v_offset[pl] + r * n].reshape(r, n)
        blobs += [encode_fibers(u), encode_fibers(v)]
    return combine_bytes([dict_to_bytes(metadata), combine_bytes(blobs)])
