"""Host-side mirror of the reference codec interface for the QMF hot path.

Same names, argument meaning, error behaviour and encoded-bytes layout as
``lrf.qmf_encode`` / ``lrf.qmf_decode`` (lrf/compression/qmf.py:116-353), so either decoder reads
either encoder's output.  The arithmetic between ``image.float()`` and the int8 factors — and between
the int8 factors and the uint8 image — runs in the sm_100a kernels behind ``liblrfb.so``; there is no
CPU fallback.  Lossless byte packing (zlib) stays on the host (lrf_b200/packing.py).

The reference has no batch API; ``qmf_encode_batch`` / ``qmf_decode_batch`` add one (single image =
batch of one).
"""
from __future__ import annotations

import ctypes as C
import math
from collections.abc import Iterable
from concurrent.futures import ThreadPoolExecutor
from typing import Optional

import numpy as np
import torch

from . import _cabi, packing

_POOL: Optional[ThreadPoolExecutor] = None


def _pool() -> ThreadPoolExecutor:
    global _POOL
    if _POOL is None:
        import os

        _POOL = ThreadPoolExecutor(max_workers=max(1, os.cpu_count() or 1))
    return _POOL


def pack_records(records: np.ndarray, cfg, lay, meta: dict, threads: int = 0) -> list[bytes]:
    """int8 factor records (B, record_bytes) on the host + metadata → B encoded ``bytes``, on the native
    thread pool of liblrfb.so (``lrfb_qmf_pack_host``: zlib level 9 per factor column and the reference's
    byte framing, lrf/compression/utils.py:246-300, :354-390; identical output to packing.pack_qmf_record)."""
    records = np.ascontiguousarray(records, np.int8)
    B = records.shape[0]
    assert records.ndim == 2 and records.shape[1] == lay.record_bytes
    mj = packing.dict_to_bytes(meta)
    lib = _cabi.lib()
    stride = int(lib.lrfb_qmf_pack_bound(C.byref(cfg), len(mj)))
    if stride <= 0:
        raise _cabi.LrfbError("lrfb_qmf_pack_bound failed")
    out = np.empty((B, stride), np.uint8)
    sizes = np.zeros(B, np.int64)
    rc = lib.lrfb_qmf_pack_host(C.byref(cfg), B, records.ctypes.data_as(C.c_void_p), mj, len(mj),
                                out.ctypes.data_as(C.c_void_p), stride, sizes.ctypes.data_as(C.c_void_p), threads)
    _cabi.check(rc, "lrfb_qmf_pack_host")
    return [out[i, : sizes[i]].tobytes() for i in range(B)]


def device_packer_supported(cfg, batch: int = 1) -> bool:
    """True when every factor column of this shape fits the device deflate (at most 65 024 bytes per column)."""
    return int(_cabi.lib().lrfb_qmf_pack_device_workspace(C.byref(cfg), batch)) > 0


def pack_records_device(records: torch.Tensor, cfg, lay, meta: dict) -> list[bytes]:
    """int8 factor records (B, record_bytes) ON THE DEVICE + metadata → B encoded ``bytes``: the zlib level-9 streams
    and the reference's framing are produced by ``lrfb_qmf_pack_device`` (one warp per factor column, byte-identical to
    zlib: lrf_b200/csrc/deflate9.cuh); only the finished streams (about a fifth of the raw factors) cross PCIe."""
    assert records.is_cuda and records.dtype == torch.int8 and records.ndim == 2 and records.shape[1] == lay.record_bytes
    records = records.contiguous()
    B = records.shape[0]
    mj = packing.dict_to_bytes(meta)
    lib = _cabi.lib()
    wsb = int(lib.lrfb_qmf_pack_device_workspace(C.byref(cfg), B))
    if wsb <= 0:
        raise _cabi.LrfbError("lrfb_qmf_pack_device: " + (lib.lrfb_last_error() or b"unsupported shape").decode())
    cap = B * int(lib.lrfb_qmf_pack_bound(C.byref(cfg), len(mj)))
    with torch.cuda.device(records.device):
        ws = torch.empty(wsb, dtype=torch.uint8, device=records.device)
        blob = torch.empty(cap, dtype=torch.uint8, device=records.device)
        offs = torch.empty(B + 1, dtype=torch.int64, device=records.device)
        rc = lib.lrfb_qmf_pack_device(C.byref(cfg), B, C.c_void_p(records.data_ptr()), mj, len(mj),
                                      C.c_void_p(blob.data_ptr()), cap, C.c_void_p(offs.data_ptr()),
                                      C.c_void_p(ws.data_ptr()), wsb, _stream_ptr())
        _cabi.check(rc, "lrfb_qmf_pack_device")
        o = offs.cpu().numpy()
        assert o[B] <= cap
        host = blob[: int(o[B])].cpu().numpy()
    mv = memoryview(host)
    return [bytes(mv[o[i] : o[i + 1]]) for i in range(B)]


class _HostPipe:
    """Per-device context of the host-buffer C-ABI calls plus grow-only pinned output buffers."""

    def __init__(self, index: int):
        self.ctx = C.c_void_p()
        _cabi.check(_cabi.lib().lrfb_ctx_create(index, C.byref(self.ctx)), "lrfb_ctx_create")
        self.blob = None
        self.offsets = None

    def buffers(self, cap: int, batch: int):
        if self.blob is None or self.blob.numel() < cap:
            self.blob = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
        if self.offsets is None or self.offsets.numel() < batch + 1:
            self.offsets = torch.empty(batch + 1, dtype=torch.int64, pin_memory=True)
        return self.blob, self.offsets


_PIPES: dict[int, _HostPipe] = {}


def encode_bytes_host(images: torch.Tensor, cfg, lay, meta: dict, device_index: Optional[int] = None) -> list[bytes]:
    """Host images (B,3,H,W) -> B encoded ``bytes`` through ``lrfb_qmf_encode_bytes_host``: chunked H2D | encode kernels |
    device deflate | D2H of the finished streams.  ``images`` should be pinned for full copy bandwidth."""
    assert not images.is_cuda and images.is_contiguous()
    index = torch.cuda.current_device() if device_index is None else device_index
    pipe = _PIPES.get(index)
    if pipe is None:
        pipe = _PIPES[index] = _HostPipe(index)
    B = images.shape[0]
    mj = packing.dict_to_bytes(meta)
    lib = _cabi.lib()
    cap = B * int(lib.lrfb_qmf_pack_bound(C.byref(cfg), len(mj)))
    blob, offs = pipe.buffers(cap, B)
    rc = lib.lrfb_qmf_encode_bytes_host(pipe.ctx, C.byref(cfg), B, C.c_void_p(images.data_ptr()), mj, len(mj),
                                        C.c_void_p(blob.data_ptr()), blob.numel(), C.c_void_p(offs.data_ptr()))
    _cabi.check(rc, "lrfb_qmf_encode_bytes_host")
    o = offs.numpy()
    mv = memoryview(blob.numpy())
    return [bytes(mv[o[i] : o[i + 1]]) for i in range(B)]


DEVICE_UNPACK = True  # qmf_decode_batch un-frames and inflates on the device (lrfb_qmf_unpack_device) instead of per image on the host


def unpack_records_device(encoded: list[bytes], cfg, lay, device) -> torch.Tensor:
    """B encoded streams of one shape -> int8 factor records (B, record_bytes) on the device: the framing is walked and
    every zlib stream inflated by ``lrfb_qmf_unpack_device`` (lrf_b200/csrc/inflate9.cuh); only the streams cross PCIe.
    Raises LrfbError naming the first malformed image."""
    B = len(encoded)
    sizes = np.fromiter((len(e) for e in encoded), np.int64, B)
    offs = np.zeros(B + 1, np.int64)
    np.cumsum(sizes, out=offs[1:])
    blob = torch.frombuffer(bytearray(b"".join(encoded)), dtype=torch.uint8)
    lib = _cabi.lib()
    wsb = int(lib.lrfb_qmf_unpack_device_workspace(C.byref(cfg), B))
    if wsb <= 0:
        raise _cabi.LrfbError("lrfb_qmf_unpack_device_workspace failed")
    d_blob = blob.to(device)
    d_offs = torch.from_numpy(offs).to(device)
    ws = torch.empty(wsb, dtype=torch.uint8, device=device)
    rec = torch.empty((B, lay.record_bytes), dtype=torch.int8, device=device)
    rc = lib.lrfb_qmf_unpack_device(C.byref(cfg), B, C.c_void_p(d_blob.data_ptr()), C.c_void_p(d_offs.data_ptr()),
                                    C.c_void_p(rec.data_ptr()), C.c_void_p(ws.data_ptr()), wsb, _stream_ptr())
    _cabi.check(rc, "lrfb_qmf_unpack_device")
    return rec


def _require_cuda() -> None:
    if not torch.cuda.is_available():
        raise _cabi.LrfbError("lrf_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def _stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def qmf_rank(size: tuple[int, int], com_ratio: float) -> int:
    """lrf/compression/qmf.py:25-40."""
    num_rows, num_cols = size
    return max(math.floor(num_rows * num_cols / (com_ratio * (num_rows + num_cols))), 1)


def _triple(value, halve):
    if isinstance(value, Iterable):
        return tuple(value)
    if value is None:
        return (None, None, None)
    return (value, halve(value), halve(value))


def resolve_plan(height, width, rank, quality, color_space, scale_factor, patch_size, bounds, num_iters,
                 input_dtype=_cabi.LRFB_U8):
    """Apply the reference's rank rule (compression/qmf.py:215-225, :244-250) → (config, layout)."""
    probe = _cabi.make_config(height, width, patch_size, color_space, input_dtype, scale_factor, (1, 1, 1),
                              bounds, num_iters)
    lay = _cabi.QmfLayout()
    _cabi.check(_cabi.lib().lrfb_qmf_layout_query(C.byref(probe), C.byref(lay)), "lrfb_qmf_layout_query")
    if color_space == "RGB":
        if rank is None:
            assert quality >= 0 and quality <= 100, "'quality' must be between 0 and 100."
            ranks = [max(round(min(lay.rows[0], lay.cols) * quality / 100), 1)]
        else:
            ranks = [rank]
    else:
        rk = _triple(rank, lambda r: max(r // 2, 1))
        ql = _triple(quality, lambda q: q / 2)
        ranks = []
        for i in range(3):
            if rk[i] is None:
                assert ql[i] >= 0 and ql[i] <= 100, "'quality' must be between 0 and 100."
                ranks.append(max(round(min(lay.rows[i], lay.cols) * ql[i] / 100), 1))
            else:
                ranks.append(rk[i])
    cfg = _cabi.make_config(height, width, patch_size, color_space, input_dtype, scale_factor, ranks, bounds,
                            num_iters)
    _cabi.check(_cabi.lib().lrfb_qmf_layout_query(C.byref(cfg), C.byref(lay)), "lrfb_qmf_layout_query")
    return cfg, lay


def _metadata(image_dtype, color_space, patch, bounds, patch_size, lay) -> dict:
    """The JSON header, key order as the reference inserts them (compression/qmf.py:157-162, :180-187,
    :233-254)."""
    meta = {"dtype": str(image_dtype).split(".")[-1], "color space": color_space, "patch": patch,
            "bounds": bounds}
    if color_space == "RGB":
        meta.update({"patch size": patch_size, "original size": [lay.orig_h[0], lay.orig_w[0]],
                     "padded size": [lay.pad_h[0], lay.pad_w[0]], "rank": lay.rank[0]})
    else:
        meta["patch size"] = patch_size
        meta["original size"] = [[lay.orig_h[i], lay.orig_w[i]] for i in range(3)]
        meta["padded size"] = [[lay.pad_h[i], lay.pad_w[i]] for i in range(3)]
        meta["rank"] = [lay.rank[i] for i in range(3)]
    return meta


class EncodePlan:
    """Resolved configuration + device buffers for a fixed (shape, batch); reusable across calls."""

    def __init__(self, cfg, lay, batch: int, device):
        self.cfg, self.lay, self.batch, self.device = cfg, lay, batch, device
        m = _cabi.QmfWorkspaceMap()
        _cabi.check(_cabi.lib().lrfb_qmf_workspace_query(C.byref(cfg), batch, C.byref(m)),
                    "lrfb_qmf_workspace_query")
        self.map = m
        self.workspace = torch.empty(m.total_bytes, dtype=torch.uint8, device=device)
        self.factors = torch.empty((batch, lay.record_bytes), dtype=torch.int8, device=device)

    def run(self, images: torch.Tensor, debug: Optional[_cabi.QmfDebug] = None) -> torch.Tensor:
        """images: (batch,3,H,W) contiguous on self.device, uint8 or float32 → int8 records on device."""
        assert images.is_cuda and images.is_contiguous() and images.shape[0] == self.batch
        rc = _cabi.lib().lrfb_qmf_encode(
            C.byref(self.cfg), self.batch, C.c_void_p(images.data_ptr()), C.c_void_p(self.factors.data_ptr()),
            C.c_void_p(self.workspace.data_ptr()), self.map.total_bytes,
            C.byref(debug) if debug is not None else None, _stream_ptr())
        _cabi.check(rc, "lrfb_qmf_encode")
        return self.factors

    def view(self, name: str, plane: int) -> torch.Tensor:
        """A typed view into the workspace (tests / stage-level inspection)."""
        lay, b = self.lay, self.batch
        shapes = {
            "x": (torch.float32, (b, lay.rows[plane], lay.cols)),
            "u": (torch.float32, (b, lay.rows[plane], lay.rank[plane])),
            "v": (torch.float32, (b, lay.cols, lay.rank[plane])),
            "gram": (torch.float64, (b, lay.cols, lay.cols)),
            "evec": (torch.float64, (b, lay.cols, lay.rank[plane])),
            "sigma": (torch.float64, (b, lay.rank[plane])),
        }
        dt, shape = shapes[name]
        off = getattr(self.map, name)[plane]
        n = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
        return self.workspace[off : off + n].view(dt).view(shape)


def _check_encode_args(rank, quality, color_space, patch, dtype, kwargs, bounds=None):
    assert (rank, quality) != (None, None), "Either 'rank' or 'quality' must be specified."
    assert color_space in ("RGB", "YCbCr"), "`color_space` must be one of 'RGB' or 'YCbCr'."
    if dtype != torch.int8:
        raise NotImplementedError("lrf_b200: only dtype=torch.int8 factors are implemented")
    if bounds is not None and (math.ceil(bounds[0]) < -128 or math.floor(bounds[1]) > 127):
        raise NotImplementedError("lrf_b200: bounds outside the int8 range are not implemented")
    extra = {k: v for k, v in kwargs.items() if k not in ("num_iters", "verbose")}
    for k, v in extra.items():
        if (k in ("l2", "l1_ratio") and v == 0) or (k == "num_levels" and v is None) or (k == "eps" and v == 1e-16):
            continue
        raise NotImplementedError(f"lrf_b200: QMF option {k}={v!r} is not implemented on the CUDA path")


def _to_device_images(images: torch.Tensor, device) -> tuple[torch.Tensor, int]:
    if images.dtype == torch.uint8:
        return images.to(device, non_blocking=True).contiguous(), _cabi.LRFB_U8
    return images.to(device).float().contiguous(), _cabi.LRFB_F32


def qmf_encode_batch(images: torch.Tensor, rank=None, quality=None, color_space: str = "YCbCr",
                     scale_factor=(0.5, 0.5), patch: bool = True, patch_size=(8, 8), bounds=(-16, 15),
                     dtype: torch.dtype = torch.int8, return_records: bool = False, **kwargs):
    """Batched ``qmf_encode``: images (B,3,H,W) → list of B encoded ``bytes`` (or, with
    ``return_records``, the raw int8 factor records on the device plus the layout / metadata)."""
    _check_encode_args(rank, quality, color_space, patch, dtype, kwargs, bounds)
    _require_cuda()
    assert images.ndim == 4 and images.shape[1] == 3, "images must be (B, 3, H, W)"
    if not patch:
        assert not return_records, "patch=False factors are not stored as patch records"
        return _qmf_encode_nopatch(images, rank, quality, color_space, scale_factor, bounds, kwargs.get("num_iters", 10))
    device = images.device if images.is_cuda else torch.device("cuda", torch.cuda.current_device())
    B, _, H, W = images.shape
    if (not images.is_cuda and not return_records and images.is_pinned() and images.is_contiguous()
            and images.dtype in (torch.uint8, torch.float32)):
        in_dtype = _cabi.LRFB_U8 if images.dtype == torch.uint8 else _cabi.LRFB_F32
        cfg, lay = resolve_plan(H, W, rank, quality, color_space, scale_factor, patch_size, bounds,
                                kwargs.get("num_iters", 10), in_dtype)
        if device_packer_supported(cfg, B):  # pinned host batch: the chunked host pipeline ends in bytes
            meta = _metadata(images.dtype, color_space, patch, bounds, patch_size, lay)
            return encode_bytes_host(images, cfg, lay, meta, device.index)
    dev_images, in_dtype = _to_device_images(images, device)
    cfg, lay = resolve_plan(H, W, rank, quality, color_space, scale_factor, patch_size, bounds,
                            kwargs.get("num_iters", 10), in_dtype)
    with torch.cuda.device(device):
        plan = EncodePlan(cfg, lay, B, device)
        records = plan.run(dev_images)
        meta = _metadata(images.dtype, color_space, patch, bounds, patch_size, lay)
        if return_records:
            return records, lay, meta
        if device_packer_supported(cfg, B):
            return pack_records_device(records, cfg, lay, meta)
        host = records.cpu().numpy()
    return pack_records(host, cfg, lay, meta)  # columns above 65 024 bytes: zlib on the host thread pool


def qmf_encode(image: torch.Tensor, rank=None, quality=None, color_space: str = "YCbCr",
               scale_factor=(0.5, 0.5), patch: bool = True, patch_size=(8, 8), bounds=(-16, 15),
               dtype: torch.dtype = torch.int8, **kwargs) -> bytes:
    """Drop-in for ``lrf.qmf_encode`` (lrf/compression/qmf.py:116-292)."""
    _check_encode_args(rank, quality, color_space, patch, dtype, kwargs)
    return qmf_encode_batch(image.unsqueeze(0), rank, quality, color_space, scale_factor, patch, patch_size,
                            bounds, dtype, **kwargs)[0]


def _rank_rule(rows: int, cols: int, quality) -> int:
    assert quality >= 0 and quality <= 100, "'quality' must be between 0 and 100."
    return max(round(min(rows, cols) * quality / 100), 1)


def _qmf_encode_nopatch(images, rank, quality, color_space, scale_factor, bounds, num_iters) -> list[bytes]:
    """The patch=False branches of ``qmf_encode`` (lrf/compression/qmf.py:195-212 RGB, :264-286 YCbCr): every channel
    is factorised as one H x W matrix.  Planes come from the front-end kernels (1 x 1 "patches" are the planes
    themselves), the factorisation from lrfb_factorize (generic kernels: N up to 1024, R up to 64); the 3-D factors
    keep their batch dimension and go through the N-D branch of encode_tensor, as in the reference."""
    from .factorization import QMF

    device = images.device if images.is_cuda else torch.device("cuda", torch.cuda.current_device())
    dev_images, in_dtype = _to_device_images(images, device)
    B, _, H, W = images.shape
    meta = {"dtype": str(images.dtype).split(".")[-1], "color space": color_space, "patch": False, "bounds": bounds}
    with torch.cuda.device(device):
        if color_space == "RGB":
            R = _rank_rule(H, W, quality) if rank is None else rank
            meta["rank"] = R
            u, v, _ = QMF(rank=R, num_iters=num_iters, bounds=bounds).decompose(dev_images.float())
            fac = [(u.to(torch.int8).cpu().numpy(), v.to(torch.int8).cpu().numpy())]  # (B,3,H,R), (B,3,W,R)
        else:
            cfg = _cabi.make_config(H, W, (1, 1), "YCbCr", in_dtype, scale_factor, (1, 1, 1), bounds, 1)
            lay, m = _cabi.QmfLayout(), _cabi.QmfWorkspaceMap()
            _cabi.check(_cabi.lib().lrfb_qmf_layout_query(C.byref(cfg), C.byref(lay)), "lrfb_qmf_layout_query")
            _cabi.check(_cabi.lib().lrfb_qmf_workspace_query(C.byref(cfg), B, C.byref(m)), "lrfb_qmf_workspace_query")
            x = torch.empty(m.u[0] - m.x[0], dtype=torch.uint8, device=device)  # the three planes, plane-major
            rc = _cabi.lib().lrfb_qmf_frontend(C.byref(cfg), B, C.c_void_p(dev_images.data_ptr()),
                                               C.c_void_p(x.data_ptr()), _stream_ptr())
            _cabi.check(rc, "lrfb_qmf_frontend")
            rk = _triple(rank, lambda r: max(r // 2, 1))
            ql = _triple(quality, lambda q: q / 2)
            meta["original size"], meta["rank"], fac = [], [], []
            for pl in range(3):
                h, w = lay.orig_h[pl], lay.orig_w[pl]
                R = _rank_rule(h, w, ql[pl]) if rk[pl] is None else rk[pl]
                meta["original size"].append([h, w])
                meta["rank"].append(R)
                off = m.x[pl] - m.x[0]
                plane = x[off : off + B * h * w * 4].view(torch.float32).view(B, 1, h, w)
                u, v, _ = QMF(rank=R, num_iters=num_iters, bounds=bounds).decompose(plane)
                fac.append((u.to(torch.int8).cpu().numpy(), v.to(torch.int8).cpu().numpy()))  # (B,1,h,R), (B,1,w,R)
    mj = packing.dict_to_bytes(meta)

    def pack(i):
        parts = []
        for u, v in fac:
            parts += [packing.encode_tensor_nd(u[i]), packing.encode_tensor_nd(v[i])]
        return packing.combine_bytes([mj, packing.combine_bytes(parts)])

    return list(_pool().map(pack, range(B)))


def _qmf_decode_nopatch(metas, bodies, device) -> torch.Tensor:
    """patch=False streams of one shape → uint8 (B,3,H,W) on the device (lrfb_qmf_decode_planes)."""
    meta = metas[0]
    ycbcr = meta["color space"] == "YCbCr"
    n_pl = 3 if ycbcr else 1
    parsed = [[packing.decode_tensor(b) for b in packing.separate_bytes(body, 2 * n_pl)] for body in bodies]
    B = len(parsed)
    us = [np.stack([p[2 * pl] for p in parsed]) for pl in range(n_pl)]
    vs = [np.stack([p[2 * pl + 1] for p in parsed]) for pl in range(n_pl)]
    if ycbcr:
        (H, W), (ch, cw) = meta["original size"][0], meta["original size"][1]
        ranks = list(meta["rank"])
    else:
        H, W, ch, cw = us[0].shape[-2], vs[0].shape[-2], 0, 0
        ranks = [meta["rank"]]
    with torch.cuda.device(device):
        du = [torch.from_numpy(np.ascontiguousarray(a, np.int8)).to(device) for a in us]
        dv = [torch.from_numpy(np.ascontiguousarray(a, np.int8)).to(device) for a in vs]
        out = torch.empty((B, 3, H, W), dtype=torch.uint8, device=device)
        pu = (C.c_void_p * 3)(*[t.data_ptr() for t in du] + [None] * (3 - n_pl))
        pv = (C.c_void_p * 3)(*[t.data_ptr() for t in dv] + [None] * (3 - n_pl))
        rk = (C.c_int32 * 3)(*(ranks + [0] * (3 - n_pl)))
        rc = _cabi.lib().lrfb_qmf_decode_planes(_cabi.LRFB_YCBCR if ycbcr else _cabi.LRFB_RGB, H, W, ch, cw, rk, B,
                                                pu, pv, C.c_void_p(out.data_ptr()), _stream_ptr())
        _cabi.check(rc, "lrfb_qmf_decode_planes")
        return out


def _parse_encoded(encoded: bytes):
    """bytes → (metadata, list of fiber-major int8 arrays [U0, V0, ...])."""
    meta_b, body = packing.separate_bytes(encoded, 2)
    meta = packing.bytes_to_dict(meta_b)
    n = 2 if meta["color space"] == "RGB" else 6
    return meta, [packing.decode_fibers(b) for b in packing.separate_bytes(body, n)]


def _decode_config(meta):
    if not meta["patch"]:
        raise NotImplementedError("lrf_b200: patch=False streams are not on the accelerated path")
    if meta["dtype"] != "uint8":
        raise NotImplementedError("lrf_b200: only uint8 images are decoded on the CUDA path")
    ycbcr = meta["color space"] == "YCbCr"
    (H, W) = meta["original size"][0] if ycbcr else meta["original size"]
    ranks = meta["rank"] if ycbcr else [meta["rank"]]
    if ycbcr:
        ch, cw = meta["original size"][1]
        # any scale with floor(H*s) == ch reproduces the geometry; the decoder only needs sizes
        scale = ((ch + 0.5) / H, (cw + 0.5) / W)
    else:
        scale = (0.5, 0.5)
    cfg = _cabi.make_config(H, W, meta["patch size"], meta["color space"], _cabi.LRFB_U8, scale, ranks,
                            meta["bounds"], 1)
    lay = _cabi.QmfLayout()
    _cabi.check(_cabi.lib().lrfb_qmf_layout_query(C.byref(cfg), C.byref(lay)), "lrfb_qmf_layout_query")
    if ycbcr:
        got = [[lay.orig_h[i], lay.orig_w[i]] for i in range(3)]
        assert got == [list(s) for s in meta["original size"]], "inconsistent plane sizes in metadata"
        assert [[lay.pad_h[i], lay.pad_w[i]] for i in range(3)] == [list(s) for s in meta["padded size"]]
    return cfg, lay


def decode_records(records: torch.Tensor, cfg) -> torch.Tensor:
    """int8 records (B, record_bytes) on the device → uint8 images (B,3,H,W) on the device."""
    B = records.shape[0]
    out = torch.empty((B, 3, cfg.height, cfg.width), dtype=torch.uint8, device=records.device)
    rc = _cabi.lib().lrfb_qmf_decode(C.byref(cfg), B, C.c_void_p(records.data_ptr()),
                                     C.c_void_p(out.data_ptr()), _stream_ptr())
    _cabi.check(rc, "lrfb_qmf_decode")
    return out


def qmf_decode_batch(encoded: list[bytes], device=None) -> torch.Tensor:
    """Batched ``qmf_decode`` of equally shaped streams → uint8 (B,3,H,W) on the device."""
    _require_cuda()
    device = device or torch.device("cuda", torch.cuda.current_device())
    heads = [packing.separate_bytes(e, 2) for e in encoded]
    metas = [packing.bytes_to_dict(h[0]) for h in heads]
    if not metas[0]["patch"]:
        if metas[0]["dtype"] != "uint8":
            raise NotImplementedError("lrf_b200: only uint8 images are decoded on the CUDA path")
        return _qmf_decode_nopatch(metas, [h[1] for h in heads], device)
    cfg, lay = _decode_config(metas[0])
    head = encoded[0][: 4 + len(heads[0][0])]
    if DEVICE_UNPACK and all(e[: len(head)] == head for e in encoded):  # one shape, one header: un-frame + inflate on the device
        with torch.cuda.device(device):
            return decode_records(unpack_records_device(encoded, cfg, lay, device), cfg)
    parsed = list(_pool().map(_parse_encoded, encoded))
    host = np.empty((len(encoded), lay.record_bytes), np.int8)
    for i, (_, fibers) in enumerate(parsed):
        for pl in range(lay.n_planes):
            u, v = fibers[2 * pl], fibers[2 * pl + 1]
            host[i, lay.u_offset[pl] : lay.u_offset[pl] + u.size] = u.reshape(-1)
            host[i, lay.v_offset[pl] : lay.v_offset[pl] + v.size] = v.reshape(-1)
    with torch.cuda.device(device):
        return decode_records(torch.from_numpy(host).to(device), cfg)


def qmf_decode(encoded_image: bytes) -> torch.Tensor:
    """Drop-in for ``lrf.qmf_decode`` (lrf/compression/qmf.py:295-353): returns a CPU uint8 (3,H,W)."""
    return qmf_decode_batch([encoded_image])[0].cpu()


def sse_u8(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Exact per-image sum of squared differences of two uint8 batches on the device."""
    assert a.shape == b.shape and a.dtype == torch.uint8 and b.dtype == torch.uint8 and a.is_cuda
    a, b = a.contiguous(), b.contiguous()
    B = a.shape[0]
    out = torch.zeros(B, dtype=torch.int64, device=a.device)
    rc = _cabi.lib().lrfb_sse_u8(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), a[0].numel(), B,
                                 C.c_void_p(out.data_ptr()), _stream_ptr())
    _cabi.check(rc, "lrfb_sse_u8")
    return out


def psnr_batch(a: torch.Tensor, b: torch.Tensor, max_value: float = 255.0) -> torch.Tensor:
    """lrf/utils/metrics.py:57-71 per image, from the exact integer SSE."""
    mse = sse_u8(a, b).double() / a[0].numel()
    return 20 * torch.log10(max_value / torch.sqrt(mse))


# ----------------------------------------------------------------------------------------------------
# SVD baseline codec (lrf/compression/svd.py), color_space="RGB", patch=True — the only branch of the
# reference that round-trips (its YCbCr branch appends "padded size" twice and fails to decode).
# ----------------------------------------------------------------------------------------------------


def _svd_plan(H, W, rank, quality, patch_size):
    probe = _cabi.make_config(H, W, patch_size, "RGB", _cabi.LRFB_U8, (0.5, 0.5), (1,), (-1, 1), 1)
    lay = _cabi.QmfLayout()
    _cabi.check(_cabi.lib().lrfb_qmf_layout_query(C.byref(probe), C.byref(lay)), "lrfb_qmf_layout_query")
    if rank is None:
        assert quality >= 0 and quality <= 100, "'quality' must be between 0 and 100."
        R = max(round(min(lay.rows[0], lay.cols) * quality / 100), 1)  # compression/svd.py:173-177
    else:
        R = rank
    cfg = _cabi.make_config(H, W, patch_size, "RGB", _cabi.LRFB_U8, (0.5, 0.5), (R,), (-1, 1), 1)
    _cabi.check(_cabi.lib().lrfb_qmf_layout_query(C.byref(cfg), C.byref(lay)), "lrfb_qmf_layout_query")
    return cfg, lay


def svd_encode_batch(images: torch.Tensor, rank=None, quality=None, color_space: str = "RGB",
                     scale_factor=(0.5, 0.5), patch: bool = True, patch_size=(8, 8), dtype=None,
                     sign_flip: Optional[torch.Tensor] = None, return_records: bool = False):
    """Batched ``svd_encode``: uint8 images (B,3,H,W) → list of encoded ``bytes``."""
    assert (rank, quality) != (None, None), "Either 'rank' or 'quality' must be specified."
    if color_space != "RGB" or not patch:
        raise NotImplementedError("lrf_b200: svd codec implements color_space='RGB', patch=True only")
    if images.dtype != torch.uint8 or (dtype is not None and dtype != torch.uint8):
        raise NotImplementedError("lrf_b200: svd codec implements uint8 images / uint8 factors only")
    _require_cuda()
    device = images.device if images.is_cuda else torch.device("cuda", torch.cuda.current_device())
    B, _, H, W = images.shape
    cfg, lay = _svd_plan(H, W, rank, quality, patch_size)
    with torch.cuda.device(device):
        dev_images = images.to(device).contiguous()
        m = _cabi.QmfWorkspaceMap()
        _cabi.check(_cabi.lib().lrfb_qmf_workspace_query(C.byref(cfg), B, C.byref(m)), "lrfb_qmf_workspace_query")
        ws = torch.empty(m.total_bytes, dtype=torch.uint8, device=device)
        codes = torch.empty((B, lay.record_bytes), dtype=torch.uint8, device=device)
        qparams = torch.empty((B, 4), dtype=torch.float32, device=device)
        dbg = _cabi.QmfDebug()
        if sign_flip is not None:
            sf = sign_flip.to(device=device, dtype=torch.int32).contiguous()
            dbg.d_sign_flip[0] = sf.data_ptr()
        rc = _cabi.lib().lrfb_svd_encode(C.byref(cfg), B, C.c_void_p(dev_images.data_ptr()),
                                         C.c_void_p(codes.data_ptr()), C.c_void_p(qparams.data_ptr()),
                                         C.c_void_p(ws.data_ptr()), m.total_bytes, C.byref(dbg), _stream_ptr())
        _cabi.check(rc, "lrfb_svd_encode")
        if return_records:
            return codes, qparams, cfg, lay
        host, qp = codes.cpu().numpy(), qparams.cpu()
    out = []
    for i in range(B):
        meta = {"dtype": "uint8", "color space": "RGB", "patch": True, "patch size": patch_size,
                "original size": [lay.orig_h[0], lay.orig_w[0]], "padded size": [lay.pad_h[0], lay.pad_w[0]],
                "quantization": {"u": [qp[i, 0].item(), qp[i, 1].item()], "v": [qp[i, 2].item(), qp[i, 3].item()]}}
        r, mm, n = lay.rank[0], lay.rows[0], lay.cols
        u = host[i, lay.u_offset[0] : lay.u_offset[0] + r * mm].reshape(r, mm)
        v = host[i, lay.v_offset[0] : lay.v_offset[0] + r * n].reshape(r, n)
        out.append(packing.combine_bytes([packing.dict_to_bytes(meta),
                                          packing.combine_bytes([packing.encode_fibers(u, "uint8"),
                                                                 packing.encode_fibers(v, "uint8")])]))
    return out


def svd_encode(image: torch.Tensor, rank=None, quality=None, color_space: str = "RGB", scale_factor=(0.5, 0.5),
               patch: bool = True, patch_size=(8, 8), dtype=None) -> bytes:
    """Drop-in for ``lrf.svd_encode`` (lrf/compression/svd.py:117-294), RGB + patch branch."""
    return svd_encode_batch(image.unsqueeze(0), rank, quality, color_space, scale_factor, patch, patch_size,
                            dtype)[0]


def svd_decode(encoded_image: bytes) -> torch.Tensor:
    """Drop-in for ``lrf.svd_decode`` (lrf/compression/svd.py:297-361), RGB + patch streams → CPU uint8."""
    _require_cuda()
    meta_b, body = packing.separate_bytes(encoded_image, 2)
    meta = packing.bytes_to_dict(meta_b)
    if meta["color space"] != "RGB" or not meta["patch"] or meta["dtype"] != "uint8" or meta["quantization"]["u"] is None:
        raise NotImplementedError("lrf_b200: svd_decode implements uint8 RGB patch streams only")
    u, v = (packing.decode_fibers(b) for b in packing.separate_bytes(body, 2))
    H, W = meta["original size"]
    cfg = _cabi.make_config(H, W, meta["patch size"], "RGB", _cabi.LRFB_U8, (0.5, 0.5), (u.shape[0],), (-1, 1), 1)
    lay = _cabi.QmfLayout()
    _cabi.check(_cabi.lib().lrfb_qmf_layout_query(C.byref(cfg), C.byref(lay)), "lrfb_qmf_layout_query")
    rec = np.empty(lay.record_bytes, np.uint8)
    rec[lay.u_offset[0] : lay.u_offset[0] + u.size] = u.reshape(-1)
    rec[lay.v_offset[0] : lay.v_offset[0] + v.size] = v.reshape(-1)
    q = meta["quantization"]
    qp = torch.tensor([[q["u"][0], q["u"][1], q["v"][0], q["v"][1], float(u.min()), float(v.min())]],
                      dtype=torch.float32)
    device = torch.device("cuda", torch.cuda.current_device())
    with torch.cuda.device(device):
        d_rec, d_qp = torch.from_numpy(rec).to(device), qp.to(device)
        out = torch.empty((1, 3, H, W), dtype=torch.uint8, device=device)
        rc = _cabi.lib().lrfb_svd_decode(C.byref(cfg), 1, C.c_void_p(d_rec.data_ptr()), C.c_void_p(d_qp.data_ptr()),
                                         C.c_void_p(out.data_ptr()), _stream_ptr())
        _cabi.check(rc, "lrfb_svd_decode")
        return out[0].cpu()
