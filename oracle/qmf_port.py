"""TEST INFRASTRUCTURE — CPU oracle for the lrf QMF / SVD codec hot path (torch-CPU port).

This module is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  ``lrf_b200`` never imports anything under ``oracle/``.

It restates, op for op, what the reference (pashtari/lrf, pure Python + torch CPU)
executes on the path ``lrf.qmf_encode`` / ``lrf.qmf_decode`` / ``lrf.svd_encode`` /
``lrf.svd_decode``, using the same third-party arithmetic the reference relies on
(torch 2.11 CPU ops: MKL sgemm/sgemv, LAPACK gesdd, adaptive_avg_pool2d,
reflection_pad2d; zlib 1.3).  None of these are pinned by the reference
(setup.py:28-39 lists unpinned names).  Because the op sequence is the same, it is
also the timing stand-in for the reference on boxes where /root/reference is absent
(``cpu_baseline.kind == "port"``).

Pinning: the reference's own tests hold no golden values for this path
(test/test_compression.py has no assertions), so the port is pinned against outputs
of the reference itself, run in the build container by ``tools/make_golden.py`` and
committed under ``tests/golden/`` (tests/test_oracle_golden.py).

Citations are relative to the reference root.
"""
from __future__ import annotations

import json
import math
import zlib
from collections.abc import Iterable
from functools import reduce

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-16

# --------------------------------------------------------------------------------------
# colour / resampling / padding / patch layout   (lrf/compression/utils.py, qmf.py)
# --------------------------------------------------------------------------------------

_RGB2YCC = [[0.299, 0.587, 0.114], [-0.168736, -0.331264, 0.5], [0.5, -0.418688, -0.081312]]
_YCC2RGB = [[1.0, 0.0, 1.40200], [1.0, -0.344136, -0.714136], [1.0, 1.77200, 0.0]]


def rgb_to_ycbcr(rgb: torch.Tensor) -> torch.Tensor:
    """lrf/compression/utils.py:24-47 — offset + T @ rgb (einsum lowers to bmm, K=3)."""
    t = torch.tensor(_RGB2YCC)
    off = torch.tensor([0, 128, 128]).view(3, 1, 1)
    return off + torch.einsum("ij, j... -> i...", t, rgb.float())


def ycbcr_to_rgb(ycc: torch.Tensor) -> torch.Tensor:
    """lrf/compression/utils.py:50-73."""
    t = torch.tensor(_YCC2RGB)
    off = torch.tensor([0.0, -128.0, -128.0]).view(3, 1, 1)
    return torch.einsum("ij, j... -> i...", t, ycc.float() + off)


def chroma_downsample(ycc: torch.Tensor, scale_factor=(0.5, 0.5)):
    """lrf/compression/utils.py:76-95 — Y untouched, Cb/Cr through F.interpolate(mode='area')."""
    planes = [ycc[0:1]]
    for c in (1, 2):
        planes.append(
            F.interpolate(ycc[None, c : c + 1], scale_factor=scale_factor, mode="area").squeeze(0)
        )
    return planes


def chroma_upsample(planes, size):
    """lrf/compression/utils.py:98-105 with mode='nearest' (compression/qmf.py:346-348)."""
    y, cb, cr = planes
    cb = F.interpolate(cb[None], size=size, mode="nearest").squeeze(0)
    cr = F.interpolate(cr[None], size=size, mode="nearest").squeeze(0)
    return torch.cat((y, cb, cr), dim=0)


def pad_amounts(h: int, w: int, patch_size):
    """lrf/compression/utils.py:123-130 → (top, bottom, left, right)."""
    p, q = patch_size
    ph = (p - h % p) % p
    pw = (q - w % q) % q
    return ph // 2, ph - ph // 2, pw // 2, pw - pw // 2


def pad_image(img: torch.Tensor, patch_size) -> torch.Tensor:
    """lrf/compression/utils.py:108-132 (mode='reflect')."""
    top, bottom, left, right = pad_amounts(img.shape[-2], img.shape[-1], patch_size)
    return F.pad(img, (left, right, top, bottom), mode="reflect")


def unpad_image(img: torch.Tensor, orig_size) -> torch.Tensor:
    """lrf/compression/utils.py:135-153."""
    hp, wp = img.shape[-2:]
    h, w = orig_size
    sh, sw = (hp - h) // 2, (wp - w) // 2
    return img[:, sh : sh + h, sw : sw + w]


def patchify(x: torch.Tensor, patch_size) -> torch.Tensor:
    """lrf/compression/qmf.py:43-56 — 'c (h p) (w q) -> (h w) (c p q)'."""
    p, q = patch_size
    c, hh, ww = x.shape
    h, w = hh // p, ww // q
    return x.reshape(c, h, p, w, q).permute(1, 3, 0, 2, 4).reshape(h * w, c * p * q)


def depatchify(x: torch.Tensor, size, patch_size) -> torch.Tensor:
    """lrf/compression/qmf.py:59-75 — '(h w) (c p q) -> c (h p) (w q)'."""
    p, q = patch_size
    h = size[0] // p
    w = x.shape[0] // h
    c = x.shape[1] // (p * q)
    return x.reshape(h, w, c, p, q).permute(2, 0, 3, 1, 4).reshape(c, h * p, w * q)


def to_dtype(t: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """lrf/compression/utils.py:156-182 — clamp to the dtype range, then truncating cast."""
    info = torch.finfo(dtype) if dtype.is_floating_point else torch.iinfo(dtype)
    return torch.clamp(t, info.min, info.max).to(dtype)


# --------------------------------------------------------------------------------------
# factorisation   (lrf/factorization/qmf.py, utils.py)
# --------------------------------------------------------------------------------------


def _safe_divide(num, den, eps=EPS):
    """lrf/factorization/utils.py:18-33 — a bitwise no-op for w=(0,1) but part of the cost."""
    small = torch.abs(den) < eps
    den = torch.where(small, eps * torch.sign(den), den)
    return num / den


def svd_init(x: torch.Tensor, rank: int):
    """lrf/factorization/qmf.py:42-71 (num_levels=None): thin SVD, top-R, scale by sqrt(s)."""
    r = min(rank, *x.shape[-2:])
    u, s, vh = torch.linalg.svd(x, full_matrices=False)
    u, s, vh = u[..., :, :r], s[..., :r], vh[..., :r, :]
    rs = torch.sqrt(s)
    u = torch.einsum("...ir, ...r -> ...ir", u, rs)
    v = torch.einsum("...rj, ...r -> ...jr", vh, rs)
    if rank > r:
        u = F.pad(u, (0, rank - r))
        v = F.pad(v, (0, rank - r))
    return u, v


def project(t: torch.Tensor, bounds):
    """lrf/factorization/qmf.py:191-195 — round half-to-even, clamp to integer bounds."""
    t = torch.round(t)
    if tuple(bounds) != (None, None):
        t = torch.clamp(t, math.ceil(bounds[0]), math.floor(bounds[1]))
    return t


def half_sweep(x, u, v, bounds, eps=EPS, faithful_cost=True):
    """lrf/factorization/qmf.py:93-126 with w=(0,1), l1=l2=0: update ``u`` given ``x`` and ``v``.

    Gauss–Seidel over the R columns: column r uses already-updated columns j<r and old
    columns j>r.  ``update_v`` (:128-139) is this function on ``x.mT`` with u, v swapped.
    """
    if faithful_cost:
        w0 = torch.zeros_like(x[..., 0:1, 0:1])
        w1 = torch.ones_like(x[..., 0:1, 0:1])
        x = _safe_divide(x - w0, w1, eps)
    r_tot = u.shape[-1]
    a, b = x @ v, v.mT @ v
    if r_tot > 1:
        new = u.clone()
        for r in range(r_tot):
            others = [j for j in range(r_tot) if j != r]
            t2 = new[..., others] @ b[..., others, r : r + 1]
            num = a[..., r : r + 1] - t2
            den = b[..., r : r + 1, r : r + 1]
            new[..., r : r + 1] = project((num + eps) / (den + eps), bounds)
        return new
    return project((a + eps) / (b + eps), bounds)


def qmf_decompose(x, rank, bounds=(-16, 15), num_iters=10, init=None, trace=None, faithful_cost=True):
    """lrf/factorization/qmf.py:197-214 with factor=(0,1): SVD init then num_iters × (u, v).

    ``init`` optionally injects (u0, v0) (teacher forcing); ``trace`` (a list) receives the
    (u, v) pair after every sweep.
    """
    x = x.float()
    u, v = svd_init(x, rank) if init is None else init
    for _ in range(num_iters):
        u = half_sweep(x, u, v, bounds, faithful_cost=faithful_cost)
        v = half_sweep(x.mT, v, u, bounds, faithful_cost=faithful_cost)
        if trace is not None:
            trace.append((u.clone(), v.clone()))
    return u, v


# --------------------------------------------------------------------------------------
# byte framing   (lrf/compression/utils.py:246-490)
# --------------------------------------------------------------------------------------


def _join2(a: bytes, b: bytes) -> bytes:
    return len(a).to_bytes(4, "big") + a + b


def combine_bytes(parts) -> bytes:
    """lrf/compression/utils.py:290-300 — left fold of BE32(len(a)) ‖ a ‖ b."""
    return reduce(_join2, parts)


def separate_bytes(blob: bytes, n: int = 2):
    """lrf/compression/utils.py:303-321."""
    out = []
    head = blob
    for _ in range(n - 1):
        k = int.from_bytes(head[:4], "big")
        head, tail = head[4 : 4 + k], head[4 + k :]
        out.insert(0, tail)
    out.insert(0, head)
    return tuple(out)


def encode_matrix(mat: torch.Tensor) -> bytes:
    """lrf/compression/utils.py:354-390 (mode='col'): one zlib-9 stream per column."""
    cols = [zlib.compress(mat[:, j : j + 1].numpy().tobytes(), level=9) for j in range(mat.shape[1])]
    meta = {"num_fibers": mat.shape[1], "mode": "col", "dtype": str(mat.dtype).split(".")[-1]}
    return combine_bytes([json.dumps(meta).encode("utf-8"), combine_bytes(cols)])


def decode_matrix(blob: bytes) -> torch.Tensor:
    """lrf/compression/utils.py:393-426."""
    meta_b, body = separate_bytes(blob)
    meta = json.loads(meta_b.decode("utf-8"))
    cols = separate_bytes(body, meta["num_fibers"])
    arrs = [np.frombuffer(zlib.decompress(c), dtype=np.dtype(meta["dtype"])) for c in cols]
    return torch.from_numpy(np.stack(arrs, axis=1))


# --------------------------------------------------------------------------------------
# codec pipelines   (lrf/compression/qmf.py:116-353, svd.py:117-361)
# --------------------------------------------------------------------------------------


def encode_tensor(t: torch.Tensor) -> bytes:
    """lrf/compression/utils.py:429-455 — 2-D tensors go column-wise through encode_matrix, anything else is one
    zlib stream of the raw buffer with a {"shape", "dtype"} header (the patch=False branches produce 3-D factors)."""
    if t.ndim == 2:
        return encode_matrix(t)
    meta = {"shape": t.shape, "dtype": str(t.dtype).split(".")[-1]}
    return combine_bytes([json.dumps(meta).encode("utf-8"), zlib.compress(t.numpy().tobytes(), 9)])


def decode_tensor(blob: bytes) -> torch.Tensor:
    """lrf/compression/utils.py:458-490."""
    meta_b, body = separate_bytes(blob, 2)
    meta = json.loads(meta_b.decode("utf-8"))
    if "num_fibers" in meta:
        return decode_matrix(blob)
    arr = np.frombuffer(zlib.decompress(body), dtype=np.dtype(meta["dtype"])).reshape(meta["shape"])
    return torch.from_numpy(arr.copy())


def _triple(v, halve):
    if isinstance(v, Iterable):
        return tuple(v)
    if v is None:
        return (None, None, None)
    return (v, halve(v), halve(v))


def rank_rule(m: int, n: int, quality) -> int:
    """lrf/compression/qmf.py:244-250 — Python banker's round."""
    assert 0 <= quality <= 100, "'quality' must be between 0 and 100."
    return max(round(min(m, n) * quality / 100), 1)


def qmf_planes(image: torch.Tensor, scale_factor=(0.5, 0.5), patch_size=(8, 8)):
    """Front half of qmf_encode (compression/qmf.py:227-242): patch matrices + sizes."""
    ycc = rgb_to_ycbcr(image.float())
    out = []
    for ch in chroma_downsample(ycc, scale_factor):
        xp = pad_image(ch, patch_size)
        out.append((patchify(xp, patch_size), tuple(ch.shape[-2:]), tuple(xp.shape[-2:])))
    return out


def qmf_encode(
    image, rank=None, quality=None, color_space="YCbCr", scale_factor=(0.5, 0.5), patch=True,
    patch_size=(8, 8), bounds=(-16, 15), dtype=torch.int8, num_iters=10, return_factors=False,
    inits=None, faithful_cost=True,
):
    """lrf/compression/qmf.py:116-292: YCbCr / RGB, patch=True (:164-193, :227-262) and patch=False (:195-212,
    :264-286: whole channels as matrices, factors keep their leading batch dimension and go through the N-D branch of
    encode_tensor)."""
    assert (rank, quality) != (None, None), "Either 'rank' or 'quality' must be specified."
    assert color_space in ("RGB", "YCbCr"), "`color_space` must be one of 'RGB' or 'YCbCr'."
    meta = {
        "dtype": str(image.dtype).split(".")[-1], "color space": color_space,
        "patch": patch, "bounds": bounds,
    }
    factors = []
    if not patch:
        if color_space == "RGB":
            x = image.float()
            r = rank_rule(*x.shape[-2:], quality) if rank is None else rank
            meta["rank"] = r
            u, v = qmf_decompose(x.unsqueeze(0), r, bounds, num_iters, faithful_cost=faithful_cost)
            factors = [u.squeeze(0).to(dtype), v.squeeze(0).to(dtype)]
        else:
            ranks = _triple(rank, lambda r: max(r // 2, 1))
            quals = _triple(quality, lambda q: q / 2)
            meta["original size"], meta["rank"] = [], []
            for i, ch in enumerate(chroma_downsample(rgb_to_ycbcr(image.float()), scale_factor)):
                r = rank_rule(*ch.shape[-2:], quals[i]) if ranks[i] is None else ranks[i]
                meta["original size"].append(ch.shape[-2:])
                meta["rank"].append(r)
                u, v = qmf_decompose(ch.unsqueeze(0), r, bounds, num_iters, faithful_cost=faithful_cost)
                factors += [u.squeeze(0).to(dtype), v.squeeze(0).to(dtype)]
        blob = combine_bytes([json.dumps(meta).encode("utf-8"), combine_bytes([encode_tensor(f) for f in factors])])
        return (blob, factors, meta) if return_factors else blob
    if color_space == "RGB":
        xp = pad_image(image.float(), patch_size)
        x = patchify(xp, patch_size)
        r = rank_rule(*x.shape[-2:], quality) if rank is None else rank
        meta.update({"patch size": patch_size, "original size": image.shape[-2:],
                     "padded size": xp.shape[-2:], "rank": r})
        u, v = qmf_decompose(x.unsqueeze(0), r, bounds, num_iters,
                             init=None if inits is None else inits[0], faithful_cost=faithful_cost)
        factors = [u.squeeze(0).to(dtype), v.squeeze(0).to(dtype)]
    else:
        ranks = _triple(rank, lambda r: max(r // 2, 1))
        quals = _triple(quality, lambda q: q / 2)
        meta["patch size"] = patch_size
        meta["original size"], meta["padded size"], meta["rank"] = [], [], []
        for i, (x, osz, psz) in enumerate(qmf_planes(image, scale_factor, patch_size)):
            r = rank_rule(*x.shape[-2:], quals[i]) if ranks[i] is None else ranks[i]
            meta["original size"].append(osz)
            meta["padded size"].append(psz)
            meta["rank"].append(r)
            u, v = qmf_decompose(x.unsqueeze(0), r, bounds, num_iters,
                                 init=None if inits is None else inits[i], faithful_cost=faithful_cost)
            factors += [u.squeeze(0).to(dtype), v.squeeze(0).to(dtype)]
    blob = combine_bytes(
        [json.dumps(meta).encode("utf-8"), combine_bytes([encode_matrix(f) for f in factors])]
    )
    return (blob, factors, meta) if return_factors else blob


def qmf_decode(blob: bytes) -> torch.Tensor:
    """lrf/compression/qmf.py:295-353."""
    meta_b, body = separate_bytes(blob, 2)
    meta = json.loads(meta_b.decode("utf-8"))
    if not meta["patch"]:
        if meta["color space"] == "RGB":
            u, v = (decode_tensor(b).float() for b in separate_bytes(body, 2))
            img = u @ v.mT
        else:
            fs = [decode_tensor(b).float() for b in separate_bytes(body, 6)]
            planes = [fs[2 * i] @ fs[2 * i + 1].mT for i in range(3)]
            img = ycbcr_to_rgb(chroma_upsample(planes, size=meta["original size"][0]))
        return to_dtype(img, getattr(torch, meta["dtype"]))
    if meta["color space"] == "RGB":
        u, v = (decode_matrix(b).float() for b in separate_bytes(body, 2))
        img = unpad_image(depatchify(u @ v.mT, meta["padded size"], meta["patch size"]),
                          meta["original size"])
    else:
        fs = [decode_matrix(b).float() for b in separate_bytes(body, 6)]
        planes = []
        for i in range(3):
            x = fs[2 * i] @ fs[2 * i + 1].mT
            ch = depatchify(x, meta["padded size"][i], meta["patch size"])
            planes.append(unpad_image(ch, meta["original size"][i]))
        img = ycbcr_to_rgb(chroma_upsample(planes, size=meta["original size"][0]))
    return to_dtype(img, getattr(torch, meta["dtype"]))


def quantize(t: torch.Tensor, dtype: torch.dtype):
    """lrf/compression/utils.py:185-220 — min/max affine map, clamp, truncating cast."""
    info = torch.finfo(dtype) if dtype.is_floating_point else torch.iinfo(dtype)
    lo, hi = t.min(), t.max()
    scale = (hi - lo) / (info.max - info.min)
    q = torch.clamp((t - lo) / scale + info.min, info.min, info.max).to(dtype)
    return q, scale.item(), lo.item()


def dequantize(q: torch.Tensor, scale: float, lo: float) -> torch.Tensor:
    """lrf/compression/utils.py:223-243 (subtracts q.min(), not qmin)."""
    q = q.to(torch.float32)
    return (q - q.min()) * scale + lo


def svd_encode(image, rank=None, quality=None, patch_size=(8, 8), dtype=None, return_factors=False):
    """lrf/compression/svd.py:117-294, color_space='RGB', patch=True (the only working baseline,
    SURVEY §3.3)."""
    assert (rank, quality) != (None, None), "Either 'rank' or 'quality' must be specified."
    dtype = image.dtype if dtype is None else dtype
    meta = {"dtype": str(image.dtype).split(".")[-1], "color space": "RGB", "patch": True}
    xp = pad_image(image.float(), patch_size)
    x = patchify(xp, patch_size)
    meta.update({"patch size": patch_size, "original size": image.shape[-2:],
                 "padded size": xp.shape[-2:]})
    r = rank_rule(*x.shape[-2:], quality) if rank is None else rank
    u, s, vh = torch.linalg.svd(x, full_matrices=False)
    u, s, vh = u[..., :, :r], s[..., :r], vh[..., :r, :]
    u = torch.einsum("...ir, ...r -> ...ir", u, torch.sqrt(s))
    v = torch.einsum("...r, ...rj -> ...jr", torch.sqrt(s), vh)
    if not dtype.is_floating_point:
        u, *qu = quantize(u, dtype)
        v, *qv = quantize(v, dtype)
    else:
        qu = qv = None
    meta["quantization"] = {"u": qu, "v": qv}
    blob = combine_bytes(
        [json.dumps(meta).encode("utf-8"), combine_bytes([encode_matrix(u), encode_matrix(v)])]
    )
    return (blob, [u, v], meta) if return_factors else blob


def svd_decode(blob: bytes) -> torch.Tensor:
    """lrf/compression/svd.py:297-361, RGB + patch branch."""
    meta_b, body = separate_bytes(blob, 2)
    meta = json.loads(meta_b.decode("utf-8"))
    u, v = (decode_matrix(b) for b in separate_bytes(body, 2))
    qz = meta["quantization"]
    if qz["u"] is not None:
        u, v = dequantize(u, *qz["u"]), dequantize(v, *qz["v"])
    x = u @ v.mT
    img = unpad_image(depatchify(x, meta["padded size"], meta["patch size"]), meta["original size"])
    return to_dtype(img, getattr(torch, meta["dtype"]))


# --------------------------------------------------------------------------------------
# metrics   (lrf/utils/metrics.py:24-35, :57-71, :149-162)
# --------------------------------------------------------------------------------------


def psnr(a: torch.Tensor, b: torch.Tensor, max_value: float = 255) -> float:
    mse = torch.mean((a.float() - b.float()) ** 2, dim=(-3, -2, -1))
    return float(20 * torch.log10(max_value / torch.sqrt(mse)))


def bits_per_pixel(hw, blob: bytes) -> float:
    return len(blob) * 8 / (hw[0] * hw[1])


# --------------------------------------------------------------------------------------
# synthetic inputs   (SURVEY §8d — host-seeded PCG64, integer-only arithmetic)
# --------------------------------------------------------------------------------------


def s_nat(seed: int, h: int = 512, w: int = 768) -> torch.Tensor:
    """Natural-image-like synthetic RGB u8 (3,h,w); blocks not aligned to the 8×8 grid."""
    rng = np.random.default_rng(seed)
    bh_, bw_, oy, ox = 24, 40, 5, 3
    bh, bw = (h + oy) // bh_ + 2, (w + ox) // bw_ + 2
    blocks = rng.integers(0, 256, size=(3, bh, bw), dtype=np.int64)
    img = np.kron(blocks, np.ones((bh_, bw_), dtype=np.int64))[:, oy : oy + h, ox : ox + w]
    ramp = (np.arange(w, dtype=np.int64)[None, None, :] * 96 // w) + (
        np.arange(h, dtype=np.int64)[None, :, None] * 64 // h
    )
    noise = rng.integers(-12, 13, size=(3, h, w), dtype=np.int64)
    return torch.from_numpy(np.clip(img * 5 // 8 + ramp * 3 // 8 + noise + 16, 0, 255).astype(np.uint8))


def s_iid(seed: int, h: int = 512, w: int = 768) -> torch.Tensor:
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.integers(0, 256, size=(3, h, w), dtype=np.uint8))
