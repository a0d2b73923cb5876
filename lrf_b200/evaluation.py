"""Evaluation harness with the reference's output keys (lrf/utils/misc.py:59-121) plus a batched variant
that keeps decoded images and the PSNR reduction on the device.  SSIM needs scikit-image in
the reference and is not part of the metric here: it is reported as NaN."""
from __future__ import annotations

import time
from typing import Callable

import numpy as np
import torch

from . import compression
from .metrics import bits_per_pixel, compression_ratio, psnr


def read_image(path, *args, **kwargs) -> torch.Tensor:
    """``lrf.read_image`` (lrf/utils/misc.py:124-134): image file -> uint8 tensor (C, H, W).  The reference goes through
    skimage.io.imread; here PIL decodes (palette / grey files are expanded to RGB the way the reference's figures are
    used) — host-side IO, nothing on the device."""
    from PIL import Image

    with Image.open(path) as im:
        arr = np.array(im.convert("RGB"))
    return torch.tensor(arr.transpose((2, 0, 1)))


def eval_compression(image, encoder: Callable, decoder: Callable, reconstruct: bool = False, **kwargs) -> dict:
    """Same contract as ``lrf.eval_compression`` (lrf/utils/misc.py:59-121): wall-clock one ``encoder(image, **kwargs)``
    and one ``decoder(encoded)`` (zlib packing included, as in the reference's published timings)."""
    if isinstance(image, str):
        image = read_image(image)
    elif isinstance(image, np.ndarray):
        image = torch.tensor(image.transpose((2, 0, 1)))
    elif not isinstance(image, torch.Tensor):
        raise ValueError("Image must be a file path, numpy array, or torch tensor.")
    t0 = time.perf_counter()
    encoded = encoder(image, **kwargs)
    t1 = time.perf_counter()
    reconstructed = decoder(encoded)
    t2 = time.perf_counter()
    out = {
        "compression ratio": compression_ratio(image, encoded),
        "bit rate (bpp)": bits_per_pixel(image.shape[-2:], encoded),
        "PSNR (dB)": psnr(image, reconstructed).item(),
        "SSIM": float("nan"),
        "encoding time (ms)": 1000 * (t1 - t0),
        "decoding time (ms)": 1000 * (t2 - t1),
    }
    if reconstruct:
        out["reconstructed"] = reconstructed
    return out


def eval_qmf_batch(images: torch.Tensor, **kwargs) -> dict:
    """Batched QMF evaluation: per-image bpp (host zlib) and PSNR (device, exact integer SSE)."""
    encoded = compression.qmf_encode_batch(images, **kwargs)
    decoded = compression.qmf_decode_batch(encoded)
    ref = images.to(decoded.device)
    hw = images.shape[-2] * images.shape[-1]
    return {
        "bit rate (bpp)": torch.tensor([len(e) * 8 / hw for e in encoded], dtype=torch.float64),
        "PSNR (dB)": compression.psnr_batch(decoded, ref).cpu(),
        "encoded": encoded,
    }


def eval_dataset(data_dir: str, qualities=(7,), pattern: str = "*.png", **kwargs) -> list[dict]:
    """Dataset sweep in the shape of experiments/comparison/eval.py:83-96 (``eval_dataset`` / the QMF arm of
    ``eval_image``): every image under ``data_dir`` at every quality, one result row each.  Images of equal shape are
    encoded together through the batch API; PSNR is reduced on the device."""
    import glob
    import os

    params = dict(color_space="YCbCr", scale_factor=(0.5, 0.5), patch=True, patch_size=(8, 8), bounds=(-16, 15),
                  dtype=torch.int8, num_iters=10)
    params.update(kwargs)
    paths = sorted(glob.glob(os.path.join(data_dir, pattern)))
    by_shape: dict[tuple, list] = {}
    for p in paths:
        img = read_image(p)
        by_shape.setdefault(tuple(img.shape), []).append((p, img))
    rows = []
    for shape, items in by_shape.items():
        batch = torch.stack([im for _, im in items])
        for q in qualities:
            t0 = time.perf_counter()
            out = eval_qmf_batch(batch, quality=(q, q / 2, q / 2), **params)
            dt = 1000 * (time.perf_counter() - t0) / len(items)
            for i, (p, _) in enumerate(items):
                rows.append({"data": os.path.splitext(os.path.basename(p))[0], "method": "QMF", "quality": q,
                             "bit rate (bpp)": float(out["bit rate (bpp)"][i]), "PSNR (dB)": float(out["PSNR (dB)"][i]),
                             "time per image (ms)": dt})
    return rows
