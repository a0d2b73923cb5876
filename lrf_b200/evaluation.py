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


def eval_compression(image, encoder: Callable, decoder: Callable, reconstruct: bool = False, **kwargs) -> dict:
    """Same contract as ``lrf.eval_compression``: wall-clock one ``encoder(image, **kwargs)`` and one
    ``decoder(encoded)`` (zlib packing included, as in the reference's published timings)."""
    if isinstance(image, np.ndarray):
        image = torch.tensor(image.transpose((2, 0, 1)))
    elif not isinstance(image, torch.Tensor):
        raise ValueError("Image must be a numpy array or a torch tensor.")
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
        "bit rate (bpp)": torch.tensor([len(e) * 8 / hw for e in encoded]),
        "PSNR (dB)": compression.psnr_batch(decoded, ref).cpu(),
        "encoded": encoded,
    }
