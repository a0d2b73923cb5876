"""Measurement helpers with the reference's definitions (lrf/utils/metrics.py:24-35, :57-71, :120-162)."""
from __future__ import annotations

import torch


def mse(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return torch.mean((a - b) ** 2, dim=(-3, -2, -1))


def psnr(img1: torch.Tensor, img2: torch.Tensor, max_value: int = 255) -> torch.Tensor:
    return 20 * torch.log10(max_value / torch.sqrt(mse(img1.float(), img2.float())))


def get_memory_usage(obj) -> int:
    if isinstance(obj, (bytes, bytearray)):
        return len(obj)
    if isinstance(obj, torch.Tensor):
        return obj.numel() * obj.element_size()
    if isinstance(obj, dict):
        return sum(get_memory_usage(v) for v in obj.values())
    if isinstance(obj, (list, tuple)):
        return sum(get_memory_usage(v) for v in obj)
    raise TypeError(f"unsupported type {type(obj)}")


def bits_per_pixel(size, compressed) -> float:
    n = 1
    for s in size:
        n *= s
    return get_memory_usage(compressed) * 8 / n


def compression_ratio(inp, compressed) -> float:
    return get_memory_usage(inp) / get_memory_usage(compressed)
