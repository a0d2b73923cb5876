"""lrf_b200 — B200 (sm_100a) implementation of the lrf QMF encode/decode hot path.

Flat namespace like the reference's ``lrf`` package for the names on that path.  All arithmetic runs in
hand-written CUDA kernels behind a C ABI (include/lrfb.h); importing works anywhere, computing needs
the built library and a CUDA device (no CPU fallback).
"""
from .compression import (  # noqa: F401
    qmf_encode, qmf_decode, qmf_encode_batch, qmf_decode_batch, qmf_rank, resolve_plan, EncodePlan,
    decode_records, sse_u8, psnr_batch, svd_encode, svd_decode, svd_encode_batch,
)
from .evaluation import eval_compression, eval_qmf_batch, eval_dataset, read_image  # noqa: F401
from .factorization import QMF  # noqa: F401
from .metrics import psnr, mse, bits_per_pixel, compression_ratio, get_memory_usage  # noqa: F401
from .packing import combine_bytes, separate_bytes, dict_to_bytes, bytes_to_dict  # noqa: F401

__all__ = [
    "qmf_encode", "qmf_decode", "svd_encode", "svd_decode", "qmf_encode_batch", "qmf_decode_batch", "qmf_rank", "QMF", "psnr", "mse",
    "bits_per_pixel", "compression_ratio", "combine_bytes", "separate_bytes", "eval_compression", "eval_dataset",
    "read_image",
]
