"""Host-side mirror of ``lrf.QMF`` (lrf/factorization/qmf.py:167-231) for the codec's use of it:
``factor=(0, 1)`` (w fixed at (0, 1)), SVD initialisation, no regularisation.  ``decompose`` runs the
FP64-Gram SVD init and the block-coordinate-descent sweeps in the sm_100a kernels (lrfb_factorize)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _cabi


class QMF(torch.nn.Module):
    """X ≈ w0 + w1 * (U @ V.T) with U, V integer-valued inside ``bounds``."""

    def __init__(self, rank: Optional[int], num_iters: int = 10, bounds=(None, None),
                 num_levels: Optional[float] = None, verbose: bool = False, **kwargs) -> None:
        super().__init__()
        if num_levels:
            raise NotImplementedError("lrf_b200: num_levels is not implemented on the CUDA path")
        factor = tuple(kwargs.pop("factor", (0, 1)))
        if factor != (0, 1):
            raise NotImplementedError("lrf_b200: only factor=(0, 1) (the codec's setting) is implemented")
        for k, v in kwargs.items():
            if (k in ("l2", "l1_ratio") and v == 0) or (k == "eps" and v == 1e-16):
                continue
            raise NotImplementedError(f"lrf_b200: QMF option {k}={v!r} is not implemented")
        if tuple(bounds) == (None, None):
            raise NotImplementedError("lrf_b200: unbounded QMF is not implemented (int8 bounds required)")
        self.rank, self.num_iters, self.bounds, self.verbose = rank, num_iters, tuple(bounds), verbose

    def decompose(self, x: torch.Tensor, init=None, sign_flip=None):
        """x (..., M, N) → (u (..., M, R), v (..., N, R), w (..., 2, 1)), float32 like the reference."""
        if not torch.cuda.is_available():
            raise _cabi.LrfbError("lrf_b200 needs a CUDA device; there is no CPU fallback")
        lead = x.shape[:-2]
        M, N = x.shape[-2:]
        R = self.rank
        device = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
        xd = x.float().reshape(-1, M, N).to(device).contiguous()
        n = xd.shape[0]
        with torch.cuda.device(device):
            wsb = _cabi.lib().lrfb_factorize_workspace_bytes(n, M, N, R)
            ws = torch.empty(max(int(wsb), 1), dtype=torch.uint8, device=device)
            u = torch.empty((n, M, R), dtype=torch.float32, device=device)
            v = torch.empty((n, N, R), dtype=torch.float32, device=device)
            iu = iv = sf = None
            if init is not None:
                iu = init[0].float().reshape(n, M, R).to(device).contiguous()
                iv = init[1].float().reshape(n, N, R).to(device).contiguous()
            if sign_flip is not None:
                sf = sign_flip.to(device=device, dtype=torch.int32).reshape(n, R).contiguous()
            p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
            rc = _cabi.lib().lrfb_factorize(p(xd), n, M, N, R, float(self.bounds[0]), float(self.bounds[1]),
                                            self.num_iters, p(u), p(v), p(iu), p(iv), p(sf), p(ws), wsb,
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream))
            _cabi.check(rc, "lrfb_factorize")
        u = u.reshape(*lead, M, R).to(x.device)
        v = v.reshape(*lead, N, R).to(x.device)
        w0, w1 = torch.zeros_like(x[..., 0:1, 0:1]).float(), torch.ones_like(x[..., 0:1, 0:1]).float()
        return u, v, torch.cat([w0, w1], dim=-2)

    @staticmethod
    def reconstruct(u: torch.Tensor, v: torch.Tensor, w: Optional[torch.Tensor] = None) -> torch.Tensor:
        out = u @ v.mT
        if w is None:
            return out
        w0, w1 = w.split(split_size=1, dim=-2)
        return w0 + w1 * out

    @staticmethod
    def loss(x, u, v, w=None, eps: float = 1e-16):
        diff = torch.norm(x - QMF.reconstruct(u, v, w), p=2, dim=(-2, -1))
        return diff / (torch.norm(x, p=2, dim=(-2, -1)) + eps)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.reconstruct(*self.decompose(x))
