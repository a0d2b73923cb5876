"""Parity checks shared by the CPU-shim tests and the GPU tests (same assertions, different backend)."""
from __future__ import annotations

import hashlib
import os

import numpy as np
import torch

from backends import config_for, lapack_sign_flips, parse_golden, reference_planes, split_record
from conftest import GOLD, golden_bytes, golden_image, golden_kwargs
from oracle import exact
from oracle import qmf_port as port


def check_frontend(backend, img, patch=(8, 8), color_space="YCbCr"):
    kw = dict(patch_size=patch, color_space=color_space, scale_factor=(0.5, 0.5), bounds=(-16, 15), num_iters=1)
    n_pl = 3 if color_space == "YCbCr" else 1
    cfg = config_for(img, kw, [1] * n_pl)
    _, view, L = backend.encode(img.numpy()[None], cfg, stop_after=1)
    xs = exact.frontend(img.numpy(), patch, color_space == "YCbCr")
    for pl in range(n_pl):
        assert np.array_equal(view("x", pl)[0], xs[pl]), f"plane {pl} differs"


def check_teacher_forced(backend, manifest, name):
    """Reference init injected → the int8 factors must equal the reference's, bit for bit."""
    e = manifest["cases"][name]
    kw = golden_kwargs(e)
    img = golden_image(e["image"])
    meta, ref = parse_golden(golden_bytes(name))
    init = np.load(os.path.join(GOLD, e["init"]))
    cfg = config_for(img, kw, meta["rank"])
    inits = [(init[f"u0_{i}"][None], init[f"v0_{i}"][None]) for i in range(3)]
    fac, _, L = backend.encode(img.numpy()[None], cfg, inits=inits)
    got = split_record(fac[0], L)
    for i, (g, r) in enumerate(zip(got, ref)):
        assert np.array_equal(g, r), f"{name}: factor {i} differs in {int((g != r).sum())} entries"


def check_free_running_sign_aligned(backend, manifest, name, allow_tie_images=0):
    """Own SVD init (FP64 Gram + eigen-solver), column signs aligned to LAPACK's → reference factors."""
    e = manifest["cases"][name]
    kw = golden_kwargs(e)
    img = golden_image(e["image"])
    meta, ref = parse_golden(golden_bytes(name))
    ycbcr = meta["color space"] == "YCbCr"
    ranks = meta["rank"] if ycbcr else [meta["rank"]]
    cfg = config_for(img, kw, ranks)
    if "init" in e:
        init = np.load(os.path.join(GOLD, e["init"]))
        ref_v0 = [init[f"v0_{i}"] for i in range(len(ranks))]
    else:  # LAPACK init recomputed by the oracle port on this machine
        ref_v0 = [port.svd_init(x.unsqueeze(0), ranks[i])[1].squeeze(0).numpy()
                  for i, x in enumerate(reference_planes(img, kw))]
    flips = lapack_sign_flips(backend, img, cfg, ref_v0)
    fac, _, L = backend.encode(img.numpy()[None], cfg, sign_flip=flips)
    got = split_record(fac[0], L)
    diffs = [int((g != r).sum()) for g, r in zip(got, ref)]
    if allow_tie_images == 0:
        assert sum(diffs) == 0, f"{name}: differing entries per factor {diffs}"
    dec = backend.decode(fac, cfg)[0]
    ref_dec = port.qmf_decode(golden_bytes(name)).numpy()
    if sum(diffs) == 0:
        assert hashlib.sha256(dec.tobytes()).hexdigest() == e["decoded_sha256"]
    psnr = port.psnr(img, torch.from_numpy(dec))
    assert abs(psnr - e["psnr"]) <= 0.01, (psnr, e["psnr"])
    return diffs, psnr, ref_dec


def check_decode(backend, manifest, name):
    e = manifest["cases"][name]
    kw = golden_kwargs(e)
    img = golden_image(e["image"])
    meta, ref = parse_golden(golden_bytes(name))
    ycbcr = meta["color space"] == "YCbCr"
    ranks = meta["rank"] if ycbcr else [meta["rank"]]
    cfg = config_for(img, kw, ranks)
    rec = np.concatenate([np.ascontiguousarray(f.T).ravel() for f in ref])[None]
    dec = backend.decode(rec, cfg)
    assert hashlib.sha256(dec[0].tobytes()).hexdigest() == e["decoded_sha256"]
    sse = backend.sse(dec, img.numpy()[None])
    assert int(sse[0]) == exact.sse_u8(dec[0], img.numpy())


def _svd_sign_flips(img, R, v0, patch=(8, 8)):
    x = port.patchify(port.pad_image(img.float(), patch), patch)
    _, S, Vh = torch.linalg.svd(x, full_matrices=False)
    vref = (Vh[:R].T * torch.sqrt(S[:R])).numpy()
    s = np.sign((v0 * vref).sum(0)).astype(np.int32)
    s[s == 0] = 1
    return s[None]
