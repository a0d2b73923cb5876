"""Generate tests/golden/ fixtures by running the UNMODIFIED reference in this container.

    python tools/make_golden.py

Needs /root/reference (build container only).  Writes, per case, the reference's encoded
bytes (they losslessly contain every int8 factor), a manifest with sha256 / bpp / PSNR /
torch + thread info, and for a few cases the LAPACK SVD initialisation (u0, v0) so GPU
tests can teacher-force the BCD sweeps without depending on the GPU box's MKL build.
It also asserts that oracle/qmf_port.py reproduces the reference byte for byte on every
case — that is what pins the oracle.
"""
import hashlib
import json
import os
import shutil
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, ROOT)
from ref_shim import import_reference  # noqa: E402

lrf = import_reference()
from oracle import qmf_port as port  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
README_KW = dict(color_space="YCbCr", scale_factor=(0.5, 0.5), quality=7, patch=True,
                 patch_size=(8, 8), bounds=(-16, 15), dtype=torch.int8, num_iters=10)


def load_png(path):
    from PIL import Image

    return torch.tensor(np.array(Image.open(path).convert("RGB")).transpose(2, 0, 1))


def jsonable(kw):
    out = {}
    for k, v in kw.items():
        out[k] = str(v).split(".")[-1] if isinstance(v, torch.dtype) else v
    return out


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    shutil.copyfile("/root/reference/figures/kodim01.png", os.path.join(GOLD, "kodim01.png"))
    kodim = load_png(os.path.join(GOLD, "kodim01.png"))

    cases = []  # (name, image-spec, codec, kwargs, save_init)

    def add(name, spec, codec="qmf", kw=None, save_init=False):
        cases.append((name, spec, codec, dict(README_KW) if kw is None else kw, save_init))

    add("kodim01_q7", ("png", "kodim01.png"))
    for s in (1000, 1001, 1002, 1003):
        add(f"snat{s}_512x768_q7", ("s_nat", s, 512, 768), save_init=(s == 1000))
    add("snat1000_1365x2048_q7", ("s_nat", 1000, 1365, 2048))
    add("siid2000_512x768_q7", ("s_iid", 2000, 512, 768))
    add("snat7_45x70_q7", ("s_nat", 7, 45, 70), save_init=True)        # M<64, chroma R=1
    add("snat8_101x131_q7", ("s_nat", 8, 101, 131), save_init=True)    # odd H and W, both padded
    add("snat9_128x192_q7", ("s_nat", 9, 128, 192), save_init=True)
    base = ("s_nat", 1000, 256, 384)
    for lo, hi in ((-8, 7), (-128, 127)):
        add(f"snat1000_256x384_b{-lo}", base, kw={**README_KW, "bounds": (lo, hi)})
    for p in (4, 16):
        add(f"snat1000_256x384_p{p}", base, kw={**README_KW, "patch_size": (p, p)})
    for it in (1, 2, 5, 20):
        add(f"snat1000_256x384_it{it}", base, kw={**README_KW, "num_iters": it})
    add("snat1000_256x384_rank", base, kw={**{k: v for k, v in README_KW.items() if k != "quality"},
                                          "rank": 6})
    add("snat1000_128x192_rgb", ("s_nat", 1000, 128, 192), kw={**README_KW, "color_space": "RGB", "quality": 3})
    # patch=False branches (lrf/compression/qmf.py:195-212, :264-286): whole channels as matrices
    add("snat1000_128x192_nopatch", ("s_nat", 1000, 128, 192), kw={**README_KW, "patch": False})
    add("snat1000_96x160_nopatch_rgb", ("s_nat", 1000, 96, 160),
        kw={**README_KW, "patch": False, "color_space": "RGB", "quality": 5})
    add("kodim01_svd_q1", ("png", "kodim01.png"), codec="svd", kw=dict(quality=1.0))
    add("kodim01_svd_q7", ("png", "kodim01.png"), codec="svd", kw=dict(quality=7))
    add("snat1000_512x768_svd_q1", ("s_nat", 1000, 512, 768), codec="svd", kw=dict(quality=1.0))

    manifest = {
        "torch": torch.__version__, "numpy": np.__version__, "threads": torch.get_num_threads(),
        "generator": "tools/make_golden.py", "cases": {},
    }
    for name, spec, codec, kw, save_init in cases:
        if spec[0] == "png":
            img = kodim
        elif spec[0] == "s_nat":
            img = port.s_nat(*spec[1:])
        else:
            img = port.s_iid(*spec[1:])
        if codec == "qmf":
            enc = lrf.qmf_encode(img, **kw)
            dec = lrf.qmf_decode(enc)
            penc = port.qmf_encode(img, **kw)
            pdec = port.qmf_decode(enc)
        else:
            enc = lrf.svd_encode(img, **kw)
            dec = lrf.svd_decode(enc)
            penc = port.svd_encode(img, **kw)
            pdec = port.svd_decode(enc)
        assert penc == enc, f"{name}: oracle port bytes differ from the reference"
        assert torch.equal(pdec, dec), f"{name}: oracle port decode differs from the reference"
        with open(os.path.join(GOLD, name + ".bin"), "wb") as f:
            f.write(enc)
        entry = {
            "image": list(spec), "codec": codec, "kwargs": jsonable(kw), "bytes": len(enc),
            "sha256": hashlib.sha256(enc).hexdigest(),
            "image_sha256": hashlib.sha256(img.numpy().tobytes()).hexdigest(),
            "decoded_sha256": hashlib.sha256(dec.numpy().tobytes()).hexdigest(),
            "psnr": float(lrf.psnr(img, dec)),
            "bpp": float(lrf.bits_per_pixel(img.shape[-2:], enc)),
        }
        if save_init and codec == "qmf":
            arrs = {}
            for i, (x, _, _) in enumerate(port.qmf_planes(img, kw["scale_factor"], kw["patch_size"])):
                meta = json.loads(port.separate_bytes(enc, 2)[0].decode())
                u0, v0 = lrf.SVDInit(rank=meta["rank"][i])(x.unsqueeze(0))[:2]
                arrs[f"u0_{i}"] = u0.squeeze(0).numpy()
                arrs[f"v0_{i}"] = v0.squeeze(0).numpy()
            np.savez_compressed(os.path.join(GOLD, name + "_init.npz"), **arrs)
            entry["init"] = name + "_init.npz"
        manifest["cases"][name] = entry
        print(f"{name:32s} {len(enc):7d} B  psnr {entry['psnr']:.5f}  bpp {entry['bpp']:.6f}")
    with open(os.path.join(GOLD, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
