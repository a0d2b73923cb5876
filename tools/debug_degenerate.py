"""Dev: print what the GPU path does on degenerate images / explicit ranks (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import parity_cases as pc
from backends import GpuBackend, config_for, split_record, reference_planes
from oracle import qmf_port as port
gpu = GpuBackend()
KW = pc.README_KW
for kind, shape in [("flat", (64, 96)), ("flat", (512, 768)), ("half_flat", (512, 768))]:
    img = pc.degenerate_image(kind, *shape)
    blob, ref, meta = port.qmf_encode(img, return_factors=True, **KW)
    cfg = config_for(img, KW, meta["rank"])
    _, view, L = gpu.encode(img.numpy()[None], cfg, stop_after=2)
    for pl, x in enumerate(reference_planes(img, KW)):
        u0, v0 = port.svd_init(x.unsqueeze(0), meta["rank"][pl])
        print(kind, shape, "plane", pl, "M", x.shape[0], "ref v0[:3]", v0[0, :3].numpy().round(4).tolist(), "gpu v0[:3]", view("v", pl)[0][:3].round(4).tolist())
    fac, _, L = gpu.encode(img.numpy()[None], cfg)
    got = split_record(fac[0], L)
    for i, (g, r) in enumerate(zip(got, ref)):
        print("   factor", i, "diffs", int((g != r.numpy()).sum()), "gpu[:2]", g[:2].tolist(), "ref[:2]", r.numpy()[:2].tolist())
