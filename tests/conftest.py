import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLD, "manifest.json")) as f:
        return json.load(f)


def golden_bytes(name: str) -> bytes:
    with open(os.path.join(GOLD, name + ".bin"), "rb") as f:
        return f.read()


def golden_image(spec):
    """Rebuild the input image of a golden case from its manifest 'image' entry."""
    from oracle import qmf_port as port

    if spec[0] == "png":
        from PIL import Image

        arr = np.array(Image.open(os.path.join(GOLD, spec[1])).convert("RGB"))
        return torch.tensor(arr.transpose(2, 0, 1))
    if spec[0] == "s_nat":
        return port.s_nat(*spec[1:])
    return port.s_iid(*spec[1:])


def golden_kwargs(entry):
    kw = dict(entry["kwargs"])
    if "dtype" in kw:
        kw["dtype"] = getattr(torch, kw["dtype"])
    for k in ("scale_factor", "patch_size", "bounds"):
        if k in kw:
            kw[k] = tuple(kw[k])
    return kw
