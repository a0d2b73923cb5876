"""Image IO and the dataset sweep driver (lrf/utils/misc.py:124-134, experiments/comparison/eval.py:83-96)."""
import hashlib
import os

import pytest
import torch

from conftest import GOLD


def test_read_image_matches_the_fixture_the_oracle_was_pinned_on(manifest):
    import lrf_b200

    img = lrf_b200.read_image(os.path.join(GOLD, "kodim01.png"))
    assert img.dtype == torch.uint8 and tuple(img.shape) == (3, 662, 992)
    assert hashlib.sha256(img.numpy().tobytes()).hexdigest() == manifest["cases"]["kodim01_q7"]["image_sha256"]


@pytest.mark.gpu
def test_eval_dataset_and_path_argument(manifest):
    import lrf_b200

    e = manifest["cases"]["kodim01_q7"]
    rows = lrf_b200.eval_dataset(GOLD, qualities=(7,))
    assert len(rows) == 1 and rows[0]["data"] == "kodim01" and rows[0]["method"] == "QMF"
    assert abs(rows[0]["PSNR (dB)"] - e["psnr"]) <= 0.01 and abs(rows[0]["bit rate (bpp)"] - e["bpp"]) <= 1e-7
    out = lrf_b200.eval_compression(os.path.join(GOLD, "kodim01.png"), lrf_b200.qmf_encode, lrf_b200.qmf_decode,
                                    quality=7, num_iters=10)
    assert abs(out["bit rate (bpp)"] - e["bpp"]) <= 1e-7
