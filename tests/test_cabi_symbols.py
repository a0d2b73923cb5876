"""CPU-only: the C-ABI library builds, loads, and exports every symbol include/lrfb.h declares; the
host-only geometry queries agree with the oracle.  No compute calls here (no GPU)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from lrf_b200 import _cabi
from lrf_b200.build import build
from oracle import exact


@pytest.fixture(scope="module")
def lib():
    build()
    return _cabi.lib()


def test_every_declared_symbol_is_exported(lib):
    header = open(os.path.join(ROOT, "include", "lrfb.h")).read()
    declared = set(re.findall(r"\b(lrfb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_cabi.PROTOTYPES), declared ^ set(_cabi.PROTOTYPES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.lrfb_abi_version() == 1


@pytest.mark.parametrize("shape,patch", [((512, 768), (8, 8)), ((662, 992), (8, 8)), ((1365, 2048), (8, 8)),
                                         ((45, 70), (8, 8)), ((256, 384), (4, 4)), ((256, 384), (16, 16))])
def test_layout_matches_oracle_plan(lib, shape, patch):
    cfg = _cabi.make_config(shape[0], shape[1], patch, "YCbCr", _cabi.LRFB_U8, (0.5, 0.5), (4, 2, 2), (-16, 15), 10)
    lay = _cabi.QmfLayout()
    assert lib.lrfb_qmf_layout_query(C.byref(cfg), C.byref(lay)) == 0
    osz, psz, rows = exact.plan(shape[0], shape[1], patch)
    assert [(lay.orig_h[i], lay.orig_w[i]) for i in range(3)] == osz
    assert [(lay.pad_h[i], lay.pad_w[i]) for i in range(3)] == psz
    assert list(lay.rows) == rows and lay.cols == patch[0] * patch[1]
    assert lay.record_bytes == sum((rows[i] + lay.cols) * lay.rank[i] for i in range(3))


def test_bad_arguments_are_rejected(lib):
    cfg = _cabi.make_config(0, 768, (8, 8), "YCbCr", _cabi.LRFB_U8, (0.5, 0.5), (4, 2, 2), (-16, 15), 10)
    lay = _cabi.QmfLayout()
    assert lib.lrfb_qmf_layout_query(C.byref(cfg), C.byref(lay)) == -1
    assert b"non-positive" in lib.lrfb_last_error()
    cfg = _cabi.make_config(64, 64, (8, 8), "YCbCr", _cabi.LRFB_U8, (0.5, 0.5), (4, 0, 2), (-16, 15), 10)
    assert lib.lrfb_qmf_layout_query(C.byref(cfg), C.byref(lay)) == -1
    assert lib.lrfb_qmf_encode(C.byref(cfg), 1, None, None, None, 0, None, None) == -1


def test_product_path_has_no_cpu_fallback():
    import torch

    import lrf_b200

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_cabi.LrfbError):
        lrf_b200.qmf_encode(torch.zeros(3, 16, 16, dtype=torch.uint8), quality=7)
    src = open(os.path.join(ROOT, "lrf_b200", "compression.py")).read() + \
        open(os.path.join(ROOT, "lrf_b200", "_cabi.py")).read()
    assert "oracle" not in src and "cpu_sim" not in src
