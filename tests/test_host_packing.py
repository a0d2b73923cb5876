"""CPU-only: the host byte packing reproduces the reference's encoded-object layout byte for byte."""
import ctypes as C
import json

import numpy as np
import pytest
import torch

from conftest import golden_bytes, golden_kwargs
from lrf_b200 import _cabi, compression, packing
from lrf_b200.build import build
from oracle import qmf_port as port

CASES = ["kodim01_q7", "snat1000_512x768_q7", "snat7_45x70_q7", "snat8_101x131_q7", "snat1000_256x384_p4",
         "snat1000_256x384_p16", "snat1000_256x384_rank", "snat1000_128x192_rgb", "snat1000_256x384_b128"]


@pytest.mark.parametrize("name", CASES)
def test_repack_golden_factors_is_byte_identical(manifest, name):
    build()
    e = manifest["cases"][name]
    kw = golden_kwargs(e)
    blob = golden_bytes(name)
    meta, fibers = compression._parse_encoded(blob)
    ycbcr = meta["color space"] == "YCbCr"
    H, W = meta["original size"][0] if ycbcr else meta["original size"]
    cfg, lay = compression.resolve_plan(H, W, kw.get("rank"), kw.get("quality"), kw["color_space"],
                                        kw["scale_factor"], kw["patch_size"], kw["bounds"], kw["num_iters"])
    assert [lay.rank[i] for i in range(lay.n_planes)] == (meta["rank"] if ycbcr else [meta["rank"]])
    rec = np.zeros(lay.record_bytes, np.int8)
    for pl in range(lay.n_planes):
        u, v = fibers[2 * pl], fibers[2 * pl + 1]
        rec[lay.u_offset[pl] : lay.u_offset[pl] + u.size] = u.reshape(-1)
        rec[lay.v_offset[pl] : lay.v_offset[pl] + v.size] = v.reshape(-1)
    new_meta = compression._metadata(torch.uint8, kw["color_space"], True, kw["bounds"], kw["patch_size"], lay)
    assert json.dumps(new_meta) == json.dumps(meta)
    assert packing.pack_qmf_record(rec, lay, new_meta) == blob
    # the native thread-pool packer of liblrfb.so (lrfb_qmf_pack_host) writes the same bytes
    assert compression.pack_records(rec[None], cfg, lay, new_meta) == [blob]


def test_native_packer_matches_python_packer_on_random_records():
    """lrfb_qmf_pack_host (C++ zlib pool) == packing.pack_qmf_record for a ragged batch of random records, any thread
    count; pure host code, so it runs without a GPU."""
    build()
    cfg, lay = compression.resolve_plan(96, 160, None, 7, "YCbCr", (0.5, 0.5), (8, 8), (-16, 15), 10)
    meta = compression._metadata(torch.uint8, "YCbCr", True, (-16, 15), (8, 8), lay)
    rng = np.random.default_rng(3)
    recs = rng.integers(-16, 16, size=(37, lay.record_bytes)).astype(np.int8)
    recs[5] = 0          # highly compressible
    recs[6, ::2] = 15
    want = [packing.pack_qmf_record(recs[i], lay, meta) for i in range(len(recs))]
    for threads in (0, 1, 3):
        assert compression.pack_records(recs, cfg, lay, meta, threads) == want
    assert port.qmf_decode(want[0]).shape == (3, 96, 160)


def test_framing_errors_match_reference_behaviour():
    with pytest.raises(TypeError):
        packing.combine_bytes([b"a", "b"])
    with pytest.raises(ValueError):
        packing.separate_bytes(b"\x00\x01", 2)
    parts = [b"x" * 5, b"", b"yz"]
    assert list(packing.separate_bytes(packing.combine_bytes(parts), 3)) == parts
    assert packing.combine_bytes(parts) == port.combine_bytes(parts)


def test_argument_errors_match_reference():
    img = torch.zeros(3, 16, 16, dtype=torch.uint8)
    with pytest.raises(AssertionError):
        compression.qmf_encode(img)
    with pytest.raises(AssertionError):
        compression.qmf_encode(img, quality=7, color_space="HSV")
    with pytest.raises(NotImplementedError):
        compression.qmf_encode(img, quality=7, dtype=torch.int16)


def test_nd_tensor_framing_matches_reference_layout(manifest):
    """patch=False streams: the N-D branch of encode_tensor / decode_tensor (lrf/compression/utils.py:429-490)."""
    blob = golden_bytes("snat1000_128x192_nopatch")
    meta_b, body = packing.separate_bytes(blob, 2)
    meta = packing.bytes_to_dict(meta_b)
    assert meta["patch"] is False and meta["rank"] == [9, 2, 2]
    parts = packing.separate_bytes(body, 6)
    facs = [packing.decode_tensor(b) for b in parts]
    assert facs[0].shape == (1, 128, 9) and facs[1].shape == (1, 192, 9) and facs[2].shape == (1, 64, 2)
    assert [packing.encode_tensor_nd(f) for f in facs] == list(parts)
    ref = [f.numpy() for f in (port.decode_tensor(b) for b in parts)]
    assert all(np.array_equal(a, b) for a, b in zip(facs, ref))
