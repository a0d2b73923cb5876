"""The device un-framing + inflate (lrf_b200/csrc/inflate9.cuh) against the records the streams were made from (zlib on the
host wrote them, or the device deflate): exact round trips, malformed input is refused."""
import ctypes as C
import zlib

import numpy as np
import pytest
import torch

from lrf_b200 import _cabi, compression, packing
from test_deflate9 import _records


def _case(H, W, quality, count, space="YCbCr", rank=None):
    cfg, lay = compression.resolve_plan(H, W, rank, quality, space, (0.5, 0.5), (8, 8), (-16, 15), 10)
    meta = compression._metadata(torch.uint8, space, True, (-16, 15), (8, 8), lay)
    recs = _records(np.random.default_rng(21), lay, count)
    if count > 4:
        recs[4] = np.random.default_rng(22).integers(-128, 128, lay.record_bytes).astype(np.int8)  # stored blocks
    return cfg, lay, meta, recs


@pytest.mark.parametrize("shape,quality,count", [((48, 64), 7, 5), ((64, 96), 25, 3)])
def test_unpack_on_shim_round_trips(shape, quality, count):
    from cpu_sim import simlib

    cfg, lay, meta, recs = _case(*shape, quality, count)
    blobs = [packing.pack_qmf_record(recs[i], lay, meta) for i in range(count)]
    assert np.array_equal(simlib.unpack_device(blobs, cfg), recs)
    bad = bytearray(blobs[1])
    bad[len(bad) // 2] ^= 0x40  # a flipped bit inside a deflate stream: caught by the decoder or by the adler32
    with pytest.raises(_cabi.LrfbError, match="image 1"):
        simlib.unpack_device([blobs[0], bytes(bad)], cfg)
    with pytest.raises(_cabi.LrfbError):
        simlib.unpack_device([blobs[0][:-3]], cfg)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,quality,count,space", [((512, 768), 7, 40, "YCbCr"), ((96, 160), 25, 33, "YCbCr"), ((8, 8), 50, 3, "YCbCr"),
                                                       ((128, 192), 30, 6, "RGB"), ((1365, 2048), 7, 5, "YCbCr")])
def test_device_unpack_round_trips(shape, quality, count, space):
    """Streams written by zlib on the host (fixed, dynamic and stored blocks; several blocks per column at CLIC size) and by
    the device deflate both inflate back to the records, bit for bit."""
    cfg, lay, meta, recs = _case(*shape, quality, count, space)
    host_blobs = compression.pack_records(recs, cfg, lay, meta)
    dev = torch.device("cuda", 0)
    assert np.array_equal(compression.unpack_records_device(host_blobs, cfg, lay, dev).cpu().numpy(), recs)
    dev_blobs = compression.pack_records_device(torch.from_numpy(recs).cuda(), cfg, lay, meta)
    assert np.array_equal(compression.unpack_records_device(dev_blobs, cfg, lay, dev).cpu().numpy(), recs)


@pytest.mark.gpu
def test_decode_batch_takes_the_device_route_and_matches_the_oracle():
    from oracle import qmf_port as port

    imgs = torch.stack([port.s_nat(1000 + i, 256, 384) for i in range(5)])
    blobs = compression.qmf_encode_batch(imgs, quality=7)
    keep = compression.DEVICE_UNPACK
    compression.DEVICE_UNPACK = True
    try:
        dec = compression.qmf_decode_batch(blobs).cpu()
        for i in range(5):
            assert torch.equal(dec[i], port.qmf_decode(blobs[i]))
        bad = bytearray(blobs[2])
        bad[-9] ^= 1
        with pytest.raises(_cabi.LrfbError, match="image 2"):
            compression.qmf_decode_batch(blobs[:2] + [bytes(bad)] + blobs[3:])
    finally:
        compression.DEVICE_UNPACK = keep


@pytest.mark.gpu
def test_decode_bytes_host_matches_decode_batch():
    """lrfb_qmf_decode_bytes_host (encoded images in host memory -> pinned uint8 images, chunked copy-back) equals the
    device route of qmf_decode_batch; 37 images over 5 chunks with a ragged tail."""
    from oracle import qmf_port as port

    H, W, B = 96, 160, 37
    pool = torch.stack([port.s_nat(2000 + i, H, W) for i in range(5)])
    imgs = pool[torch.arange(B) % 5].contiguous()
    blobs = compression.qmf_encode_batch(imgs, quality=12)
    want = compression.qmf_decode_batch(blobs).cpu()
    cfg, lay = compression._decode_config(packing.bytes_to_dict(packing.separate_bytes(blobs[0], 2)[0]))
    offs = np.zeros(B + 1, np.int64)
    np.cumsum([len(b) for b in blobs], out=offs[1:])
    blob = np.frombuffer(b"".join(blobs), np.uint8).copy()
    out = torch.empty((B, 3, H, W), dtype=torch.uint8, pin_memory=True)
    lib = _cabi.lib()
    ctx = C.c_void_p()
    _cabi.check(lib.lrfb_ctx_create(0, C.byref(ctx)), "ctx")
    try:
        _cabi.check(lib.lrfb_ctx_set_chunk_bytes(ctx, 8 * 3 * H * W), "chunk")
        for _ in range(2):
            out.zero_()
            rc = lib.lrfb_qmf_decode_bytes_host(ctx, C.byref(cfg), B, blob.ctypes.data_as(C.c_void_p),
                                                offs.ctypes.data_as(C.c_void_p), C.c_void_p(out.data_ptr()))
            _cabi.check(rc, "lrfb_qmf_decode_bytes_host")
            assert torch.equal(out, want)
        blob[int(offs[4]) - 9] ^= 0x10  # inside the last column stream of image 3
        rc = lib.lrfb_qmf_decode_bytes_host(ctx, C.byref(cfg), B, blob.ctypes.data_as(C.c_void_p),
                                            offs.ctypes.data_as(C.c_void_p), C.c_void_p(out.data_ptr()))
        assert rc != 0 and b"image 3" in lib.lrfb_last_error()
        blob[int(offs[4]) - 9] ^= 0x10
        blob[int(offs[21]) - 9] ^= 0x10  # image 20 sits in the third chunk: the index reported is the batch's, not the chunk's
        rc = lib.lrfb_qmf_decode_bytes_host(ctx, C.byref(cfg), B, blob.ctypes.data_as(C.c_void_p),
                                            offs.ctypes.data_as(C.c_void_p), C.c_void_p(out.data_ptr()))
        assert rc != 0 and b"image 20" in lib.lrfb_last_error()
        blob[int(offs[21]) - 9] ^= 0x10
        out.zero_()
        _cabi.check(lib.lrfb_qmf_decode_bytes_host(ctx, C.byref(cfg), B, blob.ctypes.data_as(C.c_void_p),
                                                   offs.ctypes.data_as(C.c_void_p), C.c_void_p(out.data_ptr())), "after errors")
        assert torch.equal(out, want)
    finally:
        lib.lrfb_ctx_destroy(ctx)
