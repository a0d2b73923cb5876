"""The device lossless stage (lrf_b200/csrc/deflate9.cuh) against zlib itself, byte for byte.

CPU part: the serial restatement (shim build only) on thousands of streams — it shares the Huffman-tree, block-choice
and header code with the kernel — and the kernel proper (warp-parallel chain walk, sort, parallel symbol coding,
framing kernels) on the SIMT shim for small records.  GPU part: lrfb_qmf_pack_device against packing.pack_qmf_record
(Python zlib) on real factor records and on adversarial ones."""
import ctypes as C
import zlib

import numpy as np
import pytest
import torch

from lrf_b200 import _cabi, compression, packing


def _streams(rng, count, max_len):
    out = []
    for it in range(count):
        n = int(rng.integers(0, 70)) if it % 5 == 0 else 64 if it % 5 == 1 else int(rng.integers(0, max_len))
        mode = it % 7
        if mode == 0:
            a = rng.integers(0, 256, n)
        elif mode == 1:
            a = rng.integers(-16, 16, n)
        elif mode == 2:
            a = np.clip(np.cumsum(rng.integers(-1, 2, n)), -16, 15)
        elif mode == 3:
            a = np.where((np.arange(n) // 50) % 3 > 0, 0, rng.integers(0, 4, n))
        elif mode == 4:
            a = np.rint(8 * np.sin(np.arange(n) * 0.01 * (1 + it % 5))) + (rng.integers(0, 3, n) == 0)
        elif mode == 5:
            a = np.zeros(n)
        else:
            base = rng.integers(-4, 5, 100)
            a = base[np.arange(n) % 100] + (rng.integers(0, 20, n) == 0)
        out.append(a.astype(np.int8).tobytes())
    return out


def test_serial_restatement_matches_zlib():
    from cpu_sim import simlib

    rng = np.random.default_rng(11)
    for data in _streams(rng, 1500, 9000) + [bytes(16382), rng.integers(-16, 16, 16382).astype(np.int8).tobytes()]:
        assert simlib.deflate9_serial(data) == zlib.compress(data, 9), len(data)
    # several deflate blocks (more than 16 383 symbols), matches limited to MAX_DIST, up to the longest column taken
    for n in (16383, 16384, 20000, 43776, 65024):
        t = np.arange(n)
        for a in (rng.integers(-16, 16, n), rng.integers(0, 256, n), np.zeros(n), np.rint(9 * np.sin(t * 0.003) + (rng.integers(0, 4, n) == 0)),
                  np.where(t < 40000, (t * 7 + t // 9) % 11, 0) + np.where(t >= 40000, ((t - 40000) * 7 + (t - 40000) // 9) % 11, 0)):
            data = a.astype(np.int8).tobytes()
            assert simlib.deflate9_serial(data) == zlib.compress(data, 9), n


def _records(rng, lay, count):
    recs = rng.integers(-16, 16, size=(count, lay.record_bytes)).astype(np.int8)
    recs[0] = 0
    if count > 1:
        recs[1, ::2] = 15
    if count > 2:  # smooth columns: long matches, long hash chains
        t = np.arange(lay.record_bytes)
        recs[2] = np.rint(10 * np.sin(t * 0.02) + 3 * np.sin(t * 0.11)).astype(np.int8)
    if count > 3:
        recs[3] = np.clip(np.cumsum(rng.integers(-1, 2, lay.record_bytes)), -16, 15).astype(np.int8)
    return recs


@pytest.mark.parametrize("shape,quality,count", [((48, 64), 7, 4), ((64, 96), 25, 3)])
def test_kernel_on_shim_matches_python_packer(shape, quality, count):
    from cpu_sim import simlib

    H, W = shape
    cfg, lay = compression.resolve_plan(H, W, None, quality, "YCbCr", (0.5, 0.5), (8, 8), (-16, 15), 10)
    meta = compression._metadata(torch.uint8, "YCbCr", True, (-16, 15), (8, 8), lay)
    recs = _records(np.random.default_rng(5), lay, count)
    want = [packing.pack_qmf_record(recs[i], lay, meta) for i in range(len(recs))]
    got = simlib.pack_device(recs, cfg, packing.dict_to_bytes(meta))
    assert got == want


@pytest.mark.gpu
@pytest.mark.parametrize("shape,quality,batch,space", [((512, 768), 7, 24, "YCbCr"), ((96, 160), 25, 33, "YCbCr"),
                                                       ((200, 328), 3, 7, "YCbCr"), ((8, 8), 50, 3, "YCbCr"),
                                                       ((16, 24), 100, 5, "YCbCr"), ((128, 192), 30, 6, "RGB")])
def test_device_packer_matches_python_packer(shape, quality, batch, space):
    """Ragged shapes down to one-row factor columns (1- and 2-byte streams), full-rank factors (64 columns), RGB planes."""
    H, W = shape
    cfg, lay = compression.resolve_plan(H, W, None, quality, space, (0.5, 0.5), (8, 8), (-16, 15), 10)
    meta = compression._metadata(torch.uint8, space, True, (-16, 15), (8, 8), lay)
    recs = _records(np.random.default_rng(7), lay, batch)
    want = [packing.pack_qmf_record(recs[i], lay, meta) for i in range(batch)]
    got = compression.pack_records_device(torch.from_numpy(recs).cuda(), cfg, lay, meta)
    assert got == want


@pytest.mark.gpu
def test_device_packer_on_real_factors_and_public_api():
    """qmf_encode_batch goes through the device packer; its bytes equal the host packers' on real factors."""
    from oracle import qmf_port as port

    imgs = torch.stack([port.s_nat(1000 + i, 256, 384) for i in range(6)])
    records, lay, meta = compression.qmf_encode_batch(imgs, quality=7, return_records=True)
    cfg, _ = compression.resolve_plan(256, 384, None, 7, "YCbCr", (0.5, 0.5), (8, 8), (-16, 15), 10)
    host = records.cpu().numpy()
    want = [packing.pack_qmf_record(host[i], lay, meta) for i in range(len(host))]
    assert compression.pack_records_device(records, cfg, lay, meta) == want
    assert compression.pack_records(host, cfg, lay, meta) == want
    assert compression.qmf_encode_batch(imgs, quality=7) == want


def test_kernel_on_shim_two_blocks():
    """A 16 512-byte luma column of noise: more than 16 383 symbols, so the kernel flushes a block in mid-parse."""
    from cpu_sim import simlib

    cfg, lay = compression.resolve_plan(1024, 1032, (1, 1, 1), None, "YCbCr", (0.5, 0.5), (8, 8), (-16, 15), 10)
    assert lay.rows[0] == 16512
    meta = compression._metadata(torch.uint8, "YCbCr", True, (-16, 15), (8, 8), lay)
    recs = np.random.default_rng(9).integers(-16, 16, size=(1, lay.record_bytes)).astype(np.int8)
    assert simlib.pack_device(recs, cfg, packing.dict_to_bytes(meta)) == [packing.pack_qmf_record(recs[0], lay, meta)]


@pytest.mark.gpu
def test_device_packer_long_columns_several_blocks():
    """CLIC-sized planes (43 776-byte luma columns, 11 008-byte chroma columns): several deflate blocks per column,
    matches limited to MAX_DIST; noise, smooth and real-factor records against the native host packer (zlib)."""
    from oracle import qmf_port as port

    H, W = 1365, 2048
    cfg, lay = compression.resolve_plan(H, W, None, 7, "YCbCr", (0.5, 0.5), (8, 8), (-16, 15), 10)
    assert lay.rows[0] == 43776
    meta = compression._metadata(torch.uint8, "YCbCr", True, (-16, 15), (8, 8), lay)
    recs = _records(np.random.default_rng(13), lay, 5)
    img = port.s_nat(1000, H, W).unsqueeze(0)
    real, _, _ = compression.qmf_encode_batch(img, quality=7, return_records=True)
    recs[4] = real[0].cpu().numpy()
    want = compression.pack_records(recs, cfg, lay, meta)
    assert compression.pack_records_device(torch.from_numpy(recs).cuda(), cfg, lay, meta) == want
    assert compression.qmf_encode_batch(img, quality=7) == [want[4]]


@pytest.mark.gpu
def test_device_packer_refuses_columns_beyond_the_window():
    cfg, lay = compression.resolve_plan(2048, 2048, None, 7, "YCbCr", (0.5, 0.5), (8, 8), (-16, 15), 10)  # 65 536 rows
    assert _cabi.lib().lrfb_qmf_pack_device_workspace(C.byref(cfg), 2) == -1


@pytest.mark.gpu
@pytest.mark.parametrize("B,chunk_images", [(37, 2), (150, 20), (5, 25)])
def test_encode_bytes_host_pipeline_ragged_chunks(B, chunk_images):
    """lrfb_qmf_encode_bytes_host (pinned host images -> finished streams) over uneven chunks equals encode + Python
    packer per image; the public batch API takes the same route for pinned inputs.  The pipeline chunk is 4 x
    chunk_images: 37 images = 4 x 8 + 5, 150 = 80 + 70, 5 = one chunk."""
    from oracle import qmf_port as port

    H, W = 96, 160
    pool = torch.stack([port.s_nat(2000 + i, H, W) for i in range(5)])
    imgs = pool[torch.arange(B) % 5].contiguous().pin_memory()
    records, lay, meta = compression.qmf_encode_batch(imgs.cuda(), quality=12, return_records=True)
    host = records.cpu().numpy()
    want = [packing.pack_qmf_record(host[i], lay, meta) for i in range(B)]
    cfg, _ = compression.resolve_plan(H, W, None, 12, "YCbCr", (0.5, 0.5), (8, 8), (-16, 15), 10)
    lib = _cabi.lib()
    ctx = C.c_void_p()
    _cabi.check(lib.lrfb_ctx_create(0, C.byref(ctx)), "ctx")
    try:
        _cabi.check(lib.lrfb_ctx_set_chunk_bytes(ctx, chunk_images * 3 * H * W), "chunk")
        mj = packing.dict_to_bytes(meta)
        cap = B * int(lib.lrfb_qmf_pack_bound(C.byref(cfg), len(mj)))
        blob = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
        offs = torch.zeros(B + 1, dtype=torch.int64, pin_memory=True)
        for _ in range(2):  # second call reuses the context's buffers
            rc = lib.lrfb_qmf_encode_bytes_host(ctx, C.byref(cfg), B, C.c_void_p(imgs.data_ptr()), mj, len(mj),
                                                C.c_void_p(blob.data_ptr()), cap, C.c_void_p(offs.data_ptr()))
            _cabi.check(rc, "lrfb_qmf_encode_bytes_host")
            o, b = offs.numpy(), blob.numpy()
            assert [b[o[i]:o[i + 1]].tobytes() for i in range(B)] == want
        rc = lib.lrfb_qmf_encode_bytes_host(ctx, C.byref(cfg), B, C.c_void_p(imgs.data_ptr()), mj, len(mj),
                                            C.c_void_p(blob.data_ptr()), 1000, C.c_void_p(offs.data_ptr()))
        assert rc == _cabi.LRFB_E_WORKSPACE if hasattr(_cabi, "LRFB_E_WORKSPACE") else rc != 0
    finally:
        lib.lrfb_ctx_destroy(ctx)
    assert compression.qmf_encode_batch(imgs, quality=12) == want


@pytest.mark.gpu
def test_device_packer_large_batch_single_warp_columns():
    """More long columns than single-warp slots: the launch takes one warp per column (the small batches above take four
    warps per long column).  Real factors of distinct images, compared with the native host packer (zlib)."""
    from oracle import qmf_port as port

    H, W, B = 512, 768, 400
    pool = torch.stack([port.s_nat(3000 + i, H, W) for i in range(8)])
    imgs = pool[torch.arange(B) % 8].cuda().contiguous()
    records, lay, meta = compression.qmf_encode_batch(imgs, quality=7, return_records=True)
    rolled = records.clone()
    rolled[B // 2:] = torch.roll(records[B // 2:], 17, dims=1)  # a second family of columns: shifted records
    cfg, _ = compression.resolve_plan(H, W, None, 7, "YCbCr", (0.5, 0.5), (8, 8), (-16, 15), 10)
    want = compression.pack_records(rolled.cpu().numpy(), cfg, lay, meta)
    assert compression.pack_records_device(rolled, cfg, lay, meta) == want
