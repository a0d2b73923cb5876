"""CPU-only, world_size 2 over gloo: the multi-GPU plumbing of bench.py — contiguous sharding of the image
batch across ranks (no data-path collective) and the single epilogue all_gather of per-image (bpp, PSNR)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lrf_b200.sharding import gather_stats, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    # stand-in for the per-image results of this rank's shard: (bpp, psnr) derived from the image index
    idx = torch.arange(lo, hi, dtype=torch.float32)
    stats = torch.stack([idx * 0.001, 20.0 + idx], dim=1)
    allstats = gather_stats(stats, total, rank, world)
    if rank == 0:
        torch.save(allstats, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 7, 2])
def test_shard_and_gather_world2(tmp_path, total):
    out = str(tmp_path / "stats.pt")
    mp.spawn(_worker, args=(2, _free_port(), total, out), nprocs=2, join=True)
    got = torch.load(out)
    idx = torch.arange(total, dtype=torch.float32)
    assert torch.equal(got, torch.stack([idx * 0.001, 20.0 + idx], dim=1))


def test_shard_ranges_cover_batch_exactly():
    for total in (1, 2, 5, 4096, 65536, 65537):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
