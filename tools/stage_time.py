"""Dev tool: stage timings (front end | +SVD init | +sweeps) of the device-resident encode for any configuration.

    python tools/stage_time.py B H W [patch] [iters] [lo hi]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from lrf_b200 import _cabi, compression
from oracle import qmf_port as port

B, H, W = (int(a) for a in sys.argv[1:4])
P = int(sys.argv[4]) if len(sys.argv) > 4 else 8
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 10
lo, hi = (int(sys.argv[6]), int(sys.argv[7])) if len(sys.argv) > 7 else (-16, 15)
pool = torch.stack([port.s_nat(1000 + i, H, W) for i in range(4)])
imgs = pool[torch.arange(B) % 4].cuda().contiguous()
cfg, lay = compression.resolve_plan(H, W, None, 7, "YCbCr", (0.5, 0.5), (P, P), (lo, hi), iters)
plan = compression.EncodePlan(cfg, lay, B, imgs.device)


def timed(stop_after, reps=3):
    dbg = _cabi.QmfDebug()
    dbg.stop_after = stop_after
    plan.run(imgs, dbg)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.run(imgs, dbg)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


t1, t2, t0 = timed(1), timed(2), timed(0)
mp = B * H * W / 1e6
print(f"B={B} {H}x{W} patch {P} ranks {[lay.rank[i] for i in range(3)]} rows {[lay.rows[i] for i in range(3)]} iters {iters}: "
      f"frontend {t1:.2f} ms | +svd-init {t2 - t1:.2f} ms | +sweeps {t0 - t2:.2f} ms | total {t0:.2f} ms -> {mp / t0 * 1e3:.0f} Mpixel/s")
