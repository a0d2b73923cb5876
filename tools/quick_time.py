"""Dev tool: stage timings of the device-resident encode on the GPU box (CUDA events)."""
import ctypes as C
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lrf_b200
from lrf_b200 import _cabi, compression
from oracle import qmf_port as port

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (512, 768)
pool = torch.stack([port.s_nat(1000 + i, H, W) for i in range(8)])
imgs = pool[torch.arange(B) % 8].cuda().contiguous()
cfg, lay = compression.resolve_plan(H, W, None, 7, "YCbCr", (0.5, 0.5), (8, 8), (-16, 15), 10)
plan = compression.EncodePlan(cfg, lay, B, imgs.device)
print("workspace GB", plan.map.total_bytes / 1e9, "record bytes", lay.record_bytes)

def timed(stop_after, reps=3):
    dbg = _cabi.QmfDebug(); dbg.stop_after = stop_after
    plan.run(imgs, dbg); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.run(imgs, dbg); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)

t1, t2, t0 = timed(1), timed(2), timed(0)
mp = B * H * W / 1e6
print(f"B={B} {H}x{W}: frontend {t1:.2f} ms | +svd-init {t2 - t1:.2f} ms | +bcd {t0 - t2:.2f} ms | total {t0:.2f} ms"
      f" -> {mp / t0 * 1e3:.0f} Mpixel/s")
rec = plan.factors
t = time.time(); dec = compression.decode_records(rec, cfg); torch.cuda.synchronize(); 
td = []
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dec = compression.decode_records(rec, cfg); e1.record(); torch.cuda.synchronize()
    td.append(e0.elapsed_time(e1))
td = min(td)
print(f"decode {td:.2f} ms -> {mp / td * 1e3:.0f} Mpixel/s; PSNR[0..3]",
      compression.psnr_batch(dec, imgs)[:4].tolist())
