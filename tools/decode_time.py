"""Dev tool: device time of the QMF decode for a batch of 768x512 images:  python tools/decode_time.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lrf_b200 import compression
from oracle import qmf_port as port
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
H, W = 512, 768
pool = torch.stack([port.s_nat(1000 + i, H, W) for i in range(8)])
imgs = pool[torch.arange(B) % 8].cuda().contiguous()
cfg, lay = compression.resolve_plan(H, W, None, 7, "YCbCr", (0.5, 0.5), (8, 8), (-16, 15), 10)
plan = compression.EncodePlan(cfg, lay, B, imgs.device)
rec = plan.run(imgs).clone()
def run():
    return compression.decode_records(rec, cfg)
for _ in range(3):
    out = run()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = min(ts)
print(f"decode B={B}: {t:.3f} ms -> {B*H*W/1e6/t*1e3:.0f} Mpixel/s, {B*H*W*3.08/t/1e6:.0f} GB/s algorithmic")
