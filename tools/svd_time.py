"""Dev tool: time lrf_b200.svd_encode_batch (device part) for a batch of 768x512 images:  python tools/svd_time.py B quality"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lrf_b200
from oracle import qmf_port as port
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
q = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
pool = torch.stack([port.s_nat(1000 + i, 512, 768) for i in range(8)])
imgs = pool[torch.arange(B) % 8].cuda().contiguous()
for _ in range(2):
    lrf_b200.svd_encode_batch(imgs, quality=q, return_records=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
lrf_b200.svd_encode_batch(imgs, quality=q, return_records=True)
e1.record()
torch.cuda.synchronize()
print(f"svd_encode B={B} q={q}: {e0.elapsed_time(e1):.1f} ms -> {B * 512 * 768 / 1e6 / e0.elapsed_time(e1) * 1e3:.0f} Mpixel/s")
