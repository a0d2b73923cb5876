"""Dev tool: device time of the lossless stage (lrfb_qmf_pack_device) on real factor records of 768x512 images, its
bytes against the host packer, and the host packer's time beside it:  python tools/pack_time.py [B] [distinct] [H W]"""
import ctypes as C
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from lrf_b200 import _cabi, compression, packing
from oracle import qmf_port as port
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
D = int(sys.argv[2]) if len(sys.argv) > 2 else 16
H, W = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (512, 768)
pool = torch.stack([port.s_nat(1000 + i, H, W) for i in range(D)])
imgs = pool[torch.arange(B) % D].cuda().contiguous()
cfg, lay = compression.resolve_plan(H, W, None, 7, "YCbCr", (0.5, 0.5), (8, 8), (-16, 15), 10)
meta = compression._metadata(torch.uint8, "YCbCr", True, (-16, 15), (8, 8), lay)
plan = compression.EncodePlan(cfg, lay, B, imgs.device)
rec = plan.run(imgs).clone()
import os as _os
lib = _cabi.bind(C.CDLL(_os.environ["LRFB_OUT"])) if _os.environ.get("LRFB_OUT") else _cabi.lib()
mj = packing.dict_to_bytes(meta)
wsb = int(lib.lrfb_qmf_pack_device_workspace(C.byref(cfg), B))
cap = B * int(lib.lrfb_qmf_pack_bound(C.byref(cfg), len(mj)))
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
blob = torch.empty(cap, dtype=torch.uint8, device="cuda")
offs = torch.empty(B + 1, dtype=torch.int64, device="cuda")
def run():
    rc = lib.lrfb_qmf_pack_device(C.byref(cfg), B, C.c_void_p(rec.data_ptr()), mj, len(mj), C.c_void_p(blob.data_ptr()), cap,
                                  C.c_void_p(offs.data_ptr()), C.c_void_p(ws.data_ptr()), wsb,
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _cabi.check(rc, "pack_device")
for _ in range(2):
    run()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = min(ts)
o = offs.cpu().numpy()
print(f"pack_device B={B}: {t:.3f} ms -> {B*H*W/1e6/t*1e3:.0f} Mpixel/s; {rec.numel()/1e6:.1f} MB of factors -> {o[B]/1e6:.1f} MB "
      f"({rec.numel()/t/1e6:.1f} GB/s of input), workspace {wsb/1e6:.0f} MB")
host = rec.cpu().numpy()
n = min(B, 256)
t0 = time.perf_counter()
want = compression.pack_records(host[:n], cfg, lay, meta)
th = time.perf_counter() - t0
got = compression.pack_records_device(rec[:n], cfg, lay, meta)
print(f"host packer: {n} images in {th*1e3:.1f} ms -> {n*H*W/1e6/th:.0f} Mpixel/s ({os.cpu_count()} hardware threads); "
      f"device bytes identical: {got == want}")
t0 = time.perf_counter()
got = compression.pack_records_device(rec, cfg, lay, meta)
torch.cuda.synchronize()
td = time.perf_counter() - t0
print(f"pack_records_device (device records -> list[bytes]): {B} images in {td*1e3:.1f} ms -> {B*H*W/1e6/td:.0f} Mpixel/s")
