"""Dev tool: device time of the un-framing + inflate (lrfb_qmf_unpack_device) on the streams of B 768x512 images, and the
whole qmf_decode_batch(list[bytes]) beside the per-image host path:  python tools/unpack_time.py [B] [distinct]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lrf_b200 import _cabi, compression, packing
from oracle import qmf_port as port
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
D = int(sys.argv[2]) if len(sys.argv) > 2 else 16
H, W = 512, 768
pool = torch.stack([port.s_nat(1000 + i, H, W) for i in range(D)])
imgs = pool[torch.arange(B) % D].cuda().contiguous()
cfg, lay = compression.resolve_plan(H, W, None, 7, "YCbCr", (0.5, 0.5), (8, 8), (-16, 15), 10)
meta = compression._metadata(torch.uint8, "YCbCr", True, (-16, 15), (8, 8), lay)
plan = compression.EncodePlan(cfg, lay, B, imgs.device)
rec = plan.run(imgs).clone()
blobs = compression.pack_records_device(rec, cfg, lay, meta)
lib = _cabi.lib()
sizes = np.fromiter((len(e) for e in blobs), np.int64, B)
offs = np.zeros(B + 1, np.int64); np.cumsum(sizes, out=offs[1:])
d_blob = torch.frombuffer(bytearray(b"".join(blobs)), dtype=torch.uint8).cuda()
d_offs = torch.from_numpy(offs).cuda()
wsb = int(lib.lrfb_qmf_unpack_device_workspace(C.byref(cfg), B))
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
out = torch.empty_like(rec)
def run():
    _cabi.check(lib.lrfb_qmf_unpack_device(C.byref(cfg), B, C.c_void_p(d_blob.data_ptr()), C.c_void_p(d_offs.data_ptr()), C.c_void_p(out.data_ptr()),
                                           C.c_void_p(ws.data_ptr()), wsb, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "unpack")
for _ in range(2):
    run()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = min(ts)
print(f"unpack_device B={B}: {t:.3f} ms -> {B*H*W/1e6/t*1e3:.0f} Mpixel/s; records identical: {bool(torch.equal(out, rec))}")
compression.DEVICE_UNPACK = True
t0 = time.perf_counter(); dec = compression.qmf_decode_batch(blobs); torch.cuda.synchronize(); t1 = time.perf_counter() - t0
n = min(B, 256)
t0 = time.perf_counter(); parsed = list(compression._pool().map(compression._parse_encoded, blobs[:n])); t2 = time.perf_counter() - t0
print(f"qmf_decode_batch(list[bytes]) -> images on the device: {B} images in {t1*1e3:.1f} ms -> {B*H*W/1e6/t1:.0f} Mpixel/s; "
      f"host un-framing + zlib of {n} images (thread pool): {t2*1e3:.1f} ms -> {n*H*W/1e6/t2:.0f} Mpixel/s")
