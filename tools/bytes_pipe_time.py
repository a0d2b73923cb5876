"""Dev tool: lrfb_qmf_encode_bytes_host (pinned host images -> byte streams) at several pipeline chunk sizes:
python tools/bytes_pipe_time.py [B]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lrf_b200 import _cabi, compression, packing
from oracle import qmf_port as port
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
H, W, D = 512, 768, 16
pool = torch.stack([port.s_nat(1000 + i, H, W) for i in range(D)])
h_in = torch.empty((B, 3, H, W), dtype=torch.uint8, pin_memory=True)
h_in.copy_(pool[torch.arange(B) % D])
cfg, lay = compression.resolve_plan(H, W, None, 7, "YCbCr", (0.5, 0.5), (8, 8), (-16, 15), 10)
mj = packing.dict_to_bytes(compression._metadata(torch.uint8, "YCbCr", True, (-16, 15), (8, 8), lay))
lib = _cabi.bind(C.CDLL(os.environ["LRFB_OUT"])) if os.environ.get("LRFB_OUT") else _cabi.lib()
cap = B * int(lib.lrfb_qmf_pack_bound(C.byref(cfg), len(mj)))
blob = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
offs = torch.zeros(B + 1, dtype=torch.int64, pin_memory=True)
recs = torch.empty((B, lay.record_bytes), dtype=torch.int8, pin_memory=True)
for mib in (128, 256, 512, 1024, 2048):
    ctx = C.c_void_p()
    _cabi.check(lib.lrfb_ctx_create(0, C.byref(ctx)), "ctx")
    _cabi.check(lib.lrfb_ctx_set_chunk_bytes(ctx, mib << 20), "chunk")
    def run():
        _cabi.check(lib.lrfb_qmf_encode_bytes_host(ctx, C.byref(cfg), B, C.c_void_p(h_in.data_ptr()), mj, len(mj), C.c_void_p(blob.data_ptr()),
                                                   cap, C.c_void_p(offs.data_ptr())), "bytes_host")
    def run_rec():
        _cabi.check(lib.lrfb_qmf_encode_host(ctx, C.byref(cfg), B, C.c_void_p(h_in.data_ptr()), C.c_void_p(recs.data_ptr())), "host")
    out = []
    for fn in (run, run_rec):
        fn()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        out.append((time.perf_counter() - t0) / 3)
    print(f"chunk {mib:5d} MiB ({(mib << 20) // (3 * H * W)} images): bytes {out[0]*1e3:7.1f} ms -> {B*H*W/1e6/out[0]:7.0f} Mpixel/s | records only "
          f"{out[1]*1e3:7.1f} ms -> {B*H*W/1e6/out[1]:7.0f} Mpixel/s")
    lib.lrfb_ctx_destroy(ctx)
