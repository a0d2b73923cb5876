import ctypes as C, sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from lrf_b200 import _cabi, compression, packing
from oracle import qmf_port as port
B, D, H, W = 1024, 16, 512, 768
pool = torch.stack([port.s_nat(1000 + i, H, W) for i in range(D)])
imgs = pool[torch.arange(B) % D].cuda().contiguous()
cfg, lay = compression.resolve_plan(H, W, None, 7, "YCbCr", (0.5, 0.5), (8, 8), (-16, 15), 10)
meta = compression._metadata(torch.uint8, "YCbCr", True, (-16, 15), (8, 8), lay)
plan = compression.EncodePlan(cfg, lay, B, imgs.device)
rec = plan.run(imgs).clone()
lib = C.CDLL(os.environ["LRFB_OUT"])
_cabi.bind(lib)
mj = packing.dict_to_bytes(meta)
wsb = int(lib.lrfb_qmf_pack_device_workspace(C.byref(cfg), B))
cap = B * int(lib.lrfb_qmf_pack_bound(C.byref(cfg), len(mj)))
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda"); blob = torch.empty(cap, dtype=torch.uint8, device="cuda"); offs = torch.empty(B + 1, dtype=torch.int64, device="cuda")
def run():
    rc = lib.lrfb_qmf_pack_device(C.byref(cfg), B, C.c_void_p(rec.data_ptr()), mj, len(mj), C.c_void_p(blob.data_ptr()), cap, C.c_void_p(offs.data_ptr()), C.c_void_p(ws.data_ptr()), wsb, None)
    assert rc == 0
run(); torch.cuda.synchronize()
lib.lrfb_d9_prof.argtypes = [C.c_void_p, C.c_int32]
out = (C.c_ulonglong * 16)()
lib.lrfb_d9_prof(None, 1)
run(); torch.cuda.synchronize()
lib.lrfb_d9_prof(out, 0)
v = list(out)
names = ["load+sort", "ranks", "parse ctl", "long walks", "short walks", "trees+hdr", "symbols+out"]
tot = sum(v[:7])
for n_, x in zip(names, v[:7]): print(f"{n_:12s} {x/1e6:10.1f} Mcycles {100*x/tot:5.1f}%")
print("master steps", v[12], "survivors", v[13], "improvements", v[14])
print("long searches", v[8], "candidates", v[9], "| short searches", v[10], "candidates", v[11])
