"""Dev tool: launch only the dominant kernel (BCD sweeps on the luma planes, via lrfb_bcd) at bench.py's
configuration, for `ncu --set full -k regex:bcd_tc` captures (profiles/)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from lrf_b200 import _cabi, compression
from oracle import qmf_port as port

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
H, W = 512, 768
pool = torch.stack([port.s_nat(1000 + i, H, W) for i in range(16)])
imgs = pool[torch.arange(B) % 16].cuda().contiguous()
cfg, lay = compression.resolve_plan(H, W, None, 7, "YCbCr", (0.5, 0.5), (8, 8), (-16, 15), 10)
plan = compression.EncodePlan(cfg, lay, B, imgs.device)
plan.run(imgs)  # fills x and leaves final factors; v holds integers: re-derive the init for the sweeps
dbg = _cabi.QmfDebug()
dbg.stop_after = 2
plan.run(imgs, dbg)
x, u, v = plan.view("x", 0), plan.view("u", 0), plan.view("v", 0)
# the init's U columns are not materialised by the encode path: U0 = X V0 / s (same values the kernel derives)
s = plan.workspace[plan.map.sigma[0] + B * lay.rank[0] * 8: plan.map.sigma[0] + B * lay.rank[0] * 12].view(torch.float32).view(B, 1, -1)
u.copy_(torch.bmm(x, v) / s)
u0, v0 = u.clone(), v.clone()
ws = torch.empty(32768 * 16 * 4, dtype=torch.uint8, device=imgs.device)
lib = _cabi.lib()
s0 = s.reshape(B, -1).contiguous()
for it in range(6):
    u.copy_(u0), v.copy_(v0)
    mode_s0 = it >= 3
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _cabi.check(lib.lrfb_bcd(C.c_void_p(x.data_ptr()), B, lay.rows[0], lay.cols, lay.rank[0], -16.0, 15.0, 10,
                             C.c_void_p(u.data_ptr()), C.c_void_p(v.data_ptr()),
                             C.c_void_p(s0.data_ptr()) if mode_s0 else None, 1, C.c_void_p(ws.data_ptr()),
                             ws.numel(), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "lrfb_bcd")
    e1.record()
    torch.cuda.synchronize()
    print(f"lrfb_bcd B={B} from_a={mode_s0}: {e0.elapsed_time(e1):.3f} ms")
