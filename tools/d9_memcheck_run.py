import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from lrf_b200 import compression, packing
from test_deflate9 import _records
for (H, W, q, B) in [(96, 160, 25, 9), (512, 768, 7, 3), (1365, 2048, 7, 2), (1024, 1032, 3, 2)]:
    cfg, lay = compression.resolve_plan(H, W, None, q, "YCbCr", (0.5, 0.5), (8, 8), (-16, 15), 10)
    meta = compression._metadata(torch.uint8, "YCbCr", True, (-16, 15), (8, 8), lay)
    recs = _records(np.random.default_rng(7), lay, B)
    want = compression.pack_records(recs, cfg, lay, meta)
    got = compression.pack_records_device(torch.from_numpy(recs).cuda(), cfg, lay, meta)
    print(H, W, q, B, got == want, flush=True)
