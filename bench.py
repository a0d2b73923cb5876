#!/usr/bin/env python
"""Headline benchmark: QMF encode Mpixel/s on BASELINE.json's configs[1]
(batch of 4096 synthetic 768x512 RGB images, quality 7, 8x8 patches, bounds (-16,15), 10 sweeps).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One process per GPU (torchrun for N>1).  A "step" is one pass of the hot path (uint8 RGB in HBM → int8
factor records in HBM) over one batch; the batch is sharded across ranks with no data-path collective
(weak scaling: every rank owns a full per-GPU batch); NCCL only gathers per-image bpp / PSNR afterwards.
Prints ONE JSON line on rank 0 (see the driver contract).  `--impl reference` times the reference's CPU
implementation (the oracle port of it — /root/reference does not exist on the GPU box) on host cores.

Besides the headline (`value`, `e2e`, `roofline`, `cpu_baseline`) the line carries, each with its own roofline
fraction: `decode` (lrf.qmf_decode's device part), `clic` (configs[4]: 2048x1365 images, every N), `svd`
(configs[2]), `ablation` (configs[3] end points), `e2e_bytes` (images on the host -> `bytes`, zlib included) and
`quality.vs_reference` (the oracle port run here on a sample of the benchmarked images: dPSNR, dbytes, identical-bytes
fraction through the public path, near-tie accounting).  `--quick` skips those extras.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 512, 768
KW = dict(color_space="YCbCr", scale_factor=(0.5, 0.5), quality=7, patch=True, patch_size=(8, 8),
          bounds=(-16, 15), dtype=torch.int8, num_iters=10)
POOL = 64  # distinct synthetic images, tiled round-robin to the batch
ALG_FLOP_PER_PIXEL = 341.6
ALG_BYTE_PER_PIXEL = 3.08    # u8 RGB in + int8 factors out


def make_pool(n, h=H, w=W):
    from oracle import qmf_port as port  # synthetic generator only (s_nat); not the codec

    return torch.stack([port.s_nat(1000 + i, h, w) for i in range(n)])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc = gpu_index, None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        self.rows = []
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 9:
                self.rows.append(f)

    def summary(self):
        rows = getattr(self, "rows", [])
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        mhz = []
        for r in rows:
            try:
                mhz.append(float(r[1]))
            except ValueError:
                pass
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        try:
            mx = float(rows[0][2])
        except ValueError:
            mx = None
        return {"sm_mhz": statistics.median(mhz) if mhz else None, "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(rows)}


def cuda_time_ms(fn, stream=None):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    fn()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1)


def time_steps_ms(fn, steps, warmup):
    """ms per call of fn: `warmup` untimed calls, then `steps` calls between CUDA events on the current stream."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / steps


def qmf_flops_per_image(lay, iters):
    """SURVEY.md §8(d) table with this geometry's M, N, R: colour + pool, Gram, projection, BCD sweeps."""
    n = lay.cols
    fl = 19.5 * lay.orig_h[0] * lay.orig_w[0]
    for pl in range(lay.n_planes):
        m, r = lay.rows[pl], lay.rank[pl]
        fl += m * n * (n + 1) + 2 * m * n * r + iters * (4 * m * n * r + (m + n) * r * (4 * r + 4))
    return fl


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def cpu_reference_mpix(n_images, threads, warmup=2):
    """Times the reference's CPU path (oracle port: same torch-CPU op sequence incl. zlib, timed the way
    lrf.eval_compression does, utils/misc.py:90-100) on `n_images` images of the workload."""
    from oracle import qmf_port as port

    torch.set_num_threads(threads)
    imgs = [port.s_nat(1000 + i, H, W) for i in range(min(n_images, 8))]
    for i in range(warmup):
        port.qmf_encode(imgs[i % len(imgs)], **KW)
    t0 = time.perf_counter()
    for i in range(n_images):
        port.qmf_encode(imgs[i % len(imgs)], **KW)
    dt = time.perf_counter() - t0
    return n_images * H * W / 1e6 / dt, dt


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = args.ref_images
    vals = []
    for _ in range(args.warmup):
        cpu_reference_mpix(max(1, per_step // 4), threads, warmup=1)
    t_all = 0.0
    for _ in range(args.steps):
        v, dt = cpu_reference_mpix(per_step, threads, warmup=0)
        vals.append(v)
        t_all += dt
    value = per_step * args.steps * H * W / 1e6 / t_all
    line = {
        "impl": "reference", "metric": "QMF encode Mpixel/s (768x512, 10 iters)", "value": value,
        "unit": "Mpixel/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_all / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"qmf_encode 768x512 q7 8x8 (-16,15) 10 iters, {per_step} images/step (bounded sample)",
                   "batch": per_step, "timing": "time.perf_counter around qmf_encode per image, zlib included"},
        "cpu_baseline": {"value": value, "unit": "Mpixel/s", "cores": threads, "kind": "port",
                         "sample": f"{per_step * args.steps} images of the workload, oracle port of the reference"},
        "e2e": {"value": value, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="images per GPU per step")
    ap.add_argument("--ref-images", type=int, default=32, help="images per step for --impl reference")
    ap.add_argument("--cpu-sample", type=int, default=512, help="images for the cpu_baseline leg")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline, e2e, roofline only")
    ap.add_argument("--clic-batch", type=int, default=256, help="2048x1365 images per GPU for the clic leg")
    ap.add_argument("--side-batch", type=int, default=1024, help="images for the ablation leg")
    ap.add_argument("--svd-batch", type=int, default=4096, help="images for the svd leg (configs[2] names the 4096 batch)")
    ap.add_argument("--bytes-images", type=int, default=4096, help="images per GPU for e2e_bytes")
    ap.add_argument("--parity-images", type=int, default=32, help="images compared with the oracle port")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    import torch.distributed as dist

    import lrf_b200
    from lrf_b200 import _cabi, compression

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from lrf_b200.sharding import bind_host_to_gpu

    # N > 1: pinned buffers and copy threads on the GPU's NUMA node (at N = 1 all cores stay available to the host legs)
    affinity = bind_host_to_gpu(local_rank) if world > 1 else {"numa_node": None, "why": "single rank: not bound"}
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.lib()
    B = args.batch

    # ---- inputs resident in HBM -------------------------------------------------------------------
    pool = make_pool(POOL)
    idx = (torch.arange(B) + rank * B) % POOL
    images = pool.to(dev)[idx.to(dev)].contiguous()          # (B,3,H,W) uint8, 4.8 GB at B=4096 (>> L2)
    cfg, lay = compression.resolve_plan(H, W, None, KW["quality"], "YCbCr", KW["scale_factor"],
                                        KW["patch_size"], KW["bounds"], KW["num_iters"])
    plan = compression.EncodePlan(cfg, lay, B, dev)

    def step():
        plan.run(images)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with ClockSampler(local_rank) as clk:  # sampling spans warm-up + timed region (the region itself is short)
        time.sleep(0.3)
        for _ in range(max(args.warmup, 3)):
            step()
        barrier()
        n0 = lib.lrfb_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
        ms_total = e0.elapsed_time(e1)
        launches = lib.lrfb_launch_count() - n0
        t_hold = time.time()
        while time.time() - t_hold < 0.4:  # keep the GPU under the same load while nvidia-smi takes samples
            step()
        torch.cuda.synchronize()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    mpix_step = world * B * H * W / 1e6
    value = mpix_step / (ms_step / 1e3)

    # ---- end to end through the host-buffer C-ABI call (H2D + kernels + D2H inside the timed region) --
    ctx = C.c_void_p()
    _cabi.check(lib.lrfb_ctx_create(local_rank, C.byref(ctx)), "lrfb_ctx_create")
    h_in = torch.empty((B, 3, H, W), dtype=torch.uint8, pin_memory=True)
    h_in.copy_(pool[idx])
    h_out = torch.empty((B, lay.record_bytes), dtype=torch.int8, pin_memory=True)

    def e2e_step():
        _cabi.check(lib.lrfb_qmf_encode_host(ctx, C.byref(cfg), B, C.c_void_p(h_in.data_ptr()),
                                             C.c_void_p(h_out.data_ptr())), "lrfb_qmf_encode_host")

    e2e_step()  # allocates the context's device buffers
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()  # synchronous: returns after the D2H of the factors
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = mpix_step / float(te.item())
    same = bool(torch.equal(h_out, plan.factors.cpu()))

    # ---- the same call ending in the reference's `bytes`: device deflate behind every chunk, only the streams come back --
    from lrf_b200 import packing as _packing

    meta_json = _packing.dict_to_bytes(compression._metadata(torch.uint8, "YCbCr", True, KW["bounds"], KW["patch_size"], lay))
    blob_cap = B * int(lib.lrfb_qmf_pack_bound(C.byref(cfg), len(meta_json)))
    h_blob = torch.empty(blob_cap, dtype=torch.uint8, pin_memory=True)
    h_offs = torch.zeros(B + 1, dtype=torch.int64, pin_memory=True)

    def e2e_bytes_step():
        _cabi.check(lib.lrfb_qmf_encode_bytes_host(ctx, C.byref(cfg), B, C.c_void_p(h_in.data_ptr()), meta_json, len(meta_json),
                                                   C.c_void_p(h_blob.data_ptr()), blob_cap, C.c_void_p(h_offs.data_ptr())),
                    "lrfb_qmf_encode_bytes_host")

    e2e_bytes_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_bytes_step()  # synchronous: returns when every stream is in host memory
    barrier()
    tb = torch.tensor([(time.perf_counter() - t0) / args.e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
    e2e_bytes_value = mpix_step / float(tb.item())
    e2e_blob_bytes = int(h_offs[B].item())

    # ---- and the way back with host buffers (N = 1 only: one more pinned image buffer): encoded images -> uint8 images ---
    decode_e2e = None
    if world == 1 and not args.quick:
        h_dec = torch.empty((B, 3, H, W), dtype=torch.uint8, pin_memory=True)

        def dec_bytes_step():
            _cabi.check(lib.lrfb_qmf_decode_bytes_host(ctx, C.byref(cfg), B, C.c_void_p(h_blob.data_ptr()), C.c_void_p(h_offs.data_ptr()),
                                                       C.c_void_p(h_dec.data_ptr())), "lrfb_qmf_decode_bytes_host")

        dec_bytes_step()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            dec_bytes_step()
        tdb = (time.perf_counter() - t0) / args.e2e_steps
        ref_dec = compression.decode_records(plan.factors[:8], cfg).cpu()
        decode_e2e = {"value": mpix_step / tdb, "unit": "Mpixel/s", "h2d_bytes_per_step": e2e_blob_bytes + 8 * (B + 1),
                      "d2h_bytes_per_step": int(h_dec.numel()), "images_equal_device_decode_sample": bool(torch.equal(h_dec[:8], ref_dec)),
                      "call": "lrfb_qmf_decode_bytes_host (C ABI): encoded images in host memory -> uint8 images in pinned host memory "
                              "(device un-framing + inflate + decode, chunked copy-back); bound by the D2H copy of the images"}
        del h_dec
    lib.lrfb_ctx_destroy(ctx)

    # ---- plain concurrent H2D copy of the same bytes: the machine's ceiling for `e2e` (PCIe / host memory) ----------
    d_probe = torch.empty_like(images)
    barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        d_probe.copy_(h_in, non_blocking=True)
    barrier()
    h2d_s = (time.perf_counter() - t0) / 2
    th = torch.tensor([h2d_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(th, op=dist.ReduceOp.MAX)
    h2d_gbs_per_gpu = int(h_in.numel()) / float(th.item()) / 1e9
    e2e_ceiling = mpix_step / float(th.item())  # Mpixel/s if the step were nothing but that copy
    del d_probe

    # ---- roofline of the dominant kernel (BCD sweeps on the luma planes), timed alone with CUDA events --
    stream = torch.cuda.current_stream()
    xy, uy, vy = plan.view("x", 0), plan.view("u", 0), plan.view("v", 0)
    dbg = _cabi.QmfDebug()
    dbg.stop_after = 2
    plan.run(images, dbg)  # leaves the SVD init in u/v
    v0 = vy.clone()
    so = plan.map.sigma[0] + B * lay.rank[0] * 8  # f32 singular values sit behind the f64 ones
    s0y = plan.workspace[so: so + B * lay.rank[0] * 4].view(torch.float32).clone()
    wsb = 32768 * 16 * 4
    bws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    My, N, Ry = lay.rows[0], lay.cols, lay.rank[0]

    def bcd_only():
        _cabi.check(lib.lrfb_bcd(C.c_void_p(xy.data_ptr()), B, My, N, Ry, -16.0, 15.0, KW["num_iters"],
                                 C.c_void_p(uy.data_ptr()), C.c_void_p(vy.data_ptr()), C.c_void_p(s0y.data_ptr()), 1,
                                 C.c_void_p(bws.data_ptr()), wsb, C.c_void_p(stream.cuda_stream)), "lrfb_bcd")

    bcd_ms = []
    for i in range(4):
        vy.copy_(v0)
        torch.cuda.synchronize()
        bcd_ms.append(cuda_time_ms(bcd_only))
    bcd_ms = statistics.median(bcd_ms[1:])
    plan.run(images)  # restore the full result
    # algorithmic work of that launch: X read once + U in/out + V in/out; flops per SURVEY §8(d) BCD row
    alg_bytes = B * (My * N * 4 + My * Ry * 4 + 2 * N * Ry * 4)  # X once, U out, V in/out
    alg_flops = B * KW["num_iters"] * (4 * My * N * Ry + (My + N) * Ry * (4 * Ry + 4))
    peaks, peak_src = measured_peaks()
    # FP32 FFMA peak measured here (not in MEASURED_PEAKS.json)
    probe = torch.empty(148 * 8 * 256 + 1024, dtype=torch.float32, device=dev)
    it = 4096
    lib.lrfb_ffma_probe(C.c_void_p(probe.data_ptr()), it, C.c_void_p(stream.cuda_stream))
    ff_ms = min(cuda_time_ms(lambda: lib.lrfb_ffma_probe(C.c_void_p(probe.data_ptr()), it,
                                                         C.c_void_p(stream.cuda_stream))) for _ in range(3))
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    fp32_peak = sms * 8 * 256 * it * 64 / (ff_ms / 1e3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tp):
        tp = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tp):  # dram__bytes_read+write of this kernel from the committed ncu --set full capture
        with open(tp) as f:
            traffic = json.load(f)["bytes_per_image"] * B
    achieved_tf = alg_flops / (bcd_ms / 1e3) / 1e12
    achieved_gbs = alg_bytes / (bcd_ms / 1e3) / 1e9
    # The kernel keeps X on chip for all ten sweeps, so it is bound by arithmetic, not by HBM: the roofline is FP32.
    roofline = {"bound": "fp32", "kernel": "bcd_tc_kernel<4,768,256>: all 10 BCD sweeps on the luma planes (lrfb_bcd)",
                "achieved": achieved_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp32_peak,
                "peak_source": "FP32 FFMA throughput measured in this run by lrfb_ffma_probe (MEASURED_PEAKS.json has no "
                               "FP32 entry; nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.5)",
                "traffic": traffic, "ms_per_launch": bcd_ms, "algorithmic_flops_per_launch": alg_flops,
                "hbm": {"achieved_gbs": achieved_gbs, "peak_gbs": peaks["hbm_gbs"], "peak_source": peak_src,
                        "frac": achieved_gbs / peaks["hbm_gbs"], "algorithmic_bytes_per_launch": alg_bytes}}
    encode_fp32_frac = value * 1e6 / world * ALG_FLOP_PER_PIXEL / 1e12 / fp32_peak

    def allmax(x):
        t_ = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return float(t_.item())

    extras = {}
    # ---- decode (lrf.qmf_decode's device part: int8 records -> uint8 RGB), HBM-bound -----------------------------
    if not args.quick:
        dec_buf = {}

        def dec_step():
            dec_buf["out"] = compression.decode_records(plan.factors, cfg)

        dms = allmax(time_steps_ms(dec_step, 5, 3))
        dec_bytes = B * (3 * H * W + lay.record_bytes)
        extras["decode"] = {"value": mpix_step / (dms / 1e3), "unit": "Mpixel/s", "ms_per_step": dms,
                            "workload": "qmf_decode device part on the %d records per GPU of the encode above" % B,
                            "roofline": {"bound": "hbm", "achieved": dec_bytes / (dms / 1e3) / 1e9,
                                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "peak_source": peak_src,
                                         "frac": dec_bytes / (dms / 1e3) / 1e9 / peaks["hbm_gbs"]}}
        dec_buf.clear()

    # ---- quality / parity epilogue: per-image PSNR on device, bpp from host zlib on a sample; NCCL gather --
    dec = compression.decode_records(plan.factors, cfg)
    psnr = compression.psnr_batch(dec, images).float()
    del dec
    n_s = min(B, POOL)
    host = plan.factors[:n_s].cpu().numpy()
    meta = compression._metadata(torch.uint8, "YCbCr", True, KW["bounds"], KW["patch_size"], lay)
    from lrf_b200 import packing

    t0 = time.perf_counter()
    blobs = compression.pack_records(host, cfg, lay, meta)
    pack_s = time.perf_counter() - t0
    bpp = torch.tensor([len(b) * 8 / (H * W) for b in blobs], dtype=torch.float32, device=dev)
    from lrf_b200.sharding import gather_stats

    stats = torch.stack([bpp, psnr[:n_s]], dim=1)
    stats = gather_stats(stats, n_s * world, rank, world).cpu()  # the only collective (NCCL all_gather)
    psnr_host = psnr[:n_s].cpu()

    # the streams the bytes-ending e2e call left in host memory against the host packer's on the same images
    ho = h_offs.numpy()
    hb = memoryview(h_blob.numpy())
    bytes_equal = all(bytes(hb[ho[i]: ho[i + 1]]) == blobs[i] for i in range(n_s))

    # ---- lossless stage alone, device resident: records of the step above -> framed zlib-9 streams ---------------------
    pws = int(lib.lrfb_qmf_pack_device_workspace(C.byref(cfg), B))
    d_pws = torch.empty(pws, dtype=torch.uint8, device=dev)
    d_blob = torch.empty(blob_cap, dtype=torch.uint8, device=dev)
    d_offs = torch.empty(B + 1, dtype=torch.int64, device=dev)

    def pack_only():
        _cabi.check(lib.lrfb_qmf_pack_device(C.byref(cfg), B, C.c_void_p(plan.factors.data_ptr()), meta_json, len(meta_json),
                                             C.c_void_p(d_blob.data_ptr()), blob_cap, C.c_void_p(d_offs.data_ptr()),
                                             C.c_void_p(d_pws.data_ptr()), pws, C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                    "lrfb_qmf_pack_device")

    barrier()
    pms = allmax(time_steps_ms(pack_only, 2, 3))
    extras["pack"] = {"value": mpix_step / (pms / 1e3), "unit": "Mpixel/s", "ms_per_step": pms,
                      "workload": "lrfb_qmf_pack_device on the %d records per GPU of the encode above: zlib level 9 per factor "
                                  "column (one warp each) + framing, byte-identical to the host packer" % B,
                      "factor_bytes_in": int(plan.factors.numel()), "stream_bytes_out": int(d_offs[B].item()),
                      "vs_host_packer": (mpix_step / world / (pms / 1e3)) / (n_s * H * W / 1e6 / pack_s),
                      "roofline": {"bound": "issue", "what": "integer / shared-memory latency work: ncu (profiles/r2_full_deflate9_b4096.txt) "
                                   "shows 40 % of the SM issue rate with 11 columns (warps) per SM and DRAM at 0.1 %; the HBM fraction of "
                                   "the algorithmic bytes (factors in + streams out) is given to show it is not the bound",
                                   "hbm_frac": (int(plan.factors.numel()) + int(d_offs[B].item())) / (pms / 1e3) / 1e9 / peaks["hbm_gbs"]}}
    # ---- and back: un-frame + inflate those streams into records on the device (the head of qmf_decode) ----------------
    uws = int(lib.lrfb_qmf_unpack_device_workspace(C.byref(cfg), B))
    d_uws = torch.empty(uws, dtype=torch.uint8, device=dev)
    d_rec2 = torch.empty_like(plan.factors)

    def unpack_only():
        _cabi.check(lib.lrfb_qmf_unpack_device(C.byref(cfg), B, C.c_void_p(d_blob.data_ptr()), C.c_void_p(d_offs.data_ptr()),
                                               C.c_void_p(d_rec2.data_ptr()), C.c_void_p(d_uws.data_ptr()), uws,
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)), "lrfb_qmf_unpack_device")

    barrier()
    ums = allmax(time_steps_ms(unpack_only, 2, 1))
    extras["unpack"] = {"value": mpix_step / (ums / 1e3), "unit": "Mpixel/s", "ms_per_step": ums,
                        "workload": "lrfb_qmf_unpack_device on the %d streams per GPU written above: framing walked, every column "
                                    "inflated (one warp each), adler32 checked" % B,
                        "records_equal": bool(torch.equal(d_rec2, plan.factors))}
    del d_pws, d_blob, d_offs, d_uws, d_rec2

    # ---- end to end to `bytes` through the Python API: pinned host uint8 images -> list[bytes] -------------------------
    if not args.quick:
        n_b = min(B, args.bytes_images)
        lrf_b200.qmf_encode_batch(h_in[:n_b], **KW)  # warm-up at the same size: the context's buffers are grow-only
        barrier()
        t0 = time.perf_counter()
        out_b = lrf_b200.qmf_encode_batch(h_in[:n_b], **KW)
        tb = allmax(time.perf_counter() - t0)
        extras["e2e_bytes"] = {"value": world * n_b * H * W / 1e6 / tb, "unit": "Mpixel/s", "images_per_gpu": n_b,
                               "call": "lrf_b200.qmf_encode_batch(pinned host uint8 images) -> list[bytes] "
                                       "(lrfb_qmf_encode_bytes_host + one bytes object per image)",
                               "seconds": tb, "bytes_equal_sample_pack": out_b[:n_s] == blobs[:min(n_s, n_b)]}
        del out_b

    # free the 768x512 working set before the other shapes
    del plan, images, xy, uy, vy, v0, s0y
    torch.cuda.empty_cache()

    # ---- configs[4]: CLIC-sized images (2048 x 1365), sharded like the headline: every rank its own batch ----------
    if not args.quick:
        Hc, Wc, Bc = 1365, 2048, args.clic_batch
        from oracle import qmf_port as port

        cpool = torch.stack([port.s_nat(1000 + i, Hc, Wc) for i in range(4)])
        cimgs = cpool.to(dev)[((torch.arange(Bc) + rank * Bc) % 4).to(dev)].contiguous()
        ccfg, clay = compression.resolve_plan(Hc, Wc, None, KW["quality"], "YCbCr", KW["scale_factor"], KW["patch_size"],
                                              KW["bounds"], KW["num_iters"])
        cplan = compression.EncodePlan(ccfg, clay, Bc, dev)
        barrier()
        cms = allmax(time_steps_ms(lambda: cplan.run(cimgs), 3, 3))
        cmpix = world * Bc * Hc * Wc / 1e6
        cfl = qmf_flops_per_image(clay, KW["num_iters"])
        ctf = world * Bc * cfl / (cms / 1e3) / 1e12 / world
        extras["clic"] = {"value": cmpix / (cms / 1e3), "unit": "Mpixel/s", "ms_per_step": cms,
                          "workload": "configs[4] shape: %d synthetic 2048x1365 images per GPU (4 distinct, %.1f GB >> L2), "
                                      "quality 7, 10 sweeps; M = (43776, 11008, 11008)" % (Bc, cimgs.numel() / 1e9),
                          "roofline": {"bound": "fp32", "achieved": ctf, "peak": fp32_peak, "unit": "TFLOP/s",
                                       "frac": ctf / fp32_peak, "flop_per_pixel": cfl / (Hc * Wc)}}
        # the lossless stage on those records (43 776-byte luma columns: several deflate blocks each)
        cmeta = _packing.dict_to_bytes(compression._metadata(torch.uint8, "YCbCr", True, KW["bounds"], KW["patch_size"], clay))
        cpws = int(lib.lrfb_qmf_pack_device_workspace(C.byref(ccfg), Bc))
        if cpws > 0:
            ccap = Bc * int(lib.lrfb_qmf_pack_bound(C.byref(ccfg), len(cmeta)))
            c_ws = torch.empty(cpws, dtype=torch.uint8, device=dev)
            c_blob = torch.empty(ccap, dtype=torch.uint8, device=dev)
            c_offs = torch.empty(Bc + 1, dtype=torch.int64, device=dev)

            def cpack():
                _cabi.check(lib.lrfb_qmf_pack_device(C.byref(ccfg), Bc, C.c_void_p(cplan.factors.data_ptr()), cmeta, len(cmeta),
                                                     C.c_void_p(c_blob.data_ptr()), ccap, C.c_void_p(c_offs.data_ptr()),
                                                     C.c_void_p(c_ws.data_ptr()), cpws,
                                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)), "lrfb_qmf_pack_device")

            barrier()
            cpms = allmax(time_steps_ms(cpack, 2, 1))
            extras["clic"]["pack"] = {"value": cmpix / (cpms / 1e3), "unit": "Mpixel/s", "ms_per_step": cpms,
                                      "stream_bytes_out": int(c_offs[Bc].item()), "factor_bytes_in": int(cplan.factors.numel())}
            del c_ws, c_blob, c_offs
        del cplan, cimgs
        torch.cuda.empty_cache()

    # ---- configs[2] (svd_encode) and configs[3] (ablation end points): N = 1 only ----------------------------------
    if not args.quick and world == 1:
        from oracle import qmf_port as port

        Bs, Bv = args.side_batch, args.svd_batch
        spool = make_pool(8).to(dev)
        vimgs = spool[(torch.arange(Bv) % 8).to(dev)].contiguous()
        svd = {}
        for q in (1.0, 7):
            sms_ = time_steps_ms(lambda: lrf_b200.svd_encode_batch(vimgs, quality=q, return_records=True), 2, 1)
            svd["quality_%g" % q] = {"value": Bv * H * W / 1e6 / (sms_ / 1e3), "unit": "Mpixel/s", "ms_per_step": sms_,
                                     "rank": max(round(192 * q / 100), 1)}
        svd["workload"] = "configs[2]: svd_encode (RGB, 8x8 patches, M x 192 matrices) on %d 768x512 images" % Bv
        extras["svd"] = svd
        del vimgs
        torch.cuda.empty_cache()
        simgs = spool[(torch.arange(Bs) % 8).to(dev)].contiguous()
        abl = {}
        cases = {"iters_1": dict(num_iters=1), "iters_50": dict(num_iters=50), "patch_4x4": dict(patch_size=(4, 4)),
                 "patch_16x16": dict(patch_size=(16, 16)), "bounds_-8_7": dict(bounds=(-8, 7)),
                 "bounds_-128_127": dict(bounds=(-128, 127))}
        for name, over in cases.items():
            kw = dict(KW, **over)
            acfg, alay = compression.resolve_plan(H, W, None, kw["quality"], "YCbCr", kw["scale_factor"], kw["patch_size"],
                                                  kw["bounds"], kw["num_iters"])
            ab = Bs if name != "patch_16x16" else max(8, Bs // 4)  # N = 256, R = 18: generic kernels, one warp per eigen-problem
            aplan = compression.EncodePlan(acfg, alay, ab, dev)
            ai = simgs[:ab].contiguous()
            ams = time_steps_ms(lambda: aplan.run(ai), 2, 2)
            afl = qmf_flops_per_image(alay, kw["num_iters"])
            abl[name] = {"value": ab * H * W / 1e6 / (ams / 1e3), "unit": "Mpixel/s", "ms_per_step": ams, "batch": ab,
                         "fp32_frac": ab * afl / (ams / 1e3) / 1e12 / fp32_peak}
            del aplan
        abl["workload"] = "configs[3] end points on 768x512 images (quality 7; everything else as the headline)"
        extras["ablation"] = abl
        del simgs
        torch.cuda.empty_cache()

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()

    if rank == 0:
        quality = {"mean_bpp": float(stats[:, 0].mean()), "mean_psnr_db": float(stats[:, 1].mean()),
                   "images": int(stats.shape[0])}
        if not args.quick:
            quality["vs_reference"] = parity_vs_reference(pool, blobs, psnr_host, host, lay, args.parity_images)
        line = {
            "metric": "QMF encode Mpixel/s (768x512, 10 iters)", "value": value, "unit": "Mpixel/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "configs[1]: batch of %d synthetic 768x512 RGB uint8 images per GPU, QMF "
                                   "quality 7, 8x8 patches, bounds (-16,15), int8, 10 sweeps" % B,
                       "batch_per_gpu": B, "distinct_images": POOL, "generator": "s_nat(seed=1000+i) SURVEY §8d",
                       "l2": "inputs %.1f GB per GPU >> 126 MB L2, no flush needed" % (B * 3 * H * W / 1e9),
                       "sharding": "images split across ranks, no data-path collective"},
            "e2e": {"value": e2e_bytes_value, "unit": "Mpixel/s", "h2d_bytes_per_step": int(h_in.numel()) * world,
                    "d2h_bytes_per_step": (e2e_blob_bytes + 8 * (B + 1)) * world,
                    "streams_equal_host_packer_sample": bytes_equal,
                    "call": "lrfb_qmf_encode_bytes_host (C ABI, pinned host buffers): host images -> the reference's "
                            "encoded byte streams in host memory (H2D, encode kernels, device zlib-9 deflate + framing, "
                            "D2H of the streams) - the same end point as the CPU arm, which includes zlib",
                    "host_affinity": affinity,
                    "records_only": {"value": e2e_value, "unit": "Mpixel/s", "d2h_bytes_per_step": int(h_out.numel()) * world,
                                     "records_equal_device_path": same,
                                     "call": "lrfb_qmf_encode_host: stops at the int8 factor records in host memory"},
                    "h2d_ceiling": {"value": e2e_ceiling, "unit": "Mpixel/s", "gbs_per_gpu": h2d_gbs_per_gpu,
                                    "what": "the same pinned input copied to the device by plain concurrent "
                                            "cudaMemcpyAsync on every rank, nothing else running",
                                    "frac_of_ceiling": e2e_bytes_value / e2e_ceiling}},
            "decode_e2e": decode_e2e,
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "roofline": roofline,
            "encode_fp32_frac": encode_fp32_frac,
            "encode_hbm_frac": value * 1e6 / world * ALG_BYTE_PER_PIXEL / 1e9 / peaks["hbm_gbs"],
            "quality": quality,
            "host_pack": {"images": n_s, "seconds": pack_s, "threads": os.cpu_count(),
                          "mpixel_per_s": n_s * H * W / 1e6 / pack_s,
                          "call": "lrfb_qmf_pack_host (native zlib-9 thread pool)"},
        }
        line.update(extras)
        if not args.no_cpu_baseline and world == 1:  # N = 1 only: the other ranks must not wait on host work
            cores = os.cpu_count() or 1
            v, dt = cpu_reference_mpix(args.cpu_sample, cores)
            line["cpu_baseline"] = {"value": v, "unit": "Mpixel/s", "cores": cores, "kind": "port",
                                    "sample": f"{args.cpu_sample} images of the workload in {dt:.1f} s "
                                              "(oracle port of the reference, zlib included)"}
        print(json.dumps(line))


def parity_vs_reference(pool, blobs, psnr_gpu, records, lay, n_img):
    """north_star's parity figures on a sample of the benchmarked images: the oracle port (the reference's op sequence
    on torch CPU) encodes the same images; compared with what the GPU path produced through the PUBLIC route (no test
    hook): bytes, PSNR, factors, and a near-tie count for every image that differs."""
    from oracle import exact
    from oracle import qmf_port as port

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from backends import split_record

    torch.set_num_threads(os.cpu_count() or 1)
    n = min(n_img, len(blobs), pool.shape[0])
    dps, dbs, ident_b, ident_f, near_imgs, unexplained = [], [], 0, 0, 0, 0
    for i in range(n):
        img = pool[i]
        blob_ref, ref, meta = port.qmf_encode(img, return_factors=True, **KW)
        p_ref = port.psnr(img, port.qmf_decode(blob_ref))
        dps.append(abs(float(psnr_gpu[i]) - p_ref))
        dbs.append(abs(len(blobs[i]) - len(blob_ref)))
        ident_b += blobs[i] == blob_ref
        got = split_record(records[i], lay)
        same = all(np.array_equal(g, r.numpy()) for g, r in zip(got, ref))
        ident_f += same
        if not same:
            near = 0
            for pl, (x, _, _) in enumerate(port.qmf_planes(img)):
                u0, v0 = port.svd_init(x.unsqueeze(0), meta["rank"][pl])
                near += exact.bcd(x.numpy(), u0.squeeze(0).numpy(), v0.squeeze(0).numpy(), KW["bounds"],
                                  KW["num_iters"])[2].near_ties
            near_imgs += near > 0
            unexplained += near == 0
    return {"images": n, "oracle": "oracle/qmf_port.py (torch CPU port of the reference) run in this process",
            "max_abs_dpsnr_db": max(dps), "mean_abs_dbytes": float(np.mean(dbs)), "max_abs_dbytes": int(max(dbs)),
            "frac_identical_bytes": ident_b / n, "frac_identical_factors": ident_f / n,
            "route": "public path, nothing injected: the SVD init carries LAPACK's signs by the closed-form rule",
            "differing_images_with_near_ties_1e-5": near_imgs, "differing_images_unexplained": unexplained}


if __name__ == "__main__":
    main()
