#!/usr/bin/env python
"""Headline benchmark: QMF encode Mpixel/s on BASELINE.json's configs[1]
(batch of 4096 synthetic 768x512 RGB images, quality 7, 8x8 patches, bounds (-16,15), 10 sweeps).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One process per GPU (torchrun for N>1).  A "step" is one pass of the hot path (uint8 RGB in HBM → int8
factor records in HBM) over one batch; the batch is sharded across ranks with no data-path collective
(weak scaling: every rank owns a full per-GPU batch); NCCL only gathers per-image bpp / PSNR afterwards.
Prints ONE JSON line on rank 0 (see the driver contract).  `--impl reference` times the reference's CPU
implementation (the oracle port of it — /root/reference does not exist on the GPU box) on host cores.
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
    lib.lrfb_ctx_destroy(ctx)

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
    tp = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tp):  # dram__bytes_read+write of this kernel from the committed ncu --set full capture
        with open(tp) as f:
            traffic = json.load(f)["bytes_per_image"] * B
    achieved_gbs = alg_bytes / (bcd_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "kernel": "bcd_tc_kernel<4,768,384>: all 10 BCD sweeps on the luma planes (lrfb_bcd)", "achieved": achieved_gbs,
                "peak": peaks["hbm_gbs"], "peak_source": peak_src, "unit": "GB/s",
                "frac": achieved_gbs / peaks["hbm_gbs"], "traffic": traffic,
                "ms_per_launch": bcd_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "fp32": {"achieved_tflops": alg_flops / (bcd_ms / 1e3) / 1e12, "peak_tflops": fp32_peak,
                         "frac": alg_flops / (bcd_ms / 1e3) / 1e12 / fp32_peak,
                         "peak_source": "lrfb_ffma_probe measured in this run"}}
    encode_fp32_frac = value * 1e6 / world * ALG_FLOP_PER_PIXEL / 1e12 / fp32_peak

    # ---- quality / parity epilogue: per-image PSNR on device, bpp from host zlib on a sample; NCCL gather --
    dec = compression.decode_records(plan.factors, cfg)
    psnr = compression.psnr_batch(dec, images).float()
    n_s = min(B, POOL)
    host = plan.factors[:n_s].cpu().numpy()
    meta = compression._metadata(torch.uint8, "YCbCr", True, KW["bounds"], KW["patch_size"], lay)
    from lrf_b200 import packing

    t0 = time.perf_counter()
    blobs = list(compression._pool().map(lambda i: packing.pack_qmf_record(host[i], lay, meta), range(n_s)))
    pack_s = time.perf_counter() - t0
    bpp = torch.tensor([len(b) * 8 / (H * W) for b in blobs], dtype=torch.float32, device=dev)
    from lrf_b200.sharding import gather_stats

    stats = torch.stack([bpp, psnr[:n_s]], dim=1)
    stats = gather_stats(stats, n_s * world, rank, world).cpu()  # the only collective (NCCL all_gather)

    if rank == 0:
        line = {
            "metric": "QMF encode Mpixel/s (768x512, 10 iters)", "value": value, "unit": "Mpixel/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "configs[1]: batch of %d synthetic 768x512 RGB uint8 images per GPU, QMF "
                                   "quality 7, 8x8 patches, bounds (-16,15), int8, 10 sweeps" % B,
                       "batch_per_gpu": B, "distinct_images": POOL, "generator": "s_nat(seed=1000+i) SURVEY §8d",
                       "l2": "inputs %.1f GB per GPU >> 126 MB L2, no flush needed" % (images.numel() / 1e9),
                       "sharding": "images split across ranks, no data-path collective"},
            "e2e": {"value": e2e_value, "unit": "Mpixel/s", "h2d_bytes_per_step": int(h_in.numel()) * world,
                    "d2h_bytes_per_step": int(h_out.numel()) * world, "records_equal_device_path": same,
                    "call": "lrfb_qmf_encode_host (C ABI, pinned host buffers)"},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "roofline": roofline,
            "encode_fp32_frac": encode_fp32_frac,
            "encode_hbm_frac": value * 1e6 / world * ALG_BYTE_PER_PIXEL / 1e9 / peaks["hbm_gbs"],
            "quality": {"mean_bpp": float(stats[:, 0].mean()), "mean_psnr_db": float(stats[:, 1].mean()),
                        "images": int(stats.shape[0])},
            "host_pack": {"images": n_s, "seconds": pack_s, "threads": os.cpu_count(),
                          "mpixel_per_s": n_s * H * W / 1e6 / pack_s},
        }
        if not args.no_cpu_baseline and world == 1:  # N = 1 only: the other ranks must not wait on host work
            cores = os.cpu_count() or 1
            v, dt = cpu_reference_mpix(args.cpu_sample, cores)
            line["cpu_baseline"] = {"value": v, "unit": "Mpixel/s", "cores": cores, "kind": "port",
                                    "sample": f"{args.cpu_sample} images of the workload in {dt:.1f} s "
                                              "(oracle port of the reference, zlib included)"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
