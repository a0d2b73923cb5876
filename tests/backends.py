"""Two ways to drive the same C ABI from the tests: the nvcc-built product library on a CUDA device
(GpuBackend, through lrf_b200's ctypes binding) and the g++/-DLRFB_SIM build of the same sources
(SimBackend, CPU SIMT shim, kernel-logic checks only)."""
from __future__ import annotations

import ctypes as C
import json

import numpy as np
import torch

from lrf_b200 import _cabi
from oracle import qmf_port as port


class SimBackend:
    name = "sim"

    def __init__(self):
        from cpu_sim import simlib

        self.s = simlib

    def encode(self, images, cfg, inits=None, sign_flip=None, stop_after=0):
        fac, ws, m, L = self.s.encode(images, cfg, inits, sign_flip, stop_after)
        B = images.shape[0]

        def view(name, pl):
            shp = {"x": (B, L.rows[pl], L.cols), "u": (B, L.rows[pl], L.rank[pl]), "v": (B, L.cols, L.rank[pl])}[name]
            return self.s.ws_view(ws, getattr(m, name)[pl], np.float32, shp)

        return fac, view, L

    def decode(self, fac, cfg):
        return self.s.decode(fac, cfg)

    def sse(self, a, b):
        return self.s.sse(a, b)


class GpuBackend:
    name = "gpu"

    def encode(self, images, cfg, inits=None, sign_flip=None, stop_after=0):
        from lrf_b200 import compression

        dev = torch.device("cuda", 0)
        lay = _cabi.QmfLayout()
        _cabi.check(_cabi.lib().lrfb_qmf_layout_query(C.byref(cfg), C.byref(lay)), "layout")
        B = images.shape[0]
        plan = compression.EncodePlan(cfg, lay, B, dev)
        dbg = _cabi.QmfDebug()
        dbg.stop_after = stop_after
        keep = []
        if inits is not None:
            for pl, (u0, v0) in enumerate(inits):
                tu = torch.from_numpy(np.ascontiguousarray(u0, np.float32)).to(dev)
                tv = torch.from_numpy(np.ascontiguousarray(v0, np.float32)).to(dev)
                keep += [tu, tv]
                dbg.d_init_u[pl], dbg.d_init_v[pl] = tu.data_ptr(), tv.data_ptr()
        if sign_flip is not None:
            for pl, s in enumerate(sign_flip):
                ts = torch.from_numpy(np.ascontiguousarray(s, np.int32)).to(dev)
                keep.append(ts)
                dbg.d_sign_flip[pl] = ts.data_ptr()
        plan.workspace.zero_()
        fac = plan.run(torch.from_numpy(np.ascontiguousarray(images)).to(dev), dbg)
        torch.cuda.synchronize()
        self._plan = plan
        return fac.cpu().numpy(), (lambda name, pl: plan.view(name, pl).cpu().numpy()), lay

    def decode(self, fac, cfg):
        from lrf_b200 import compression

        out = compression.decode_records(torch.from_numpy(np.ascontiguousarray(fac)).cuda(), cfg)
        return out.cpu().numpy()

    def sse(self, a, b):
        from lrf_b200 import compression

        return compression.sse_u8(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu().numpy().astype(np.uint64)


def split_record(rec, L):
    out = []
    for pl in range(L.n_planes):
        r = L.rank[pl]
        out.append(rec[L.u_offset[pl]: L.u_offset[pl] + L.rows[pl] * r].reshape(r, L.rows[pl]).T)
        out.append(rec[L.v_offset[pl]: L.v_offset[pl] + L.cols * r].reshape(r, L.cols).T)
    return out


def parse_golden(blob):
    meta_b, body = port.separate_bytes(blob, 2)
    meta = json.loads(meta_b)
    ycbcr = meta["color space"] == "YCbCr"
    fs = [port.decode_matrix(b).numpy() for b in port.separate_bytes(body, 6 if ycbcr else 2)]
    return meta, fs


def config_for(img, kw, ranks):
    in_dt = _cabi.LRFB_U8 if img.dtype == torch.uint8 else _cabi.LRFB_F32
    return _cabi.make_config(img.shape[-2], img.shape[-1], kw["patch_size"], kw["color_space"], in_dt,
                             kw["scale_factor"], ranks, kw["bounds"], kw["num_iters"])


def reference_planes(img, kw):
    if kw["color_space"] == "YCbCr":
        return [p[0] for p in port.qmf_planes(img, kw["scale_factor"], kw["patch_size"])]
    return [port.patchify(port.pad_image(img.float(), kw["patch_size"]), kw["patch_size"])]


def lapack_sign_flips(backend, img, cfg, ref_v0):
    """Run only the SVD init, compare each v0 column with the reference's and return ±1 per column
    (LAPACK's signs for components >= 2 follow no rule, the harness aligns them)."""
    _, view, L = backend.encode(img.numpy()[None], cfg, stop_after=2)
    flips = []
    for pl in range(L.n_planes):
        v0 = view("v", pl)[0]
        s = np.sign((v0 * ref_v0[pl]).sum(0)).astype(np.int32)
        s[s == 0] = 1
        flips.append(s[None])
    return flips
