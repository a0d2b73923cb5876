"""Build liblrfb.so in-tree with nvcc for sm_100a:  python -m lrf_b200.build [--force]"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OUT = os.environ.get("LRFB_OUT") or os.path.join(CSRC, "build", "liblrfb.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # every multiply/add that must stay separately rounded is written as such; FMAs are explicit
    "-fmad=false", "-diag-suppress", "128,177",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared",
]


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = glob.glob(os.path.join(CSRC, "*.cu*")) + [os.path.join(os.path.dirname(_HERE), "include", "lrfb.h")]
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(s) for s in srcs):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, os.path.join(CSRC, "lrfb_api.cu"), "-o", OUT, "-lz"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    if os.environ.get("LRFB_DEV"):  # development build: the library then honours its LRFB_* environment knobs
        cmd.insert(1, "-DLRFB_DEV")
    for d in filter(None, os.environ.get("LRFB_DEFS", "").split(",")):  # experiment builds: extra -D switches
        cmd.insert(1, "-D" + d)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
