"""Per-kernel counts of the SASS opcodes that prove which hardware units a kernel uses (B200_PROFILING.md table).

    python tools/sass_opcodes.py [lib.so] > profiles/r2_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "lrf_b200", "csrc", "build", "liblrfb.so")
OPS = [("UTC*MMA (tcgen05.mma)", r"\bUTC\w*MMA"), ("LDTM (tcgen05.ld)", r"\bLDTM"), ("STTM (tcgen05.st)", r"\bSTTM"),
       ("UTMALDG (TMA load)", r"\bUTMALDG"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("FFMA2", r"\bFFMA2"), ("FFMA", r"\bFFMA\b"),
       ("DFMA", r"\bDFMA"), ("DMMA", r"\bDMMA"), ("IDP.4A", r"\bIDP\.4A"), ("LDGSTS (cp.async)", r"\bLDGSTS"),
       ("UCGABAR (cluster barrier)", r"\bUCGABAR"), ("C?REDUX", r"\bC?REDUX"), ("MATCH (match.any)", r"\bMATCH"),
       ("VOTE (ballot)", r"\bVOTE"), ("ATOMS (shared atomics)", r"\bATOMS")]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name).replace("void ", "")
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for label, pat in OPS:
        if re.search(pat, line):
            counts[cur][label] += 1
print(f"SASS opcode counts per kernel of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a)")
print(f"{'kernel':58s} " + " ".join(f"{l.split(' ')[0]:>8s}" for l, _ in OPS))
for k, c in counts.items():
    if not k.startswith("lrfb::"):
        continue
    print(f"{k[:58]:58s} " + " ".join(f"{c[l]:8d}" for l, _ in OPS))
print("\nlegend: " + "; ".join(l for l, _ in OPS))
