"""Dev tool: samples and executed instructions per CUDA source line of one kernel launch in an ncu report (needs -lineinfo).

    python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [launch index] [top]
"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Line No" and len(r) > 8]
# a launch = a run of tables (one per source file); a new launch starts when a "Kernel Name" row precedes
first = [i for i, r in enumerate(rows) if r and r[0] == "File Path"]
launch_starts = [i for i in first if rows[i][1] == rows[first[0]][1]]  # every launch lists its files in the same order
lo = launch_starts[which]
hi = launch_starts[which + 1] if which + 1 < len(launch_starts) else len(rows)
agg = {}
fname = "?"
for i in range(lo, hi):
    r = rows[i]
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if r[0] in ("Line No", "File Path", "Function Name") or len(r) < 9 or not r[0].isdigit():
        continue
    hdr = rows[max(h for h in heads if h < i)]
    ix = {h: k for k, h in enumerate(hdr)}
    try:
        s, n = int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0)
    except ValueError:
        continue
    key = (fname, int(r[0]))
    a = agg.setdefault(key, [0, 0, r[1].strip()[:100]])
    a[0] += s
    a[1] += n
tot_s = sum(a[0] for a in agg.values()) or 1
tot_n = sum(a[1] for a in agg.values()) or 1
print(f"samples {tot_s}, warp instructions {tot_n}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a[0] / tot_s * 100:5.1f}% smp {a[1] / tot_n * 100:5.1f}% ins  {f}:{ln:4d}  {a[2]}")
