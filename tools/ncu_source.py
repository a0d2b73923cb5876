"""Dev tool: stall samples per SASS instruction range of one kernel from an ncu report (source page).

    python tools/ncu_source.py <report.ncu-rep> [bucket] [min_pct]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 100
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hdr_i], rows[hdr_i + 1:]
end = next((i for i, r in enumerate(data) if len(r) < len(hdr)), len(data))  # a report with several launches: the first
data = data[:end]
ix = {h: i for i, h in enumerate(hdr)}
S = ix["# Samples"]
tot = sum(int(r[S] or 0) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("kernel:", rows[hdr_i - 1][1] if hdr_i else "?", "| samples", tot, "| instructions", len(data))
agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stalls}
print("stall reasons:", ", ".join(f"{h[6:]} {v / tot * 100:.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for b in range(0, len(data), bucket):
    s = sum(int(r[S] or 0) for r in data[b:b + bucket])
    ops = {}
    for r in data[b:b + bucket]:
        t = r[ix["Source"]].split()
        if not t:
            continue
        op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        ops[op] = ops.get(op, 0) + 1
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:4]
    print(f"{b:5d} {s / tot * 100:5.1f}%  {top}")
print("hot instructions:")
for i, r in enumerate(data):
    s = int(r[S] or 0)
    if s > tot * min_pct / 100:
        top = max(stalls, key=lambda h: int(r[ix[h]] or 0))
        print(f"{i:5d} {s / tot * 100:5.1f}% {top[6:]:14s} {r[ix['Source']][:90]}")
