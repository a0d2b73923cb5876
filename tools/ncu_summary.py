"""Dev tool: summarise ncu outputs into small text files for profiles/.

    python tools/ncu_summary.py launches <launches.csv>          # per-kernel launch list summary
    python tools/ncu_summary.py full <report.ncu-rep>            # key metrics + stall reasons per kernel
"""
import collections
import csv
import subprocess
import sys


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = (row["Kernel Name"].split("(")[0][:60], row["Grid Size"], row["Block Size"])
        agg.setdefault(k, []).append(float(row["Metric Value"].replace(",", "")) / 1e3)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':60s} {'grid':>18s} {'block':>14s} {'n':>4s} {'avg us':>10s} {'share':>7s}")
    for (k, g, b), v in agg.items():
        print(f"{k:60s} {g:>18s} {b:>14s} {len(v):4d} {sum(v) / len(v):10.1f} {sum(v) / tot * 100:6.1f}%")


WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum",
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("==", name.split("(")[0])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"   {w:80s} {r[i]:>18s} {units[i]}")
        st = [(float(r[i].replace(",", "")), h) for i, h in enumerate(hdr)
              if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and r[i]]
        print("   top stall reasons (warps stalled per issue):",
              ", ".join(f"{h[34:-23]}={v:.2f}" for v, h in sorted(st, reverse=True)[:6]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
