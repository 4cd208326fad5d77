#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

  ncu_summary.py launches <launches.csv> <out.md> [title]    per-kernel totals / shares of a `--metrics gpu__time_duration.sum` pass
  ncu_summary.py report   <x.ncu-rep>   <out.md> [title]     key counters of every profiled launch of a `--set full` capture
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "smsp__cycles_active.avg",
]


def launches(path, out, title):
    rows = []
    with open(path, newline="") as f:
        text = "".join(line for line in f if line.startswith('"'))
    for r in csv.DictReader(io.StringIO(text)):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"]), r["Grid Size"], r["Block Size"]))
    agg = OrderedDict()
    for name, ns, grid, block in rows:
        short = name.split("(")[0].replace("void ", "")
        a = agg.setdefault(short, [0, 0.0, 1e30, 0.0])
        a[0] += 1
        a[1] += ns
        a[2] = min(a[2], ns)
        a[3] = max(a[3], ns)
    total = sum(a[1] for a in agg.values()) or 1.0
    with open(out, "w") as f:
        f.write(f"# {title}\n\nSource: `{path}` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised launches: read SHARES).\n\n")
        f.write(f"{len(rows)} launches, {total / 1e6:.3f} ms summed device time.\n\n| kernel | launches | total us | share | min us | max us | avg us |\n|---|---:|---:|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1] / 1e3:.1f} | {100 * a[1] / total:.1f}% | {a[2] / 1e3:.2f} | {a[3] / 1e3:.2f} | {a[1] / a[0] / 1e3:.2f} |\n")


def report(path, out, title):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [(k, hdr.index(k)) for k in KEYS if k in hdr]
    kn = hdr.index("Kernel Name")
    with open(out, "w") as f:
        f.write(f"# {title}\n\nSource: `{path}` (ncu --set full --clock-control none --import-source on), read with `ncu -i ... --page raw --csv`.\n\n")
        f.write("| launch | kernel | " + " | ".join(f"{k} [{units[i]}]" for k, i in idx) + " |\n|---|---|" + "---:|" * len(idx) + "\n")
        for n, r in enumerate(data):
            f.write(f"| {n} | `{r[kn].split('(')[0].replace('void ', '')}` | " + " | ".join(r[i] for _, i in idx) + " |\n")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else src
    (launches if mode == "launches" else report)(src, dst, title)
