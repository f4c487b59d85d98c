#!/usr/bin/env python
"""Turns ncu outputs (brought back in gpurun_out/) into the small text summaries committed under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md "title"
  python scripts/summarize_ncu.py report   gpurun_out/prof.ncu-rep     profiles/r1_kernel.md   "title"
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def launches(src, dst, title):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(row["Metric Unit"], v)
        name = row["Kernel Name"]
        name = name if len(name) < 110 else name[:107] + "..."
        agg[name][0] += 1
        agg[name][1] += v
        n += 1
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# %s\n\nSource: `ncu --metrics gpu__time_duration.sum --clock-control none` launch list (%d launches, "
                "%.1f ms of kernel time; cold-cache, serialised: compare SHARES, not absolutes).\n\n" % (title, n, tot / 1e3))
        f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write("| `%s` | %d | %.1f | %.1f | %.3f |\n" % (k.replace("|", "\\|"), v[0], v[1], v[1] / v[0], v[1] / tot))
    print("wrote", dst)


def report(src, dst, title):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("# %s\n\nSource: `ncu --set full --clock-control none --import-source on` (%s).\n\n" % (title, src))
        for row in rows[2:]:
            d = dict(zip(hdr, row))
            f.write("## %s\n\n| metric | value | unit |\n|---|---:|---|\n" % d.get("Kernel Name", "?")[:150])
            for i, h in enumerate(hdr):
                stall = "issue_stalled" in h and h.endswith("per_issue_active.ratio")
                if h in KEYS or (stall and row[i] and float(row[i].replace(",", "")) > 0.2):
                    f.write("| %s | %s | %s |\n" % (h, row[i], units[i]))
            f.write("\n")
    print("wrote", dst)


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4])
