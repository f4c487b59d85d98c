#!/usr/bin/env python
"""Condenses `ncu -i rep --page raw --csv` into one markdown row per kernel launch (duration, DRAM bytes and
throughput, occupancy, issue rate, dominant stall reasons).  python scripts/ncu_raw_table.py in.csv out.md "title" """
import csv
import sys


def main(src, dst, title):
    rows = list(csv.reader(open(src)))
    hdr = rows[0]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, k, default=float("nan")):
        i = col.get(k)
        if i is None or r[i] in ("", "n/a"):
            return default
        try:
            return float(r[i].replace(",", ""))
        except ValueError:
            return default
    units = rows[1]
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    out = ["# %s\n" % title, "Source: `%s` (per-launch values; cold-cache, serialised under the profiler).\n" % src,
           "| # | kernel | us | DRAM rd MB | DRAM wr MB | DRAM TB/s | DRAM %peak | L2 %peak | occ % | regs | issue % | tensor % | top stalls (cycles per issue) |",
           "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|"]
    for n, r in enumerate(data):
        name = r[col["Kernel Name"]]
        name = name.replace("cartnet::", "").replace("__nv_bfloat16", "bf16")[:70]
        dur = get(r, "gpu__time_duration.sum")
        du = units[col["gpu__time_duration.sum"]]
        dur_us = dur * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(du, 1.0)

        def byt(k):
            v = get(r, k, 0.0)
            u = units[col[k]] if k in col else "byte"
            return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        rd, wr = byt("dram__bytes_read.sum"), byt("dram__bytes_write.sum")
        st = sorted(((get(r, s, 0.0), s[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for s in stalls), reverse=True)[:3]
        out.append("| %d | `%s` | %.1f | %.0f | %.0f | %.2f | %.0f | %.0f | %.0f | %.0f | %.0f | %.0f | %s |" % (
            n, name, dur_us, rd / 1e6, wr / 1e6, (rd + wr) / 1e12 / (dur_us * 1e-6) if dur_us else 0,
            get(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), get(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
            get(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), get(r, "launch__registers_per_thread"),
            get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in col else get(r, "sm__inst_executed_pipe_tensor.sum", 0.0) * 0,
            ", ".join("%s %.1f" % (b, a) for a, b in st)))
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main(*sys.argv[1:4])
