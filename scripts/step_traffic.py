#!/usr/bin/env python
"""One training step out of an ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per
launch): per-kernel launches / time / DRAM bytes inside ONE step, the step totals, and the roofline.traffic entry that
bench.py reads from profiles/r2_traffic.json.

  python scripts/step_traffic.py gpurun_out/r2_launches_x3.csv profiles/r2_step_kernels_bf16x3.md adp_train/bf16x3/batch64 "title"

Steps are delimited by the fused Adam launches (torch's multi_tensor_apply kernel with FusedOptimizerTensorListMetadata);
the LAST complete step before the instrumented passes is used (warm caches, steady-state allocator)."""
import collections
import csv
import json
import os
import sys


def load(src):
    lines = [l for l in open(src) if not l.startswith("==")]
    launches = collections.OrderedDict()
    for row in csv.DictReader(lines):
        lid = int(row["ID"])
        d = launches.setdefault(lid, {"name": row["Kernel Name"]})
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        m = row["Metric Name"]
        if m == "gpu__time_duration.sum":
            d["us"] = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(unit, v)
        elif m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            d[m] = {"byte": v, "Kbyte": v * 1e3, "Mbyte": v * 1e6, "Gbyte": v * 1e9}.get(unit, v)
    return list(launches.values())


def main(src, dst, key, title):
    L = load(src)
    adam = [i for i, d in enumerate(L) if "FusedOptimizer" in d["name"]]
    # consecutive Adam launches belong to one optimiser step: keep the last index of every group
    ends = [a for k, a in enumerate(adam) if k + 1 == len(adam) or adam[k + 1] != a + 1]
    assert len(ends) >= 5, "not enough optimiser steps in the capture"
    # bench.py --steps 2 --warmup 3: steps 0-2 warm-up, 3-4 timed ("value" loop) -> the window between ends[3] and ends[4]
    a, b = ends[3] + 1, ends[4] + 1
    step = L[a:b]
    agg = collections.OrderedDict()
    for d in step:
        name = d["name"].replace("cartnet::", "")
        name = name if len(name) < 100 else name[:97] + "..."
        x = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
        x[0] += 1
        x[1] += d.get("us", 0.0)
        x[2] += d.get("dram__bytes_read.sum", 0.0)
        x[3] += d.get("dram__bytes_write.sum", 0.0)
    tot_us = sum(v[1] for v in agg.values())
    tot_rd = sum(v[2] for v in agg.values())
    tot_wr = sum(v[3] for v in agg.values())
    mine = sum(v[1] for k, v in agg.items() if not k.startswith("void native") and not k.startswith("void at") and "nccl" not in k)
    with open(dst, "w") as f:
        f.write("# %s\n\n" % title)
        f.write("Source: `%s` -- `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none` over "
                "`python bench.py --steps 2 --warmup 3`; ONE timed step (launches %d..%d of %d). Per-launch times under ncu are serialised "
                "and cold-cache: compare SHARES, not absolutes.\n\n" % (os.path.basename(src), a, b - 1, len(L)))
        f.write("Step totals: **%d launches, %.2f ms of kernel time, DRAM read %.2f GB + write %.2f GB = %.2f GB** "
                "(this library's kernels: %.1f %% of the kernel time).\n\n" % (len(step), tot_us / 1e3, tot_rd / 1e9, tot_wr / 1e9, (tot_rd + tot_wr) / 1e9, 100 * mine / tot_us))
        f.write("| kernel | launches | total us | avg us | share | DRAM read GB | DRAM write GB | GB/s |\n|---|---:|---:|---:|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
            gbs = (v[2] + v[3]) / (v[1] * 1e-6) / 1e9 if v[1] > 0 else 0.0
            f.write("| `%s` | %d | %.1f | %.1f | %.3f | %.3f | %.3f | %.0f |\n" % (k, v[0], v[1], v[1] / v[0], v[1] / tot_us, v[2] / 1e9, v[3] / 1e9, gbs))
    tp = os.path.join(os.path.dirname(os.path.abspath(dst)), "r2_traffic.json")
    t = json.load(open(tp)) if os.path.isfile(tp) else {}
    t[key] = int(tot_rd + tot_wr)
    json.dump(t, open(tp, "w"), indent=1, sort_keys=True)
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    main(*sys.argv[1:5])
