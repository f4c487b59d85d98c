#!/usr/bin/env python
"""Key metrics of an `ncu --set full` capture (raw-page CSV) as one markdown table, one column per kernel launch.
python scripts/ncu_full_summary.py in_raw.csv out.md "title" """
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU data-pipe wavefronts %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "LSU wavefronts, shared"),
    ("l1tex__data_pipe_lsu_wavefronts.sum", "LSU wavefronts, total"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (HMMA) active %"),
    ("sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "tcgen05 issue pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard (cyc/issue)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall sleeping"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
]


def main(src, dst, title):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    col = {h: i for i, h in enumerate(hdr)}
    names = [r[col["Kernel Name"]].replace("cartnet::", "").replace("__nv_bfloat16", "bf16")[:48] for r in data]
    out = ["# %s\n" % title, "Source: `%s` (`ncu --set full --clock-control none --import-source on`, raw page).\n" % src,
           "| metric | " + " | ".join("`%s`" % n for n in names) + " |", "|---|" + "---:|" * len(names)]
    for k, label in KEYS:
        if k not in col:
            continue
        u = units[col[k]]
        out.append("| %s%s | " % (label, (" [%s]" % u) if u and u != "%" else "") + " | ".join(r[col[k]] for r in data) + " |")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main(*sys.argv[1:4])
