#!/usr/bin/env python
"""Text summary of an `ncu --set full --import-source on` report, as committed under profiles/.

usage: python scripts/ncu_summary.py report.ncu-rep [top_lines]

Prints the headline raw metrics of the (single) captured kernel, the shared-memory data pipe
counters, and the source lines with the most warp-state samples (L long scoreboard, S short
scoreboard, W wait, B barrier).  Needs `ncu` on PATH (it only reads the report).
"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    # instruction pipes: is any of them saturated, or is the kernel latency bound?
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
]
STALLS = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_barrier", "stall_mio",
          "stall_lg", "stall_not_selected", "stall_selected", "stall_math",
          "stall_branch_resolving", "stall_dispatch"]


def ncu(*args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    rows = list(csv.reader(ncu("-i", rep, "--page", "raw", "--csv").splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for name in WANT:
        if name in hdr:
            i = hdr.index(name)
            print(name, units[i], [r[i] for r in data])
    print()
    rows = list(csv.reader(ncu("-i", rep, "--page", "source", "--csv", "--print-source",
                               "cuda,sass").splitlines()))
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
    seen, lines = set(), []
    for hi, end in zip(heads, heads[1:] + [len(rows)]):
        name = rows[hi - 2][1].split("/")[-1] if hi >= 2 and len(rows[hi - 2]) > 1 else "?"
        if name in seen:
            continue
        seen.add(name)
        h = rows[hi]
        isamp = h.index("# Samples")
        cols = [h.index(s) for s in STALLS]
        for r in rows[hi + 1:end]:
            if len(r) < len(h) or not r[0].strip().isdigit() or r[2].strip() not in ("-", ""):
                continue
            try:
                lines.append((name, int(r[0]), r[1].strip()[:70], int(r[isamp]),
                              [int(r[c]) for c in cols]))
            except ValueError:
                pass
    total = sum(x[3] for x in lines)
    sums = [sum(x[4][k] for x in lines) for k in range(len(STALLS))]
    print("# per-line warp-state samples: L long scoreboard, S short scoreboard, W wait, B barrier")
    print("samples", total, " ".join(f"{s[6:]} {v}" for s, v in zip(STALLS, sums)))
    for x in sorted(lines, key=lambda x: -x[3])[:top]:
        st = x[4]
        print(f"{x[0][:8]:8s}{x[1]:5d} {100 * x[3] / max(total, 1):5.1f}% L{st[0]:5d} S{st[1]:5d} "
              f"W{st[2]:5d} B{st[3]:4d}  {x[2]}")


if __name__ == "__main__":
    main()
