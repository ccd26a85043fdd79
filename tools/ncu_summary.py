#!/usr/bin/env python3
"""Summarise an ncu report (--set full) into the handful of counters the design argues from.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_x_ncu_summary.txt

One block per distinct kernel (first captured launch of each): duration, registers, occupancy, the
integer-multiply pipe (fmaheavy: IMAD / IMAD.WIDE issue there), the alu pipe, issue-slot use, the
top stall reasons and DRAM bytes.
"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "blocks/SM (register limit)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fmaheavy pipe (IMAD) cycles active % of peak"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipes (heavy+lite) cycles active %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu pipe cycles active %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe cycles active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots used %"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC (per SM)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % (max of pipes)"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("smsp__sass_inst_executed_op_local_ld.sum", "local loads"),
    ("smsp__sass_inst_executed_op_local_st.sum", "local stores"),
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = set()
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        print("== %s" % name)
        for k, label in KEYS:
            if k in idx and r[idx[k]] not in ("", "n/a"):
                print("  %-52s %s %s" % (label, r[idx[k]], units[idx[k]]))
        st = []
        for h, i in idx.items():
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(r[i]), h[len(STALL):].replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print("  stall reasons (warps per issue): " + ", ".join("%s %.2f" % (n, v) for v, n in st[:6]))
        print()


if __name__ == "__main__":
    main()
