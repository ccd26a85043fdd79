#!/usr/bin/env python3
"""Summarise an ncu report (--set full) into the handful of counters the design argues from.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_x_ncu_summary.txt
    python tools/ncu_summary.py prof.ncu-rep --json WORKLOAD "launch description" SOURCE >> fragment.jsonl
        one JSON object {WORKLOAD: {...}} for profiles/ncu_static.json (merged by tools/merge_ncu_static.py)

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


def as_json(rows, idx, units, workload, launch, source):
    import json

    r = rows[0]

    def num(k):
        try:
            v = float(r[idx[k]].replace(",", ""))
        except (KeyError, ValueError):
            return None
        u = units[idx[k]]
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}
        if k.startswith("dram__bytes"):
            return int(v * scale.get(u, 1.0))
        if k == "gpu__time_duration.sum":
            return v * scale.get(u, 1.0)
        return v

    st = []
    for h, i in idx.items():
        if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(r[i]), h[len(STALL):].replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    st.sort(reverse=True)
    fm, alu = num("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"), num("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active")
    top = st[0][1] if st else None
    # what bounds the kernel, from the top stall: the multiplier pipe itself, or latencies it cannot hide
    bound = "imad" if top in ("math_pipe_throttle",) or (fm or 0) >= 70 else ("latency (%s)" % top if top else "imad")
    out = {workload: {
        "kernel": r[idx["Kernel Name"]].split("(")[0], "launch": launch, "duration_ms_under_ncu": num("gpu__time_duration.sum"),
        "grid": num("launch__grid_size"), "block": num("launch__block_size"), "regs": num("launch__registers_per_thread"),
        "occupancy_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "fmaheavy_pct": fm, "alu_pct": alu, "fp64_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "dram_read_bytes": num("dram__bytes_read.sum"), "dram_write_bytes": num("dram__bytes_write.sum"),
        "dram_throughput_pct": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), "l2_hit_pct": num("lts__t_sector_hit_rate.pct"),
        "local_loads": num("smsp__sass_inst_executed_op_local_ld.sum"), "local_stores": num("smsp__sass_inst_executed_op_local_st.sum"),
        "top_stall": top, "stalls": {n: round(v, 2) for v, n in st[:4]}, "bound": bound, "source": source}}
    print(json.dumps(out))


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    if len(sys.argv) > 2 and sys.argv[2] == "--json":
        return as_json(data, idx, units, sys.argv[3], sys.argv[4], sys.argv[5])
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
