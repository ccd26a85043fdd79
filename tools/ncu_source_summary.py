#!/usr/bin/env python3
"""Per-opcode view of an `ncu --page source --csv` dump (tools/gpu_profile_p256.sh writes one): share of the executed warp
instructions, share of the warp-stall samples and the three top stall reasons of every opcode.
    python tools/ncu_source_summary.py gpurun_out/TAG_source.csv > profiles/TAG_source_page_summary.txt"""
import collections, csv, re, sys

rows = list(csv.reader(open(sys.argv[1])))
print(rows[0][1] if rows and len(rows[0]) > 1 else "")
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ix = {k: i for i, k in enumerate(h)}
data = rows[hi + 1:]


def f(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except Exception:
        return 0.0


tot_exec = sum(f(r, "Instructions Executed") for r in data)
tot_samp = sum(f(r, "# Samples") for r in data)
print("%d SASS instructions, %.4g warp instructions executed, %d stall samples" % (len(data), tot_exec, tot_samp))
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(f(r, k) for r in data) for k in stalls}
print("stall samples by reason: " + ", ".join("%s %.1f %%" % (k.replace("stall_", ""), 100 * v / tot_samp) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
byop = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
    op = m.group(2) if m else "?"
    op = ".".join(op.split(".")[:3])
    byop[op][0] += f(r, "Instructions Executed")
    byop[op][1] += f(r, "# Samples")
    for k in stalls:
        byop[op][2][k] += f(r, k)
print("%-22s %8s %8s  %s" % ("opcode", "exec %", "samples %", "top stall reasons of its samples"))
for op, (e, s, c) in sorted(byop.items(), key=lambda kv: -kv[1][1])[:20]:
    top = ", ".join("%s %.0f %%" % (k.replace("stall_", ""), 100 * v / max(s, 1)) for k, v in c.most_common(3))
    print("%-22s %7.1f%% %8.1f%%  %s" % (op, 100 * e / tot_exec, 100 * s / tot_samp, top))
