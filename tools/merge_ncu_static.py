#!/usr/bin/env python3
"""Merge the JSON fragments written by `tools/ncu_summary.py REP --json ...` into profiles/ncu_static.json
(bench.py copies these per-kernel counters into roofline.traffic / hbm_frac / pipe_util_ncu).

    python tools/merge_ncu_static.py gpurun_out/r02_k_static.jsonl
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(ROOT, "profiles", "ncu_static.json")
cur = json.load(open(path))
for ln in open(sys.argv[1]):
    ln = ln.strip()
    if ln.startswith("{"):
        for k, v in json.loads(ln).items():
            cur[k] = v
json.dump(cur, open(path, "w"), indent=1)
print("merged into", path, ":", ", ".join(k for k in cur if not k.startswith("_")))
