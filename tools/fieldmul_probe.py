#!/usr/bin/env python3
"""GF(2^255-19) products per second with the warps of a kernel split between the integer-multiply pipe
(fe25519.cuh) and the FP64 pipe (fe43.cuh): ecb_fieldmul_probe.  One JSON line per (split, occupancy)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eccoxide_b200 import Context

with Context() as c:
    base = {}
    for bps in (4, 8, 12, 16):
        for num, den in ((0, 1), (1, 1), (1, 2), (1, 4), (3, 4), (1, 3), (2, 3)):
            v, chk = c.fieldmul_probe(num, den, bps, 512)
            if (num, den) == (0, 1):
                base[bps] = v
            print(json.dumps({"fp64_warps": "%d/%d" % (num, den), "blocks_per_sm": bps, "warps_per_sm": bps * 4, "gmul_per_s": round(v / 1e9, 2),
                              "vs_integer_only": round(v / base[bps], 3), "checksum": chk}), flush=True)
