#!/usr/bin/env python3
"""Dependent-chain latencies of the GF(2^255-19) building blocks on one warp per SM (ecb_latency_probe)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eccoxide_b200 import Context

NAMES = {0: "field_mul", 1: "field_sqr", 2: "invert_safegcd", 3: "invert_fermat", 4: "block_invert", 5: "shuffle_8_words", 6: "ge_madd", 7: "ge_add_p3", 8: "field_mul2_pair", 9: "invert_warp"}
with Context() as c:
    for v, reps in ((0, 64), (8, 64), (1, 64), (5, 64), (6, 16), (7, 16), (2, 4), (9, 4), (3, 4)):
        cyc, mhz = c.latency_probe(v, 32, reps)
        print(json.dumps({"op": NAMES[v], "threads": 32, "cycles": round(cyc, 1), "sm_mhz": round(mhz), "us": round(cyc / mhz, 3)}), flush=True)
    for th in (32, 128, 256, 448, 480):
        cyc, mhz = c.latency_probe(4, th, 4)
        print(json.dumps({"op": NAMES[4], "threads": th, "cycles": round(cyc, 1), "sm_mhz": round(mhz), "us": round(cyc / mhz, 3)}), flush=True)
    for v, reps in ((0, 64), (2, 4)):
        for th in (128, 448):
            cyc, mhz = c.latency_probe(v, th, reps)
            print(json.dumps({"op": NAMES[v], "threads": th, "cycles": round(cyc, 1), "sm_mhz": round(mhz), "us": round(cyc / mhz, 3)}), flush=True)
